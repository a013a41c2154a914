// dcn_tc.cu — DCNv2 forward: bilinear sampling fused into a tcgen05 GEMM.
//
// One CTA = a TH x TW (8 x 16) tile of 128 output pixels = the M of a 128 x Cout UMMA (one TMEM
// lane per pixel). 384 threads: thread (pixel p, row r) handles taps r*TPR + s. The input is first
// re-laid out once per call into the group-blocked form [b][group][y][x][8 channels]
// (nchw_to_blocked, 2 x 16.8 MB of L2-resident traffic at the benchmark shape), so that the 8
// channels of a bilinear corner are ONE 32-byte sector: a corner costs two 128-bit loads instead of
// eight scalar loads from eight planes, and neighbouring pixels share cache lines.
// Per deformable group and stage s, each thread reads offset/mask of its tap once, gathers the 4
// corners, applies the bilinear weights and the mask, splits into TF32 hi + lo and writes the
// values straight into the K-major A operand in shared memory (K order of a stage: row r major,
// channel minor) — the reference's 151 MB column buffer (dcn_v2_cuda.cu:68) lives only there. The
// matching pre-split weight image arrives by one bulk (TMA) copy; one thread issues the 3xTF32
// tcgen05.mma chain into TMEM and commits to the stage buffer's mbarrier. Operand stages are double
// buffered and two CTAs share an SM, so sampling overlaps the MMAs.
// Epilogue: tcgen05.ld (lane = pixel), sum of the split accumulators, + bias, NCHW stores.
//
// History (profiles/): gathering from the NCHW planes was bound by L1 misses / LSU wavefronts
// (455 wavefronts per pixel); staging an input box in shared memory did not pay either (extra
// phases and barriers, latency-bound at 12 warps per CTA). The blocked layout cuts load
// instructions 4x and L2 sectors ~8x.
//
// fp32 parity: products are 3xTF32 (umma.cuh); the hi*hi chain is spread over several TMEM
// accumulators because the tensor core's fp32 accumulate rounds toward zero.
#include "dcn_common.cuh"
#include "umma.cuh"

#include <algorithm>
#include <climits>

namespace ebfi_dcn {

namespace {

using ebfi::ceil_div;

constexpr int TM = 128;          // pixels per CTA tile
constexpr int TH = 8, TW = 16;   // tile shape
constexpr int NR = 3;            // thread rows: thread (p, r)
constexpr int NSAMP = TM * NR;   // 384 sampler threads
constexpr int NTHR = NSAMP + 32; // + one controller warp (weight prefetch, MMA issue)
constexpr int TMEM_COLS = 256;

struct FwdPlan {
    int cs, ncs;         // channels per chunk (4 or 8), chunks per deformable group
    int TPR;             // taps per thread row = stages per chunk: tap = r * TPR + s
    int Ks, Ksp, kch;    // K of a stage (NR * cs), padded to 8, 16-byte chunks per operand row
    int nacc;            // hi*hi accumulators (+1 for the cross terms)
    int tiles_x, tiles_y;
    int a_bytes, b_bytes, smem;
};


template <int TPR, bool PACKED>
__global__ void __launch_bounds__(NTHR, 2)
dcn_fwd_tc_kernel(const float *__restrict__ in_blk,
                  const float *__restrict__ bias, const float *__restrict__ offset,
                  const float *__restrict__ mask, float *__restrict__ output,
                  const float *__restrict__ wimg, DcnDims d, FwdPlan pl)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *opnd = smem_raw;                                        // [2 buffers][a_hi | a_lo]
    unsigned char *wring = smem_raw + 4 * pl.a_bytes;                      // [4 slots][b_hi | b_lo], filled two stages ahead
    __shared__ __align__(8) uint64_t bar_free[2];   // MMAs that read A buffer i are complete (tcgen05.commit)
    __shared__ __align__(8) uint64_t bar_full[2];   // all 12 sampler warps have written A buffer i
    __shared__ __align__(8) uint64_t bar_w[4];      // weight image in ring slot i has landed
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total_stages = d.dg * pl.ncs * TPR;
    const uint32_t sbo = (uint32_t)pl.kch * 128u;

    if (warp == 0) umma::tmem_alloc<TMEM_COLS>(&tmem_slot);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { umma::mbar_init(&bar_free[i], 1); umma::mbar_init(&bar_full[i], NSAMP / 32); }
        for (int i = 0; i < 4; ++i) umma::mbar_init(&bar_w[i], 1);
        umma::mbar_fence_init();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;

    if (warp == NSAMP / 32) {
        // ================= controller warp: weight prefetch + MMA issue, one elected lane =================
        // Keeping this off the sampler warps takes the serial issue of 9 MMAs + descriptors per stage out
        // of their critical path (with it on thread 0, a stage cost T_sample + T_issue; now max of the two).
        // All lanes run the loop with warp-uniform values (uniform datapath); one elected lane issues.
        const bool leader = umma::elect_one();
        const uint32_t idesc = umma::instr_desc_tf32(TM, d.Co, 0, 0);
        const uint32_t wb = 2u * (uint32_t)pl.b_bytes;
        const int ksteps = pl.Ksp / 8;
        int step = 0;
        for (int n = 0; n < total_stages; ++n) {
            const int bi = n & 1;
            if (n == 0) {
                if (leader)
                    for (int ns = 0; ns < 3 && ns < total_stages; ++ns) {
                        umma::mbar_expect_tx(&bar_w[ns & 3], wb);
                        umma::bulk_g2s(wring + (ns & 3) * wb, wimg + (size_t)ns * (wb / 4), wb, &bar_w[ns & 3]);
                    }
            } else if (n + 2 < total_stages) {
                // ring slot (n+2) & 3 was read by stage n-2: wait for its MMAs, then refill it
                if (n >= 2) umma::mbar_wait(&bar_free[bi], (uint32_t)(((n - 2) >> 1) & 1));
                const int ns = n + 2;
                if (leader) {
                    umma::mbar_expect_tx(&bar_w[ns & 3], wb);
                    umma::bulk_g2s(wring + (ns & 3) * wb, wimg + (size_t)ns * (wb / 4), wb, &bar_w[ns & 3]);
                }
            }
            umma::mbar_wait(&bar_w[n & 3], (uint32_t)((n >> 2) & 1));
            umma::mbar_wait(&bar_full[bi], (uint32_t)((n >> 1) & 1));
            umma::fence_after_sync();
            const uint32_t ah = umma::smem_u32(opnd + bi * 2 * pl.a_bytes), bh = umma::smem_u32(wring + (n & 3) * wb);
            // descriptors of K-step ks = descriptor of step 0 + ks * 256 bytes (start-address field, 16-byte units)
            uint64_t dah = umma::smem_desc(ah, 128, sbo), dal = umma::smem_desc(ah + (uint32_t)pl.a_bytes, 128, sbo);
            uint64_t dbh = umma::smem_desc(bh, 128, sbo), dbl = umma::smem_desc(bh + (uint32_t)pl.b_bytes, 128, sbo);
            for (int ks = 0; ks < ksteps; ++ks, ++step, dah += 16, dal += 16, dbh += 16, dbl += 16) {
                const uint32_t d_x = tmem + pl.nacc * d.Co, d_h = tmem + (step % pl.nacc) * d.Co;
                if (leader) {
                    umma::mma_tf32(d_x, dal, dbh, idesc, step > 0);
                    umma::mma_tf32(d_x, dah, dbl, idesc, true);
                    umma::mma_tf32(d_h, dah, dbh, idesc, step >= pl.nacc);
                }
            }
            if (leader) umma::commit(&bar_free[bi]);
            __syncwarp();
        }
    } else {
        // ================= sampler warps: thread (pixel p, row r) =================
        const int p = tid % TM, r = tid / TM;
        int tile = blockIdx.x;
        const int tx0 = (tile % pl.tiles_x) * TW; tile /= pl.tiles_x;
        const int ty0 = (tile % pl.tiles_y) * TH;
        const int b = tile / pl.tiles_y;
        const int ho = ty0 + p / TW, wo = tx0 + p % TW;
        const bool valid = ho < d.Ho && wo < d.Wo;
        const int pix = ho * d.Wo + wo;
        const size_t plane = (size_t)d.Ho * d.Wo, in_plane = (size_t)d.H * d.W;
        const int a_row = (p >> 3) * (pl.kch * 32) + (p & 7) * 4;         // floats, row of pixel p in an A image
        const unsigned uplane = (unsigned)plane, upix = (unsigned)pix;

        // undeformed sampling positions of this thread's TPR taps (group independent: no divisions in the loops)
        float by[TPR], bx[TPR];
        bool tv[TPR];
#pragma unroll
        for (int s = 0; s < TPR; ++s) {
            const int t = r * TPR + s, i = t / d.kw, j = t - i * d.kw;
            tv[s] = valid && t < d.KK;
            by[s] = (float)(ho * d.sh - d.ph + i * d.dh);
            bx[s] = (float)(wo * d.sw - d.pw + j * d.dw);
        }
        // Offsets / masks are read exactly once, i.e. every read misses to DRAM (~4 k cycles with the
        // dependent corner loads behind it): they are fetched one deformable group AHEAD into registers.
        float ndy[TPR], ndx[TPR], nm[TPR];
        auto fetch_group = [&](int g) {
            const float *off_bg = off_ptr(d, offset, b, g, plane);
            const float *mask_bg = mask_ptr(d, mask, b, g, plane);
#pragma unroll
            for (int s = 0; s < TPR; ++s) {
                ndy[s] = 0.f; ndx[s] = 0.f; nm[s] = 0.f;
                if (tv[s]) tap_read(off_bg, mask_bg, uplane, (unsigned)(r * TPR + s), upix, ndy[s], ndx[s], nm[s]);
            }
        };
        fetch_group(0);
        float asum = 0.f;                          // sum |offset| of this thread's taps (packed entry)

        int n = 0;
        for (int g = 0; g < d.dg; ++g) {
            // sampling positions of this group's taps; (-2,-2) = outside the window
            float sy[TPR], sx[TPR], sm[TPR];
#pragma unroll
            for (int s = 0; s < TPR; ++s) {
                sy[s] = tv[s] ? by[s] + ndy[s] : -2.f;
                sx[s] = tv[s] ? bx[s] + ndx[s] : -2.f;
                sm[s] = mask_act_t<PACKED>(nm[s]);
                if (PACKED) asum += fabsf(ndy[s]) + fabsf(ndx[s]);       // invalid taps hold zeros
            }
            if (g + 1 < d.dg) fetch_group(g + 1);
            for (int ci = 0; ci < pl.ncs; ++ci) {
                // blocked input of this chunk: [y][x][8 channels]
                const float *ibf = in_blk + (((size_t)b * d.dg + g) * pl.ncs + ci) * in_plane * 8;
#pragma unroll 1
                for (int s = 0; s < TPR; ++s, ++n) {
                    const int bi = n & 1;
                    float *a_hi = reinterpret_cast<float *>(opnd + bi * 2 * pl.a_bytes);
                    float *a_lo = reinterpret_cast<float *>(opnd + bi * 2 * pl.a_bytes + pl.a_bytes);
                    // this thread's tap of the stage -> 8 values of row p, columns r*8 .. r*8+7
                    float y = sy[0], x = sx[0], m = sm[0];               // register select, no local-memory indexing
#pragma unroll
                    for (int q = 1; q < TPR; ++q)
                        if (s == q) { y = sy[q]; x = sx[q]; m = sm[q]; }
                    const Tap tp = make_tap(y, x, d.H, d.W);
                    const float w1 = tp.hy * tp.hx, w2 = tp.hy * tp.lx, w3 = tp.ly * tp.hx, w4 = tp.ly * tp.lx;
                    const f8 a = ldg_f8(ibf + (size_t)tp.i00 * 8, tp.c00), bq = ldg_f8(ibf + (size_t)tp.i01 * 8, tp.c01);
                    const f8 c = ldg_f8(ibf + (size_t)tp.i10 * 8, tp.c10), e8 = ldg_f8(ibf + (size_t)tp.i11 * 8, tp.c11);
                    float hi[8], lo[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        umma::split_tf32((w1 * a.v[j] + w2 * bq.v[j] + w3 * c.v[j] + w4 * e8.v[j]) * m, hi[j], lo[j]);
                    if (n >= 2) umma::mbar_wait(&bar_free[bi], (uint32_t)(((n - 2) >> 1) & 1));   // MMAs of stage n-2 done
                    const int off = a_row + ((r * pl.cs) >> 2) * 32;
                    *reinterpret_cast<float4 *>(a_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4 *>(a_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    *reinterpret_cast<float4 *>(a_hi + off + 32) = make_float4(hi[4], hi[5], hi[6], hi[7]);
                    *reinterpret_cast<float4 *>(a_lo + off + 32) = make_float4(lo[4], lo[5], lo[6], lo[7]);
                    if (r == 0 && pl.Ksp > pl.Ks) {                        // zero the K padding of this row
                        const int offp = a_row + (pl.Ks >> 2) * 32;
                        *reinterpret_cast<float4 *>(a_hi + offp) = make_float4(0.f, 0.f, 0.f, 0.f);
                        *reinterpret_cast<float4 *>(a_lo + offp) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    umma::fence_smem_to_async();
                    __syncwarp();
                    if (lane == 0) umma::mbar_arrive(&bar_full[bi]);
                }
            }
        }
        if (PACKED && d.abs_sum) warp_atomic_sum(d.abs_sum, asum);
        // ---- all MMAs complete (commits complete in order: the last one covers everything)
        umma::mbar_wait(&bar_free[(n - 1) & 1], (uint32_t)(((n - 1) >> 1) & 1));
        umma::fence_after_sync();

        // ---- epilogue: thread (p, r) takes every NR-th block of 8 output channels of pixel p
        const int total_steps = n * (pl.Ksp / 8), nused = min(pl.nacc, total_steps);
        const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;
        for (int cb = r * 8; cb < d.Co; cb += NR * 8) {
            float v[8], u[8];
            umma::tmem_ld8(umma::tmem_addr(tmem, lane_base, pl.nacc * d.Co + cb), v);
            umma::tmem_ld_wait();
            for (int j = 0; j < nused; ++j) {
                umma::tmem_ld8(umma::tmem_addr(tmem, lane_base, j * d.Co + cb), u);
                umma::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] += u[i];
            }
            if (valid) {
#pragma unroll
                for (int i = 0; i < 8; ++i) output[((size_t)b * d.Co + cb + i) * plane + pix] = v[i] + __ldg(bias + cb + i);
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<TMEM_COLS>(tmem);
}

// Weight images for the bulk copies: [chunk][stage][hi | lo][Co * Ksp] in operand (shared-memory)
// order, k'' = rr * cs + cc  <-  weight[co][(c0 + cc) * KK + rr * TPR + s]; zero where the tap or
// the K padding does not exist. 2 x 147 KB at the benchmark shape, rebuilt on every call
// (weights change every training step).
__global__ void dcn_prep_weights(const float *__restrict__ weight, float *__restrict__ wimg, DcnDims d, FwdPlan pl)
{
    const int per_img = d.Co * pl.Ksp;
    const int nimg = d.dg * pl.ncs * pl.TPR;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nimg * per_img; i += gridDim.x * blockDim.x) {
        const int img = i / per_img, e = i - img * per_img;
        const int s = img % pl.TPR, chunk = img / pl.TPR;
        const int c0 = (chunk / pl.ncs) * d.cpg + (chunk % pl.ncs) * pl.cs;
        const int kk = e & 3, cor = (e >> 2) & 7, rest = e >> 5;
        const int kc = rest % pl.kch, cog = rest / pl.kch;
        const int co = cog * 8 + cor, kp = kc * 4 + kk;
        const int rr = kp / pl.cs, cc = kp - rr * pl.cs, t = rr * pl.TPR + s;
        float hi = 0.f, lo = 0.f;
        if (kp < pl.Ks && t < d.KK)
            umma::split_tf32(__ldg(weight + (size_t)co * d.C * d.KK + (size_t)(c0 + cc) * d.KK + t), hi, lo);
        wimg[(size_t)img * 2 * per_img + e] = hi;
        wimg[(size_t)img * 2 * per_img + per_img + e] = lo;
    }
}

bool make_plan(const DcnDims &d, FwdPlan &pl)
{
    if (d.Co % 16 != 0 || d.Co > 128 || d.cpg % 8 != 0) return false;   // blocked layout: 8-channel chunks
    if ((long)2 * d.KK * d.Ho * d.Wo >= (1L << 31)) return false;        // 32-bit offsets inside one group
    pl.nacc = std::min(3, TMEM_COLS / d.Co - 1);
    if (pl.nacc < 1) return false;
    pl.TPR = ceil_div(d.KK, NR);
    if (pl.TPR > 4) return false;                       // kernels up to 12 taps (3x3, 1x1, 3x4, ...)
    pl.cs = 8;
    pl.ncs = d.cpg / pl.cs;
    pl.Ks = NR * pl.cs;
    pl.Ksp = ebfi::round_up(pl.Ks, 8);
    pl.kch = pl.Ksp / 4;
    pl.a_bytes = TM * pl.Ksp * 4;
    pl.b_bytes = d.Co * pl.Ksp * 4;
    pl.tiles_x = ceil_div(d.Wo, TW);
    pl.tiles_y = ceil_div(d.Ho, TH);
    pl.smem = 4 * pl.a_bytes + 8 * pl.b_bytes;                // 2 A buffers (hi, lo) + 4 weight slots (hi | lo)
    return pl.smem <= 110 * 1024;
}

}  // namespace

size_t forward_tc_workspace(const DcnDims &d)
{
    FwdPlan pl{};
    if (!make_plan(d, pl)) return 0;
    return ebfi::round_up((size_t)d.dg * pl.ncs * pl.TPR * 2 * pl.b_bytes, (size_t)256) +
           (size_t)d.B * d.C * d.H * d.W * sizeof(float);
}

int forward_tc(cudaStream_t st, const DcnDims &d, const float *input, const float *weight, const float *bias,
               const float *offset, const float *mask, float *output, void *workspace, size_t workspace_bytes)
{
    FwdPlan pl{};
    if (!make_plan(d, pl)) return EBFI_ERR_UNSUPPORTED;
    const size_t wbytes = (size_t)d.dg * pl.ncs * pl.TPR * 2 * pl.b_bytes;
    const size_t need = ebfi::round_up(wbytes, (size_t)256) + (size_t)d.B * d.C * d.H * d.W * sizeof(float);
    if (!workspace || workspace_bytes < need || !ebfi::aligned16(workspace)) return EBFI_ERR_UNSUPPORTED;
    float *wimg = static_cast<float *>(workspace);
    float *in_blk = reinterpret_cast<float *>(static_cast<char *>(workspace) + ebfi::round_up(wbytes, (size_t)256));
    dcn_prep_weights<<<ceil_div((int)(wbytes / 8), 256), 256, 0, st>>>(weight, wimg, d, pl);
    EBFI_LAUNCH_OK("dcn_prep_weights");
    if (int rc = launch_nchw_to_blocked(st, input, in_blk, d.B * d.C / 8, d.H * d.W)) return rc;
    const unsigned grid = (unsigned)(d.B * pl.tiles_x * pl.tiles_y);
#define EBFI_FWD_TC(T)                                                                                        \
    do {                                                                                                      \
        if (d.packed) {                                                                                       \
            EBFI_CUDA_OK(cudaFuncSetAttribute(dcn_fwd_tc_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem)); \
            dcn_fwd_tc_kernel<T, true><<<grid, NTHR, pl.smem, st>>>(in_blk, bias, offset, mask, output, wimg, d, pl); \
        } else {                                                                                              \
            EBFI_CUDA_OK(cudaFuncSetAttribute(dcn_fwd_tc_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem)); \
            dcn_fwd_tc_kernel<T, false><<<grid, NTHR, pl.smem, st>>>(in_blk, bias, offset, mask, output, wimg, d, pl); \
        }                                                                                                     \
    } while (0)
    switch (pl.TPR) {
    case 1: EBFI_FWD_TC(1); break;
    case 2: EBFI_FWD_TC(2); break;
    case 3: EBFI_FWD_TC(3); break;
    default: EBFI_FWD_TC(4); break;
    }
#undef EBFI_FWD_TC
    EBFI_LAUNCH_OK("dcn_fwd_tc_kernel");
    return EBFI_OK;
}

}  // namespace ebfi_dcn
