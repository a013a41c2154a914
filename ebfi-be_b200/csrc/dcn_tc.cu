// dcn_tc.cu — DCNv2 forward with the weight contraction on the sm_100a tensor cores.
//
// One CTA = one tile of 128 output pixels (the M of a 128 x Cout UMMA, one TMEM lane per pixel).
// It walks the deformable groups (split into channel chunks if a group does not fit); per chunk
//   1. all 384 threads sample: thread (pixel p, tap-row r) reads offset/mask of its taps once,
//      gathers the chunk's channels at the 4 bilinear corners (NCHW planes through L1), applies
//      the mask, splits every value into TF32 hi + lo and writes them as 16-byte chunks straight
//      into the UMMA A operand in shared memory (K order inside a chunk: tap-major, channel-minor,
//      so one (pixel, tap) item is one contiguous K run) — the column buffer of the reference
//      (dcn_v2_cuda.cu:68) lives only here;
//   2. the same threads stage the matching weight slice W[co][chunk] (hi / lo, same K order);
//   3. one thread issues the 3xTF32 tcgen05.mma chain into TMEM accumulators and commits to an
//      mbarrier; the CTA's threads wait on it before overwriting the operands.
// Two CTAs are resident per SM (launch bounds 384 x 2, ~110 KB smem, 256 TMEM columns each), so
// one CTA samples while the other's MMAs drain: the overlap is across CTAs, no double buffering.
// Epilogue: tcgen05.ld (lane = pixel), sum of the split accumulators, + bias, NCHW stores that are
// coalesced across the 32 pixels of a warp.
//
// fp32 parity: products are 3xTF32 (umma.cuh) and the hi*hi chain is spread over several TMEM
// accumulators because the tensor core's fp32 accumulate rounds toward zero (error grows with the
// length of a dependent chain); measured max-rel error vs fp64 is < 1e-6 at K = 576.
#include "dcn_common.cuh"
#include "umma.cuh"

#include <algorithm>

namespace ebfi_dcn {

namespace {

using ebfi::ceil_div;

constexpr int TM = 128;          // pixels per CTA tile
constexpr int NR = 3;            // tap-rows of threads: thread (p, r) handles taps r, r+NR, ...
constexpr int NTHR = TM * NR;    // 384
constexpr int TMEM_COLS = 256;

struct TcPlan {
    int cs;          // channels per chunk (multiple of 4, divides cpg)
    int ncs;         // chunks per deformable group
    int Kc;          // cs * KK, K extent of a chunk
    int Kp;          // Kc rounded up to 8 (MMA K-step)
    int kch;         // Kp / 4: 16-byte chunks per operand row
    int nacc;        // hi*hi accumulators (plus one for the cross terms)
    int smem;        // dynamic shared memory bytes
};

__global__ void __launch_bounds__(NTHR, 2)
dcn_fwd_tc_kernel(const float *__restrict__ input, const float *__restrict__ weight,
                  const float *__restrict__ bias, const float *__restrict__ offset,
                  const float *__restrict__ mask, float *__restrict__ output, DcnDims d, TcPlan pl)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int a_bytes = TM * pl.Kp * 4, b_bytes = d.Co * pl.Kp * 4;
    float *a_hi = reinterpret_cast<float *>(smem_raw);
    float *a_lo = reinterpret_cast<float *>(smem_raw + a_bytes);
    float *b_hi = reinterpret_cast<float *>(smem_raw + 2 * a_bytes);
    float *b_lo = reinterpret_cast<float *>(smem_raw + 2 * a_bytes + b_bytes);
    __shared__ __align__(8) uint64_t bar_mma;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int p = tid % TM, r = tid / TM;
    const int npix = d.Ho * d.Wo;
    const int b = blockIdx.x / d.ntile, pix = (blockIdx.x % d.ntile) * TM + p;
    const bool valid = pix < npix;
    const size_t plane = (size_t)npix, in_plane = (size_t)d.H * d.W;
    const int Kdim = d.C * d.KK;
    const uint32_t sbo = (uint32_t)pl.kch * 128u;            // bytes between 8-row groups

    if (warp == 0) umma::tmem_alloc<TMEM_COLS>(&tmem_slot);
    if (tid == 0) { umma::mbar_init(&bar_mma, 1); umma::mbar_fence_init(); }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = umma::instr_desc_tf32(TM, d.Co, 0, 0);

    // row base of this thread's pixel inside a K-major operand image (floats)
    const int a_row = (p >> 3) * (pl.kch * 32) + (p & 7) * 4;
    uint32_t phase = 0;
    int step = 0;                                            // k-step counter (thread 0)
    const int nchunks = d.dg * pl.ncs;
    for (int c = 0; c < nchunks; ++c) {
        const int g = c / pl.ncs, c0 = g * d.cpg + (c % pl.ncs) * pl.cs;
        if (c > 0) {                                         // previous chunk's MMAs have read the operands
            umma::mbar_wait(&bar_mma, phase);
            phase ^= 1;
        }
        // ---- weights of the chunk: b[co][k'] with k' = tap * cs + cc  <-  weight[co][(c0 + cc) * KK + tap]
        for (int e = tid; e < d.Co * pl.Kp; e += NTHR) {
            const int co = e / pl.Kp, k = e - co * pl.Kp;    // k in natural (channel-major) order
            float hi = 0.f, lo = 0.f;
            int kp = k;
            if (k < pl.Kc) {
                const int cc = k / d.KK, t = k - cc * d.KK;
                kp = t * pl.cs + cc;
                umma::split_tf32(__ldg(weight + (size_t)co * Kdim + (size_t)c0 * d.KK + k), hi, lo);
            }
            const int off = (co >> 3) * (pl.kch * 32) + (kp >> 2) * 32 + (co & 7) * 4 + (kp & 3);
            b_hi[off] = hi; b_lo[off] = lo;
        }
        // ---- sample this thread's taps of the chunk into the A operand
        if (valid) {
            const float *off_bg = offset + ((size_t)b * d.dg + g) * 2 * d.KK * plane;
            const float *mask_bg = mask + ((size_t)b * d.dg + g) * d.KK * plane;
            const float *ip0 = input + ((size_t)b * d.C + c0) * in_plane;
            for (int t = r; t < d.KK; t += NR) {
                float y, x, xq, m;
                tap_coords(d, off_bg, mask_bg, t, pix, y, x, xq, m);
                const Tap tp = make_tap(y, x, d.H, d.W);
                const float w1 = tp.hy * tp.hx, w2 = tp.hy * tp.lx, w3 = tp.ly * tp.hx, w4 = tp.ly * tp.lx;
                for (int cc = 0; cc < pl.cs; cc += 4) {
                    float hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float *ip = ip0 + (size_t)(cc + j) * in_plane;
                        const float v1 = tp.c00 ? __ldg(ip + tp.i00) : 0.f;
                        const float v2 = tp.c01 ? __ldg(ip + tp.i01) : 0.f;
                        const float v3 = tp.c10 ? __ldg(ip + tp.i10) : 0.f;
                        const float v4 = tp.c11 ? __ldg(ip + tp.i11) : 0.f;
                        umma::split_tf32((w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4) * m, hi[j], lo[j]);
                    }
                    const int off = a_row + ((t * pl.cs + cc) >> 2) * 32;
                    *reinterpret_cast<float4 *>(a_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4 *>(a_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            if (r == 0 && pl.Kp > pl.Kc) {                   // zero the K padding of this row
                const int off = a_row + (pl.Kc >> 2) * 32;
                *reinterpret_cast<float4 *>(a_hi + off) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4 *>(a_lo + off) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else if (c == 0) {                                  // rows of pixels past the image: finite zeros
            for (int ch = r; ch < pl.kch; ch += NR) {
                *reinterpret_cast<float4 *>(a_hi + a_row + ch * 32) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4 *>(a_lo + a_row + ch * 32) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        umma::fence_smem_to_async();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after_sync();
            const uint32_t ah = umma::smem_u32(a_hi), al = umma::smem_u32(a_lo);
            const uint32_t bh = umma::smem_u32(b_hi), bl = umma::smem_u32(b_lo);
            for (int ks = 0; ks < pl.Kp / 8; ++ks, ++step) {
                const uint32_t ko = (uint32_t)ks * 256u;     // two 16-byte K chunks per k-step
                const uint64_t dah = umma::smem_desc(ah + ko, 128, sbo), dal = umma::smem_desc(al + ko, 128, sbo);
                const uint64_t dbh = umma::smem_desc(bh + ko, 128, sbo), dbl = umma::smem_desc(bl + ko, 128, sbo);
                const uint32_t d_x = tmem + pl.nacc * d.Co, d_h = tmem + (step % pl.nacc) * d.Co;
                umma::mma_tf32(d_x, dal, dbh, idesc, step > 0);
                umma::mma_tf32(d_x, dah, dbl, idesc, true);
                umma::mma_tf32(d_h, dah, dbh, idesc, step >= pl.nacc);
            }
            umma::commit(&bar_mma);
        }
    }
    umma::mbar_wait(&bar_mma, phase);
    umma::fence_after_sync();

    // ---- epilogue: thread (p, r) takes every NR-th block of 8 output channels of pixel p
    const int total_steps = nchunks * (pl.Kp / 8), nused = min(pl.nacc, total_steps);
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
            for (int i = 0; i < 8; ++i)
                if (cb + i < d.Co) output[((size_t)b * d.Co + cb + i) * plane + pix] = v[i] + __ldg(bias + cb + i);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<TMEM_COLS>(tmem);
}

bool make_plan(const DcnDims &d, TcPlan &pl)
{
    if (d.Co % 16 != 0 || d.Co > 128 || d.cpg % 4 != 0) return false;
    pl.nacc = std::min(3, TMEM_COLS / d.Co - 1);
    if (pl.nacc < 1) return false;
    // largest channel chunk (multiple of 4, dividing cpg) whose operands fit ~110 KB (2 CTAs per SM)
    const int budget = 112 * 1024;
    pl.cs = 0;
    for (int cs = d.cpg; cs >= 4; cs -= 4) {
        if (d.cpg % cs) continue;
        const int Kp = ebfi::round_up(cs * d.KK, 8);
        if (2 * (TM + d.Co) * Kp * 4 <= budget) { pl.cs = cs; break; }
    }
    if (pl.cs == 0) return false;
    pl.ncs = d.cpg / pl.cs;
    pl.Kc = pl.cs * d.KK;
    pl.Kp = ebfi::round_up(pl.Kc, 8);
    pl.kch = pl.Kp / 4;
    if (pl.kch * 128 >= (1 << 18)) return false;
    pl.smem = 2 * (TM + d.Co) * pl.Kp * 4;
    return true;
}

}  // namespace

int forward_tc(cudaStream_t st, const DcnDims &d, const float *input, const float *weight, const float *bias,
               const float *offset, const float *mask, float *output)
{
    TcPlan pl{};
    if (!make_plan(d, pl)) return EBFI_ERR_UNSUPPORTED;
    const int ntile = ceil_div(d.Ho * d.Wo, TM);
    DcnDims dd = d;
    dd.ntile = ntile;
    EBFI_CUDA_OK(cudaFuncSetAttribute(dcn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem));
    dcn_fwd_tc_kernel<<<(unsigned)(d.B * ntile), NTHR, pl.smem, st>>>(input, weight, bias, offset, mask, output, dd, pl);
    EBFI_LAUNCH_OK("dcn_fwd_tc_kernel");
    return EBFI_OK;
}

}  // namespace ebfi_dcn
