// dcn_tc.cu — DCNv2 forward: shared-memory-staged bilinear sampling fused into a tcgen05 GEMM.
//
// One CTA = a TH x TW (8 x 16) tile of 128 output pixels = the M of a 128 x Cout UMMA (one TMEM
// lane per pixel). 384 threads: thread (pixel p, row r) handles taps r*TPR + s. Per deformable group:
//   1. every thread computes the sampling positions of its taps (offset / mask are read once per
//      (pixel, tap, group)); a block-wide min/max gives the bounding box of all bilinear corners;
//   2. the box (up to RH x RW input pixels, all channels of the chunk) is copied ONCE into shared
//      memory with coalesced loads, stored as 16-byte units of 4 channels — the NHWC staging tile;
//   3. per stage s, each thread gathers its tap's 4 corners x cs channels with 128-bit shared
//      loads, applies the bilinear weights and the mask, splits into TF32 hi + lo and writes the
//      values straight into the K-major A operand (K order of a stage: row r major, channel minor);
//      the weight slice of the stage is staged next to it; one thread issues the 3xTF32
//      tcgen05.mma chain into TMEM and commits to the stage buffer's mbarrier.
//      Operand stages are double buffered, so sampling of stage n+1 overlaps the MMAs of stage n.
//   Taps whose corners fall outside the staged box (offsets beyond the box capacity — nothing bounds
//   offsets in DCNv2, dcn_v2.py:221-223 only warns) read from global memory instead; the result is
//   the same, only slower.
// Epilogue: tcgen05.ld (lane = pixel), sum of the split accumulators, + bias, NCHW stores.
//
// Why: with L1 gathers the kernel was bound by L1 misses and LSU wavefronts (ncu: 455 wavefronts
// per pixel, 70 % L1 hit rate with the L1 squeezed by the operand buffers); staged, each input line
// is fetched once per (tile, group) and the gather runs at shared-memory speed.
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
constexpr int NTHR = TM * NR;    // 384
constexpr int TMEM_COLS = 256;

struct FwdPlan {
    int cs, ncs;         // channels per chunk (4 or 8), chunks per deformable group
    int TPR;             // taps per thread row = stages per chunk: tap = r * TPR + s
    int Ks, Ksp, kch;    // K of a stage (NR * cs), padded to 8, 16-byte chunks per operand row
    int nacc;            // hi*hi accumulators (+1 for the cross terms)
    int RH, RW;          // staged box capacity (input pixels); 0 = staging disabled
    int J;               // offset magnitude the box is sized for
    int tiles_x, tiles_y;
    int reg_bytes, a_bytes, b_bytes, smem;
};

__device__ __forceinline__ float4 ld_corner(const float4 *reg, bool ok, int unit)
{
    return ok ? reg[unit] : make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(NTHR, 2)
dcn_fwd_tc_kernel(const float *__restrict__ input, const float *__restrict__ weight,
                  const float *__restrict__ bias, const float *__restrict__ offset,
                  const float *__restrict__ mask, float *__restrict__ output,
                  const float *__restrict__ wimg, DcnDims d, FwdPlan pl)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *region = reinterpret_cast<float4 *>(smem_raw);                 // [cs/4][RH][RW] units of 4 channels
    unsigned char *opnd = smem_raw + pl.reg_bytes;                         // [2 buffers][a_hi, a_lo, b_hi, b_lo]
    const int buf_bytes = 2 * pl.a_bytes + 2 * pl.b_bytes;
    __shared__ __align__(8) uint64_t bar[2];        // MMAs that read stage buffer i are complete
    __shared__ __align__(8) uint64_t bar_w[2];      // weight image of stage buffer i has landed
    __shared__ uint32_t tmem_slot;
    __shared__ int bbox[4];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p = tid % TM, r = tid / TM;
    int tile = blockIdx.x;
    const int tx0 = (tile % pl.tiles_x) * TW; tile /= pl.tiles_x;
    const int ty0 = (tile % pl.tiles_y) * TH;
    const int b = tile / pl.tiles_y;
    const int ho = ty0 + p / TW, wo = tx0 + p % TW;
    const bool valid = ho < d.Ho && wo < d.Wo;
    const int pix = ho * d.Wo + wo;
    const size_t plane = (size_t)d.Ho * d.Wo, in_plane = (size_t)d.H * d.W;
    const int Kdim = d.C * d.KK;
    const uint32_t sbo = (uint32_t)pl.kch * 128u;

    if (warp == 0) umma::tmem_alloc<TMEM_COLS>(&tmem_slot);
    if (tid == 0) {
        umma::mbar_init(&bar[0], 1); umma::mbar_init(&bar[1], 1);
        umma::mbar_init(&bar_w[0], 1); umma::mbar_init(&bar_w[1], 1);
        umma::mbar_fence_init();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = umma::instr_desc_tf32(TM, d.Co, 0, 0);

    const int a_row = (p >> 3) * (pl.kch * 32) + (p & 7) * 4;             // floats, row of pixel p in an A image
    uint32_t phase[2] = {0, 0}, phase_w[2] = {0, 0};
    int nstage = 0;                                                        // stages issued so far (all threads)
    int step = 0;                                                          // MMA k-steps issued (thread 0)
    const bool staging = pl.RH > 0;

    for (int g = 0; g < d.dg; ++g) {
        const float *off_bg = offset + ((size_t)b * d.dg + g) * 2 * d.KK * plane;
        const float *mask_bg = mask + ((size_t)b * d.dg + g) * d.KK * plane;
        // ---- sampling positions of this thread's taps; bounding box of their corners
        float sy[4], sx[4], sm[4];                                         // TPR <= 4 (see make_plan)
        int ymin = INT_MAX, ymax = INT_MIN, xmin = INT_MAX, xmax = INT_MIN;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            sy[s] = -2.f; sx[s] = -2.f; sm[s] = 0.f;                       // (-2,-2): outside the window
            const int t = r * pl.TPR + s;
            if (s < pl.TPR && valid && t < d.KK) {
                float xq;
                tap_coords(d, off_bg, mask_bg, t, pix, sy[s], sx[s], xq, sm[s]);
                if (sy[s] > -1.f && sx[s] > -1.f && sy[s] < (float)d.H && sx[s] < (float)d.W) {
                    const int y0 = (int)floorf(sy[s]), x0 = (int)floorf(sx[s]);
                    ymin = min(ymin, max(y0, 0)); ymax = max(ymax, min(y0 + 1, d.H - 1));
                    xmin = min(xmin, max(x0, 0)); xmax = max(xmax, min(x0 + 1, d.W - 1));
                }
            }
        }
        int oy = 0, ox = 0;
        if (staging) {
            if (tid == 0) { bbox[0] = INT_MAX; bbox[1] = INT_MIN; bbox[2] = INT_MAX; bbox[3] = INT_MIN; }
            __syncthreads();                                               // also: previous group's gathers done
            ymin = __reduce_min_sync(0xffffffffu, ymin); ymax = __reduce_max_sync(0xffffffffu, ymax);
            xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
            if (lane == 0) {
                atomicMin(&bbox[0], ymin); atomicMax(&bbox[1], ymax);
                atomicMin(&bbox[2], xmin); atomicMax(&bbox[3], xmax);
            }
            __syncthreads();
            // Box origin: the bounding box if it fits, else centred on the tile's nominal footprint.
            const int ny = ty0 * d.sh - d.ph - pl.J, nx = tx0 * d.sw - d.pw - pl.J;
            if (bbox[0] <= bbox[1]) {                                      // else: no tap of the tile is inside the window
                oy = (bbox[1] - bbox[0] < pl.RH) ? bbox[0] : max(0, min(ny, d.H - pl.RH));
                ox = (bbox[3] - bbox[2] < pl.RW) ? bbox[2] : max(0, min(nx, d.W - pl.RW));
            }
        }
        for (int ci = 0; ci < pl.ncs; ++ci) {
            const int c0 = g * d.cpg + ci * pl.cs;
            const float *ip0 = input + ((size_t)b * d.C + c0) * in_plane;
            if (staging) {
                if (ci > 0) __syncthreads();                               // gathers of the previous chunk done
                // rows of the box are walked by (warp, lane) without divisions: warp -> (q, ry), lane -> rx
                const int nrow = (pl.cs >> 2) * pl.RH;
                for (int row = warp; row < nrow; row += NTHR / 32) {
                    const int q = row / pl.RH, ry = row - q * pl.RH, gy = oy + ry;
                    const float *ip = ip0 + (size_t)(4 * q) * in_plane + (size_t)gy * d.W + ox;
                    for (int rx = lane; rx < pl.RW; rx += 32) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (gy < d.H && ox + rx < d.W)
                            v = make_float4(__ldg(ip + rx), __ldg(ip + in_plane + rx), __ldg(ip + 2 * in_plane + rx),
                                            __ldg(ip + 3 * in_plane + rx));
                        region[row * pl.RW + rx] = v;
                    }
                }
                __syncthreads();
            }
            for (int s = 0; s < pl.TPR; ++s, ++nstage) {
                const int bi = nstage & 1;
                float *a_hi = reinterpret_cast<float *>(opnd + bi * buf_bytes);
                float *a_lo = reinterpret_cast<float *>(opnd + bi * buf_bytes + pl.a_bytes);
                float *b_hi = reinterpret_cast<float *>(opnd + bi * buf_bytes + 2 * pl.a_bytes);
                float *b_lo = reinterpret_cast<float *>(opnd + bi * buf_bytes + 2 * pl.a_bytes + pl.b_bytes);
                if (nstage >= 2) {                                         // MMAs that read this buffer are done
                    umma::mbar_wait(&bar[bi], phase[bi]);
                    phase[bi] ^= 1;
                }
                // ---- weights of the stage: one bulk (TMA) copy of the pre-split hi|lo image that
                //      dcn_prep_weights wrote in exactly this shared-memory order
                if (tid == 0) {
                    const int img = ((g * pl.ncs + ci) * pl.TPR + s);
                    umma::mbar_expect_tx(&bar_w[bi], 2u * (uint32_t)pl.b_bytes);
                    umma::bulk_g2s(b_hi, wimg + (size_t)img * (2 * pl.b_bytes / 4), 2u * (uint32_t)pl.b_bytes, &bar_w[bi]);
                }
                // ---- this thread's tap of the stage -> cs values of row p, columns r*cs .. r*cs+cs-1
                {
                    const float y = s == 0 ? sy[0] : s == 1 ? sy[1] : s == 2 ? sy[2] : sy[3];
                    const float x = s == 0 ? sx[0] : s == 1 ? sx[1] : s == 2 ? sx[2] : sx[3];
                    const float m = s == 0 ? sm[0] : s == 1 ? sm[1] : s == 2 ? sm[2] : sm[3];
                    const Tap tp = make_tap(y, x, d.H, d.W);
                    const float w1 = tp.hy * tp.hx, w2 = tp.hy * tp.lx, w3 = tp.ly * tp.hx, w4 = tp.ly * tp.lx;
                    const bool any = tp.c00 || tp.c01 || tp.c10 || tp.c11;
                    const int y0 = any ? (int)floorf(y) : 0, x0 = any ? (int)floorf(x) : 0;   // corner 00
                    // corners that are read must lie inside the staged box
                    const int ry = y0 - oy, rx = x0 - ox;
                    const bool in_box = staging && any &&
                        (!(tp.c00 || tp.c01) || (ry >= 0 && ry < pl.RH)) && (!(tp.c10 || tp.c11) || (ry + 1 >= 0 && ry + 1 < pl.RH)) &&
                        (!(tp.c00 || tp.c10) || (rx >= 0 && rx < pl.RW)) && (!(tp.c01 || tp.c11) || (rx + 1 >= 0 && rx + 1 < pl.RW));
                    for (int q = 0; q < (pl.cs >> 2); ++q) {
                        float v[4];
                        if (in_box) {
                            const int u = (q * pl.RH + ry) * pl.RW + rx;
                            const float4 a = ld_corner(region, tp.c00, u), bq = ld_corner(region, tp.c01, u + 1);
                            const float4 c = ld_corner(region, tp.c10, u + pl.RW), e4 = ld_corner(region, tp.c11, u + pl.RW + 1);
                            v[0] = (w1 * a.x + w2 * bq.x + w3 * c.x + w4 * e4.x) * m;
                            v[1] = (w1 * a.y + w2 * bq.y + w3 * c.y + w4 * e4.y) * m;
                            v[2] = (w1 * a.z + w2 * bq.z + w3 * c.z + w4 * e4.z) * m;
                            v[3] = (w1 * a.w + w2 * bq.w + w3 * c.w + w4 * e4.w) * m;
                        } else if (any) {                                  // outside the box: global gather
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float *ip = ip0 + (size_t)(4 * q + j) * in_plane;
                                const float v1 = tp.c00 ? __ldg(ip + tp.i00) : 0.f;
                                const float v2 = tp.c01 ? __ldg(ip + tp.i01) : 0.f;
                                const float v3 = tp.c10 ? __ldg(ip + tp.i10) : 0.f;
                                const float v4 = tp.c11 ? __ldg(ip + tp.i11) : 0.f;
                                v[j] = (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4) * m;
                            }
                        } else {
                            v[0] = v[1] = v[2] = v[3] = 0.f;
                        }
                        float hi[4], lo[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) umma::split_tf32(v[j], hi[j], lo[j]);
                        const int off = a_row + ((r * pl.cs + 4 * q) >> 2) * 32;
                        *reinterpret_cast<float4 *>(a_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<float4 *>(a_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    }
                    if (r == 0 && pl.Ksp > pl.Ks) {                        // zero the K padding of this row
                        const int off = a_row + (pl.Ks >> 2) * 32;
                        *reinterpret_cast<float4 *>(a_hi + off) = make_float4(0.f, 0.f, 0.f, 0.f);
                        *reinterpret_cast<float4 *>(a_lo + off) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                umma::fence_smem_to_async();
                __syncthreads();
                if (tid == 0) {
                    umma::mbar_wait(&bar_w[bi], phase_w[bi]);
                    phase_w[bi] ^= 1;
                    umma::fence_after_sync();
                    const uint32_t ah = umma::smem_u32(a_hi), al = umma::smem_u32(a_lo);
                    const uint32_t bh = umma::smem_u32(b_hi), bl = umma::smem_u32(b_lo);
                    for (int ks = 0; ks < pl.Ksp / 8; ++ks, ++step) {
                        const uint32_t ko = (uint32_t)ks * 256u;
                        const uint64_t dah = umma::smem_desc(ah + ko, 128, sbo), dal = umma::smem_desc(al + ko, 128, sbo);
                        const uint64_t dbh = umma::smem_desc(bh + ko, 128, sbo), dbl = umma::smem_desc(bl + ko, 128, sbo);
                        const uint32_t d_x = tmem + pl.nacc * d.Co, d_h = tmem + (step % pl.nacc) * d.Co;
                        umma::mma_tf32(d_x, dal, dbh, idesc, step > 0);
                        umma::mma_tf32(d_x, dah, dbl, idesc, true);
                        umma::mma_tf32(d_h, dah, dbh, idesc, step >= pl.nacc);
                    }
                    umma::commit(&bar[bi]);
                }
            }
        }
    }
    // ---- drain: the last commit on each buffer has not been waited for yet
    if (nstage >= 2) umma::mbar_wait(&bar[nstage & 1], phase[nstage & 1]);
    umma::mbar_wait(&bar[(nstage - 1) & 1], phase[(nstage - 1) & 1]);
    umma::fence_after_sync();

    // ---- epilogue: thread (p, r) takes every NR-th block of 8 output channels of pixel p
    const int total_steps = nstage * (pl.Ksp / 8), nused = min(pl.nacc, total_steps);
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
    if (d.Co % 16 != 0 || d.Co > 128 || d.cpg % 4 != 0) return false;
    pl.nacc = std::min(3, TMEM_COLS / d.Co - 1);
    if (pl.nacc < 1) return false;
    pl.TPR = ceil_div(d.KK, NR);
    if (pl.TPR > 4) return false;                       // kernels up to 12 taps (3x3, 1x1, 3x4, ...)
    pl.cs = (d.cpg % 8 == 0) ? 8 : 4;
    pl.ncs = d.cpg / pl.cs;
    pl.Ks = NR * pl.cs;
    pl.Ksp = ebfi::round_up(pl.Ks, 8);
    pl.kch = pl.Ksp / 4;
    pl.a_bytes = TM * pl.Ksp * 4;
    pl.b_bytes = d.Co * pl.Ksp * 4;
    pl.tiles_x = ceil_div(d.Wo, TW);
    pl.tiles_y = ceil_div(d.Ho, TH);
    // Staged box: the tile's nominal footprint plus J pixels of offset on every side, J as large as
    // ~40 KB allows (J = 7 covers |offset| <= 7, i.e. 3.5 sigma of the benchmark's 2*randn offsets).
    const char *env = getenv("EBFI_DCN_STAGE");
    pl.RH = pl.RW = 0; pl.J = 0;
    if (!(env && env[0] == '0')) {
        for (int J = 7; J >= 1; --J) {
            const int RH = (TH - 1) * d.sh + (d.kh - 1) * d.dh + 2 * J + 2;
            const int RW = (TW - 1) * d.sw + (d.kw - 1) * d.dw + 2 * J + 2;
            if (RH * RW * pl.cs * 4 <= 40 * 1024) { pl.RH = std::min(RH, d.H); pl.RW = std::min(RW, d.W); pl.J = J; break; }
        }
    }
    pl.reg_bytes = ebfi::round_up(pl.RH * pl.RW * pl.cs * 4, 128);
    pl.smem = pl.reg_bytes + 2 * (2 * pl.a_bytes + 2 * pl.b_bytes);
    return pl.smem <= 110 * 1024;
}

}  // namespace

size_t forward_tc_workspace(const DcnDims &d)
{
    FwdPlan pl{};
    if (!make_plan(d, pl)) return 0;
    return (size_t)d.dg * pl.ncs * pl.TPR * 2 * pl.b_bytes;
}

int forward_tc(cudaStream_t st, const DcnDims &d, const float *input, const float *weight, const float *bias,
               const float *offset, const float *mask, float *output, void *workspace, size_t workspace_bytes)
{
    FwdPlan pl{};
    if (!make_plan(d, pl)) return EBFI_ERR_UNSUPPORTED;
    const size_t need = (size_t)d.dg * pl.ncs * pl.TPR * 2 * pl.b_bytes;
    if (!workspace || workspace_bytes < need || !ebfi::aligned16(workspace)) return EBFI_ERR_UNSUPPORTED;
    float *wimg = static_cast<float *>(workspace);
    dcn_prep_weights<<<ceil_div((int)(need / 8), 256), 256, 0, st>>>(weight, wimg, d, pl);
    EBFI_LAUNCH_OK("dcn_prep_weights");
    EBFI_CUDA_OK(cudaFuncSetAttribute(dcn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem));
    const unsigned grid = (unsigned)(d.B * pl.tiles_x * pl.tiles_y);
    dcn_fwd_tc_kernel<<<grid, NTHR, pl.smem, st>>>(input, weight, bias, offset, mask, output, wimg, d, pl);
    EBFI_LAUNCH_OK("dcn_fwd_tc_kernel");
    return EBFI_OK;
}

}  // namespace ebfi_dcn
