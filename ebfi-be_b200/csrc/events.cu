// events.cu — event -> image / voxel-grid / polarity-stack encoders for sm_100a.
//
// GPU scatter kernels for the bins of /root/reference/dataloader/encodings.py
// (events_to_image :243-268, events_to_voxel :271-286, events_to_stack :307-350,
// events_to_mask :353-377). The reference runs these on the CPU in DataLoader workers as
// B (or 2B) full passes of index_put_(accumulate=True) over all events; here ONE pass reads
// each event once (struct-of-arrays, coalesced) and issues at most 2 (voxel) / 2 per
// containing bin (stack) reductions into the L2-resident grid.
//
// Reference side effects that are observable in the OUTPUT are reproduced:
//   * events_to_image zeroes out-of-range events in the caller's tensors (:254-256). Because
//     events_to_voxel calls it once per bin with the same xs/ys, an out-of-range event is
//     dropped from bin 0 but lands on pixel (0,0) in bins >= 1; events_to_stack shows the
//     same effect between its positive and negative pass and between overlapping bins.
//   * events_to_stack finds bin boundaries with binary_search_torch_tensor (:77-99), whose
//     early exits decide how duplicate / boundary-equal timestamps are split; the same
//     search runs here on the device.
// Polarity-count outputs (stack, channels, image of +-1) are sums of small integers and are
// therefore exact and order-independent in fp32; the temporally-weighted voxel grid is
// accumulated with fp32 reductions whose order is not fixed (|diff| vs the CPU's sequential
// order is a few ulp; see tests for the bound).
#include "common.cuh"

#include <algorithm>

namespace {

using ebfi::ceil_div;

template <typename T> __device__ __forceinline__ bool out_of_range(T x, T y, int H, int W)
{
    // (xs >= W) + (xs < 0) + (ys >= H) + (ys < 0), encodings.py:251-253
    return (x >= (T)W) || (x < (T)0) || (y >= (T)H) || (y < (T)0);
}

__device__ __forceinline__ bool out_of_range(int16_t x, int16_t y, int H, int W)
{
    return ((int)x >= W) || (x < 0) || ((int)y >= H) || (y < 0);
}

__device__ __forceinline__ void red_add(float *p, float v)
{
    if (v != 0.f) atomicAdd(p, v);     // result unused -> RED.ADD.F32; adding +-0 is a no-op
}

template <typename T>
__global__ void events_image_kernel(T *__restrict__ xs, T *__restrict__ ys, float *__restrict__ ps,
                                    int64_t n, int H, int W, float *__restrict__ img, int write_back)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const T x = xs[i], y = ys[i];
        if (out_of_range(x, y, H, W)) {
            if (write_back) { xs[i] = (T)0; ys[i] = (T)0; ps[i] = 0.f; }
            continue;
        }
        red_add(img + (int64_t)y * W + (int64_t)x, ps[i]);   // .long() truncation, :262-265
    }
}

// events_to_mask: index_put_(accumulate=False) on the CPU = the LAST event of a pixel wins.
// Pass 1 records the largest event index per pixel, pass 2 lets exactly that event write.
// Out-of-range events take part at pixel (0,0) with value 0, as in the reference.
template <typename T>
__global__ void events_mask_last_kernel(const T *__restrict__ xs, const T *__restrict__ ys, int64_t n,
                                        int H, int W, long long *__restrict__ last)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const T x = xs[i], y = ys[i];
        const int64_t pix = out_of_range(x, y, H, W) ? 0 : (int64_t)y * W + (int64_t)x;
        atomicMax(last + pix, (long long)i);
    }
}

template <typename T>
__global__ void events_mask_write_kernel(T *__restrict__ xs, T *__restrict__ ys, float *__restrict__ ps,
                                         int64_t n, int H, int W, const long long *__restrict__ last,
                                         float *__restrict__ img, int write_back)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const T x = xs[i], y = ys[i];
        const bool oob = out_of_range(x, y, H, W);
        const int64_t pix = oob ? 0 : (int64_t)y * W + (int64_t)x;
        if (last[pix] == (long long)i) img[pix] = oob ? 0.f : fabsf(ps[i]);
        if (oob && write_back) { xs[i] = (T)0; ys[i] = (T)0; ps[i] = 0.f; }
    }
}

// Temporal-bilinear weight of bin b, evaluated operation by operation in the dtype of ts like
// the tensor expression `ps * max(0, 1 - |ts*(bins-1) - b|)` (:279-283); no FMA contraction.
__device__ __forceinline__ float voxel_value(float t_scaled, int b, float p)
{
    const float k = __fsub_rn(1.0f, fabsf(__fsub_rn(t_scaled, (float)b)));
    return __fmul_rn(p, k > 0.f ? k : 0.f);
}
__device__ __forceinline__ float voxel_value(double t_scaled, int b, float p)
{
    const double k = __dsub_rn(1.0, fabs(__dsub_rn(t_scaled, (double)b)));
    return (float)__dmul_rn((double)p, k > 0.0 ? k : 0.0);
}
__device__ __forceinline__ float scale_ts(float t, int bins) { return __fmul_rn(t, (float)(bins - 1)); }
__device__ __forceinline__ double scale_ts(double t, int bins) { return __dmul_rn(t, (double)(bins - 1)); }

template <typename T>
__global__ void events_voxel_kernel(T *__restrict__ xs, T *__restrict__ ys, const T *__restrict__ ts,
                                    const float *__restrict__ ps, int64_t n, int bins, int H, int W,
                                    float *__restrict__ voxel, int write_back)
{
    const int64_t plane = (int64_t)H * W;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const T x = xs[i], y = ys[i];
        const bool oob = out_of_range(x, y, H, W);
        const int64_t pix = oob ? 0 : (int64_t)y * W + (int64_t)x;
        const T tsc = scale_ts(ts[i], bins);
        const float p = ps[i];
        // only bins floor(t), floor(t)+1 can carry a non-zero weight
        const int b0 = (int)floor((double)tsc);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int b = b0 + k;
            if (b < 0 || b >= bins) continue;
            if (oob && b == 0) continue;            // zeroed by the first bin's events_to_image call
            red_add(voxel + (int64_t)b * plane + pix, voxel_value(tsc, b, p));
        }
        if (oob && write_back) { xs[i] = (T)0; ys[i] = (T)0; }
    }
}

// Timestamps as the datasets hand them to events_to_stack: GetEventsIndex (dataloader/h5dataset.py:327-336)
// normalises the on-disk float64 seconds with `ts = (ts - ts[0]) / (ts[-1] - ts[0] + 1e-6)` in float64.
// The same two IEEE operations are evaluated on access, so the raw array never has to be rewritten.
struct NormalizedTs {
    const double *raw;
    double t0, den;
    __device__ __forceinline__ double operator[](int64_t i) const { return __ddiv_rn(__dsub_rn(raw[i], t0), den); }
};
__device__ __forceinline__ NormalizedTs normalized(const double *raw, int64_t n)
{
    return NormalizedTs{raw, raw[0], __dadd_rn(__dsub_rn(raw[n - 1], raw[0]), 1e-6)};
}

// binary_search_torch_tensor (:77-99) for all 2*bins boundaries, one WARP each. The search path itself is the
// reference's (its result for duplicate timestamps depends on which probe hits first), but a warp runs it three levels
// at a time: the 7 (l, r) ranges reachable within three steps depend only on the current range, not on the data, so
// the warp fetches their 21 probes (ts[l], ts[r], ts[mid] each) in ONE round of loads and then replays the three
// comparisons from registers. 24 dependent DRAM round trips become 8 (21-27 us -> ~8 us at 10 M events).
template <typename T, typename TS>
__device__ void stack_bounds(const TS &ts, int64_t n, int bins, int64_t *__restrict__ bounds, int e)
{
    const int bi = e >> 1, right = e & 1, lane = threadIdx.x & 31;
    // dt = ts[-1]-ts[0]+1e-6; delta = dt/B; tstart = ts[0]+delta*bi; tend = tstart+delta (:324-329),
    // every step rounded in the dtype of ts
    const T t0 = ts[0];
    const T dt = (T)((T)(ts[n - 1] - t0) + (T)1e-6);
    const T delta = dt / (T)bins;
    T target;
    if (sizeof(T) == 4) {
        const float a = __fadd_rn((float)t0, __fmul_rn((float)delta, (float)bi));
        target = (T)(right ? __fadd_rn(a, (float)delta) : a);
    } else {
        const double a = __dadd_rn((double)t0, __dmul_rn((double)delta, (double)bi));
        target = (T)(right ? __dadd_rn(a, (double)delta) : a);
    }
    int64_t l = 0, r = n - 1, res = -2;
    bool done = false;
    while (!done) {
        // lane = node * 3 + probe; node k of the implicit tree: 0 = (l, r), 2k+1 = its left child (r = mid - 1),
        // 2k+2 = its right child (l = mid + 1)
        const int node = lane / 3, probe = lane - node * 3;
        int64_t nl = l, nr = r;
        bool live = lane < 21;
        if (live) {
            // path from the root: bits of (node + 1) below its leading one, most significant first
            const int depth = 31 - __clz(node + 1);
            for (int dlev = depth - 1; dlev >= 0 && nl <= nr; --dlev) {
                const int64_t mid = nl + (nr - nl) / 2;
                if (((node + 1) >> dlev) & 1) nl = mid + 1; else nr = mid - 1;
            }
            live = nl <= nr;
        }
        T v = (T)0;
        if (live) v = ts[probe == 0 ? nl : (probe == 1 ? nr : nl + (nr - nl) / 2)];
        int k = 0;
        for (int step = 0; step < 3; ++step) {
            if (l > r) { done = true; break; }
            const T tl = __shfl_sync(0xffffffffu, v, k * 3), tr = __shfl_sync(0xffffffffu, v, k * 3 + 1), tm = __shfl_sync(0xffffffffu, v, k * 3 + 2);
            if (tl == target) { res = l; done = true; break; }
            if (tr == target) { res = r; done = true; break; }
            const int64_t mid = l + (r - l) / 2;
            if (tm == target) { res = mid; done = true; break; }
            if (tm < target) { l = mid + 1; k = 2 * k + 2; } else { r = mid - 1; k = 2 * k + 1; }
        }
        if (!done && l > r) done = true;
    }
    if (res == -2) res = right ? r : l;
    if (lane == 0) bounds[e] = right ? res + 1 : res;
}

template <typename T>
__global__ void stack_bounds_kernel(const T *__restrict__ ts, int64_t n, int bins, int64_t *__restrict__ bounds,
                                    const unsigned char *__restrict__ skip)
{
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;        // one warp per boundary
    if (e >= 2 * bins) return;
    if (skip && *skip) { bounds[e] = 0; return; }     // empty slices: the scatter kernel adds nothing
    stack_bounds<T>(ts, n, bins, bounds, e);
}

// Raw (on-disk) timestamps. The early-out `ts.sum() == 0 or len(ts) <= 3` (encodings.py:319-320) on the
// normalised, non-decreasing timestamps is `ts[-1] == ts[0]` (every term is >= 0).
__global__ void stack_bounds_raw_kernel(const double *__restrict__ ts, int64_t n, int bins, int64_t *__restrict__ bounds)
{
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;        // one warp per boundary
    if (e >= 2 * bins) return;
    if (n <= 3 || ts[n - 1] == ts[0]) { bounds[e] = 0; return; }
    stack_bounds<double>(normalized(ts, n), n, bins, bounds, e);
}

// T: coordinate type (float / double tensors, or the on-disk int16, never written back); P: polarity type
// (float, or the on-disk int8 -> `.float()`, h5dataset.py:349). pol_stride / bin_stride: element strides of
// the two leading output dimensions, (bins*plane, plane) for the reference's (2, B, H, W) and
// (plane, 2*plane) for the (B, 2, H, W) the datasets transpose it to.
template <typename T, typename P>
__global__ void events_stack_kernel(T *__restrict__ xs, T *__restrict__ ys, const P *__restrict__ ps,
                                    int64_t n, int bins, int H, int W, const int64_t *__restrict__ bounds,
                                    float *__restrict__ stack, int write_back, int64_t pol_stride, int64_t bin_stride)
{
    extern __shared__ int64_t s_bounds[];
    for (int e = threadIdx.x; e < 2 * bins; e += blockDim.x) s_bounds[e] = bounds[e];
    __syncthreads();
    // The slices are windows of a sorted timestamp array: both their lower and their upper bounds are non-decreasing in b,
    // so the slices that hold event i form one contiguous range of b — and the 32 consecutive events of a warp almost
    // always share it. Two warp-uniform binary searches per 32 events replace a scan of all bins per event
    // (ncu r2f: the scan made the kernel instruction-bound, 8.6 warp instructions per event).
    bool monotone = true;
    for (int b = 1; b < bins; ++b) monotone &= s_bounds[2 * b] >= s_bounds[2 * b - 2] && s_bounds[2 * b + 1] >= s_bounds[2 * b - 1];
    for (int64_t i0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) & ~(int64_t)31; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = i0 + (threadIdx.x & 31);
        int b_lo = 0, b_hi = bins;                   // candidate slices [b_lo, b_hi) of the warp's events i0 .. i0 + 31
        if (monotone) {
            int l = 0, r = bins;                     // first slice whose upper bound exceeds the warp's first event
            while (l < r) { const int m = (l + r) >> 1; if (s_bounds[2 * m + 1] > i0) r = m; else l = m + 1; }
            b_lo = l;
            l = b_lo; r = bins;                      // first slice whose lower bound exceeds the warp's last event
            while (l < r) { const int m = (l + r) >> 1; if (s_bounds[2 * m] > i0 + 31) r = m; else l = m + 1; }
            b_hi = l;
        }
        if (i >= n) continue;
        const T x = xs[i], y = ys[i];
        const bool oob = out_of_range(x, y, H, W);
        const int64_t pix = oob ? 0 : (int64_t)y * W + (int64_t)x;
        const float p = (float)ps[i];
        const float vpos = p * (p < 0.f ? 0.f : p);    // ps * mask_pos, :333-336
        const float vneg = p * (p > 0.f ? 0.f : p);    // ps * mask_neg
        bool seen = false;
        for (int b = b_lo; b < b_hi; ++b) {
            if (i < s_bounds[2 * b] || i >= s_bounds[2 * b + 1]) continue;
            // first slice that holds an out-of-range event: its positive pass sees value 0
            if (!(oob && !seen)) red_add(stack + b * bin_stride + pix, vpos);
            red_add(stack + pol_stride + b * bin_stride + pix, vneg);
            seen = true;
        }
        if (oob && seen && write_back) { xs[i] = (T)0; ys[i] = (T)0; }
    }
}

// flag <- (sum of ts == 0): the other half of events_to_stack's early-out (:319-320). Fixed summation order (per-thread
// strided partials -> block tree -> the last block adds the block partials in index order), so the value compared with
// zero is the same on every run; for the non-negative timestamps of the datasets any order gives the same answer.
constexpr int SUM_BLOCKS = 592, SUM_THREADS = 256;
template <typename T>
__global__ void __launch_bounds__(SUM_THREADS)
events_ts_sum_kernel(const T *__restrict__ ts, int64_t n, double *__restrict__ partial, unsigned *__restrict__ ticket,
                     unsigned char *__restrict__ flag)
{
    __shared__ double sh[SUM_THREADS];
    __shared__ bool last;
    // per-thread partial in the dtype of ts (fp64 adds are scarce on this GPU), four independent chains in flight
    T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    const int64_t stride = (int64_t)gridDim.x * SUM_THREADS;
    int64_t i = blockIdx.x * (int64_t)SUM_THREADS + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) { a0 += ts[i]; a1 += ts[i + stride]; a2 += ts[i + 2 * stride]; a3 += ts[i + 3 * stride]; }
    for (; i < n; i += stride) a0 += ts[i];
    sh[threadIdx.x] = ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
    __syncthreads();
    for (int o = SUM_THREADS / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = sh[0];
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {                                          // the last block adds the block partials: fixed tree again
        __threadfence();
        double t = 0.0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += SUM_THREADS) t += reinterpret_cast<volatile double *>(partial)[b];
        sh[threadIdx.x] = t;
        __syncthreads();
        for (int o = SUM_THREADS / 2; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) { *flag = sh[0] == 0.0 ? 1 : 0; *ticket = 0u; }
    }
}

unsigned grid_for(int64_t n)
{
    const int64_t want = ceil_div(n, (int64_t)256);
    const int64_t cap = (int64_t)ebfi::sm_count() * 16;
    return (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
}

int check_common(const void *xs, const void *ys, int dtype, int64_t n, int H, int W)
{
    EBFI_REQUIRE(dtype == EBFI_F32 || dtype == EBFI_F64, "events: dtype must be EBFI_F32 or EBFI_F64");
    EBFI_REQUIRE(n >= 0 && H > 0 && W > 0, "events: bad sizes n=%lld H=%d W=%d", (long long)n, H, W);
    EBFI_REQUIRE(n == 0 || (xs && ys), "events: null coordinate array");
    return EBFI_OK;
}

}  // namespace

extern "C" {

int ebfi_events_to_image(void *stream, void *xs, void *ys, float *ps, int coord_dtype, int64_t n,
                         int height, int width, float *img, int write_back)
{
    if (int rc = check_common(xs, ys, coord_dtype, n, height, width)) return rc;
    EBFI_REQUIRE(img && (n == 0 || ps), "events_to_image: null pointer");
    if (n == 0) return EBFI_OK;
    cudaStream_t st = ebfi::as_stream(stream);
    if (coord_dtype == EBFI_F32)
        events_image_kernel<float><<<grid_for(n), 256, 0, st>>>((float *)xs, (float *)ys, ps, n, height, width, img, write_back);
    else
        events_image_kernel<double><<<grid_for(n), 256, 0, st>>>((double *)xs, (double *)ys, ps, n, height, width, img, write_back);
    EBFI_LAUNCH_OK("events_image_kernel");
    return EBFI_OK;
}

int ebfi_events_to_mask(void *stream, void *xs, void *ys, float *ps, int coord_dtype, int64_t n,
                        int height, int width, float *img, int64_t *last_index_scratch, int write_back)
{
    if (int rc = check_common(xs, ys, coord_dtype, n, height, width)) return rc;
    EBFI_REQUIRE(img && last_index_scratch && (n == 0 || ps), "events_to_mask: null pointer");
    if (n == 0) return EBFI_OK;
    cudaStream_t st = ebfi::as_stream(stream);
    long long *last = reinterpret_cast<long long *>(last_index_scratch);
    EBFI_CUDA_OK(cudaMemsetAsync(last, 0xFF, (size_t)height * width * sizeof(long long), st));   // -1
    if (coord_dtype == EBFI_F32) {
        events_mask_last_kernel<float><<<grid_for(n), 256, 0, st>>>((float *)xs, (float *)ys, n, height, width, last);
        events_mask_write_kernel<float><<<grid_for(n), 256, 0, st>>>((float *)xs, (float *)ys, ps, n, height, width, last, img, write_back);
    } else {
        events_mask_last_kernel<double><<<grid_for(n), 256, 0, st>>>((double *)xs, (double *)ys, n, height, width, last);
        events_mask_write_kernel<double><<<grid_for(n), 256, 0, st>>>((double *)xs, (double *)ys, ps, n, height, width, last, img, write_back);
    }
    EBFI_LAUNCH_OK("events_mask kernels");
    return EBFI_OK;
}

int ebfi_events_to_voxel(void *stream, void *xs, void *ys, const void *ts, const float *ps, int dtype,
                         int64_t n, int num_bins, int height, int width, float *voxel, int write_back)
{
    if (int rc = check_common(xs, ys, dtype, n, height, width)) return rc;
    EBFI_REQUIRE(num_bins > 0, "events_to_voxel: num_bins must be positive");
    EBFI_REQUIRE(voxel && (n == 0 || (ts && ps)), "events_to_voxel: null pointer");
    if (n == 0) return EBFI_OK;
    cudaStream_t st = ebfi::as_stream(stream);
    if (dtype == EBFI_F32)
        events_voxel_kernel<float><<<grid_for(n), 256, 0, st>>>((float *)xs, (float *)ys, (const float *)ts, ps, n, num_bins, height, width, voxel, write_back);
    else
        events_voxel_kernel<double><<<grid_for(n), 256, 0, st>>>((double *)xs, (double *)ys, (const double *)ts, ps, n, num_bins, height, width, voxel, write_back);
    EBFI_LAUNCH_OK("events_voxel_kernel");
    return EBFI_OK;
}

int ebfi_events_ts_sum_is_zero(void *stream, const void *ts, int dtype, int64_t n, void *scratch, unsigned char *flag)
{
    EBFI_REQUIRE(dtype == EBFI_F32 || dtype == EBFI_F64, "events_ts_sum_is_zero: dtype must be EBFI_F32 or EBFI_F64");
    EBFI_REQUIRE(n >= 0 && flag && scratch && (n == 0 || ts), "events_ts_sum_is_zero: null pointer");
    EBFI_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 7u) == 0, "events_ts_sum_is_zero: scratch must be 8-byte aligned");
    cudaStream_t st = ebfi::as_stream(stream);
    double *partial = static_cast<double *>(scratch);
    unsigned *ticket = reinterpret_cast<unsigned *>(partial + SUM_BLOCKS);
    EBFI_CUDA_OK(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
    const unsigned grid = (unsigned)std::min<int64_t>(SUM_BLOCKS, std::max<int64_t>(1, ceil_div(n, (int64_t)SUM_THREADS)));
    if (dtype == EBFI_F32)
        events_ts_sum_kernel<float><<<grid, SUM_THREADS, 0, st>>>((const float *)ts, n, partial, ticket, flag);
    else
        events_ts_sum_kernel<double><<<grid, SUM_THREADS, 0, st>>>((const double *)ts, n, partial, ticket, flag);
    EBFI_LAUNCH_OK("events_ts_sum_kernel");
    return EBFI_OK;
}

int ebfi_events_to_stack(void *stream, void *xs, void *ys, const void *ts, const float *ps, int dtype,
                         int64_t n, int num_bins, int height, int width, float *stack, int64_t *bounds,
                         int write_back, const unsigned char *skip_flag)
{
    const unsigned char *skip = skip_flag;
    if (int rc = check_common(xs, ys, dtype, n, height, width)) return rc;
    EBFI_REQUIRE(num_bins > 0 && num_bins <= 2048, "events_to_stack: num_bins must be in [1, 2048]");
    EBFI_REQUIRE(stack && bounds && (n == 0 || (ts && ps)), "events_to_stack: null pointer");
    if (n == 0) return EBFI_OK;
    cudaStream_t st = ebfi::as_stream(stream);
    const int nb2 = 2 * num_bins;
    const size_t smem = (size_t)nb2 * sizeof(int64_t);
    const int64_t plane = (int64_t)height * width;
    if (dtype == EBFI_F32) {
        stack_bounds_kernel<float><<<ceil_div(nb2, 2), 64, 0, st>>>((const float *)ts, n, num_bins, bounds, skip);
        events_stack_kernel<float, float><<<grid_for(n), 256, smem, st>>>((float *)xs, (float *)ys, ps, n, num_bins, height, width, bounds, stack, write_back, num_bins * plane, plane);
    } else {
        stack_bounds_kernel<double><<<ceil_div(nb2, 2), 64, 0, st>>>((const double *)ts, n, num_bins, bounds, skip);
        events_stack_kernel<double, float><<<grid_for(n), 256, smem, st>>>((double *)xs, (double *)ys, ps, n, num_bins, height, width, bounds, stack, write_back, num_bins * plane, plane);
    }
    EBFI_LAUNCH_OK("events_stack kernels");
    return EBFI_OK;
}

int ebfi_events_raw_to_stack(void *stream, const int16_t *xs, const int16_t *ys, const double *ts, const int8_t *ps,
                             int64_t n, int num_bins, int height, int width, float *stack, int64_t *bounds,
                             int bins_major)
{
    EBFI_REQUIRE(n >= 0 && height > 0 && width > 0, "events_raw_to_stack: bad sizes n=%lld H=%d W=%d", (long long)n, height, width);
    EBFI_REQUIRE(num_bins > 0 && num_bins <= 2048, "events_raw_to_stack: num_bins must be in [1, 2048]");
    EBFI_REQUIRE(stack && bounds && (n == 0 || (xs && ys && ts && ps)), "events_raw_to_stack: null pointer");
    if (n <= 3) return EBFI_OK;                        // encodings.py:319-320 (and h5dataset.py:332-333 for n == 0)
    cudaStream_t st = ebfi::as_stream(stream);
    const int nb2 = 2 * num_bins;
    const int64_t plane = (int64_t)height * width;
    stack_bounds_raw_kernel<<<ceil_div(nb2, 2), 64, 0, st>>>(ts, n, num_bins, bounds);
    // the raw arrays are read-only: the in-place zeroing of out-of-range events is reproduced in the output only
    events_stack_kernel<int16_t, int8_t><<<grid_for(n), 256, (size_t)nb2 * sizeof(int64_t), st>>>(
        const_cast<int16_t *>(xs), const_cast<int16_t *>(ys), ps, n, num_bins, height, width, bounds, stack, 0,
        bins_major ? plane : num_bins * plane, bins_major ? 2 * plane : plane);
    EBFI_LAUNCH_OK("events_raw_to_stack kernels");
    return EBFI_OK;
}

}  // extern "C"
