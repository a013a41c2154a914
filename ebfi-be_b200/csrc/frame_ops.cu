// frame_ops.cu — the two per-frame "blurry level" maps the model computes on the HOST in every forward
// (SURVEY §8f rank 4): myutils/utils.py:15-49, called from models/Ours/model_singleframe.py:311-326.
//   Frame2Lap  (utils.py:34-49): (im*255).astype(uint8) -> cv2.cvtColor(BGR2GRAY) -> cv2.Laplacian(CV_64F) -> float32
//   Frame2DCP  (utils.py:15-31): min over the 3 channels -> cv2.erode with a sz x sz rectangle (35 x 35)
// The reference moves every frame GPU -> CPU -> GPU for them (a device synchronisation per forward). Both are
// integer / min arithmetic, so the results are bit-identical to OpenCV's:
//   gray  = (3735*c0 + 19235*c1 + 9798*c2 + 2^14) >> 15      (OpenCV >= 4 RGB2Gray<uchar>, channel 0 = "B")
//   lap   = 4-neighbour Laplacian [0 1 0; 1 -4 1; 0 1 0] with BORDER_REFLECT_101
//   erode = window minimum, anchor = sz/2, pixels outside the image ignored (morphologyDefaultBorderValue)
#include "common.cuh"

namespace {

using ebfi::ceil_div;

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// (im * 255).astype(np.uint8): fp32 product, truncation toward zero, low 8 bits (utils.py:44)
__device__ __forceinline__ int to_u8(float v) { return (int)__fmul_rn(v, 255.f) & 0xFF; }

__device__ __forceinline__ int gray_at(const float *__restrict__ im, size_t plane, int y, int x, int W)
{
    const size_t o = (size_t)y * W + x;
    return (3735 * to_u8(__ldg(im + o)) + 19235 * to_u8(__ldg(im + plane + o)) + 9798 * to_u8(__ldg(im + 2 * plane + o)) +
            (1 << 14)) >> 15;
}

__global__ void frame_lap_kernel(const float *__restrict__ frames, float *__restrict__ lap, int B, int H, int W)
{
    const size_t plane = (size_t)H * W, n = (size_t)B * plane;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / plane), y = (int)((i % plane) / W), x = (int)(i % W);
        const float *im = frames + (size_t)b * 3 * plane;
        const int c = gray_at(im, plane, y, x, W);
        const int s = gray_at(im, plane, reflect101(y - 1, H), x, W) + gray_at(im, plane, reflect101(y + 1, H), x, W) +
                      gray_at(im, plane, y, reflect101(x - 1, W), W) + gray_at(im, plane, y, reflect101(x + 1, W), W);
        lap[i] = (float)(s - 4 * c);
    }
}

// pass 1: dark channel (min of the 3 planes) + horizontal window minimum; pass 2: vertical window minimum
__global__ void frame_dcp_rows(const float *__restrict__ frames, float *__restrict__ tmp, int B, int H, int W, int sz)
{
    const size_t plane = (size_t)H * W, n = (size_t)B * plane;
    const int a = sz / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / plane), y = (int)((i % plane) / W), x = (int)(i % W);
        const float *row = frames + (size_t)b * 3 * plane + (size_t)y * W;
        float m = INFINITY;
        for (int xx = max(0, x - a); xx <= min(W - 1, x - a + sz - 1); ++xx)
            m = fminf(m, fminf(fminf(__ldg(row + 2 * plane + xx), __ldg(row + plane + xx)), __ldg(row + xx)));
        tmp[i] = m;
    }
}

__global__ void frame_dcp_cols(const float *__restrict__ tmp, float *__restrict__ dark, int B, int H, int W, int sz)
{
    const size_t plane = (size_t)H * W, n = (size_t)B * plane;
    const int a = sz / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / plane), y = (int)((i % plane) / W), x = (int)(i % W);
        const float *col = tmp + (size_t)b * plane + x;
        float m = INFINITY;
        for (int yy = max(0, y - a); yy <= min(H - 1, y - a + sz - 1); ++yy) m = fminf(m, __ldg(col + (size_t)yy * W));
        dark[i] = m;
    }
}

unsigned grid_for(size_t n)
{
    return (unsigned)std::min<size_t>(ceil_div(n, (size_t)256), (size_t)ebfi::sm_count() * 16);
}

}  // namespace

extern "C" {

int ebfi_frame_to_lap(void *stream, const float *frames, float *lap, int batch, int height, int width)
{
    EBFI_REQUIRE(frames && lap, "frame_to_lap: null pointer");
    EBFI_REQUIRE(batch > 0 && height > 0 && width > 0, "frame_to_lap: bad sizes");
    frame_lap_kernel<<<grid_for((size_t)batch * height * width), 256, 0, ebfi::as_stream(stream)>>>(frames, lap, batch, height, width);
    EBFI_LAUNCH_OK("frame_lap_kernel");
    return EBFI_OK;
}

int ebfi_frame_to_dcp(void *stream, const float *frames, float *dark, float *scratch, int batch, int height, int width,
                      int window)
{
    EBFI_REQUIRE(frames && dark && scratch, "frame_to_dcp: null pointer");
    EBFI_REQUIRE(batch > 0 && height > 0 && width > 0 && window > 0, "frame_to_dcp: bad sizes");
    cudaStream_t st = ebfi::as_stream(stream);
    const size_t n = (size_t)batch * height * width;
    frame_dcp_rows<<<grid_for(n), 256, 0, st>>>(frames, scratch, batch, height, width, window);
    EBFI_LAUNCH_OK("frame_dcp_rows");
    frame_dcp_cols<<<grid_for(n), 256, 0, st>>>(scratch, dark, batch, height, width, window);
    EBFI_LAUNCH_OK("frame_dcp_cols");
    return EBFI_OK;
}

}  // extern "C"
