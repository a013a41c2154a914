// Test infrastructure: host driver for the reference's UNMODIFIED CUDA kernels
// (models/DCNv2/src/cuda/dcn_v2_im2col_cuda.cu, compiled where it lies). The reference's
// own driver, src/cuda/dcn_v2_cuda.cu, does not compile against torch >= 1.11 (THCState,
// THArgCheck, :11,:110); this file restates its call sequence (:64-95, :138-211) with ATen
// so the reference kernels can be timed and compared on the B200 box.
#include <torch/extension.h>
#include <c10/cuda/CUDAStream.h>
#include "cuda/dcn_v2_im2col_cuda.h"

namespace {
struct Geo { int B, C, H, W, Co, Ho, Wo; };
Geo geo(const at::Tensor &x, const at::Tensor &w, int kh, int kw, int sh, int sw, int ph, int pw,
        int dh, int dw) {
  Geo g{(int)x.size(0), (int)x.size(1), (int)x.size(2), (int)x.size(3), (int)w.size(0), 0, 0};
  g.Ho = (g.H + 2 * ph - (dh * (kh - 1) + 1)) / sh + 1;
  g.Wo = (g.W + 2 * pw - (dw * (kw - 1) + 1)) / sw + 1;
  return g;
}
}  // namespace

at::Tensor ref_forward(const at::Tensor &x, const at::Tensor &w, const at::Tensor &bias,
                       const at::Tensor &off, const at::Tensor &msk, int kh, int kw, int sh,
                       int sw, int ph, int pw, int dh, int dw, int dg) {
  const Geo g = geo(x, w, kh, kw, sh, sw, ph, pw, dh, dw);
  auto cols = at::empty({g.B, g.C * kh * kw, g.Ho * g.Wo}, x.options());
  modulated_deformable_im2col_cuda(c10::cuda::getCurrentCUDAStream(), x.data_ptr<float>(),
                                   off.data_ptr<float>(), msk.data_ptr<float>(), g.B, g.C, g.H, g.W,
                                   g.Ho, g.Wo, kh, kw, ph, pw, sh, sw, dh, dw, dg,
                                   cols.data_ptr<float>());
  auto out = at::matmul(w.view({g.Co, g.C * kh * kw}), cols).view({g.B, g.Co, g.Ho, g.Wo});
  return out + bias.view({1, g.Co, 1, 1});
}

std::vector<at::Tensor> ref_backward(const at::Tensor &x, const at::Tensor &w,
                                     const at::Tensor &bias, const at::Tensor &off,
                                     const at::Tensor &msk, const at::Tensor &gout, int kh, int kw,
                                     int sh, int sw, int ph, int pw, int dh, int dw, int dg) {
  const Geo g = geo(x, w, kh, kw, sh, sw, ph, pw, dh, dw);
  auto gx = at::zeros_like(x), gw = at::zeros_like(w), gb = at::zeros_like(bias);
  auto goff = at::zeros_like(off), gmsk = at::zeros_like(msk);
  auto wt = w.view({g.Co, g.C * kh * kw}).t();
  auto stream = c10::cuda::getCurrentCUDAStream();
  for (int b = 0; b < g.B; ++b) {
    auto go_b = gout.select(0, b).reshape({g.Co, g.Ho * g.Wo});
    auto cols = at::matmul(wt, go_b).contiguous();
    auto xb = x.select(0, b), ob = off.select(0, b), mb = msk.select(0, b);
    modulated_deformable_col2im_coord_cuda(stream, cols.data_ptr<float>(), xb.data_ptr<float>(),
                                           ob.data_ptr<float>(), mb.data_ptr<float>(), 1, g.C, g.H,
                                           g.W, g.Ho, g.Wo, kh, kw, ph, pw, sh, sw, dh, dw, dg,
                                           goff.select(0, b).data_ptr<float>(),
                                           gmsk.select(0, b).data_ptr<float>());
    modulated_deformable_col2im_cuda(stream, cols.data_ptr<float>(), ob.data_ptr<float>(),
                                     mb.data_ptr<float>(), 1, g.C, g.H, g.W, g.Ho, g.Wo, kh, kw, ph,
                                     pw, sh, sw, dh, dw, dg, gx.select(0, b).data_ptr<float>());
    modulated_deformable_im2col_cuda(stream, xb.data_ptr<float>(), ob.data_ptr<float>(),
                                     mb.data_ptr<float>(), 1, g.C, g.H, g.W, g.Ho, g.Wo, kh, kw, ph,
                                     pw, sh, sw, dh, dw, dg, cols.data_ptr<float>());
    gw += at::matmul(go_b, cols.t()).view_as(w);
    gb += go_b.sum(1);
  }
  return {gx, goff, gmsk, gw, gb};
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("dcn_v2_forward", &ref_forward, "reference CUDA kernels, forward");
  m.def("dcn_v2_backward", &ref_backward, "reference CUDA kernels, backward");
}
