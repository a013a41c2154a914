// Test infrastructure: binds the reference's OWN CPU entry points
// (models/DCNv2/src/cpu/vision.h: dcn_v2_cpu_forward / dcn_v2_cpu_backward) under the
// names the reference's Python wrapper expects from `_ext` (models/DCNv2/dcn_v2.py:27,50).
// No reference source is copied; the .cpp files are compiled where they lie.
#include <torch/extension.h>

at::Tensor dcn_v2_cpu_forward(const at::Tensor &, const at::Tensor &, const at::Tensor &,
                              const at::Tensor &, const at::Tensor &, const int, const int,
                              const int, const int, const int, const int, const int, const int,
                              const int);
std::vector<at::Tensor> dcn_v2_cpu_backward(const at::Tensor &, const at::Tensor &,
                                            const at::Tensor &, const at::Tensor &,
                                            const at::Tensor &, const at::Tensor &, int, int, int,
                                            int, int, int, int, int, int);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("dcn_v2_forward", &dcn_v2_cpu_forward, "reference CPU forward");
  m.def("dcn_v2_backward", &dcn_v2_cpu_backward, "reference CPU backward");
}
