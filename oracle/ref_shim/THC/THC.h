// Build-side shim (test infrastructure): models/DCNv2/src/cuda/dcn_v2_im2col_cuda.cu
// includes THC headers it does not use. Empty on purpose.
#pragma once
