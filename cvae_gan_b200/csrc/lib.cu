// cvaegan_b200 - single translation unit of the shared library (keeps the __global__ templates in
// gemm.cuh / misc_kernels.cuh defined exactly once).
#include "layout.cu"
#include "mega.cu"
#include "train.cu"
#include "generate.cu"
#include "eval_tc.cu"
#include "api.cu"
