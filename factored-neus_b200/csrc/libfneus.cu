// Unity build of libfneus_b200.so (one translation unit: the GEMM engine's kernels live in a header).
#include "mlp_fp32.cu"
#include "mlp_relu.cu"
#include "nerf.cu"
#include "sampling.cu"
#include "composite.cu"
#include "pack.cu"
#include "loss.cu"
#include "api.cu"
