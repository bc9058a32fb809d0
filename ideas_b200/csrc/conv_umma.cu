// tcgen05 / TMA implicit-GEMM convolution (placeholder until the kernels land).
#include "common.cuh"
#include "conv_geom.h"
namespace ideas {
int umma_conv_launch(const ConvGeom&, float*, const float*, const float*, const float*, const float*, int, float, float,
                     cudaStream_t, bool) { return IDEAS_ERR_UNSUPPORTED; }
int umma_wgrad_launch(const ConvGeom&, float*, const float*, const float*, cudaStream_t, bool) { return IDEAS_ERR_UNSUPPORTED; }
}  // namespace ideas
