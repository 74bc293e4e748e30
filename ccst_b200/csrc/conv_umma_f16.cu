// tcgen05 convolution kernels, f16-operand instantiations (see conv_umma_impl.cuh)
#define CCST_INST_BF16 0
#define CCST_INST_F16 1
#include "conv_umma_impl.cuh"
