// tcgen05 convolution kernels, bf16-operand instantiations (see conv_umma_impl.cuh)
#define CCST_INST_BF16 1
#define CCST_INST_F16 0
#include "conv_umma_impl.cuh"
