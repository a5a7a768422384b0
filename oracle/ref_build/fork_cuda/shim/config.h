/* Build configuration for compiling the fork's CUDA FEP kernels in place (oracle/ref_build/fork_cuda):
 * the CPU shim's values with the GPU switches turned on. */
#ifndef FEPFORK_CUDA_CONFIG_H
#define FEPFORK_CUDA_CONFIG_H
#include "../../shim/config.h"
#undef GMX_GPU
#undef GMX_GPU_CUDA
#define GMX_GPU 1
#define GMX_GPU_CUDA 1
#endif
