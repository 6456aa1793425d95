// Stub for <ATen/cuda/CUDAContext.h>; see ref_stubs/torch/serialize/tensor.h.
#pragma once
#include <cuda_runtime.h>
