// Stub standing in for <torch/serialize/tensor.h> when compiling the reference's *_kernel.cu files
// with plain nvcc (oracle/Makefile `ref`): the kernel headers only need at::Tensor as a name.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
namespace at { class Tensor; }
using std::max;
using std::min;
