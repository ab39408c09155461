// TEST INFRASTRUCTURE: stub of the few CUDA runtime names the emulated sources touch (see ../cuda_emu.h).
#pragma once
#include "../cuda_emu.h"
typedef int cudaError_t;
typedef void* cudaStream_t;
#define cudaSuccess 0
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
