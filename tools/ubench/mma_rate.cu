// Micro-benchmark: tcgen05.mma issue/throughput/latency for the shapes the attention kernel uses (1-CTA, fp16, SW128).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I rmem_b200/csrc tools/ubench/mma_rate.cu -o gpurun_out/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tcgen05.cuh"
using namespace rmem::tc;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: SS M128 N=NN, nmma per group ; mode 1: TS
template <int NN, int MODE>
__global__ void __launch_bounds__(128) k(long long* out, int groups, int nmma) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_before(); __syncthreads(); fence_after();
  const uint32_t tmem = slot;
  fence_async_smem();
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = make_idesc(128, NN);
    const uint64_t da = make_desc_sw128(smem_u32(smem)), db = make_desc_sw128(smem_u32(smem + 16384));
    long long t0 = clock64(), t_issue = 0;
    if (elect_one()) {
      for (int g = 0; g < groups; ++g) {
        for (int i = 0; i < nmma; ++i) {
          if (MODE == 0) umma_ss(tmem, da + (uint64_t)((i & 3) * 2), db + (uint64_t)((i & 3) * 2), idesc, 1u);
          else umma_ts(tmem, tmem + 384 + (i & 3) * 8, db + (uint64_t)((i & 3) * 2), idesc, 1u);
        }
      }
      t_issue = clock64();
      commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0, nullptr, 1);
    long long t1 = clock64();
    if (threadIdx.x == 0 || t_issue) {
      if (t_issue) { out[blockIdx.x * 2 + 0] = t_issue - t0; out[blockIdx.x * 2 + 1] = t1 - t0; }
    }
  }
  fence_before(); __syncthreads();
  if (threadIdx.x < 32) { fence_after(); tmem_dealloc<512>(tmem); }
}

template <int NN, int MODE>
void run(const char* name, int grid, int groups, int nmma) {
  long long* d; cudaMalloc(&d, grid * 16); cudaMemset(d, 0, grid * 16);
  cudaFuncSetAttribute(k<NN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  k<NN, MODE><<<grid, 128, 60000>>>(d, groups, nmma);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  int n = groups * nmma;
  printf("%-28s grid=%3d mmas=%5d issue=%8lld cyc total=%8lld cyc  -> %7.1f cyc/mma  (ideal %d)  %s\n", name, grid, n, h[0], h[1],
         (double)h[1] / n, 128 * NN / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<64, 0>("SS N=64", grid, 1, 8);
    run<64, 0>("SS N=64", grid, 64, 8);
    run<128, 0>("SS N=128", grid, 64, 8);
    run<256, 0>("SS N=256", grid, 1, 4);
    run<256, 0>("SS N=256", grid, 64, 4);
    run<256, 1>("TS N=256", grid, 1, 4);
    run<256, 1>("TS N=256", grid, 64, 4);
  }
  return 0;
}
