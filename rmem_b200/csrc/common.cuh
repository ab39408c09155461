// Shared device/host helpers for the rmem_b200 kernels (sm_100a only).
#pragma once
#include <cstdlib>
#include <exception>
#include <new>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

// 16-bit tensor-core operand type of the whole path.  fp16 by default: same tcgen05 / mma.sync rate as bf16 with
// 3 more mantissa bits, which is what keeps the logits within tolerance of the fp32 reference through ~90 layers
// (the reference's own reduced-precision mode, `--amp`, is fp16 autocast as well).  -DRMEM_OPERAND_BF16 switches.
#ifdef RMEM_OPERAND_BF16
typedef __nv_bfloat16 t16;
typedef __nv_bfloat162 t162;
#define RMEM_OPERAND_NAME "bf16"
#define RMEM_MMA_TYPE "bf16"
#define RMEM_UMMA_FORMAT 1u
#else
typedef __half t16;
typedef __half2 t162;
#define RMEM_OPERAND_NAME "fp16"
#define RMEM_MMA_TYPE "f16"
#define RMEM_UMMA_FORMAT 0u
#endif

namespace rmem {

// ---- error plumbing (C-ABI never throws / exits; see include/rmem_b200.h) ----
void set_error(const char* fmt, ...);
const char* get_error();

#define RMEM_OK 0
#define RMEM_ERR_ARG (-1)
#define RMEM_ERR_CUDA (-2)
#define RMEM_ERR_WEIGHT (-3)
#define RMEM_ERR_ARENA (-4)
#define RMEM_ERR_STATE (-5)

#define RMEM_CUDA_CHECK(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      rmem::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return RMEM_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

// Every kernel launch goes through this macro; the thread-local counter backs bench.py's gpu_launches.
long long& launch_counter();
#define RMEM_LAUNCH_CHECK()                 \
  do {                                      \
    ++rmem::launch_counter();               \
    RMEM_CUDA_CHECK(cudaGetLastError());    \
  } while (0)

#define RMEM_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      rmem::set_error(__VA_ARGS__);    \
      return RMEM_ERR_ARG;             \
    }                                  \
  } while (0)

// Every multi-line entry point is wrapped: no C++ exception (std::bad_alloc from the host-side containers, ...) may cross
// the extern "C" boundary; it becomes a status code + message like every other failure.
#define RMEM_API_BEGIN try {
#define RMEM_API_END                                                              \
  }                                                                               \
  catch (const std::bad_alloc&) {                                                 \
    rmem::set_error("out of host memory");                                        \
    return RMEM_ERR_STATE;                                                        \
  }                                                                               \
  catch (const std::exception& ex) {                                              \
    rmem::set_error("unexpected C++ exception: %s", ex.what());                   \
    return RMEM_ERR_STATE;                                                        \
  }                                                                               \
  catch (...) {                                                                   \
    rmem::set_error("unexpected C++ exception");                                  \
    return RMEM_ERR_STATE;                                                        \
  }

#define RMEM_TRY(expr)          \
  do {                          \
    int _rc = (expr);           \
    if (_rc != RMEM_OK) return _rc; \
  } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return cdiv(a, b) * b; }

// ---- device helpers ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32). `red` needs 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
    r = warp_sum(r);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  return red[0];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : -INFINITY;
  if (w == 0) {
    r = warp_max(r);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  return red[0];
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }

#ifdef RMEM_OPERAND_BF16
__device__ __forceinline__ t16 f2t(float x) { return __float2bfloat16(x); }
__device__ __forceinline__ float t2f(t16 x) { return __bfloat162float(x); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  t162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  t162 v = *reinterpret_cast<t162*>(&u);
  return __bfloat1622float2(v);
}
#else
// fp16 stores saturate at +-65504 instead of overflowing to inf
__device__ __forceinline__ float sat16(float x) { return fminf(fmaxf(x, -65504.f), 65504.f); }
__device__ __forceinline__ t16 f2t(float x) { return __float2half_rn(sat16(x)); }
__device__ __forceinline__ float t2f(t16 x) { return __half2float(x); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  t162 v = __floats2half2_rn(sat16(lo), sat16(hi));
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  t162 v = *reinterpret_cast<t162*>(&u);
  return __half22float2(v);
}
#endif

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_SILU = 2 };

// ---- programmatic dependent launch (sm_90+) ----
// Every kernel of the path starts with pdl_prologue(): it lets the NEXT kernel in the stream be scheduled early (its
// launch latency and prologue overlap this kernel), then blocks until the PREVIOUS kernel has completed and its memory
// is visible.  Nothing may touch global memory before it.  Kernels are launched with launch_pdl() (the PDL attribute);
// a kernel launched without it simply behaves as usual.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_launch_dependents();
  pdl_wait();
}
// Per-thread switch: the engine issues the prefetched encoder (side stream) without the attribute -- an early-launched
// grid holds shared memory and TMEM on its SMs while it waits, which starves the stream that owns the critical path.
inline bool& pdl_enabled() {
  static thread_local bool on = [] { const char* e = getenv("RMEM_NO_PDL"); return !(e && e[0] == '1'); }();
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace rmem
