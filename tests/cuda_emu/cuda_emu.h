// TEST INFRASTRUCTURE: a minimal host emulation of the CUDA execution model, enough to run the simple (non-tensor-core)
// kernels of rmem_b200/csrc on the CPU, unmodified, in the build container that has no GPU.  tests/cuda_emu/build.py
// rewrites `kernel<<<grid, block, 0, stream>>>(args)` into emu::launch(kernel, grid, block, args) and compiles the file
// with g++ against this header and the stub <cuda_runtime.h> / <cuda_fp16.h> / <cuda_bf16.h> next to it.
//
// Model: the blocks of a grid run one after the other; the threads of a block are real std::threads, so __syncthreads
// is a barrier over the block and the warp primitives (__shfl_xor_sync, __ballot_sync) exchange through a per-warp
// buffer between two warp barriers -- the lock-step the kernels rely on.  __shared__ becomes `static` (one block at a
// time).  A thread that returns early leaves both barriers (arrive_and_drop), as on the hardware.  "Device" memory is host
// memory.  Nothing here is linked into the product.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)

struct uint3 { unsigned int x, y, z; };
struct dim3 {
  unsigned int x, y, z;
  dim3(unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) uint4 { unsigned int x, y, z, w; };
struct alignas(8) uint2 { unsigned int x, y; };
inline uint4 make_uint4(unsigned int x, unsigned int y, unsigned int z, unsigned int w) { return uint4{x, y, z, w}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
#define __align__(n) alignas(n)

namespace emu {
struct Warp {
  std::barrier<> bar;
  uint64_t buf[32];
  explicit Warp(int n) : bar(n) {}
};
struct Ctx {
  std::barrier<>* block = nullptr;
  Warp* warp = nullptr;
  int lane = 0, warp_lanes = 32;
  unsigned char* dyn_smem = nullptr;      // `extern __shared__` of the running block (build.py rewrites the declaration)
};
inline thread_local Ctx ctx;
}  // namespace emu

inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

inline void __syncthreads() { emu::ctx.block->arrive_and_wait(); }

template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int o) {
  static_assert(sizeof(T) <= 8, "shuffle of up to 8 bytes");
  emu::Warp* w = emu::ctx.warp;
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  w->buf[emu::ctx.lane] = raw;
  w->bar.arrive_and_wait();
  const int src = emu::ctx.lane ^ o;
  uint64_t got = src < emu::ctx.warp_lanes ? w->buf[src] : raw;
  w->bar.arrive_and_wait();
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
inline unsigned int __ballot_sync(unsigned, int pred) {
  emu::Warp* w = emu::ctx.warp;
  w->buf[emu::ctx.lane] = pred ? 1u : 0u;
  w->bar.arrive_and_wait();
  unsigned int b = 0;
  for (int l = 0; l < emu::ctx.warp_lanes; ++l) b |= (unsigned int)(w->buf[l] & 1u) << l;
  w->bar.arrive_and_wait();
  return b;
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, int o) {
  static_assert(sizeof(T) <= 8, "shuffle of up to 8 bytes");
  emu::Warp* w = emu::ctx.warp;
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  w->buf[emu::ctx.lane] = raw;
  w->bar.arrive_and_wait();
  const int src = emu::ctx.lane + o;
  uint64_t got = src < emu::ctx.warp_lanes ? w->buf[src] : raw;
  w->bar.arrive_and_wait();
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
inline unsigned int __match_any_sync(unsigned, unsigned int v) {
  emu::Warp* w = emu::ctx.warp;
  w->buf[emu::ctx.lane] = v;
  w->bar.arrive_and_wait();
  unsigned int peers = 0;
  for (int l = 0; l < emu::ctx.warp_lanes; ++l) peers |= (unsigned int)(w->buf[l] == (uint64_t)v) << l;
  w->bar.arrive_and_wait();
  return peers;
}
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
template <typename T> inline T __ldcg(const T* p) { return *p; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __popc(unsigned int v) { return __builtin_popcount(v); }
inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned int __float_as_uint(float f) { unsigned int u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned int u) { float f; std::memcpy(&f, &u, 4); return f; }
// compiled with -ffp-contract=off: plain operators are the round-to-nearest, unfused operations
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __expf(float x) { return expf(x); }
inline float rsqrtf(float x) { return 1.f / sqrtf(x); }
template <typename T> inline T min(T a, T b) { return a < b ? a : b; }
template <typename T> inline T max(T a, T b) { return a > b ? a : b; }
inline long long min(long long a, int b) { return a < b ? a : b; }
inline long long min(int a, long long b) { return a < b ? a : b; }
inline unsigned int min(unsigned int a, int b) { return a < (unsigned int)b ? a : (unsigned int)b; }
inline unsigned int min(int a, unsigned int b) { return (unsigned int)a < b ? (unsigned int)a : b; }

namespace emu {
// One pool of block-size threads per launch; the blocks of the grid run one after the other on it.  The block / warp
// barriers are rebuilt between two blocks (a thread that returned early has dropped out of them), framed by a pool-wide
// barrier that nobody ever leaves.
template <typename... KArgs, typename... Args>
inline void launch_smem(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem_bytes, Args... args) {
  std::vector<unsigned char> dyn(smem_bytes + 16);
  unsigned char* dyn_base = dyn.data() + ((16 - reinterpret_cast<uintptr_t>(dyn.data()) % 16) % 16);
  const int nthreads = (int)(block.x * block.y * block.z);
  const int nwarps = (nthreads + 31) / 32;
  const long long nblocks = (long long)grid.x * grid.y * grid.z;
  std::barrier<> pool_bar(nthreads);
  std::unique_ptr<std::barrier<>> block_bar;
  std::vector<std::unique_ptr<Warp>> warps(nwarps);
  auto rebuild = [&] {
    block_bar.reset(new std::barrier<>(nthreads));
    for (int w = 0; w < nwarps; ++w) warps[w].reset(new Warp(std::min(32, nthreads - 32 * w)));
  };
  rebuild();
  std::vector<std::thread> th;
  th.reserve(nthreads);
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&, t] {
      threadIdx = {(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
      blockDim = block;
      gridDim = grid;
      ctx.lane = t % 32;
      ctx.warp_lanes = std::min(32, nthreads - 32 * (t / 32));
      ctx.dyn_smem = dyn_base;
      for (long long b = 0; b < nblocks; ++b) {
        blockIdx = {(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((long long)grid.x * grid.y))};
        ctx.block = block_bar.get();
        ctx.warp = warps[t / 32].get();
        kernel(static_cast<KArgs>(args)...);
        ctx.warp->bar.arrive_and_drop();
        ctx.block->arrive_and_drop();
        pool_bar.arrive_and_wait();            // every thread is out of this block
        if (t == 0) rebuild();
        pool_bar.arrive_and_wait();            // fresh barriers are in place
      }
    });
  for (auto& x : th) x.join();
}
template <typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, Args... args) {
  launch_smem(kernel, grid, block, 0, args...);
}
}  // namespace emu

// the launcher of common.cuh (programmatic dependent launch) and its device-side prologue: plain launches here
namespace rmem {
inline void pdl_prologue() {}
inline bool& pdl_enabled() { static thread_local bool on = true; return on; }
template <typename... KArgs, typename... Args>
inline int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, void*, Args&&... args) {
  emu::launch_smem(kernel, grid, block, smem, static_cast<KArgs>(args)...);
  return 0;
}
}  // namespace rmem
