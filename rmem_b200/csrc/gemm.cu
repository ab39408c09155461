// t16 GEMM / implicit-GEMM conv, mma.sync.m16n8k16 + ldmatrix + 3-stage cp.async pipeline.
// See gemm.cuh for the contract.
#include "gemm.cuh"

namespace rmem {

namespace {

constexpr int BK = 32;       // k elements per stage
constexpr int BKP = 40;      // padded smem pitch (80 B): conflict-free ldmatrix
constexpr int STAGES = 3;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(s));
}
__device__ __forceinline__ void mma_t16(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32." RMEM_MMA_TYPE "." RMEM_MMA_TYPE ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int BM, int BN, int WARPS_M, int WARPS_N>
__global__ void __launch_bounds__(WARPS_M* WARPS_N * 32) gemm_kernel(const GemmParams p) {
  pdl_prologue();
  constexpr int NT = WARPS_M * WARPS_N * 32;
  constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
  constexpr int MI = WM / 16, NI = WN / 8;
  constexpr int A_ITERS = (BM * 4) / NT, B_ITERS = (BN * 4) / NT;
  static_assert((BM * 4) % NT == 0 && (BN * 4) % NT == 0, "tile/threads mismatch");
  static_assert(NI % 2 == 0, "NI must be even");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  t16* sA = reinterpret_cast<t16*>(smem_raw);              // [STAGES][BM][BKP]
  t16* sB = sA + STAGES * BM * BKP;                         // [STAGES][BN][BKP]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const t16* const Ab = p.A + (long long)blockIdx.z * p.sA;     // batched mode: one problem per blockIdx.z
  const t16* const Bb = p.B + (long long)blockIdx.z * p.sB;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int vec = tid & 3;             // which 8-element k vector of the 32-wide stage
  const int row0 = tid >> 2;           // first tile row handled by this thread

  // per-thread A row bookkeeping (fixed across k)
  const t16* a_base[A_ITERS];
  int a_iy0[A_ITERS], a_ix0[A_ITERS];
  bool a_ok[A_ITERS];
#pragma unroll
  for (int j = 0; j < A_ITERS; ++j) {
    int m = m0 + row0 + j * (NT / 4);
    a_ok[j] = m < p.M;
    if (p.conv) {
      int oy = m / p.Wout, ox = m - oy * p.Wout;
      a_iy0[j] = oy * p.stride - p.pad;
      a_ix0[j] = ox * p.stride - p.pad;
      a_base[j] = Ab;
    } else {
      a_iy0[j] = a_ix0[j] = 0;
      a_base[j] = Ab + (long long)(a_ok[j] ? m : 0) * p.lda;
    }
  }
  const t16* b_base[B_ITERS];
  bool b_ok[B_ITERS];
#pragma unroll
  for (int j = 0; j < B_ITERS; ++j) {
    int n = n0 + row0 + j * (NT / 4);
    b_ok[j] = n < p.N;
    b_base[j] = Bb + (long long)(b_ok[j] ? n : 0) * p.ldb;
  }

  auto load_tile = [&](int kt, int stage) {
    const int k = kt * BK + vec * 8;
    const bool k_ok = k < p.K;
    int ky = 0, kx = 0, ci = 0;
    if (p.conv) {
      int tap = k / p.Cin;
      ci = k - tap * p.Cin;
      ky = tap / p.kw;
      kx = tap - ky * p.kw;
    }
#pragma unroll
    for (int j = 0; j < A_ITERS; ++j) {
      t16* dst = sA + ((stage * BM) + row0 + j * (NT / 4)) * BKP + vec * 8;
      const t16* src = Ab;
      bool ok = a_ok[j] && k_ok;
      if (p.conv) {
        int iy = a_iy0[j] + ky, ix = a_ix0[j] + kx;
        ok = ok && (unsigned)iy < (unsigned)p.Hin && (unsigned)ix < (unsigned)p.Win;
        if (ok) src = Ab + ((long long)iy * p.Win + ix) * p.Cin + ci;
      } else if (ok) {
        src = a_base[j] + k;
      }
      cp_async16(dst, src, ok ? 16 : 0);
    }
#pragma unroll
    for (int j = 0; j < B_ITERS; ++j) {
      t16* dst = sB + ((stage * BN) + row0 + j * (NT / 4)) * BKP + vec * 8;
      bool ok = b_ok[j] && k_ok;
      cp_async16(dst, ok ? (b_base[j] + k) : Bb, ok ? 16 : 0);
    }
  };

  float acc[MI][NI][4];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;

  const int KT = (p.K + BK - 1) / BK;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_tile(s, s);
    cp_async_commit();
  }

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nk = kt + STAGES - 1;
      if (nk < KT) load_tile(nk, nk % STAGES);
      cp_async_commit();
    }
    const int stage = kt % STAGES;
    const t16* tA = sA + stage * BM * BKP;
    const t16* tB = sB + stage * BN * BKP;
#pragma unroll
    for (int ks = 0; ks < BK / 16; ++ks) {
      uint32_t af[MI][4], bfr[NI][2];
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const t16* ptr = tA + (wm * WM + i * 16 + (lane & 15)) * BKP + ks * 16 + (lane >> 4) * 8;
        ldmatrix_x4(af[i][0], af[i][1], af[i][2], af[i][3], ptr);
      }
#pragma unroll
      for (int j = 0; j < NI; j += 2) {
        const t16* ptr = tB + (wn * WN + j * 8 + (lane & 7) + (lane >> 4) * 8) * BKP + ks * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4(bfr[j][0], bfr[j][1], bfr[j + 1][0], bfr[j + 1][1], ptr);
      }
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) mma_t16(acc[i][j], af[i], bfr[j]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue: registers -> global ----
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int i = 0; i < MI; ++i) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = m0 + wm * WM + i * 16 + g + h * 8;
      if (m >= p.M) continue;
      const float bm = (p.bias && p.bias_m) ? p.bias[m] : 0.f;
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int n = n0 + wn * WN + j * 8 + tq * 2;
        if (n >= p.N) continue;
        const bool two = (n + 1) < p.N;
        float v0 = acc[i][j][h * 2 + 0] * p.alpha + bm;
        float v1 = acc[i][j][h * 2 + 1] * p.alpha + bm;
        if (p.bias && !p.bias_m) {
          v0 += p.bias[n];
          if (two) v1 += p.bias[n + 1];
        }
        if (p.res) {
          const t16* r = p.res + (long long)m * p.ldr + n;
          v0 += t2f(r[0]);
          if (two) v1 += t2f(r[1]);
        }
        if (p.act == ACT_RELU) {
          if (n >= p.act_from) v0 = fmaxf(v0, 0.f);
          if (n + 1 >= p.act_from) v1 = fmaxf(v1, 0.f);
        } else if (p.act == ACT_SILU) {
          if (n >= p.act_from) v0 = silu_f(v0);
          if (n + 1 >= p.act_from) v1 = silu_f(v1);
        }
        if (p.gate) {
          const t16* gp = p.gate + (long long)m * p.ldg + n;
          v0 *= t2f(gp[0]);
          if (two) v1 *= t2f(gp[1]);
        }
        // destination (n and n+1 never straddle n_split: n is even, n_split is even)
        void* base = p.C;
        long long ld = p.ldc;
        int nn = n, f32 = p.c_fp32;
        if (n >= p.n_split) {
          base = p.C2; ld = p.ldc2; nn = n - p.n_split; f32 = p.c2_fp32;
        }
        if (f32) {
          float* o = reinterpret_cast<float*>(base) + (long long)blockIdx.z * p.sC + (long long)m * ld + nn;
          if (p.accumulate) {
            v0 += o[0];
            if (two) v1 += o[1];
          }
          o[0] = v0;
          if (two) o[1] = v1;
        } else {
          t16* o = reinterpret_cast<t16*>(base) + (long long)blockIdx.z * p.sC + (long long)m * ld + nn;
          if (two && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) {
            *reinterpret_cast<uint32_t*>(o) = pack2(v0, v1);
          } else {
            o[0] = f2t(v0);
            if (two) o[1] = f2t(v1);
          }
        }
      }
    }
  }
}

template <int BM, int BN, int WARPS_M, int WARPS_N>
int launch_cfg(const GemmParams& p, cudaStream_t stream) {
  constexpr int smem = STAGES * (BM + BN) * BKP * (int)sizeof(t16);
  static bool attr_done = false;
  if (!attr_done) {
    RMEM_CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel<BM, BN, WARPS_M, WARPS_N>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  dim3 grid(cdiv(p.N, BN), cdiv(p.M, BM), p.batch > 1 ? p.batch : 1);
  RMEM_CUDA_CHECK(launch_pdl(gemm_kernel<BM, BN, WARPS_M, WARPS_N>, dim3(grid), dim3(WARPS_M * WARPS_N * 32), smem, stream, p));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

}  // namespace

int& gemm_impl_switch() {
  static thread_local int impl = 0;
  return impl;
}

int gemm_launch(const GemmParams& p, cudaStream_t stream) {
  RMEM_REQUIRE(p.A && p.B && p.C, "gemm: null operand");
  RMEM_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm: empty shape M=%d N=%d K=%d", p.M, p.N, p.K);
  RMEM_REQUIRE(!(p.accumulate && !p.c_fp32), "gemm: accumulate needs an fp32 destination");
  if (p.n_split < p.N) RMEM_REQUIRE(p.C2 != nullptr, "gemm: n_split without C2");
  if (p.conv == 2) {      // stem mode exists on the tcgen05 kernel only
    RMEM_REQUIRE(gemm_tc_supported(p), "gemm(stem): unsupported shape (Cin=%d kw=%d stride=%d K=%d)", p.Cin, p.kw, p.stride, p.K);
    return gemm_tc_launch(p, stream);
  }
  if (gemm_impl_switch() == 0 && p.batch <= 1 && gemm_tc_supported(p)) return gemm_tc_launch(p, stream);
  return gemm_legacy_launch(p, stream);
}

int gemm_legacy_launch(const GemmParams& p, cudaStream_t stream) {
  RMEM_REQUIRE(p.A && p.B && p.C, "gemm: null operand");
  RMEM_REQUIRE(p.nimg == 1, "gemm: stacked images (nimg=%d) exist on the tcgen05 kernel only", p.nimg);
  RMEM_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm: empty shape M=%d N=%d K=%d", p.M, p.N, p.K);
  RMEM_REQUIRE(p.K % 8 == 0, "gemm: K=%d must be a multiple of 8", p.K);
  RMEM_REQUIRE(p.ldb % 8 == 0 && (reinterpret_cast<uintptr_t>(p.B) & 15) == 0, "gemm: B not 16B aligned");
  RMEM_REQUIRE((reinterpret_cast<uintptr_t>(p.A) & 15) == 0, "gemm: A not 16B aligned");
  if (p.conv) {
    RMEM_REQUIRE(p.Cin % 8 == 0, "gemm(conv): Cin=%d must be a multiple of 8", p.Cin);
    RMEM_REQUIRE(p.K % (p.kw * p.Cin) == 0, "gemm(conv): K=%d is not kh*kw*Cin", p.K);
  } else {
    RMEM_REQUIRE(p.lda % 8 == 0, "gemm: lda=%lld must be a multiple of 8", p.lda);
  }
  RMEM_REQUIRE(!(p.accumulate && !p.c_fp32), "gemm: accumulate needs an fp32 destination");
  RMEM_REQUIRE(p.n_split % 2 == 0, "gemm: n_split must be even");
  if (p.n_split < p.N) RMEM_REQUIRE(p.C2 != nullptr, "gemm: n_split without C2");

  // Tile choice: fill the 148 SMs.  Large problems take 128x128, mid 128x64, small 64x64.
  const int nb = p.batch > 1 ? p.batch : 1;
  const long long t128 = (long long)cdiv(p.M, 128) * cdiv(p.N, 128) * nb;
  const long long t12864 = (long long)cdiv(p.M, 128) * cdiv(p.N, 64) * nb;
  if (t128 >= 2 * 148 && p.N >= 128) return launch_cfg<128, 128, 2, 4>(p, stream);
  if (t12864 >= 2 * 148) return launch_cfg<128, 64, 4, 2>(p, stream);
  return launch_cfg<64, 64, 2, 2>(p, stream);
}

}  // namespace rmem
