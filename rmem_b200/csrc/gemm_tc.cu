// tcgen05 GEMM / implicit-GEMM convolution for sm_100a with the fused epilogue contract of gemm.cuh:
//   C[M,N] = epilogue( alpha * A[M,K] . B[N,K]^T ),  fp16 operands, fp32 accumulation in TMEM.
//
//   block = 320 threads:  warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-9 = epilogue
//   tile  = 128 (M) x BN (64 | 128 | 256) x 64 (K, one 128-byte swizzle atom), STAGES-deep mbarrier ring
//   A tile: linear mode  -> 2-D TMA box [64 k, 128 rows] of the row-major matrix
//           conv mode    -> 3-D TMA box [64 ch, BW, BH] (BW*BH = 128 output pixels of a rectangular patch) of the
//                           NHWC map at the tap's offset; out-of-bounds pixels (the padding halo and the ragged tile
//                           edge) are zero-filled by TMA, the traversal stride is the conv stride.  One k-block per
//                           (tap, 64-channel slice): the im2col matrix is never materialised.
//   B tile: 2-D TMA box [64 k, BN rows] of the [N,K] weight.
//   Epilogue: tcgen05.ld 32 lanes x 32 columns -> bias / residual / activation / gate / split / accumulate -> global
//   (each thread owns an output row: 64-128 contiguous bytes per store burst).
//   Split-K: a GEMM with few output tiles and a long K (layer3's 3x3 convolutions: 56 tiles x 36 k-blocks) is bound by
//   the ~40 B/clk one SM can pull from L2, not by the tensor pipe.  Such a launch runs as clusters of S = 2 | 4 CTAs
//   along K: every CTA accumulates its k-range in its own TMEM, ranks 1.. push their fp32 partial tile into rank 0's
//   shared memory (DSMEM stores, column-major so lanes are contiguous) and rank 0 reduces in rank order inside its
//   normal epilogue -- no global workspace, no atomics, bit-reproducible.
#include "gemm.cuh"
#include "tcgen05.cuh"

#include <cstdlib>

namespace rmem {

namespace {

using namespace tc;

constexpr int TBM = 128;
constexpr int TBK = 64;
constexpr int kEpiWarps = 8;                  // two per TMEM lane quadrant, splitting the tile's columns
constexpr int kTcThreads = (2 + kEpiWarps) * 32;
constexpr int SMEM_A_STAGE = TBM * TBK * 2;   // 16 KB

struct TcGemmParams {
  int M, N, nk, stages;
  int splitk;                 // CTAs per cluster along K (1 = no split)
  int conv, taps_w, cin_blocks, pad, stride, BW, BH, tiles_x, Hout, Wout;
  int nimg, tiles_img;        // conv modes: images stacked along M (4-D tensor map), row tiles per image
  int sx, sy;                 // coordinate steps of the A tensor map per output pixel (conv: stride, stride; stem: 1, 2)
  float alpha;
  const float* bias;
  int bias_m, act, act_from;
  const t16* res;
  long long ldr;
  const t16* gate;
  long long ldg;
  int accumulate;
  void* C;
  long long ldc;
  int c_fp32;
  void* C2;
  long long ldc2;
  int c2_fp32, n_split;
  int* err;
  // GroupNorm statistics of the output, produced by the epilogue (the FPN decoder's conv -> GroupNorm pairs): per
  // (row tile, 32-column chunk, lane quadrant) fp32 partial sums of the fp32 results, folded in a fixed order and in double
  // by the last CTA to finish -> gn_stats[2 * group] = (sum, sum of squares).  gn_cpg = channels per group (16 | 32).
  // TMA-store epilogue (one t16 destination): the tile is staged in the (by then idle) operand ring as 64-column panels in
  // the 128-byte swizzle and leaves as full 128-byte lines through map_c; rows / columns outside the output are dropped by
  // the tensor map.  Replaces 16-byte stores at a row stride per lane (half-used sectors, partial-line writes).
  int tma_store;
  double* gn_stats;
  float4* gn_part;
  unsigned int* gn_counter;
  int gn_cpg;
};

constexpr int kMaxStages = 6;
// Optional per-CTA event trace (clock64; 8 slots per CTA, first 64 CTAs); null in production.
__device__ long long* g_gemm_trace = nullptr;
#define GTRACE(k)                                                                                  \
  do {                                                                                             \
    if (gtrace && (threadIdx.x & 31) == 0) gtrace[(blockIdx.y * gridDim.x + blockIdx.x) * 8 + (k)] = clock64(); \
  } while (0)
template <int BN>
struct TcSmem {
  static constexpr int kB = BN * TBK * 2;
  static constexpr int kStage = SMEM_A_STAGE + kB;
  // [1024B-aligned] stages x (A | B) tiles, then the barriers
  static constexpr int kRecv = TBM * BN * 4;     // one fp32 partial tile (split-K)
  // + barriers + alignment slack; split-K adds (S-1) receive tiles behind the barrier area
  static constexpr int total(int stages, int splitk = 1) { return stages * kStage + 256 + 1024 + (splitk - 1) * kRecv; }
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  long long spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (!done && ++spins > (1ll << 26)) __trap();
  }
}

__device__ __forceinline__ void load8h(const t16* p, float* v) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack2(u.x), b = unpack2(u.y), c = unpack2(u.z), d = unpack2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

template <int BN>
__global__ void __launch_bounds__(kTcThreads, BN <= 128 ? 2 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_c, const TcGemmParams p) {
  using L = TcSmem<BN>;
  const int STAGES = p.stages;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * L::kStage);
  uint64_t* full = bars;                  // [STAGES]
  uint64_t* empty = bars + STAGES;        // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES;
  uint64_t* recv_full = bars + 2 * STAGES + 1;   // split-K: every epilogue warp of every peer has pushed its partial
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);
  float* recv = reinterpret_cast<float*>(smem + STAGES * L::kStage + 256);   // [(S-1)][BN][128] fp32, rank 0 only
  const int S = BN == 64 ? p.splitk : 1;         // split-K exists for the narrow tile only (compiled out elsewhere)
  const uint32_t krank = S > 1 ? cluster_ctarank() : 0u;
  // this CTA's k-blocks: [kb0, kb0 + nkl)
  const int kq = p.nk / S, kr = p.nk - kq * S;
  const int kb0 = (int)krank * kq + ((int)krank < kr ? (int)krank : kr);
  const int nkl = kq + ((int)krank < kr ? 1 : 0);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* const gtrace = (blockIdx.y * gridDim.x + blockIdx.x) < 64 ? g_gemm_trace : nullptr;
  if (warp == 0) GTRACE(0);
  const int n0 = blockIdx.x * BN;
  int m0 = blockIdx.y * TBM, x0 = 0, y0 = 0, img = 0;
  if (p.conv) {
    int t = blockIdx.y;
    if (p.nimg > 1) { img = t / p.tiles_img; t -= img * p.tiles_img; }
    const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
    x0 = tx * p.BW;
    y0 = ty * p.BH;
  }

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(recv_full, (S - 1) * kEpiWarps);
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.tma_store) tma_prefetch_desc(&map_c);
  }
  if (warp == 1) tmem_alloc<BN>(tmem_slot);
  __shared__ __align__(16) float s_bias[BN];
  fence_before();
  __syncthreads();
  fence_after();
  if (S > 1) cluster_sync_all();   // rank 0's receive barrier is initialised before any peer can signal it
  const uint32_t tmem = *tmem_slot;
  // Programmatic dependent launch: the next kernel in the stream may start its own prologue now (this CTA already
  // holds its TMEM), and nothing above touched global memory, so the previous kernel's tail overlapped our prologue.
  pdl_prologue();
  if (warp == 0) GTRACE(1);

  if (warp == 0) {
    // ================================ TMA producer ================================
    for (int kl = 0; kl < nkl; ++kl) {
      const int st = kl % STAGES, kb = kb0 + kl;
      if (kl >= STAGES) mbar_wait(&empty[st], ((kl / STAGES) - 1) & 1, p.err, 1);
      if (elect_one()) {
        unsigned char* sa = smem + st * L::kStage;
        unsigned char* sb = sa + SMEM_A_STAGE;
        mbar_expect_tx(&full[st], L::kStage);
        if (p.conv) {
          const int tap = kb / p.cin_blocks, cb = kb - tap * p.cin_blocks;
          const int ky = tap / p.taps_w, kx = tap - ky * p.taps_w;
          if (p.nimg > 1)
            tma_load_4d(sa, &map_a, &full[st], cb * TBK, x0 * p.sx - p.pad + kx, y0 * p.sy - p.pad + ky, img);
          else
            tma_load_3d(sa, &map_a, &full[st], cb * TBK, x0 * p.sx - p.pad + kx, y0 * p.sy - p.pad + ky);
        } else {
          tma_load_2d(sa, &map_a, &full[st], kb * TBK, m0);
        }
        tma_load_2d(sb, &map_b, &full[st], kb * TBK, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc = make_idesc(TBM, BN);
    const uint32_t smem_base = smem_u32(smem);
    for (int kb = 0; kb < nkl; ++kb) {
      const int st = kb % STAGES;
      mbar_wait(&full[st], (kb / STAGES) & 1, p.err, 2);
      fence_after();
      if (kb == 0) GTRACE(2);
      if (kb == nkl - 1) GTRACE(3);
      if (elect_one()) {
        const uint32_t a_addr = smem_base + st * L::kStage;
        const uint64_t da = make_desc_sw128(a_addr), db = make_desc_sw128(a_addr + SMEM_A_STAGE);
#pragma unroll
        for (int kk = 0; kk < TBK / 16; ++kk)
          umma_ss(tmem, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (kb > 0 || kk > 0) ? 1u : 0u);
        commit(&empty[st]);
        if (kb == nkl - 1) commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // ================================ epilogue (warps 2-9) ================================
    // A single warp per scheduler runs the ~150-instruction chunk body at ~6 cycles per instruction (event trace,
    // profiles/r01), so the tile's columns are split over two warps per TMEM lane quadrant.
    // per-column bias of this tile -> shared memory, read back as a broadcast (weights are not produced by the
    // preceding kernel, but the PDL wait has passed anyway)
    for (int i = threadIdx.x - 64; i < BN; i += kEpiWarps * 32)
      s_bias[i] = (p.bias && !p.bias_m && n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;                // which half of the tile's columns
    const int r = quad * 32 + lane;                  // tile row
    long long gm;
    bool valid;
    if (p.conv) {
      const int yy = y0 + r / p.BW, xx = x0 + r % p.BW;
      valid = yy < p.Hout && xx < p.Wout;
      gm = ((long long)img * p.Hout + yy) * p.Wout + xx;
    } else {
      gm = m0 + r;
      valid = gm < p.M;
    }
    const float bias_row = (p.bias && p.bias_m && valid) ? p.bias[gm] : 0.f;
    // Prefetch this row's residual for the whole tile while the main loop runs (full 32-column chunks only): the
    // epilogue would otherwise serialise one L2 round trip per chunk.
    constexpr bool kPre = BN == 128;
    constexpr int kHalfCols = BN / 2;
    const int cbeg = half * kHalfCols;
    uint4 rpre[kPre ? kHalfCols / 8 : 1];
    const bool res_pre = kPre && p.res != nullptr && valid;
    if (res_pre) {
      const t16* rp = p.res + gm * p.ldr + n0 + cbeg;
#pragma unroll
      for (int c = 0; c < kHalfCols / 8; ++c)
        if (n0 + cbeg + c * 8 < p.N) rpre[c] = *reinterpret_cast<const uint4*>(rp + c * 8);
    }
    mbar_wait(acc_full, 0, p.err, 3);
    fence_after();
    if (warp == 2) GTRACE(4);
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    if (krank != 0) {
      // ---- split-K peer: fp32 partial tile -> rank 0's receive buffer [BN][128] (lanes = consecutive rows) ----
      const uint32_t dst0 = map_to_cta(smem_u32(recv + ((size_t)(krank - 1) * BN) * TBM + r), 0);
#pragma unroll 1
      for (int cc = 0; cc < kHalfCols; cc += 32) {
        const int c0 = cbeg + cc;
        if (n0 + c0 >= p.N) break;
        float v[32];
        tmem_ld32(lane_addr + c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) st_cluster_f32(dst0 + (uint32_t)((c0 + j) * TBM * 4), v[j]);
      }
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(recv_full), 0));
    } else {
    if (S > 1) mbar_wait_cluster(recv_full, 0);
    auto chunk = [&](const int c0) {
      const int n = n0 + c0;
      if (n >= p.N) return;                          // warp-uniform
      // destination of this 32-column chunk (n_split is a multiple of 32)
      void* base = p.C;
      long long ld = p.ldc;
      int nn = n, f32 = p.c_fp32;
      if (n >= p.n_split) { base = p.C2; ld = p.ldc2; nn = n - p.n_split; f32 = p.c2_fp32; }
      // ---- issue every global load of the chunk first, so their L2 round trips overlap ----
      uint4 rr[4], gg[4];
      const bool do_res = p.res != nullptr && valid, do_gate = p.gate != nullptr && valid;
      const bool do_acc = f32 && p.accumulate && valid;
      if (do_res && !kPre) {
        const t16* rp = p.res + gm * p.ldr + n;
#pragma unroll
        for (int j = 0; j < 4; ++j) rr[j] = *reinterpret_cast<const uint4*>(rp + j * 8);
      }
      if (do_gate) {
        const t16* gp = p.gate + gm * p.ldg + n;
#pragma unroll
        for (int j = 0; j < 4; ++j) gg[j] = *reinterpret_cast<const uint4*>(gp + j * 8);
      }
      float v[32];
      tmem_ld32(lane_addr + c0, v);
      if (warp == 2 && c0 == 0) GTRACE(5);
      for (int sp = 1; sp < S; ++sp) {               // peers' partials, in rank order
        const float* rv = recv + ((size_t)(sp - 1) * BN + c0) * TBM + r;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += rv[j * TBM];
      }
      if (!valid && !p.gn_stats) return;             // (statistics: every lane takes part in the warp reductions below)
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(&s_bias[c0 + j]);
        v[j] = fmaf(v[j], p.alpha, bias_row) + b.x;
        v[j + 1] = fmaf(v[j + 1], p.alpha, bias_row) + b.y;
        v[j + 2] = fmaf(v[j + 2], p.alpha, bias_row) + b.z;
        v[j + 3] = fmaf(v[j + 3], p.alpha, bias_row) + b.w;
      }
      if (do_res) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 u = kPre ? rpre[kPre ? (c0 - cbeg) / 8 + j : 0] : rr[j];
          const float2 a = unpack2(u.x), b = unpack2(u.y), c = unpack2(u.z), d = unpack2(u.w);
          v[j * 8] += a.x; v[j * 8 + 1] += a.y; v[j * 8 + 2] += b.x; v[j * 8 + 3] += b.y;
          v[j * 8 + 4] += c.x; v[j * 8 + 5] += c.y; v[j * 8 + 6] += d.x; v[j * 8 + 7] += d.y;
        }
      }
      if (n >= p.act_from) {                         // act_from is a multiple of 32 (gemm_tc_supported)
        if (p.act == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (p.act == ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
        }
      }
      if (do_gate) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 u = gg[j];
          const float2 a = unpack2(u.x), b = unpack2(u.y), c = unpack2(u.z), d = unpack2(u.w);
          v[j * 8] *= a.x; v[j * 8 + 1] *= a.y; v[j * 8 + 2] *= b.x; v[j * 8 + 3] *= b.y;
          v[j * 8 + 4] *= c.x; v[j * 8 + 5] *= c.y; v[j * 8 + 6] *= d.x; v[j * 8 + 7] *= d.y;
        }
      }
      if (p.gn_stats) {
        float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
        if (valid) {
          if (p.gn_cpg >= 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { s0 += v[j]; q0 = fmaf(v[j], v[j], q0); }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) { s0 += v[j]; q0 = fmaf(v[j], v[j], q0); s1 += v[16 + j]; q1 = fmaf(v[16 + j], v[16 + j], q1); }
          }
        }
        s0 = warp_sum(s0); q0 = warp_sum(q0); s1 = warp_sum(s1); q1 = warp_sum(q1);
        if (lane == 0)
          p.gn_part[((size_t)blockIdx.y * (p.N >> 5) + (n >> 5)) * 4 + quad] = make_float4(s0, q0, s1, q1);
        if (!valid) return;
      }
      if (warp == 2 && c0 == 0) GTRACE(6);
      if (p.tma_store) {
        // row r of the 64-column panel c0 / 64: 128 bytes, 16-byte piece q stored at piece q ^ (r & 7) (128-byte swizzle)
        unsigned char* panel = smem + (size_t)(c0 >> 6) * (TBM * 128) + (size_t)r * 128;
        const int q0 = (c0 & 63) >> 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          u.x = pack2(v[j * 8], v[j * 8 + 1]); u.y = pack2(v[j * 8 + 2], v[j * 8 + 3]);
          u.z = pack2(v[j * 8 + 4], v[j * 8 + 5]); u.w = pack2(v[j * 8 + 6], v[j * 8 + 7]);
          *reinterpret_cast<uint4*>(panel + (((q0 + j) ^ (r & 7)) << 4)) = u;
        }
        return;
      }
      if (f32) {
        float* o = reinterpret_cast<float*>(base) + gm * ld + nn;
        if (do_acc) {                                // one batch of 8 loads = one L2 round trip
          float4 aa[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) aa[j] = *reinterpret_cast<const float4*>(o + j * 4);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[j * 4] += aa[j].x; v[j * 4 + 1] += aa[j].y; v[j * 4 + 2] += aa[j].z; v[j * 4 + 3] += aa[j].w;
          }
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
        t16* o = reinterpret_cast<t16*>(base) + gm * ld + nn;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 u;
          u.x = pack2(v[j], v[j + 1]); u.y = pack2(v[j + 2], v[j + 3]);
          u.z = pack2(v[j + 4], v[j + 5]); u.w = pack2(v[j + 6], v[j + 7]);
          *reinterpret_cast<uint4*>(o + j) = u;
        }
      }
      if (warp == 2 && c0 == 0) GTRACE(7);
    };
    if constexpr (kPre) {
#pragma unroll
      for (int cc = 0; cc < kHalfCols; cc += 32) chunk(cbeg + cc);
    } else {
#pragma unroll 1
      for (int cc = 0; cc < kHalfCols; cc += 32) chunk(cbeg + cc);
    }
    if (p.tma_store) {
      fence_async_smem();                              // the panels (generic-proxy writes) are visible to the TMA unit
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
      if (warp == 2 && elect_one()) {
#pragma unroll
        for (int pn = 0; pn < BN / 64; ++pn) {
          if (n0 + pn * 64 >= p.N) break;
          const unsigned char* src = smem + (size_t)pn * (TBM * 128);
          if (p.conv) tma_store_4d(&map_c, src, n0 + pn * 64, x0, y0, img);
          else tma_store_2d(&map_c, src, n0 + pn * 64, m0);
        }
        tma_store_commit_and_wait_read();              // the panels are read before this CTA's shared memory goes away
      }
    }
    }
    if (p.gn_stats) __threadfence();                 // this warp's partial sums are visible before the CTA signs off
    fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    fence_after();
    tmem_dealloc<BN>(tmem);
  }
  if (p.gn_stats && krank == 0) {
    // last CTA to finish: fold the partials -- thread -> (output o = (group, sum | sum of squares), slice j); slice j adds
    // slots j, j + J, ... in order, the J slices are then added in order: a fixed summation order, bit-reproducible
    __shared__ bool gn_last;
    __shared__ double gn_sh[kTcThreads];
    if (threadIdx.x == 0) gn_last = atomicAdd(p.gn_counter, 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (gn_last) {
      __threadfence();
      const int chunks = p.N >> 5, gpc = p.gn_cpg >= 32 ? 1 : 2;      // groups per 32-column chunk
      const int n_out = chunks * gpc * 2, J = kTcThreads / n_out;
      const int o = threadIdx.x % n_out, j = threadIdx.x / n_out;
      const int grp = o >> 1, which = o & 1;
      const int chunk = grp / gpc, comp = (grp % gpc) * 2 + which;    // float4 component holding this output
      double a = 0.0;
      if (j < J) {
        const int nslots = (int)gridDim.y * 4;                        // (row tile, quadrant) pairs
        for (int k = j; k < nslots; k += J) {
          const float4 f = p.gn_part[((size_t)(k >> 2) * chunks + chunk) * 4 + (k & 3)];
          a += (double)(comp == 0 ? f.x : comp == 1 ? f.y : comp == 2 ? f.z : f.w);
        }
      }
      gn_sh[threadIdx.x] = a;
      __syncthreads();
      if (threadIdx.x < n_out) {
        double t = 0.0;
        for (int jj = 0; jj < J; ++jj) t += gn_sh[jj * n_out + threadIdx.x];
        p.gn_stats[threadIdx.x] = t;
      }
      if (threadIdx.x == 0) *p.gn_counter = 0u;                       // re-arm for the next launch on this stream
    }
  }
}

constexpr int kMaxSplitK = 4;
// Tuning aids (tools/tune_gemm.py), thread-local: a forced (BN, split-K, stages) for every following launch (0 = the
// launcher's own choice) and a host-side log of the shapes launched.
struct TcForce { int bn = 0, splitk = 0, stages = 0; };
TcForce& tc_force() {
  static thread_local TcForce f;
  return f;
}
struct TcLog { int* buf = nullptr; int cap = 0, n = 0; };
TcLog& tc_log() {
  static thread_local TcLog l;
  return l;
}
int tc_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}
// RMEM_GEMM_SPLITK=0 disables the split (A/B measurements)
bool gemm_splitk_switch() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("RMEM_GEMM_SPLITK"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

template <int BN>
int launch_tc(const CUtensorMap* ma, const CUtensorMap* mb, const CUtensorMap* mc, TcGemmParams& p, int m_tiles,
              int max_stages, cudaStream_t s) {
  using L = TcSmem<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    // most the launcher below can ask for: BN = 256 runs 4 stages (6 x 48 KB would not fit in 227 KB)
    const int a = L::total(BN == 256 ? 4 : kMaxStages), b = BN == 64 ? L::total(4, kMaxSplitK) : 0;
    RMEM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         a > b ? a : b));
    attr_done = true;
  }
  // Split-K (BN = 64 only): few output tiles and a long K -> spread the k-blocks over a cluster of S CTAs so that
  // every SM pulls a share of the operands (the per-SM L2 ingest rate is the bound, see the header).
  const int tiles = cdiv(p.N, BN) * m_tiles;
  int S = 1;
  if (BN == 64 && gemm_splitk_switch() && tiles * 2 <= tc_sm_count() && p.nk >= 8) {
    S = tc_sm_count() / tiles;
    if (S > kMaxSplitK) S = kMaxSplitK;
    if (S > p.nk / 4) S = p.nk / 4;
    if (S < 2) S = 1;
  }
  // One wave at two CTAs per SM: a long-K convolution with more tiles than half the SMs (layer2's 3x3 at 1/8 resolution:
  // 102 tiles x 36 k-blocks) still halves its per-SM operand ingest with S = 2 when both CTAs of every cluster stay
  // resident (shallow ring, see below): 20.7 -> 15.5 us (tools/tune_gemm.py).
  bool shallow = false;
  if (BN == 64 && S == 1 && gemm_splitk_switch() && p.conv == 1 && p.nk >= 32 && tiles <= tc_sm_count() &&
      tiles * 2 > tc_sm_count()) {
    S = 2;
    shallow = true;
  }
  if (BN == 64 && tc_force().splitk > 0) {
    S = tc_force().splitk;
    if (S > kMaxSplitK) S = kMaxSplitK;
    if (S > p.nk) S = p.nk;
  }
  p.splitk = S;
  // Only as many stages as there are k-blocks: short-K GEMMs (most of the encoder) are latency-bound, and a small
  // shared-memory footprint lets several CTAs share an SM so one CTA's epilogue overlaps another's loads.
  const int nkl = cdiv(p.nk, S);
  if (shallow) max_stages = 2;
  // A grid that fits the SMs one CTA each is operand-ingest bound per SM (header): give it the deepest ring -- whose
  // footprint (> half of the 227 KB) also keeps the block scheduler from stacking two CTAs on one SM while others idle
  // (K = 2048 tail projection, 112 CTAs: 22.8 -> 16.6 us; tools/tune_gemm.py).
  static const bool deep = [] { const char* e = getenv("RMEM_GEMM_DEEP"); return !(e && e[0] == '0'); }();   // A/B switch
  if (deep && S == 1 && tiles <= tc_sm_count() && nkl >= 6) max_stages = BN == 256 ? 4 : kMaxStages;
  if (tc_force().stages > 0) {
    max_stages = tc_force().stages;
    const int cap = BN == 256 ? 4 : (S > 1 ? 4 : kMaxStages);     // what the shared-memory attribute above covers
    if (max_stages > cap) max_stages = cap;
  }
  p.stages = nkl < max_stages ? nkl : max_stages;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cdiv(p.N, BN), m_tiles, S);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = L::total(p.stages, S);
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (S > 1) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 1;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = (unsigned)S;
    cfg.numAttrs = 2;
  }
  // the output panels are staged in the operand ring: BN = 256 needs two stages' worth of it
  if (p.tma_store && BN == 256 && p.stages < 2) p.tma_store = 0;
  RMEM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN>, *ma, *mb, *(p.tma_store ? mc : ma),
                                     static_cast<const TcGemmParams&>(p)));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

}  // namespace

int gemm_tc_set_force(int bn, int splitk, int stages) {
  RMEM_REQUIRE(bn == 0 || bn == 64 || bn == 128 || bn == 256, "gemm_tc force: BN must be 0, 64, 128 or 256 (got %d)", bn);
  RMEM_REQUIRE(splitk >= 0 && splitk <= kMaxSplitK && stages >= 0 && stages <= kMaxStages,
               "gemm_tc force: split-K 0..%d, stages 0..%d", kMaxSplitK, kMaxStages);
  tc_force().bn = bn; tc_force().splitk = splitk; tc_force().stages = stages;
  return RMEM_OK;
}
int gemm_tc_set_log(int* host_buf, int cap_records) {
  tc_log().buf = host_buf; tc_log().cap = host_buf ? cap_records : 0; tc_log().n = 0;
  return RMEM_OK;
}
int gemm_tc_log_count() { return tc_log().n; }

int gemm_tc_set_trace(long long* dev_buf) {
  RMEM_CUDA_CHECK(cudaMemcpyToSymbol(g_gemm_trace, &dev_buf, sizeof(dev_buf)));
  return RMEM_OK;
}

bool gemm_tc_supported(const GemmParams& p) {
  if (p.K % TBK != 0) return false;
  // the epilogue stores whole 32-column chunks: ragged N only when the caller says the padding columns are writable
  if (p.N % 32 != 0 && !(p.pad_n_ok && p.n_split >= p.N && p.ldc >= round_up(p.N, 32) && !p.res && !p.gate &&
                         (!p.bias || p.bias_m)))
    return false;
  if ((reinterpret_cast<uintptr_t>(p.A) & 15) || (reinterpret_cast<uintptr_t>(p.B) & 15)) return false;
  if (p.ldb % 8 != 0) return false;
  if (p.n_split < p.N && p.n_split % 32 != 0) return false;
  if (p.act != ACT_NONE && p.act_from % 32 != 0) return false;
  if (reinterpret_cast<uintptr_t>(p.C) & 15) return false;
  if (p.ldc % (p.c_fp32 ? 4 : 8) != 0) return false;
  if (p.n_split < p.N && ((reinterpret_cast<uintptr_t>(p.C2) & 15) || p.ldc2 % (p.c2_fp32 ? 4 : 8) != 0)) return false;
  if (p.res && ((reinterpret_cast<uintptr_t>(p.res) & 15) || p.ldr % 8 != 0)) return false;
  if (p.gate && ((reinterpret_cast<uintptr_t>(p.gate) & 15) || p.ldg % 8 != 0)) return false;
  if (p.bias && !p.bias_m && (reinterpret_cast<uintptr_t>(p.bias) & 15)) return false;
  if (p.nimg < 1 || (p.nimg > 1 && !p.conv)) return false;
  if (p.conv == 2) {                       // stem mode (see gemm.cuh): K = kw rows of 64 contiguous elements
    if (p.K != p.kw * TBK || p.stride != 2 || p.Cin != 8) return false;
    if (p.M % (p.Wout * p.nimg) != 0) return false;
    if ((long long)(p.Wout - 1) * 2 * p.Cin + TBK > (long long)p.Win * p.Cin) return false;   // last window inside the row
  } else if (p.conv) {
    if (p.Cin % TBK != 0) return false;
    if (p.K % (p.kw * p.Cin) != 0) return false;
    if (p.stride < 1 || p.stride > 2) return false;
    if (p.M % (p.Wout * p.nimg) != 0) return false;
  } else {
    if (p.lda % 8 != 0) return false;
  }
  return true;
}

int gemm_tc_launch(const GemmParams& g, cudaStream_t stream) {
  RMEM_REQUIRE(gemm_tc_supported(g), "gemm_tc: unsupported shape/alignment (M=%d N=%d K=%d)", g.M, g.N, g.K);
  TcGemmParams p;
  p.M = g.M; p.N = round_up(g.N, 32); p.nk = g.K / TBK;
  p.conv = g.conv; p.taps_w = g.kw; p.cin_blocks = g.conv == 1 ? g.Cin / TBK : 1; p.pad = g.pad; p.stride = g.stride;
  p.sx = p.sy = g.stride;
  if (g.conv == 2) { p.taps_w = 1; p.pad = 0; p.sx = 1; p.sy = 2; }   // k-block kb = window row ky; padding is physical
  p.BW = 128; p.BH = 1; p.tiles_x = 1; p.Hout = 0; p.Wout = g.Wout; p.nimg = g.nimg; p.tiles_img = 0;
  p.alpha = g.alpha; p.bias = g.bias; p.bias_m = g.bias_m; p.act = g.act; p.act_from = g.act_from;
  p.res = g.res; p.ldr = g.ldr; p.gate = g.gate; p.ldg = g.ldg; p.accumulate = g.accumulate;
  p.C = g.C; p.ldc = g.ldc; p.c_fp32 = g.c_fp32; p.C2 = g.C2; p.ldc2 = g.ldc2; p.c2_fp32 = g.c2_fp32;
  p.n_split = g.n_split;
  p.err = nullptr;   // watchdog traps without a flag word (the library never allocates device memory)
  p.gn_stats = g.gn_stats; p.gn_cpg = 0; p.gn_part = nullptr; p.gn_counter = nullptr;
  if (g.gn_stats) {
    RMEM_REQUIRE(g.gn_groups > 0 && g.N % g.gn_groups == 0 && (g.N / g.gn_groups == 16 || g.N / g.gn_groups == 32) &&
                 g.N % 32 == 0 && g.N <= 256 && !g.c_fp32 && g.n_split >= g.N && g.act == ACT_NONE && !g.gate,
                 "gemm_tc: GroupNorm statistics need 16 or 32 channels per group, N <= 256, one t16 destination (N=%d G=%d)",
                 g.N, g.gn_groups);
    p.gn_cpg = g.N / g.gn_groups;
    // scratch layout shared with the stand-alone statistics kernel (ops.cuh): [0,64) stats | [64] counter | [72,..) partials
    p.gn_counter = reinterpret_cast<unsigned int*>(g.gn_stats + 64);
    p.gn_part = reinterpret_cast<float4*>(g.gn_stats + 72);
  }

  // ---- tile width: fill the 148 SMs, then prefer wide tiles (fewer A re-reads) ----
  int m_tiles;
  const CUtensorMap *ma = nullptr, *mb = nullptr;
  if (g.conv) {
    const int Hout = g.M / (g.Wout * g.nimg);
    p.Hout = Hout;
    const int rank = g.nimg > 1 ? 4 : 3;   // images = a fourth tensor-map dimension (box 1): padding stays per image
    // rectangular 128-pixel patch that wastes the fewest tile slots
    int best = 1 << 30;
    for (int bw = 128; bw >= 8; bw >>= 1) {
      const int bh = 128 / bw;
      if (bw * g.stride > 256 || bh * g.stride > 256) continue;
      const int n = cdiv(g.Wout, bw) * cdiv(Hout, bh);
      if (n < best) { best = n; p.BW = bw; p.BH = bh; }
    }
    p.tiles_x = cdiv(g.Wout, p.BW);
    p.tiles_img = best;
    m_tiles = best * g.nimg;
    if (g.conv == 2) {
      // Stem: the A row of output pixel (oy, ox) and window row ky is the 64 contiguous elements (8 pixels x 8 channels,
      // the 8th pixel meets zero weights) that start at padded pixel (oy*2 + ky, ox*2): a tensor map whose second
      // dimension steps by TWO pixels (32 bytes) while the first one spans 64 elements -- overlapping rows.
      uint64_t dims[4] = {(uint64_t)TBK, (uint64_t)g.Wout, (uint64_t)g.Hin, (uint64_t)g.nimg};
      uint64_t strides[3] = {(uint64_t)2 * g.Cin * 2, (uint64_t)g.Win * g.Cin * 2, (uint64_t)g.Hin * g.Win * g.Cin * 2};
      uint32_t box[4] = {(uint32_t)TBK, (uint32_t)p.BW, (uint32_t)(p.BH * 2), 1};
      uint32_t estr[4] = {1, 1, 2, 1};
      RMEM_TRY(tma_encode_cached(&ma, g.A, rank, dims, strides, box, estr));
    } else {
    uint64_t dims[4] = {(uint64_t)g.Cin, (uint64_t)g.Win, (uint64_t)g.Hin, (uint64_t)g.nimg};
    uint64_t strides[3] = {(uint64_t)g.Cin * 2, (uint64_t)g.Win * g.Cin * 2, (uint64_t)g.Hin * g.Win * g.Cin * 2};
    uint32_t box[4] = {(uint32_t)TBK, (uint32_t)(p.BW * g.stride), (uint32_t)(p.BH * g.stride), 1};
    uint32_t estr[4] = {1, (uint32_t)g.stride, (uint32_t)g.stride, 1};
    RMEM_TRY(tma_encode_cached(&ma, g.A, rank, dims, strides, box, estr));
    }
  } else {
    m_tiles = cdiv(g.M, TBM);
    uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)g.M};
    uint64_t strides[1] = {(uint64_t)g.lda * 2};
    uint32_t box[2] = {(uint32_t)TBK, (uint32_t)TBM};
    RMEM_TRY(tma_encode_cached(&ma, g.A, 2, dims, strides, box, nullptr));
  }
  int BN = 128;
  // (a wide tile only when it still gives every SM a CTA: N = 1152 x 14 row tiles = 126 wide tiles ran 10.7 us, 252 narrow
  // ones 8.5 us, and the narrow CTAs share an SM with the other streams' kernels)
  static const int wide_min = [] { const char* e = getenv("RMEM_GEMM_WIDE_MIN"); return e ? atoi(e) : 148; }();
  if (g.N <= 64 || m_tiles * cdiv(g.N, 128) < wide_min) BN = 64;
  else if (g.N >= 256 && g.K >= 512 && m_tiles * cdiv(g.N, 256) >= 148) BN = 256;
  if (tc_force().bn) BN = tc_force().bn;
  if (tc_log().buf && tc_log().n < tc_log().cap) {
    // record = 16 ints: M N K conv Hin Win Cin Wout kw stride pad | act has_res has_gate c_fp32 n_split<N
    int* r = tc_log().buf + (size_t)tc_log().n * 16;
    r[0] = g.M; r[1] = g.N; r[2] = g.K; r[3] = g.conv; r[4] = g.Hin; r[5] = g.Win; r[6] = g.Cin; r[7] = g.Wout;
    r[8] = g.kw; r[9] = g.stride; r[10] = g.pad; r[11] = g.act; r[12] = g.res != nullptr; r[13] = g.gate != nullptr;
    r[14] = g.c_fp32 | (g.accumulate << 1) | (g.bias_m << 2); r[15] = g.n_split < g.N ? g.n_split : 0;
    ++tc_log().n;
  }
  {
    uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)g.N};
    uint64_t strides[1] = {(uint64_t)g.ldb * 2};
    uint32_t box[2] = {(uint32_t)TBK, (uint32_t)BN};
    RMEM_TRY(tma_encode_cached(&mb, g.B, 2, dims, strides, box, nullptr));
  }
  // TMA-store epilogue: one t16 destination, whole 64-column panels inside the allocation (ldc >= round_up(N, 64) or
  // N % 64 == 0 -- the tensor map drops what lies beyond N anyway, but a panel's box must not straddle the row pitch)
  static const int tma_store_on = [] { const char* e = getenv("RMEM_GEMM_TMA_STORE"); return e ? atoi(e) : 1; }();
  const CUtensorMap* mc = nullptr;
  p.tma_store = 0;
  if (tma_store_on && !g.c_fp32 && g.n_split >= g.N && !g.bias_m && g.N % 8 == 0 && (g.ldc * 2) % 16 == 0) {
    if (g.conv) {
      uint64_t dims[4] = {(uint64_t)g.N, (uint64_t)g.Wout, (uint64_t)p.Hout, (uint64_t)g.nimg};
      uint64_t strides[3] = {(uint64_t)g.ldc * 2, (uint64_t)g.Wout * g.ldc * 2, (uint64_t)p.Hout * g.Wout * g.ldc * 2};
      uint32_t box[4] = {64, (uint32_t)p.BW, (uint32_t)p.BH, 1};
      RMEM_TRY(tma_encode_cached(&mc, g.C, 4, dims, strides, box, nullptr));
    } else {
      uint64_t dims[2] = {(uint64_t)g.N, (uint64_t)g.M};
      uint64_t strides[1] = {(uint64_t)g.ldc * 2};
      uint32_t box[2] = {64, (uint32_t)TBM};
      RMEM_TRY(tma_encode_cached(&mc, g.C, 2, dims, strides, box, nullptr));
    }
    p.tma_store = 1;
  }
  if (BN == 64) return launch_tc<64>(ma, mb, mc, p, m_tiles, 4, stream);
  if (BN == 128) return launch_tc<128>(ma, mb, mc, p, m_tiles, 3, stream);
  return launch_tc<256>(ma, mb, mc, p, m_tiles, 4, stream);
}

}  // namespace rmem
