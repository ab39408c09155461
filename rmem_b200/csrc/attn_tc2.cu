// Long-term / self attention v2 for sm_100a  (K1 + K1b + K8 of SURVEY.md; attention.py:174-193, transformer.py:1140-1197).
//
// What changed against attn_tc.cu (kept as RMEM_ATTN_TC for comparison), driven by the round-1 ncu capture:
//   * stream-K style static schedule: the (query tile, Dv chunk, KV tile) space is cut into one contiguous range per
//     SM (grid = #SMs), so there is no partial last wave and each CTA flushes at most two segments;
//   * eight softmax warps (two per TMEM lane quadrant, splitting the 64 score columns) with a ~170-instruction tile
//     body: scale, temporal-PE bias and the running max folded into one FFMA feeding EX2, no saturating packs, the
//     ragged-tile mask only on the last tile of a frame;
//   * P goes back to the tensor core through TMEM (tcgen05.st, A operand of the P.V MMA read from TMEM): no shared-memory
//     round trip, no proxy fence, and the freed shared memory buys a 4th K/V stage;
//   * partial results leave the CTA as normalised fp16 rows (64 B bursts per thread) + fp32 (m, l): 4x less partial
//     traffic than the fp32 un-normalised partials of v1; per-frame (m, l) pieces give the attention mass.
//
//   block = 320 threads: warps 0-7 softmax + epilogue, warp 8 TMA producer, warp 9 MMA issuer + TMEM owner
//   TMEM (512 cols): O[256] | S0[64] S1[64] | P0[32] P1[32] (fp16 pairs)
//   smem: Q 32 KB + 4 x (K 16 KB + V^T 32 KB) + barriers + exchange
#include "attn.cuh"
#include "tcgen05.cuh"

namespace rmem {

namespace {

using namespace tc;

constexpr int BM = 128;        // query rows per CTA
constexpr int BN = 64;         // keys per KV tile
constexpr int DK = 128;
constexpr int DVC = 256;       // Dv columns per unit
constexpr int STAGES = 4;
constexpr int kSoftmaxWarps = 8;
constexpr int kThreads = (kSoftmaxWarps + 2) * 32;

constexpr int SMEM_Q = BM * DK * 2;            // 32 KB (two 64-col swizzle atoms)
constexpr int SMEM_K = BN * DK * 2;            // 16 KB
constexpr int SMEM_V = DVC * BN * 2;           // 32 KB
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + SMEM_Q;
constexpr int OFF_V = OFF_K + STAGES * SMEM_K;
constexpr int OFF_XCH = OFF_V + STAGES * SMEM_V;          // float [2][4][2][32]
constexpr int OFF_BAR = OFF_XCH + 2 * 4 * 2 * 32 * 4;
constexpr int SMEM_TOTAL = OFF_BAR + 256;   // no alignment slack: 4 stages only fit in 227 KB if the dynamic base is
                                            // 1024B-aligned already (it follows the 1 KB driver reservation); checked below

constexpr int TMEM_COLS = 512;
constexpr int TMEM_O = 0;
constexpr int TMEM_S = 256;    // 2 x 64
constexpr int TMEM_P = 384;    // 2 x 32 (packed fp16 pairs)

constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 8.0f;      // log2 units: P <= 2^8 before a lazy rescale is forced

struct Tc2Params {
  int HW, HWp, T, tpf, TPU, n_units, n_dv, nCTA, Dv;
  long long L;                 // n_units * TPU
  int slot[kMaxBankFrames];
  float scale_log2;            // scale * log2(e)
  const float* qbias;          // [HW, T] or null (already multiplied by scale)
  t16* part_o;                 // [nCTA][2][BM][DVC]  normalised partial O
  float* part_ml;              // [nCTA][2][BM][2]    (m in log2 units, l)
  float* pieces;               // [nCTA][2][T][BM][2] per-frame (m, l) of the segment (Dv chunk 0 units only) or null
};

// D[tmem] (+)= A[tmem] . B[smem]^T   (A = P as packed fp16 pairs, one TMEM lane per row)
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// unsaturated fp32x2 -> t16x2 (values are bounded by 2^8 here)
__device__ __forceinline__ uint32_t pack2_fast(float lo, float hi) {
#ifdef RMEM_OPERAND_BF16
  t162 v = __floats2bfloat162_rn(lo, hi);
#else
  t162 v = __floats2half2_rn(lo, hi);
#endif
  return *reinterpret_cast<uint32_t*>(&v);
}

struct Seg { int unit, lo, hi; };   // tiles [lo, hi) of the unit

__device__ __forceinline__ void cta_range(const Tc2Params& p, int cta, long long& lo, long long& hi) {
  lo = (p.L * cta) / p.nCTA;
  hi = (p.L * (cta + 1)) / p.nCTA;
}

__global__ void __launch_bounds__(kThreads, 1)
long_attn_tc2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const Tc2Params p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();       // 128B-swizzled TMA / UMMA tiles need 1024B alignment
  float* xch = reinterpret_cast<float*>(smem + OFF_XCH);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* q_free = bars + 1;
  uint64_t* o_drained = bars + 2;
  uint64_t* kv_full = bars + 3;                 // [STAGES]
  uint64_t* kv_empty = kv_full + STAGES;        // [STAGES]
  uint64_t* s_full = kv_empty + STAGES;         // [2]
  uint64_t* s_free = s_full + 2;                // [2]
  uint64_t* p_full = s_free + 2;                // [2]
  uint64_t* p_free = p_full + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- this CTA's work: a contiguous range of (unit, tile) steps -> at most two segments ----
  long long lo, hi;
  cta_range(p, blockIdx.x, lo, hi);
  Seg seg[2];
  int nseg = 0;
  {
    long long x = lo;
    while (x < hi && nseg < 2) {
      const int u = (int)(x / p.TPU);
      const long long ue = (long long)(u + 1) * p.TPU;
      const long long e = hi < ue ? hi : ue;
      seg[nseg].unit = u;
      seg[nseg].lo = (int)(x - (long long)u * p.TPU);
      seg[nseg].hi = (int)(e - (long long)u * p.TPU);
      ++nseg;
      x = e;
    }
  }
  const int n0 = nseg > 0 ? seg[0].hi - seg[0].lo : 0;
  const int ntot = n0 + (nseg > 1 ? seg[1].hi - seg[1].lo : 0);
  // units are ordered (query tile major, Dv chunk minor): the Q tile changes only when unit / n_dv changes
  const bool q_reload1 = nseg > 1 && (seg[1].unit / p.n_dv != seg[0].unit / p.n_dv);

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_free, 1);
    mbar_init(o_drained, kSoftmaxWarps);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&s_free[i], kSoftmaxWarps);
      mbar_init(&p_full[i], kSoftmaxWarps); mbar_init(&p_free[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == kSoftmaxWarps + 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kSoftmaxWarps) {
    // ================================ TMA producer (converged warp, one elected lane issues) ================
    if (ntot > 0) {
      if (elect_one()) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_v);
      }
      __syncwarp();
      int j = 0;
      for (int s = 0; s < nseg; ++s) {
        const int qt = seg[s].unit / p.n_dv, dv0 = (seg[s].unit % p.n_dv) * DVC;
        if (s == 0 || q_reload1) {
          if (s > 0) mbar_wait(q_free, 0, nullptr, 1);
          if (elect_one()) {
            mbar_expect_tx(q_full, SMEM_Q);
            tma_load_2d(smem + OFF_Q, &map_q, q_full, 0, qt * BM);
            tma_load_2d(smem + OFF_Q + BM * 128, &map_q, q_full, 64, qt * BM);
          }
          __syncwarp();
        }
        for (int g = seg[s].lo; g < seg[s].hi; ++g, ++j) {
          const int st = j % STAGES;
          if (j >= STAGES) mbar_wait(&kv_empty[st], ((j / STAGES) - 1) & 1, nullptr, 2);
          const int t = g / p.tpf, jt = g - t * p.tpf;
          const int key0 = p.slot[t] * p.HWp + jt * BN;
          if (elect_one()) {
            mbar_expect_tx(&kv_full[st], SMEM_K + SMEM_V);
            unsigned char* sk = smem + OFF_K + st * SMEM_K;
            tma_load_2d(sk, &map_k, &kv_full[st], 0, key0);
            tma_load_2d(sk + BN * 128, &map_k, &kv_full[st], 64, key0);
            tma_load_2d(smem + OFF_V + st * SMEM_V, &map_v, &kv_full[st], key0, dv0);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kSoftmaxWarps + 1) {
    // ================================ MMA issuer (converged warp, one elected lane issues) ==================
    if (ntot > 0) {
      constexpr uint32_t idesc_s = make_idesc(BM, BN);
      constexpr uint32_t idesc_o = make_idesc(BM, DVC);
      const uint32_t smem_base = smem_u32(smem);
      const uint64_t dq = make_desc_sw128(smem_base + OFF_Q);
      auto issue_s = [&](int j) {
        const int st = j % STAGES, b = j & 1;
        if (j == 0) mbar_wait(q_full, 0, nullptr, 3);
        if (j == n0 && q_reload1) mbar_wait(q_full, 1, nullptr, 4);
        mbar_wait(&kv_full[st], (j / STAGES) & 1, nullptr, 5);
        if (j >= 2) mbar_wait(&s_free[b], ((j - 2) >> 1) & 1, nullptr, 6);
        fence_after();
        if (elect_one()) {
          const uint64_t dk = make_desc_sw128(smem_base + OFF_K + st * SMEM_K);
          const uint32_t d = tmem + TMEM_S + b * BN;
#pragma unroll
          for (int kk = 0; kk < DK / 16; ++kk) {
            // 16 k-elements = 32 B inside the 128 B swizzle atom; the second 64-wide atom starts one tile-half later
            const uint64_t oa = (uint64_t)(((kk >> 2) * (BM * 128) + (kk & 3) * 32) >> 4);
            const uint64_t ob = (uint64_t)(((kk >> 2) * (BN * 128) + (kk & 3) * 32) >> 4);
            umma_ss(d, dq + oa, dk + ob, idesc_s, kk > 0);
          }
          commit(&s_full[b]);
          if (j == n0 - 1 && q_reload1) commit(q_free);      // last read of the old Q tile
        }
        __syncwarp();
      };
      issue_s(0);
      for (int j = 0; j < ntot; ++j) {
        const bool defer = (j + 1 == n0) && q_reload1;     // next S needs a new Q tile: do not block P.V(j) on it
        if (j + 1 < ntot && !defer) issue_s(j + 1);
        const int st = j % STAGES, b = j & 1;
        mbar_wait(&p_full[b], (j >> 1) & 1, nullptr, 7);
        const bool first = (j == 0) || (j == n0);
        if (j == n0 && n0 > 0 && nseg > 1) mbar_wait(o_drained, 0, nullptr, 8);
        fence_after();
        if (elect_one()) {
          const uint64_t dv = make_desc_sw128(smem_base + OFF_V + st * SMEM_V);
          const uint32_t pa = tmem + TMEM_P + b * 32;
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk)
            umma_ts(tmem + TMEM_O, pa + kk * 8, dv + (uint64_t)(kk * 2), idesc_o, (first && kk == 0) ? 0u : 1u);
          commit(&kv_empty[st]);
          commit(&p_free[b]);
        }
        __syncwarp();
        if (j + 1 < ntot && defer) issue_s(j + 1);
      }
    }
  } else if (ntot > 0) {
    // ================================ softmax + epilogue (warps 0-7) ================================
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;                       // tile row == TMEM lane
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    const int bar_id = 1 + quad;
    int j = 0, xp = 0;
    // exchange slot of this thread / its partner (other half of the same row)
    auto xslot = [&](int par, int h) { return xch + ((par * 4 + quad) * 2 + h) * 32 + lane; };

    for (int s = 0; s < nseg; ++s) {
      const int unit = seg[s].unit;
      const int qt = unit / p.n_dv, dvc = unit % p.n_dv;
      const int qi = qt * BM + row;
      const bool row_ok = qi < p.HW;
      float m_used = -INFINITY, l_tot = 0.f, l_piece = 0.f, bias2 = 0.f;
      int cur_t = -1;
      const int j_first = j;
      float* piece_base = (p.pieces && dvc == 0)
                              ? p.pieces + ((long long)(blockIdx.x * 2 + s) * p.T) * (BM * 2) + row * 2
                              : nullptr;
      auto flush_piece = [&](int t) {
        // l of this frame's piece, both column halves
        *xslot(xp, half) = l_piece;
        named_bar_sync(bar_id, 64);
        const float other = *xslot(xp, half ^ 1);
        xp ^= 1;
        if (piece_base && half == 0) {
          float* d = piece_base + (long long)t * (BM * 2);
          d[0] = m_used;
          d[1] = l_piece + other;
        }
      };

      for (int g = seg[s].lo; g < seg[s].hi; ++g, ++j) {
        const int t = g / p.tpf, jt = g - t * p.tpf;
        if (t != cur_t) {
          if (cur_t >= 0) flush_piece(cur_t);
          cur_t = t;
          l_piece = 0.f;
          bias2 = (p.qbias && row_ok) ? p.qbias[(long long)qi * p.T + t] * LOG2E : 0.f;
        }
        const int b = j & 1;
        mbar_wait(&s_full[b], (j >> 1) & 1, nullptr, 9);
        fence_after();
        float sc[32];
        tmem_ld32(lane_addr + TMEM_S + b * BN + half * 32, sc);
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[b]);
        if (jt == p.tpf - 1) {                               // ragged last tile of the frame
          const int key0 = jt * BN + half * 32;
#pragma unroll
          for (int c = 0; c < 32; ++c) sc[c] = (key0 + c < p.HW) ? sc[c] : -INFINITY;
        }
        float mx = sc[0];
#pragma unroll
        for (int c = 1; c < 32; ++c) mx = fmaxf(mx, sc[c]);
        *xslot(xp, half) = mx;
        named_bar_sync(bar_id, 64);
        mx = fmaxf(mx, *xslot(xp, half ^ 1));
        xp ^= 1;
        const float mt = fmaf(mx, p.scale_log2, bias2);      // scale > 0: max commutes with the affine map
        // lazy rescale (the two warps of a quadrant see identical values -> identical decisions)
        const bool need = mt > m_used + RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, need)) {
          if (j > j_first) {
            mbar_wait(&p_free[(j - 1) & 1], ((j - 1) >> 1) & 1, nullptr, 10);   // P.V(j-1) retired
            fence_after();
            const float f = need ? exp2f(m_used - mt) : 1.f;
            l_tot *= f;
            l_piece *= f;
#pragma unroll 1
            for (int c = 0; c < DVC / 2; c += 32) {
              float o[32];
              tmem_ld32(lane_addr + TMEM_O + half * (DVC / 2) + c, o);
#pragma unroll
              for (int e = 0; e < 32; ++e) o[e] *= f;
              tmem_st32(lane_addr + TMEM_O + half * (DVC / 2) + c, o);
            }
            fence_before();
          }
          if (need) m_used = mt;
        }
        const float c0 = bias2 - m_used;
        float lsum = 0.f;
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float e0 = exp2f(fmaf(sc[c], p.scale_log2, c0));
          const float e1 = exp2f(fmaf(sc[c + 1], p.scale_log2, c0));
          lsum += e0 + e1;
          pk[c >> 1] = pack2_fast(e0, e1);
        }
        l_tot += lsum;
        l_piece += lsum;
        if (j >= 2) mbar_wait(&p_free[b], ((j - 2) >> 1) & 1, nullptr, 11);    // P.V(j-2) done reading P[b]
        fence_after();
        tmem_st16(lane_addr + TMEM_P + b * 32 + half * 16, pk);
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[b]);
      }
      flush_piece(cur_t);

      // ---- segment epilogue: normalised fp16 partial O + (m, l) ----
      *xslot(xp, half) = l_tot;
      named_bar_sync(bar_id, 64);
      const float l_row = l_tot + *xslot(xp, half ^ 1);
      xp ^= 1;
      const float inv = 1.f / l_row;
      const int last = j - 1;
      mbar_wait(&p_free[last & 1], (last >> 1) & 1, nullptr, 12);
      fence_after();
      t16* po = p.part_o + ((long long)(blockIdx.x * 2 + s) * BM + row) * DVC + half * (DVC / 2);
#pragma unroll 1
      for (int c = 0; c < DVC / 2; c += 32) {
        float o[32];
        tmem_ld32(lane_addr + TMEM_O + half * (DVC / 2) + c, o);
        if (row_ok) {
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            uint4 u;
            u.x = pack2(o[e] * inv, o[e + 1] * inv);
            u.y = pack2(o[e + 2] * inv, o[e + 3] * inv);
            u.z = pack2(o[e + 4] * inv, o[e + 5] * inv);
            u.w = pack2(o[e + 6] * inv, o[e + 7] * inv);
            *reinterpret_cast<uint4*>(po + c + e) = u;
          }
        }
      }
      if (half == 0) {
        float* ml = p.part_ml + ((long long)(blockIdx.x * 2 + s) * BM + row) * 2;
        ml[0] = m_used;
        ml[1] = l_row;
      }
      fence_before();
      __syncwarp();
      if (lane == 0 && s + 1 < nseg) mbar_arrive(o_drained);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == kSoftmaxWarps + 1) {
    fence_after();
    tmem_dealloc<TMEM_COLS>(tmem);
  }
}

// Merge the segments of every unit: out = (sum_s w_s O_s) * gate with w_s = l_s 2^(m_s - M) / L;
// mass[i,t] = sum_{pieces of frame t} l_p 2^(m_p - M) / L  (from the Dv-chunk-0 units).
__global__ void __launch_bounds__(256) combine2_kernel(const Tc2Params p, const t16* __restrict__ gate, long long ldg,
                                                       t16* __restrict__ out, long long ldo,
                                                       float* __restrict__ mass) {
  const int i = blockIdx.x;
  const int qt = i / BM, r = i - qt * BM;
  const int col = threadIdx.x * 4;
  if (col < p.Dv) {
    const int k = col / DVC, cc = col - k * DVC;
    const int unit = qt * p.n_dv + k;
    const long long u_lo = (long long)unit * p.TPU, u_hi = u_lo + p.TPU;
    const int c_first = (int)(((u_lo + 1) * p.nCTA - 1) / p.L);
    float M = -INFINITY;
    for (int c = c_first; c < p.nCTA; ++c) {
      long long lo, hi;
      cta_range(p, c, lo, hi);
      if (lo >= u_hi) break;
      if (hi <= lo) continue;
      const int s = lo < u_lo ? 1 : 0;
      M = fmaxf(M, p.part_ml[((long long)(c * 2 + s) * BM + r) * 2]);
    }
    float L = 0.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = c_first; c < p.nCTA; ++c) {
      long long lo, hi;
      cta_range(p, c, lo, hi);
      if (lo >= u_hi) break;
      if (hi <= lo) continue;
      const int s = lo < u_lo ? 1 : 0;
      const float* ml = p.part_ml + ((long long)(c * 2 + s) * BM + r) * 2;
      const float w = exp2f(ml[0] - M) * ml[1];
      L += w;
      const uint2 u = *reinterpret_cast<const uint2*>(p.part_o + ((long long)(c * 2 + s) * BM + r) * DVC + cc);
      const float2 a = unpack2(u.x), b = unpack2(u.y);
      acc.x += w * a.x; acc.y += w * a.y; acc.z += w * b.x; acc.w += w * b.y;
    }
    const float inv = 1.f / L;
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    if (gate) {
      const uint2 g = *reinterpret_cast<const uint2*>(gate + (long long)i * ldg + col);
      const float2 g0 = unpack2(g.x), g1 = unpack2(g.y);
      acc.x *= g0.x; acc.y *= g0.y; acc.z *= g1.x; acc.w *= g1.y;
    }
    uint2 o;
    o.x = pack2(acc.x, acc.y);
    o.y = pack2(acc.z, acc.w);
    *reinterpret_cast<uint2*>(out + (long long)i * ldo + col) = o;
  }
  if (mass && threadIdx.x < p.T) {
    const int t = threadIdx.x;
    const int unit = qt * p.n_dv;
    const long long u_lo = (long long)unit * p.TPU, u_hi = u_lo + p.TPU;
    const int c_first = (int)(((u_lo + 1) * p.nCTA - 1) / p.L);
    float M = -INFINITY;
    for (int c = c_first; c < p.nCTA; ++c) {
      long long lo, hi;
      cta_range(p, c, lo, hi);
      if (lo >= u_hi) break;
      if (hi <= lo) continue;
      const int s = lo < u_lo ? 1 : 0;
      M = fmaxf(M, p.part_ml[((long long)(c * 2 + s) * BM + r) * 2]);
    }
    float L = 0.f, a = 0.f;
    const long long f_lo = u_lo + (long long)t * p.tpf, f_hi = f_lo + p.tpf;
    for (int c = c_first; c < p.nCTA; ++c) {
      long long lo, hi;
      cta_range(p, c, lo, hi);
      if (lo >= u_hi) break;
      if (hi <= lo) continue;
      const int s = lo < u_lo ? 1 : 0;
      const float* ml = p.part_ml + ((long long)(c * 2 + s) * BM + r) * 2;
      L += exp2f(ml[0] - M) * ml[1];
      const long long a_lo = lo > u_lo ? lo : u_lo, a_hi = hi < u_hi ? hi : u_hi;   // this CTA's tiles of the unit
      if (a_lo < f_hi && f_lo < a_hi) {
        const float* pc = p.pieces + (((long long)(c * 2 + s) * p.T + t) * BM + r) * 2;
        a += exp2f(pc[0] - M) * pc[1];
      }
    }
    mass[(long long)i * p.T + t] = a / L;
  }
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

void schedule(int HW, int T, int Dv, int* n_units, int* tpf, int* TPU, int* nCTA) {
  const int qtiles = cdiv(HW, BM), n_dv = Dv / DVC;
  *n_units = qtiles * n_dv;
  *tpf = cdiv(HW, BN);
  *TPU = T * *tpf;
  const long long L = (long long)*n_units * *TPU;
  int n = sm_count();
  if (n < *n_units) n = *n_units;          // a CTA never spans more than two units
  if ((long long)n > L) n = (int)L;
  *nCTA = n;
}

size_t part_bytes(int nCTA, int T, size_t* off_ml, size_t* off_pieces) {
  size_t o = (size_t)nCTA * 2 * BM * DVC * sizeof(t16);
  o = (o + 255) & ~size_t(255);
  *off_ml = o;
  o += (size_t)nCTA * 2 * BM * 2 * sizeof(float);
  o = (o + 255) & ~size_t(255);
  *off_pieces = o;
  o += (size_t)nCTA * 2 * T * BM * 2 * sizeof(float);
  return o + 256;
}

}  // namespace

size_t long_attn_tc2_workspace(int HW, int HWp, int nslots, int Dv) {
  (void)HWp;
  size_t best = 0;
  for (int T = 1; T <= nslots && T <= kMaxBankFrames; ++T) {
    int n_units, tpf, TPU, nCTA;
    schedule(HW, T, Dv, &n_units, &tpf, &TPU, &nCTA);
    size_t a, b;
    size_t n = part_bytes(nCTA, T, &a, &b);
    if (n > best) best = n;
  }
  return best;
}

int long_attn_tc2(const LongAttnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  RMEM_REQUIRE(a.Dk == DK, "long_attn_tc2: Dk=%d (built for 128)", a.Dk);
  RMEM_REQUIRE(a.Dv % DVC == 0 && a.Dv <= 1024, "long_attn_tc2: Dv=%d must be a multiple of 256, <= 1024", a.Dv);
  RMEM_REQUIRE(a.HWp % BN == 0 && a.HWp >= a.HW, "long_attn_tc2: HWp=%d must be a multiple of 64", a.HWp);
  RMEM_REQUIRE(a.T >= 1 && a.T <= kMaxBankFrames && a.T <= a.nslots, "long_attn_tc2: T=%d nslots=%d", a.T, a.nslots);
  RMEM_REQUIRE(a.ldo % 4 == 0 && (!a.gate || a.ldg % 4 == 0), "long_attn_tc2: ldo/ldg alignment");
  Tc2Params p;
  p.HW = a.HW; p.HWp = a.HWp; p.T = a.T; p.Dv = a.Dv; p.n_dv = a.Dv / DVC;
  schedule(a.HW, a.T, a.Dv, &p.n_units, &p.tpf, &p.TPU, &p.nCTA);
  p.L = (long long)p.n_units * p.TPU;
  size_t off_ml, off_pieces;
  const size_t need = part_bytes(p.nCTA, a.T, &off_ml, &off_pieces);
  RMEM_REQUIRE(workspace_bytes >= need, "long_attn_tc2: workspace %zu < %zu", workspace_bytes, need);
  RMEM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "long_attn_tc2: workspace alignment");
  for (int t = 0; t < kMaxBankFrames; ++t) p.slot[t] = t < a.T ? a.slot[t] : 0;
  p.scale_log2 = a.scale * LOG2E;
  p.qbias = a.qbias;
  char* ws = reinterpret_cast<char*>(workspace);
  p.part_o = reinterpret_cast<t16*>(ws);
  p.part_ml = reinterpret_cast<float*>(ws + off_ml);
  p.pieces = a.mass ? reinterpret_cast<float*>(ws + off_pieces) : nullptr;

  const CUtensorMap *mq, *mk, *mv;
  {
    uint64_t dims[2] = {(uint64_t)DK, (uint64_t)a.HW};
    uint64_t str[1] = {(uint64_t)DK * 2};
    uint32_t box[2] = {64, (uint32_t)BM};
    RMEM_TRY(tma_encode_cached(&mq, a.qt, 2, dims, str, box, nullptr));
  }
  {
    uint64_t dims[2] = {(uint64_t)DK, (uint64_t)a.nslots * a.HWp};
    uint64_t str[1] = {(uint64_t)DK * 2};
    uint32_t box[2] = {64, (uint32_t)BN};
    RMEM_TRY(tma_encode_cached(&mk, a.kbank, 2, dims, str, box, nullptr));
  }
  {
    uint64_t dims[2] = {(uint64_t)a.nslots * a.HWp, (uint64_t)a.Dv};
    uint64_t str[1] = {(uint64_t)a.nslots * a.HWp * 2};
    uint32_t box[2] = {(uint32_t)BN, (uint32_t)DVC};
    RMEM_TRY(tma_encode_cached(&mv, a.vtbank, 2, dims, str, box, nullptr));
  }
  static bool attr_done = false;
  if (!attr_done) {
    RMEM_CUDA_CHECK(cudaFuncSetAttribute(long_attn_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    attr_done = true;
  }
  long_attn_tc2_kernel<<<p.nCTA, kThreads, SMEM_TOTAL, s>>>(*mq, *mk, *mv, p);
  RMEM_LAUNCH_CHECK();
  combine2_kernel<<<a.HW, 256, 0, s>>>(p, a.gate, a.ldg, a.out, a.ldo, a.mass);
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

}  // namespace rmem
