// Windowed short-term attention on tensor cores (LocalGatedPropagation core, attention.py:289-353, 363-413), sm_100a.
//
//   s[i,d] = scale * <q_i, k_{i+d}> + rel[i,d],  d in [-7,7]^2 (zero-padded K, out-of-bounds -> excluded)
//   out_i  = (sum_d softmax_d(s)[i,d] * v_{i+d}) * gate_i
//
// The CUDA-core kernel (attn_dense.cu) reads 225 x (128 + 1024) x 2 B per query from L2: 860 MB per c3 layer, 115 us.
// Here a CTA owns a patch of 4 x 32 = 128 queries and one 256-column Dv chunk; the keys any query of the patch can see
// form an 18 x 46 halo, walked as 18 tiles of one key row x 64 key columns (a 64-key row is exactly one 128-byte
// swizzle row of the value-major V copy; narrower boxes would need a 64-byte swizzle mode).  S = Q.K^T and O += P.V are
// tcgen05 MMAs exactly as in attn_tc2.cu (same warp roles, score look-ahead, row-max hand-over between two softmax
// groups, P through TMEM); the window mask and the learned relative bias are applied to S in registers.  K tiles come
// straight from the token-major previous-frame K with a 3-D TMA box [64 ch, 64 x, 1 y] (halo outside the frame is
// zero-filled by TMA and masked); V^T tiles come from a value-major, row-padded copy [Dv][h][wp] with a 3-D box
// [64 x, 1 y, 256 dv].  One CTA sees the whole window of its queries, so it writes the final gated output: no partials.
//
//   block = 384 threads: warps 0-3 / 4-7 softmax groups (even / odd tiles), 8 Q+K producer, 9 V producer,
//                        10 S issuer + TMEM owner, 11 P.V issuer
#include "attn.cuh"
#include "tcgen05.cuh"

namespace rmem {

namespace {

using namespace tc;

constexpr int QH = 4, QW = 32;     // query patch
constexpr int BM = QH * QW;        // 128
constexpr int KW = 64, KR = 1;     // key tile: one row of 64 columns (128 B of keys: one 128B-swizzle row of V^T)
constexpr int KX_OFF = 16;         // key columns of a tile start at x0 - 16: [x0-16, x0+48) covers [x0-7, x0+38]
constexpr int BN = KW * KR;        // 64
constexpr int DK = 128;
constexpr int DVC = 256;
constexpr int MD = 7;              // max displacement (15 x 15 window)
// tile columns that can fall inside some query's window, widened to whole fp16 pairs
constexpr int C_MIN = (KX_OFF - MD) & ~1, C_MAX = (KX_OFF + QW - 1 + MD) | 1;
static_assert(C_MIN >= 0 && C_MAX < KW, "key tile covers every window");
constexpr int NT = (QH + 2 * MD + KR - 1) / KR;   // 18 key tiles per patch
constexpr int KS = 4, VS = 3, NSB = 4;
constexpr int kWarpK = 8, kWarpV = 9, kWarpMmaS = 10, kWarpMmaPV = 11;
constexpr int kThreads = 12 * 32;

constexpr int SMEM_Q = BM * DK * 2;
constexpr int SMEM_K = BN * DK * 2;
constexpr int SMEM_V = DVC * BN * 2;
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + SMEM_Q;
constexpr int OFF_V = OFF_K + KS * SMEM_K;
constexpr int OFF_MSH = OFF_V + VS * SMEM_V;
constexpr int OFF_LX = OFF_MSH + BM * 4;
constexpr int OFF_BAR = OFF_LX + 2 * BM * 2 * 4;
constexpr int SMEM_TOTAL = OFF_BAR + 256 + 1024;
static_assert(SMEM_TOTAL <= 232448, "shared memory budget");

constexpr int TMEM_COLS = 512;
constexpr int TMEM_O = 0;
constexpr int TMEM_S = 256;

constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 12.0f;
constexpr float M_FLOOR = -1.0e30f;   // finite stand-in for "no key seen yet" (masked scores are -inf)

__device__ long long* g_ltrace = nullptr;   // debug event trace (rmem_debug_attn_trace), null in production

struct LocalParams {
  int h, w, n_dv, tiles_x;
  float scale_log2;
  const float* rel;      // [HW, ldrel] fp32: 15 window rows of rel_pitch floats (15 = reference order, 16 = one aligned
  long long ldrel;       //   64-byte line per window row, last float unused)
  int rel_pitch;
  const t16* gate;       // [HW, ldg] or null
  long long ldg;
  t16* out;              // [HW, ldo]
  long long ldo;
};

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st32u(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
      "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int threads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ uint32_t pack2_fast(float lo, float hi) {
#ifdef RMEM_OPERAND_BF16
  t162 v = __floats2bfloat162_rn(lo, hi);
#else
  t162 v = __floats2half2_rn(lo, hi);
#endif
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(kThreads, 1)
local_attn_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const LocalParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* m_sh = reinterpret_cast<float*>(smem + OFF_MSH);
  float* lx = reinterpret_cast<float*>(smem + OFF_LX);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;                      // [1]
  uint64_t* k_full = q_full + 1;                // [KS]
  uint64_t* k_empty = k_full + KS;
  uint64_t* v_full = k_empty + KS;              // [VS]
  uint64_t* v_empty = v_full + VS;
  uint64_t* s_full = v_empty + VS;              // [NSB]
  uint64_t* p_full = s_full + NSB;
  uint64_t* sp_free = p_full + NSB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sp_free + NSB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* const trace = blockIdx.x == 0 ? g_ltrace : nullptr;   // clock64 events of CTA 0 (debug), 16 slots per tile
#define LTRACE(j, k) do { if (trace && lane == 0) trace[(long long)(j) * 16 + (k)] = clock64(); } while (0)
  if (warp == 0) LTRACE(40, 0);
  const int patch = blockIdx.x / p.n_dv, dvc = blockIdx.x - patch * p.n_dv;
  const int ty = patch / p.tiles_x, tx = patch - ty * p.tiles_x;
  const int y0 = ty * QH, x0 = tx * QW;
  const int ky_base = y0 - MD, kx_base = x0 - KX_OFF;   // key tile j: row ky_base + j; columns kx_base .. +63

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < VS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < NSB; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&sp_free[i], 1); }
    mbar_fence_init();
  }
  if (warp == kWarpMmaS) tmem_alloc<TMEM_COLS>(tmem_slot);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_prologue();   // barriers, TMEM and descriptors are set up; global memory is touched only from here on

  if (warp == kWarpK) {
    // ================================ Q + K producer ================================
    if (elect_one()) {
      tma_prefetch_desc(&map_q);
      tma_prefetch_desc(&map_k);
      mbar_expect_tx(q_full, SMEM_Q);
      // Q patch: 3-D box [64 ch, 32 x, 4 y] of the token-major query map -> rows ordered (y, x)
      tma_load_3d(smem + OFF_Q, &map_q, q_full, 0, x0, y0);
      tma_load_3d(smem + OFF_Q + BM * 128, &map_q, q_full, 64, x0, y0);
    }
    __syncwarp();
    for (int i = 0; i < NT; ++i) {
      const int st = i % KS;
      if (i >= KS) mbar_wait(&k_empty[st], ((i / KS) - 1) & 1, nullptr, 1);
      if (elect_one()) {
        unsigned char* sk = smem + OFF_K + st * SMEM_K;
        mbar_expect_tx(&k_full[st], SMEM_K);
        tma_load_3d(sk, &map_k, &k_full[st], 0, kx_base, ky_base + KR * i);
        tma_load_3d(sk + BN * 128, &map_k, &k_full[st], 64, kx_base, ky_base + KR * i);
      }
      __syncwarp();
    }
  } else if (warp == kWarpV) {
    // ================================ V^T producer ================================
    if (elect_one()) tma_prefetch_desc(&map_v);
    __syncwarp();
    for (int i = 0; i < NT; ++i) {
      const int st = i % VS;
      if (i >= VS) mbar_wait(&v_empty[st], ((i / VS) - 1) & 1, nullptr, 2);
      if (elect_one()) {
        mbar_expect_tx(&v_full[st], SMEM_V);
        tma_load_3d(smem + OFF_V + st * SMEM_V, &map_v, &v_full[st], kx_base, ky_base + KR * i, dvc * DVC);
      }
      __syncwarp();
    }
  } else if (warp == kWarpMmaS) {
    // ================================ S = Q.K^T issuer ================================
    constexpr uint32_t idesc_s = make_idesc(BM, BN);
    const uint32_t smem_base = smem_u32(smem);
    for (int i = 0; i < NT; ++i) {
      const int st = i % KS, b = i % NSB;
      if (i == 0) mbar_wait(q_full, 0, nullptr, 3);
      mbar_wait(&k_full[st], (i / KS) & 1, nullptr, 5);
      if (i >= NSB) mbar_wait(&sp_free[b], ((i - NSB) / NSB) & 1, nullptr, 6);
      fence_after();
      LTRACE(i, 2);
      if (elect_one()) {
        const uint64_t dq = make_desc_sw128(smem_base + OFF_Q);
        const uint64_t dk = make_desc_sw128(smem_base + OFF_K + st * SMEM_K);
        const uint32_t d = tmem + TMEM_S + b * BN;
#pragma unroll
        for (int kk = 0; kk < DK / 16; ++kk) {
          const uint64_t oa = (uint64_t)(((kk >> 2) * (BM * 128) + (kk & 3) * 32) >> 4);
          const uint64_t ob = (uint64_t)(((kk >> 2) * (BN * 128) + (kk & 3) * 32) >> 4);
          umma_ss(d, dq + oa, dk + ob, idesc_s, kk > 0);
        }
        commit(&k_empty[st]);
        commit(&s_full[b]);
      }
      __syncwarp();
      LTRACE(i, 3);
    }
  } else if (warp == kWarpMmaPV) {
    // ================================ O += P.V issuer ================================
    constexpr uint32_t idesc_o = make_idesc(BM, DVC);
    const uint32_t smem_base = smem_u32(smem);
    for (int j = 0; j < NT; ++j) {
      const int b = j % NSB, sv = j % VS;
      mbar_wait(&p_full[b], (j / NSB) & 1, nullptr, 7);
      LTRACE(j, 0);
      mbar_wait(&v_full[sv], (j / VS) & 1, nullptr, 9);
      fence_after();
      LTRACE(j, 10);
      if (elect_one()) {
        const uint64_t dv = make_desc_sw128(smem_base + OFF_V + sv * SMEM_V);
        const uint32_t pa = tmem + TMEM_S + b * BN;
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk)
          umma_ts(tmem + TMEM_O, pa + kk * 8, dv + (uint64_t)(kk * 2), idesc_o, (j == 0 && kk == 0) ? 0u : 1u);
        commit(&v_empty[sv]);
        commit(&sp_free[b]);
      }
      __syncwarp();
      LTRACE(j, 1);
    }
  } else {
    // ================================ softmax + epilogue (warps 0-7) ================================
    const int quad = warp & 3, grp = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    const int id_in = grp == 1 ? 1 + quad : 5 + quad;
    const int id_out = grp == 0 ? 1 + quad : 5 + quad;
    const int id_ex = 9 + quad;
    const int qy = y0 + row / QW, qx = x0 + row % QW;
    const bool row_ok = qy < p.h && qx < p.w;
    const long long qi = (long long)qy * p.w + qx;
    const float* relrow = p.rel + (row_ok ? qi : 0) * p.ldrel;
    // valid key columns of this query inside a key tile: kx = kx_base + c, |kx - qx| <= 7, 0 <= kx < w
    const int c_lo = max(qx - MD, 0) - kx_base, c_hi = min(qx + MD, p.w - 1) - kx_base;      // inclusive, in [0, 31]
    const int dx0 = kx_base - qx + MD;                                                     // window column of c = 0
    float m_ref = M_FLOOR, l_tot = 0.f;
    // rel_pitch 16: the 15 biases of one window row are one 64-byte line per query: four 16-byte loads, issued one
    // own-tile ahead.  (One scalar load per (query, key) -- 32 different lines per instruction -- made this kernel
    // LSU-bound at ~4300 cycles per key tile.)
    const bool rel16 = p.rel_pitch == 16;
    // Key column c of a tile is window column c + dx0, so the thread wants its 16-float line rotated left by
    // rot = dx0 mod 16 (then bias(c) = line[c & 15], a compile-time register).  rot is fixed per thread: whole
    // float4s are rotated for free by the load addresses, the remaining 0..3 floats by two select stages.
    const int rot = dx0 & 15, rot4 = rot >> 2;
    float4 rnext[4];
    auto load_rel = [&](int jj) {
      const int dy = ky_base + KR * jj - qy + MD;
      const bool in = jj < NT && row_ok && (unsigned)dy <= (unsigned)(2 * MD);
      const float4* src = reinterpret_cast<const float4*>(relrow + (in ? dy : 0) * 16);
#pragma unroll
      for (int e = 0; e < 4; ++e) rnext[e] = in ? src[(e + rot4) & 3] : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    if (rel16) load_rel(grp);

    for (int j = 0; j < NT; ++j) {
      if ((j & 1) != grp) continue;
      const int b = j % NSB;
      if (quad == 0) LTRACE(j, 4);
      mbar_wait(&s_full[b], (j / NSB) & 1, nullptr, 10);
      fence_after();
      if (quad == 0) LTRACE(j, 5);
      float sc[64];
      {
        uint32_t r0[32], r1[32];
        tmem_ld32_nowait(lane_addr + TMEM_S + b * BN, r0);
        tmem_ld32_nowait(lane_addr + TMEM_S + b * BN + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) { sc[c] = __uint_as_float(r0[c]); sc[32 + c] = __uint_as_float(r1[c]); }
      }
      // x = scale*s + rel (log2 units) inside the window, -inf outside; rows of the patch beyond the frame see x = 0
      float mt = -INFINITY;
      if (rel16) {
        static_assert(KR == 1, "one key row per tile");
        float R[16];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          R[4 * e] = rnext[e].x * LOG2E; R[4 * e + 1] = rnext[e].y * LOG2E;
          R[4 * e + 2] = rnext[e].z * LOG2E; R[4 * e + 3] = rnext[e].w * LOG2E;
        }
        load_rel(j + 2);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const bool on = (rot >> b) & 1;
          float T[16];
#pragma unroll
          for (int m = 0; m < 16; ++m) T[m] = on ? R[(m + (1 << b)) & 15] : R[m];
#pragma unroll
          for (int m = 0; m < 16; ++m) R[m] = T[m];
        }
        const int ky = ky_base + j;
        const int dy = ky - qy + MD;
        const bool row_in = row_ok && (unsigned)dy <= (unsigned)(2 * MD) && (unsigned)ky < (unsigned)p.h;
        // Columns outside [C_MIN, C_MAX] are outside every query's window (the exp loop skips them as well).  Rows of
        // the patch beyond the frame are fully masked: M_FLOOR keeps their exponentials at 0, nothing is stored.
#pragma unroll
        for (int c = C_MIN; c <= C_MAX; ++c) {
          const bool ok = row_in && c >= c_lo && c <= c_hi;
          const float x = ok ? fmaf(sc[c], p.scale_log2, R[c & 15]) : -INFINITY;
          sc[c] = x;
          mt = fmaxf(mt, x);
        }
      } else
#pragma unroll
      for (int kr = 0; kr < KR; ++kr) {
        const int ky = ky_base + KR * j + kr;
        const int dy = ky - qy + MD;
        const bool row_in = row_ok && (unsigned)dy <= (unsigned)(2 * MD) && (unsigned)ky < (unsigned)p.h;
        const float* rr = relrow + dy * p.rel_pitch + dx0;
#pragma unroll
        for (int c = 0; c < KW; ++c) {
          const bool ok = row_in && c >= c_lo && c <= c_hi;
          float x = -INFINITY;
          if (ok) x = fmaf(sc[kr * KW + c], p.scale_log2, rr[c] * LOG2E);
          if (!row_ok) x = 0.f;
          sc[kr * KW + c] = x;
          mt = fmaxf(mt, x);
        }
      }
      if (quad == 0) LTRACE(j, 6);
      // ---- hand-over of the lazily updated row maximum from the other group's tile j-1 ----
      float m_prev = M_FLOOR;
      if (j > 0) {
        named_bar_sync(id_in, 64);
        m_prev = m_sh[row];
      }
      const bool need = mt > m_prev + RESCALE_THRESHOLD;
      float m_new = m_prev;
      if (__any_sync(0xffffffffu, need)) {
        if (j > 0) {
          mbar_wait(&sp_free[(j - 1) % NSB], ((j - 1) / NSB) & 1, nullptr, 11);
          fence_after();
          const float f = need ? exp2f(m_prev - mt) : 1.f;
#pragma unroll 1
          for (int c = 0; c < DVC; c += 32) {
            float o[32];
            tmem_ld32(lane_addr + TMEM_O + c, o);
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] *= f;
            tmem_st32(lane_addr + TMEM_O + c, o);
          }
          fence_before();
        }
        if (need) m_new = mt;
      }
      if (j + 1 < NT) {
        m_sh[row] = m_new;
        named_bar_arrive(id_out, 64);
      }
      if (m_new != m_ref) {
        l_tot *= exp2f(m_ref - m_new);
        m_ref = m_new;
      }
      if (quad == 0) LTRACE(j, 7);
      uint32_t pk[32];
      float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
      for (int c = 0; c < 64; c += 2) {
        if (c + 1 < C_MIN || c > C_MAX) { pk[c >> 1] = 0u; continue; }   // never inside a window (compile time)
        const float e0 = exp2f(sc[c] - m_new);          // exp2(-inf) = 0 for masked keys
        const float e1 = exp2f(sc[c + 1] - m_new);
        ls0 += e0; ls1 += e1;
        pk[c >> 1] = pack2_fast(e0, e1);
      }
      l_tot += ls0 + ls1;
      if (quad == 0) LTRACE(j, 8);
      tmem_st32u(lane_addr + TMEM_S + b * BN, pk);
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[b]);
      if (quad == 0) LTRACE(j, 9);
    }
    if (warp == 0) LTRACE(40, 1);

    // ---- epilogue: out = O / l * gate ----
    // The gate rows this lane will write (two rows per store instruction, see below) are fetched now, all sixteen at
    // once: their L2 latency overlaps the wait for the last P.V instead of being paid once per row pair.
    const int ecol = dvc * DVC + grp * (DVC / 2) + (lane & 15) * 8;
    uint4 gq[16];
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      const int prow = quad * 32 + it * 2 + (lane >> 4);
      const int oy = y0 + prow / QW, ox = x0 + prow % QW;
      gq[it] = make_uint4(0u, 0u, 0u, 0u);
      if (p.gate && oy < p.h && ox < p.w)
        gq[it] = *reinterpret_cast<const uint4*>(p.gate + ((long long)oy * p.w + ox) * p.ldg + ecol);
    }
    lx[(grp * BM + row) * 2 + 0] = m_ref;
    lx[(grp * BM + row) * 2 + 1] = l_tot;
    named_bar_sync(id_ex, 64);
    const float m_o = lx[((grp ^ 1) * BM + row) * 2 + 0], l_o = lx[((grp ^ 1) * BM + row) * 2 + 1];
    const float M = fmaxf(m_ref, m_o);
    const float l_row = l_tot * exp2f(m_ref - M) + l_o * exp2f(m_o - M);
    const float inv = 1.f / l_row;
    mbar_wait(&sp_free[(NT - 1) % NSB], ((NT - 1) / NSB) & 1, nullptr, 12);
    fence_after();
    if (warp == 0) LTRACE(40, 2);
    const int col0 = dvc * DVC + grp * (DVC / 2);
    // Every MMA has retired and every TMA tile has been consumed: the K/V rings are free.  Each warp stages its
    // 32 rows x 128 columns (fp32, normalised) there and writes them out two rows per instruction -- 16 lanes x 16 bytes
    // per row, whole lines -- with the gate read the same way.  (One row per lane costs a line per lane per access:
    // ~10000 cycles for this epilogue.)
    constexpr int ROWP = (DVC / 2) * 4 + 16;                 // staged row pitch in bytes (+16: conflict-free float4 phases)
    static_assert(8 * 32 * ROWP <= KS * SMEM_K + VS * SMEM_V, "staging fits in the rings");
    unsigned char* stg = smem + OFF_K + warp * (32 * ROWP);
#pragma unroll 1
    for (int c = 0; c < DVC / 2; c += 32) {
      float o[32];
      tmem_ld32(lane_addr + TMEM_O + grp * (DVC / 2) + c, o);
      float4* d = reinterpret_cast<float4*>(stg + lane * ROWP + c * 4);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        d[e] = make_float4(o[4 * e] * inv, o[4 * e + 1] * inv, o[4 * e + 2] * inv, o[4 * e + 3] * inv);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      const int rr = it * 2 + (lane >> 4), cc = (lane & 15) * 8;
      const int prow = quad * 32 + rr;
      const int oy = y0 + prow / QW, ox = x0 + prow % QW;
      if (oy < p.h && ox < p.w) {
        const long long oi = (long long)oy * p.w + ox;
        const float4* sp = reinterpret_cast<const float4*>(stg + rr * ROWP + cc * 4);
        const float4 a = sp[0], b = sp[1];
        float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        if (p.gate) {
          const uint4 g = gq[it];
          const float2 g0 = unpack2(g.x), g1 = unpack2(g.y), g2 = unpack2(g.z), g3 = unpack2(g.w);
          v[0] *= g0.x; v[1] *= g0.y; v[2] *= g1.x; v[3] *= g1.y; v[4] *= g2.x; v[5] *= g2.y; v[6] *= g3.x; v[7] *= g3.y;
        }
        uint4 u;
        u.x = pack2(v[0], v[1]); u.y = pack2(v[2], v[3]); u.z = pack2(v[4], v[5]); u.w = pack2(v[6], v[7]);
        *reinterpret_cast<uint4*>(p.out + oi * p.ldo + col0 + cc) = u;
      }
    }
    fence_before();
    if (warp == 0) LTRACE(40, 3);
  }
  fence_before();
  __syncthreads();
  if (warp == kWarpMmaS) {
    fence_after();
    tmem_dealloc<TMEM_COLS>(tmem);
  }
}

// v token-major [h*w, Dv] (row stride ldv) -> value-major, row-padded [Dv][h][wp] (pad columns zero)
__global__ void transpose_pad_kernel(const t16* __restrict__ x, long long ldx, t16* __restrict__ y, int h, int w, int wp,
                                     int C) {
  pdl_prologue();
  __shared__ t16 tile[64][66];
  const int yy = blockIdx.z;                       // frame row
  const int x0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  for (int i = threadIdx.y; i < 64; i += blockDim.y) {
    const int px = x0 + i, c = c0 + threadIdx.x * 2;
    t16 a = f2t(0.f), b = a;
    if (px < w && c < C) {
      const t16* s = x + ((long long)yy * w + px) * ldx + c;
      a = s[0];
      if (c + 1 < C) b = s[1];
    }
    tile[i][threadIdx.x * 2] = a;
    tile[i][threadIdx.x * 2 + 1] = b;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 64; i += blockDim.y) {
    const int c = c0 + i;
    if (c >= C) continue;
    const int px = x0 + threadIdx.x * 2;
    t16* d = y + ((long long)c * h + yy) * wp;
    if (px < wp) d[px] = tile[threadIdx.x * 2][i];
    if (px + 1 < wp) d[px + 1] = tile[threadIdx.x * 2 + 1][i];
  }
}

}  // namespace

size_t local_attn_tc_workspace(int h, int w, int Dv) {
  const int wp = round_up(w, 8);
  return (size_t)Dv * h * wp * sizeof(t16) + 256;
}

int local_attn_tc_prepare_v(const t16* v, long long ldv, int h, int w, int Dv, void* workspace, size_t workspace_bytes,
                            cudaStream_t s) {
  RMEM_REQUIRE(workspace && workspace_bytes >= local_attn_tc_workspace(h, w, Dv), "local_attn_tc: workspace too small");
  RMEM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "local_attn_tc: workspace alignment");
  const int wp = round_up(w, 8);
  dim3 grid(cdiv(wp, 64), cdiv(Dv, 64), h), block(32, 8);
  RMEM_CUDA_CHECK(launch_pdl(transpose_pad_kernel, dim3(grid), dim3(block), 0, s, v, ldv, reinterpret_cast<t16*>(workspace),
                             h, w, wp, Dv));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int local_attn_tc(const t16* q, long long ldq, const t16* k, long long ldk, const t16* v, long long ldv,
                  const float* rel, long long ldrel, int rel_pitch, const t16* gate, long long ldg, t16* out, long long ldo, int h,
                  int w, int Dv, float scale, void* workspace, size_t workspace_bytes, cudaStream_t s, bool v_prepared) {
  RMEM_REQUIRE(Dv % DVC == 0, "local_attn_tc: Dv=%d must be a multiple of 256", Dv);
  RMEM_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && (!gate || ldg % 8 == 0),
               "local_attn_tc: row strides must keep 16B alignment");
  RMEM_REQUIRE(workspace && workspace_bytes >= local_attn_tc_workspace(h, w, Dv), "local_attn_tc: workspace too small");
  RMEM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "local_attn_tc: workspace alignment");
  const int wp = round_up(w, 8);
  t16* vt = reinterpret_cast<t16*>(workspace);
  if (!v_prepared) RMEM_TRY(local_attn_tc_prepare_v(v, ldv, h, w, Dv, workspace, workspace_bytes, s));
  const CUtensorMap *mq, *mk, *mv;
  {
    uint64_t dims[3] = {(uint64_t)DK, (uint64_t)w, (uint64_t)h};
    uint64_t str[2] = {(uint64_t)ldq * 2, (uint64_t)ldq * 2 * w};
    uint32_t box[3] = {64, (uint32_t)QW, (uint32_t)QH};
    RMEM_TRY(tma_encode_cached(&mq, q, 3, dims, str, box, nullptr));
  }
  {
    uint64_t dims[3] = {(uint64_t)DK, (uint64_t)w, (uint64_t)h};
    uint64_t str[2] = {(uint64_t)ldk * 2, (uint64_t)ldk * 2 * w};
    uint32_t box[3] = {64, (uint32_t)KW, (uint32_t)KR};
    RMEM_TRY(tma_encode_cached(&mk, k, 3, dims, str, box, nullptr));
  }
  {
    uint64_t dims[3] = {(uint64_t)wp, (uint64_t)h, (uint64_t)Dv};
    uint64_t str[2] = {(uint64_t)wp * 2, (uint64_t)wp * 2 * h};
    uint32_t box[3] = {(uint32_t)KW, (uint32_t)KR, (uint32_t)DVC};
    RMEM_TRY(tma_encode_cached(&mv, vt, 3, dims, str, box, nullptr));
  }
  LocalParams p;
  p.h = h; p.w = w; p.n_dv = Dv / DVC; p.tiles_x = cdiv(w, QW);
  p.scale_log2 = scale * LOG2E;
  RMEM_REQUIRE(rel_pitch == 15 || (rel_pitch == 16 && ldrel % 4 == 0 && (reinterpret_cast<uintptr_t>(rel) & 15) == 0),
               "local_attn_tc: rel_pitch must be 15, or 16 with 16-byte aligned rows");
  p.rel = rel; p.ldrel = ldrel; p.rel_pitch = rel_pitch; p.gate = gate; p.ldg = ldg; p.out = out; p.ldo = ldo;
  static bool attr_done = false;
  if (!attr_done) {
    RMEM_CUDA_CHECK(cudaFuncSetAttribute(local_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    attr_done = true;
  }
  const int patches = cdiv(h, QH) * p.tiles_x;
  RMEM_CUDA_CHECK(launch_pdl(local_attn_tc_kernel, dim3(patches * p.n_dv), dim3(kThreads), SMEM_TOTAL, s, *mq, *mk, *mv, p));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int local_attn_tc_set_trace(long long* dev_buf) {
  RMEM_CUDA_CHECK(cudaMemcpyToSymbol(g_ltrace, &dev_buf, sizeof(dev_buf)));
  return RMEM_OK;
}

}  // namespace rmem
