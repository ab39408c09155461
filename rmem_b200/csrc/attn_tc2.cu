// Long-term / self attention for sm_100a, third iteration  (K1 + K1b + K8 of SURVEY.md; attention.py:174-193,
// transformer.py:1140-1197).  RMEM_ATTN_TC2.
//
// Why it looks like this (profiles/r01_*): the first two versions were bound by the LATENCY of the dependency loop
// S = Q.K^T -> softmax -> P.V, not by any pipe: tcgen05.mma groups take ~400 cycles from issue to a visible commit on
// top of their throughput (tools/ubench/mma_rate.cu: 8 x N=64 MMAs = 757 cycles alone, 49 cycles each in a stream),
// and one set of softmax warps needed ~1600 cycles per 64-key tile.  This version removes the loop:
//   * S runs up to three tiles ahead of P.V in three TMEM score buffers, issued by its own warp (tcgen05.mma issue blocks
//     the issuing thread for roughly the MMA's execution time, so one issuing warp per MMA stream), so softmax never
//     waits for scores;
//   * two softmax warp groups (4 warps each, one query row per thread) alternate KV tiles; the only value that has to
//     travel from tile j to tile j+1 is the lazily-updated row maximum, handed over through 512 B of shared memory and
//     a 64-thread named barrier per TMEM lane quadrant -- the exp / pack / store phase of tile j overlaps the load /
//     max phase of tile j+1;
//   * P is written back as packed fp16 over the first 32 columns of its own score buffer (each thread overwrites only
//     the row it has already read) and is the TMEM A operand of the P.V MMA: no shared-memory round trip;
//   * stream-K static schedule: the (query tile, Dv chunk, KV tile) space is cut into one contiguous range per SM, each
//     CTA flushes at most two segments as normalised fp16 rows + fp32 (m, l); per-frame (m, l) pieces give the mass;
//   * shared-memory bandwidth (MMA operand reads + TMA fills, 128 B/cycle) turned out to be the binding resource, so the
//     query tile is kept in TMEM as the A operand of S (tcgen05.mma with A in TMEM): S reads only K from shared memory;
//   * K and V^T tiles travel in separate TMA rings (6 x 16 KB, 4 x 32 KB; K is needed earlier than V);
//   * what bounds the steady state now is the ~48 B/clk one SM can pull from L2: 48 KB of K + V^T per 64-key step is
//     ~1000 cycles, against ~780 cycles of tensor work -- the next step is cta_group::2;
//   * the static schedule equalises tiles + segment epilogues across CTAs (make_bounds), partial O is stored in
//     16-column groups so the row-per-lane TMEM read-out writes whole lines.
//
//   block = 384 threads: warps 0-3 softmax group 0 (even tiles), 4-7 group 1 (odd tiles), 8 K producer,
//                        9 V producer, 10 S issuer + TMEM owner, 11 P.V issuer
//   TMEM (512 cols): O[256] | 3 x S/P[64] | Q[64]
//   smem: K 6 x 16 KB | V^T 4 x 32 KB | row-max hand-over | barriers
#include <cstdlib>

#include "attn.cuh"
#include "tcgen05.cuh"

namespace rmem {

namespace {

using namespace tc;

constexpr int BM = 128;        // query rows per CTA
constexpr int BN = 64;         // keys per KV tile
constexpr int DK = 128;
constexpr int DVC = 256;       // Dv columns per unit
constexpr int KS = 6;          // K ring depth
constexpr int VS = 4;          // V ring depth
constexpr int NSB = 3;         // score buffers in TMEM; P (packed fp16) aliases the first 32 columns of its S buffer
constexpr int kSoftmaxWarps = 8;
constexpr int kWarpK = 8, kWarpV = 9, kWarpMmaS = 10, kWarpMmaPV = 11;
constexpr int kThreads = 12 * 32;

constexpr int SMEM_K = BN * DK * 2;            // 16 KB
constexpr int SMEM_V = DVC * BN * 2;           // 32 KB
constexpr int OFF_K = 0;
constexpr int OFF_V = OFF_K + KS * SMEM_K;
constexpr int OFF_MSH = OFF_V + VS * SMEM_V;              // float [128]        row-max hand-over
constexpr int OFF_LX = OFF_MSH + BM * 4;                  // float [2][128][2]  (m, l) exchange at segment end
constexpr int OFF_BAR = OFF_LX + 2 * BM * 2 * 4;
constexpr int SMEM_TOTAL = OFF_BAR + 256;   // no alignment slack: the rings only fit in 227 KB if the dynamic base is
                                            // 1024B-aligned already (it follows the 1 KB driver reservation); checked below
static_assert(SMEM_TOTAL <= 232448, "shared memory budget");

constexpr int TMEM_COLS = 512;
constexpr int TMEM_O = 0;
constexpr int TMEM_S = 256;    // 3 x 64 fp32 score columns; P(j) overwrites the first 32 of S(j) = A operand of O += P.V
constexpr int TMEM_Q = 448;    // 64 columns: the 128 x 128 fp16 query tile as packed pairs = TMEM A operand of S = Q.K^T

constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 12.0f;     // log2 units: P <= 2^12 (fp16 max 2^16) before a lazy rescale is forced

constexpr int kMaxCTA = 192;   // SMs a launch may use (B200: 148)
struct Tc2Params {
  int HW, HWp, T, tpf, TPU, n_units, n_dv, nCTA, Dv;
  int bounds[kMaxCTA + 1];     // CTA c owns steps [bounds[c], bounds[c+1]) of the (unit, tile) sequence, see make_bounds()
  long long L;                 // n_units * TPU
  int slot[kMaxBankFrames];
  float scale_log2;            // scale * log2(e)
  const t16* q;                // [HW, 128] query (cur PE already added), row-major
  const float* qbias;          // [HW, T] or null (already multiplied by scale)
  t16* part_o;                 // [nCTA][2][BM][DVC]     normalised partial O
  float* part_ml;              // [nCTA][2][BM][2]       (m in log2 units, l)
  float* pieces;               // [nCTA][2][T][2][BM][2] per-frame (m, l) of each softmax group (Dv chunk 0 units) or null
  // direct mode (one CTA = one whole unit, e.g. the T = 1 self-attention): the kernel normalises, gates and writes the
  // final rows itself; no partials, no combine launch
  int direct;
  const t16* gate;
  long long ldg;
  t16* out;
  long long ldo;
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

// 2^x on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial, relative error 7.5e-5 -- below the fp16
// rounding of P).  The MUFU unit evaluates 16 ex2 per cycle and SM: 8192 per 64-key tile = 512 cycles, more than the
// tile's MMA time; a quarter of the exponentials can bypass it (kPolyExp).
constexpr bool kPolyExp = false;   // measured: no gain (the tile period is set by the MMA pipe, not by MUFU)
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;                         // 1.5 * 2^23: nearest integer lands in the low mantissa bits
  const float f = x - (t - 12582912.f);                   // [-0.5, 0.5]
  float r = fmaf(f, 0.0551716648f, 0.2426111251f);
  r = fmaf(r, f, 0.6932609677f);
  r = fmaf(r, f, 0.9999280572f);
  return __int_as_float(__float_as_int(r) + (__float_as_int(t) << 23));
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

// unsaturated fp32x2 -> t16x2 (values are bounded by 2^8 here)
__device__ __forceinline__ uint32_t pack2_fast(float lo, float hi) {
#ifdef RMEM_OPERAND_BF16
  t162 v = __floats2bfloat162_rn(lo, hi);
#else
  t162 v = __floats2half2_rn(lo, hi);
#endif
  return *reinterpret_cast<uint32_t*>(&v);
}

// Optional event trace of CTA 0 (clock64 per pipeline event, 16 slots per tile); null in production.
__device__ long long* g_trace = nullptr;
__device__ int g_trace_cta = 0;
#define TRACE(j, k)                                                                  \
  do {                                                                               \
    if (trace && lane == 0) trace[(long long)(j) * 16 + (k)] = clock64();            \
  } while (0)

struct Seg { int unit, lo, hi; };   // tiles [lo, hi) of the unit

__device__ __forceinline__ void cta_range(const Tc2Params& p, int cta, long long& lo, long long& hi) {
  lo = p.bounds[cta];
  hi = p.bounds[cta + 1];
}

__global__ void __launch_bounds__(kThreads, 1)
long_attn_tc2_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                     const Tc2Params p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();       // 128B-swizzled TMA / UMMA tiles need 1024B alignment
  float* m_sh = reinterpret_cast<float*>(smem + OFF_MSH);
  float* lx = reinterpret_cast<float*>(smem + OFF_LX);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_ready = bars;                     // Q tile stored to TMEM (one phase per segment that (re)loads it)
  uint64_t* k_full = q_ready + 1;               // [KS]
  uint64_t* k_empty = k_full + KS;              // [KS]
  uint64_t* v_full = k_empty + KS;              // [VS]
  uint64_t* v_empty = v_full + VS;              // [VS]
  uint64_t* s_full = v_empty + VS;              // [NSB]  S(j) complete
  uint64_t* p_full = s_full + NSB;              // [NSB]  P(j) stored by its softmax group
  uint64_t* sp_free = p_full + NSB;             // [NSB]  P.V(j) complete: score buffer (and everything before) retired
  uint64_t* o_drained = sp_free + NSB;          // segment epilogue has read O
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_drained + 1);
  static_assert((1 + 2 * KS + 2 * VS + 3 * NSB + 1) * 8 + 4 <= 256, "barrier area");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* const trace = (int)blockIdx.x == g_trace_cta ? g_trace : nullptr;
  // per-CTA wall times (rows 100.. of the trace buffer): globaltimer start/end, tiles, segments, SM id
  long long* const cta_times = (g_trace && blockIdx.x < 148) ? g_trace + (100 + blockIdx.x) * 16 : nullptr;
  if (cta_times && threadIdx.x == 0) {
    unsigned long long gt; unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    cta_times[0] = (long long)gt; cta_times[4] = smid; cta_times[5] = clock64();
  }

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
    mbar_init(q_ready, 4);
    for (int i = 0; i < KS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < VS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < NSB; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&sp_free[i], 1);
    }
    mbar_init(o_drained, kSoftmaxWarps);
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
    if (ntot > 0) {
      if (elect_one()) {
        tma_prefetch_desc(&map_k);
      }
      __syncwarp();
      int i = 0;
      for (int s = 0; s < nseg; ++s) {
        int t = seg[s].lo / p.tpf, jt = seg[s].lo - t * p.tpf;
        for (int g = seg[s].lo; g < seg[s].hi; ++g, ++i, ++jt) {
          if (jt == p.tpf) { jt = 0; ++t; }
          const int st = i % KS;
          if (i >= KS) mbar_wait(&k_empty[st], ((i / KS) - 1) & 1, nullptr, 1);
          if (elect_one()) {
            const int key0 = p.slot[t] * p.HWp + jt * BN;
            unsigned char* sk = smem + OFF_K + st * SMEM_K;
            mbar_expect_tx(&k_full[st], SMEM_K);
            tma_load_2d(sk, &map_k, &k_full[st], 0, key0);
            tma_load_2d(sk + BN * 128, &map_k, &k_full[st], 64, key0);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kWarpV) {
    // ================================ V^T producer ================================
    if (ntot > 0) {
      if (elect_one()) tma_prefetch_desc(&map_v);
      __syncwarp();
      int i = 0;
      for (int s = 0; s < nseg; ++s) {
        const int dv0 = (seg[s].unit % p.n_dv) * DVC;
        int t = seg[s].lo / p.tpf, jt = seg[s].lo - t * p.tpf;
        for (int g = seg[s].lo; g < seg[s].hi; ++g, ++i, ++jt) {
          if (jt == p.tpf) { jt = 0; ++t; }
          const int st = i % VS;
          if (i >= VS) mbar_wait(&v_empty[st], ((i / VS) - 1) & 1, nullptr, 2);
          if (elect_one()) {
            const int key0 = p.slot[t] * p.HWp + jt * BN;
            mbar_expect_tx(&v_full[st], SMEM_V);
            tma_load_2d(smem + OFF_V + st * SMEM_V, &map_v, &v_full[st], key0, dv0);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kWarpMmaS) {
    // ================================ S = Q.K^T issuer ================================
    // tcgen05.mma issue blocks the issuing thread for about the MMA's execution time, and every mbarrier wait costs
    // ~90 cycles even when already complete (trace in profiles/r01_attn_trace.txt), so the two MMA streams get one
    // warp each: this one runs up to four score tiles ahead, bounded only by the K ring and the score buffers.
    if (ntot > 0) {
      constexpr uint32_t idesc_s = make_idesc(BM, BN);
      const uint32_t smem_base = smem_u32(smem);
      for (int i = 0; i < ntot; ++i) {
        const int st = i % KS, b = i % NSB;
        if (i == 0) mbar_wait(q_ready, 0, nullptr, 3);
        if (i == n0 && q_reload1) mbar_wait(q_ready, 1, nullptr, 4);
        mbar_wait(&k_full[st], (i / KS) & 1, nullptr, 5);
        if (i >= NSB) mbar_wait(&sp_free[b], ((i - NSB) / NSB) & 1, nullptr, 6);    // P.V(i-3) retired its buffer
        fence_after();
        TRACE(i, 2);
        if (elect_one()) {
          const uint64_t dk = make_desc_sw128(smem_base + OFF_K + st * SMEM_K);
          const uint32_t d = tmem + TMEM_S + b * BN;
#pragma unroll
          for (int kk = 0; kk < DK / 16; ++kk) {
            // A = Q from TMEM (16 k-elements = 8 packed columns); B: 32 B inside the 128 B swizzle atom, the second
            // 64-wide atom of the K tile starts one tile-half later
            const uint64_t ob = (uint64_t)(((kk >> 2) * (BN * 128) + (kk & 3) * 32) >> 4);
            umma_ts(d, tmem + TMEM_Q + kk * 8, dk + ob, idesc_s, kk > 0);
          }
          commit(&k_empty[st]);
          commit(&s_full[b]);
        }
        __syncwarp();
        TRACE(i, 3);
      }
    }
  } else if (warp == kWarpMmaPV) {
    // ================================ O += P.V issuer ================================
    if (ntot > 0) {
      constexpr uint32_t idesc_o = make_idesc(BM, DVC);
      const uint32_t smem_base = smem_u32(smem);
      for (int j = 0; j < ntot; ++j) {
        const int b = j % NSB, sv = j % VS;
        mbar_wait(&p_full[b], (j / NSB) & 1, nullptr, 7);
        TRACE(j, 0);
        const bool first = (j == 0) || (j == n0);
        if (j == n0 && nseg > 1) mbar_wait(o_drained, 0, nullptr, 8);
        mbar_wait(&v_full[sv], (j / VS) & 1, nullptr, 9);
        fence_after();
        if (elect_one()) {
          const uint64_t dv = make_desc_sw128(smem_base + OFF_V + sv * SMEM_V);
          const uint32_t pa = tmem + TMEM_S + b * BN;
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk)
            umma_ts(tmem + TMEM_O, pa + kk * 8, dv + (uint64_t)(kk * 2), idesc_o, (first && kk == 0) ? 0u : 1u);
          commit(&v_empty[sv]);
          commit(&sp_free[b]);
        }
        __syncwarp();
        TRACE(j, 1);
      }
    }
  } else if (ntot > 0) {
    // ================================ softmax + epilogue (warps 0-7) ================================
    const int quad = warp & 3, grp = warp >> 2;
    const int row = quad * 32 + lane;                       // tile row == TMEM lane
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    // named barriers: group 0 -> group 1 hand-over, group 1 -> group 0 hand-over, segment-end exchange
    const int id_in = grp == 1 ? 1 + quad : 5 + quad;
    const int id_out = grp == 0 ? 1 + quad : 5 + quad;
    const int id_ex = 9 + quad;
    int j = 0;
    // The 128 x 128 fp16 query tile lives in TMEM as the A operand of S = Q.K^T (one row per lane, packed pairs): the
    // score MMAs then read only K from shared memory, whose bandwidth (operand reads + TMA fills) bounds this kernel.
    auto store_q = [&](int qt, uint64_t* after, uint32_t parity) {
      const int qr = qt * BM + row;
      uint32_t w[64];
      if (qr < p.HW) {
        const uint4* src = reinterpret_cast<const uint4*>(p.q + (long long)qr * DK);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const uint4 u = src[c];
          w[c * 4] = u.x; w[c * 4 + 1] = u.y; w[c * 4 + 2] = u.z; w[c * 4 + 3] = u.w;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 64; ++c) w[c] = 0u;
      }
      if (after) {                                          // the rows are in registers; now wait until Q may be replaced
        mbar_wait(after, parity, nullptr, 14);
        fence_after();
      }
      tmem_st32u(lane_addr + TMEM_Q, w);
      tmem_st32u(lane_addr + TMEM_Q + 32, w + 32);
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_ready);
    };
    if (grp == 0) store_q(seg[0].unit / p.n_dv, nullptr, 0);

    for (int s = 0; s < nseg; ++s) {
      const int unit = seg[s].unit;
      const int qt = unit / p.n_dv, dvc = unit % p.n_dv;
      const int qi = qt * BM + row;
      const bool row_ok = qi < p.HW;
      // m_ref: the running maximum this group's l / l_piece are relative to
      float m_ref = -INFINITY, l_tot = 0.f, l_piece = 0.f, bias2 = 0.f;
      int cur_t = -1;
      const int j_first = j;
      float* piece_base = (p.pieces && dvc == 0)
                              ? p.pieces + (((long long)(blockIdx.x * 2 + s) * p.T) * 2 + grp) * (BM * 2) + row * 2
                              : nullptr;
      auto flush_piece = [&](int t) {
        if (piece_base) {
          float* d = piece_base + (long long)t * (2 * BM * 2);
          d[0] = m_ref;
          d[1] = l_piece;
        }
      };
      int t = seg[s].lo / p.tpf, jt = seg[s].lo - t * p.tpf;
      for (int g = seg[s].lo; g < seg[s].hi; ++g, ++j, ++jt) {
        if (jt == p.tpf) { jt = 0; ++t; }
        if (t != cur_t) {                                   // both groups walk every tile's frame index
          if (cur_t >= 0) flush_piece(cur_t);
          cur_t = t;
          l_piece = 0.f;
          bias2 = (p.qbias && row_ok) ? p.qbias[(long long)qi * p.T + t] * LOG2E : 0.f;
        }
        if ((j & 1) != grp) {
          // The other group owns the first segment's last tile; this one is idle until the epilogue, so it brings in
          // the next segment's query tile: rows to registers, wait for S(n0-1) (= every score MMA of the segment has
          // read Q), then into TMEM.  Off the critical path of the last tile.
          if (q_reload1 && j == n0 - 1) store_q(seg[1].unit / p.n_dv, &s_full[j % NSB], (j / NSB) & 1);
          continue;
        }
        const int b = j % NSB;
        long long* const trace_s = quad == 0 ? trace : nullptr;
#define TRACE_S(k) do { if (trace_s && lane == 0) trace_s[(long long)j * 16 + (k)] = clock64(); } while (0)
        TRACE_S(4);
        mbar_wait(&s_full[b], (j / NSB) & 1, nullptr, 10);
        fence_after();
        TRACE_S(5);
        float sc[64];
        {
          uint32_t r0[32], r1[32];
          tmem_ld32_nowait(lane_addr + TMEM_S + b * BN, r0);
          tmem_ld32_nowait(lane_addr + TMEM_S + b * BN + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) { sc[c] = __uint_as_float(r0[c]); sc[32 + c] = __uint_as_float(r1[c]); }
        }
        if (jt == p.tpf - 1) {                               // ragged last tile of the frame
          const int key0 = jt * BN;
#pragma unroll
          for (int c = 0; c < 64; ++c) sc[c] = (key0 + c < p.HW) ? sc[c] : -INFINITY;
        }
        // row maximum: balanced tree
        float mx;
        {
          float a[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) a[c] = fmaxf(fmaxf(sc[c], sc[16 + c]), fmaxf(sc[32 + c], sc[48 + c]));
#pragma unroll
          for (int c = 0; c < 4; ++c) a[c] = fmaxf(fmaxf(a[c], a[4 + c]), fmaxf(a[8 + c], a[12 + c]));
          mx = fmaxf(fmaxf(a[0], a[1]), fmaxf(a[2], a[3]));
        }
        const float mt = fmaf(mx, p.scale_log2, bias2);      // scale > 0: max commutes with the affine map
        TRACE_S(6);
        // ---- hand-over of the lazily updated row maximum from the other group's tile j-1 ----
        float m_prev = -INFINITY;
        if (j > j_first) {
          named_bar_sync(id_in, 64);
          m_prev = m_sh[row];
        }
        const bool need = mt > m_prev + RESCALE_THRESHOLD;
        float m_new = m_prev;
        if (__any_sync(0xffffffffu, need)) {
          if (j > j_first) {
            mbar_wait(&sp_free[(j - 1) % NSB], ((j - 1) / NSB) & 1, nullptr, 11);   // P.V(<= j-1) retired
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
        if (g + 1 < seg[s].hi) {
          m_sh[row] = m_new;
          named_bar_arrive(id_out, 64);
        }
        TRACE_S(7);
        if (m_new != m_ref) {                                // bring this group's sums to the current reference
          const float f2 = exp2f(m_ref - m_new);
          l_tot *= f2;
          l_piece *= f2;
          m_ref = m_new;
        }
        const float c0 = bias2 - m_new;
        uint32_t pk[32];
        float ls0 = 0.f, ls1 = 0.f, ls2 = 0.f, ls3 = 0.f;
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
          const float e0 = exp2f(fmaf(sc[c], p.scale_log2, c0));
          const float e1 = exp2f(fmaf(sc[c + 1], p.scale_log2, c0));
          const float e2 = exp2f(fmaf(sc[c + 2], p.scale_log2, c0));
          const float a3 = fmaf(sc[c + 3], p.scale_log2, c0);
          const float e3 = kPolyExp ? exp2_poly(a3) : exp2f(a3);
          ls0 += e0; ls1 += e1; ls2 += e2; ls3 += e3;
          pk[c >> 1] = pack2_fast(e0, e1);
          pk[(c >> 1) + 1] = pack2_fast(e2, e3);
        }
        const float lsum = (ls0 + ls1) + (ls2 + ls3);
        l_tot += lsum;
        l_piece += lsum;
        TRACE_S(8);
        // P(j) over the first 32 columns of S(j): this thread's row was fully read above
        tmem_st32u(lane_addr + TMEM_S + b * BN, pk);
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[b]);
        TRACE_S(9);
      }
      flush_piece(cur_t);

      // ---- segment epilogue: normalised fp16 partial O + (m, l) ----
      if (quad == 0 && grp == 0 && trace && lane == 0) trace[(long long)(j - 1) * 16 + 10] = clock64();
      lx[(grp * BM + row) * 2 + 0] = m_ref;
      lx[(grp * BM + row) * 2 + 1] = l_tot;
      named_bar_sync(id_ex, 64);
      if (quad == 0 && grp == 0 && trace && lane == 0) trace[(long long)(j - 1) * 16 + 12] = clock64();
      const float m_o = lx[((grp ^ 1) * BM + row) * 2 + 0], l_o = lx[((grp ^ 1) * BM + row) * 2 + 1];
      const float M = fmaxf(m_ref, m_o);                    // == the maximum O is relative to (m is monotone)
      const float l_row = l_tot * exp2f(m_ref - M) + l_o * exp2f(m_o - M);
      const float inv = 1.f / l_row;
      const int last = j - 1;
      mbar_wait(&sp_free[last % NSB], (last / NSB) & 1, nullptr, 12);
      fence_after();
      if (quad == 0 && grp == 0 && trace && lane == 0) trace[(long long)(j - 1) * 16 + 13] = clock64();
      if (p.direct) {
        // Whole unit in this CTA: out = O / l * gate, written here.  Every MMA has retired and every TMA tile has been
        // consumed, so the K/V rings are free: each warp stages its 32 rows x 128 columns there (fp32) and writes two
        // rows per instruction (16 lanes x 16 bytes per row = whole lines), the gate read the same way.
        constexpr int ROWP = (DVC / 2) * 4 + 16;
        static_assert(kSoftmaxWarps * 32 * ROWP <= KS * SMEM_K + VS * SMEM_V, "staging fits in the rings");
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
        const int ecol = dvc * DVC + grp * (DVC / 2) + (lane & 15) * 8;
#pragma unroll 4
        for (int it = 0; it < 16; ++it) {
          const int rr = it * 2 + (lane >> 4);
          const int oi = qt * BM + quad * 32 + rr;
          if (oi < p.HW) {
            const float4* sp = reinterpret_cast<const float4*>(stg + rr * ROWP + (lane & 15) * 32);
            const float4 a = sp[0], b = sp[1];
            float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            if (p.gate) {
              const uint4 g = *reinterpret_cast<const uint4*>(p.gate + (long long)oi * p.ldg + ecol);
              const float2 g0 = unpack2(g.x), g1 = unpack2(g.y), g2 = unpack2(g.z), g3 = unpack2(g.w);
              v[0] *= g0.x; v[1] *= g0.y; v[2] *= g1.x; v[3] *= g1.y; v[4] *= g2.x; v[5] *= g2.y; v[6] *= g3.x; v[7] *= g3.y;
            }
            uint4 u;
            u.x = pack2(v[0], v[1]); u.y = pack2(v[2], v[3]); u.z = pack2(v[4], v[5]); u.w = pack2(v[6], v[7]);
            *reinterpret_cast<uint4*>(p.out + (long long)oi * p.ldo + ecol) = u;
          }
        }
      } else {
      // part_o: [slot][16-column group][row][16] -- the 32 rows of a warp are contiguous per group, so every store
      // instruction covers whole lines (row-major rows 512 B apart cost one line per lane and ~5000 cycles per segment)
      t16* po = p.part_o + (((long long)(blockIdx.x * 2 + s) * (DVC / 16) + grp * (DVC / 32)) * BM + row) * 16;
#pragma unroll 1
      for (int c = 0; c < DVC / 2; c += 32) {
        float o[32];
        tmem_ld32(lane_addr + TMEM_O + grp * (DVC / 2) + c, o);
        if (row_ok) {
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            uint4 u;
            u.x = pack2(o[e] * inv, o[e + 1] * inv);
            u.y = pack2(o[e + 2] * inv, o[e + 3] * inv);
            u.z = pack2(o[e + 4] * inv, o[e + 5] * inv);
            u.w = pack2(o[e + 6] * inv, o[e + 7] * inv);
            *reinterpret_cast<uint4*>(po + (long long)((c + e) >> 4) * (BM * 16) + ((c + e) & 8)) = u;
          }
        }
      }
      }
      if (grp == 0 && !p.direct) {
        float* ml = p.part_ml + ((long long)(blockIdx.x * 2 + s) * BM + row) * 2;
        ml[0] = M;
        ml[1] = l_row;
      }
      fence_before();
      __syncwarp();
      if (quad == 0 && grp == 0 && trace && lane == 0) trace[(long long)(j - 1) * 16 + 11] = clock64();
      if (lane == 0 && s + 1 < nseg) mbar_arrive(o_drained);
    }
  }
  fence_before();
  __syncthreads();
  if (cta_times && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    cta_times[1] = (long long)gt; cta_times[2] = ntot; cta_times[3] = nseg; cta_times[6] = clock64();
  }
  if (warp == kWarpMmaS) {
    fence_after();
    tmem_dealloc<TMEM_COLS>(tmem);
  }
}

// Merge the segments of every unit: out = (sum_s w_s O_s) * gate with w_s = l_s 2^(m_s - M) / L;
// mass[i,t] = sum_{pieces of frame t} l_p 2^(m_p - M) / L  (from the Dv-chunk-0 units).
// Four query rows per block, 64 threads x 16 columns per row: part_o is stored as [slot][16-column group][row][16], so
// a thread reads one full 32-byte sector per segment (and the attention epilogue's row-per-lane stores coalesce).
constexpr int kMaxSegsPerUnit = 24;
constexpr int kCombRows = 4;
constexpr int kColGroups = DVC / 16;
static_assert(BM % kCombRows == 0, "the rows of a combine block share a query tile");
__global__ void __launch_bounds__(256) combine2_kernel(const Tc2Params p, const t16* __restrict__ gate, long long ldg,
                                                       t16* __restrict__ out, long long ldo,
                                                       float* __restrict__ mass) {
  pdl_prologue();
  __shared__ int s_n[kCombRows][4];
  __shared__ int s_slot[kCombRows][4][kMaxSegsPerUnit];        // (cta * 2 + seg)
  __shared__ float s_w[kCombRows][4][kMaxSegsPerUnit];         // l_s 2^(m_s - M) / L
  __shared__ float s_M0[kCombRows], s_L0[kCombRows];
  __shared__ int s_alo[kCombRows][kMaxSegsPerUnit], s_ahi[kCombRows][kMaxSegsPerUnit];   // unit-0 tile ranges
  const int rr = threadIdx.x >> 6, tc = threadIdx.x & 63;
  const int i = blockIdx.x * kCombRows + rr;
  const bool live = i < p.HW;
  const int qt = (blockIdx.x * kCombRows) / BM, r = i - qt * BM;
  if (live && tc < p.n_dv) {
    const int k = tc;
    const int unit = qt * p.n_dv + k;
    const long long u_lo = (long long)unit * p.TPU, u_hi = u_lo + p.TPU;
    int c = 0;                                       // first CTA whose range reaches past u_lo (bounds are ascending)
    for (int step = 128; step > 0; step >>= 1)
      if (c + step < p.nCTA && p.bounds[c + step] <= u_lo) c += step;
    if (p.bounds[c + 1] <= u_lo) ++c;
    int n = 0;
    float M = -INFINITY;
    for (; c < p.nCTA && n < kMaxSegsPerUnit; ++c) {
      long long lo, hi;
      cta_range(p, c, lo, hi);
      if (lo >= u_hi) break;
      if (hi <= lo) continue;
      const int slot = c * 2 + (lo < u_lo ? 1 : 0);
      s_slot[rr][k][n] = slot;
      if (k == 0) {
        s_alo[rr][n] = (int)((lo > u_lo ? lo : u_lo) - u_lo);
        s_ahi[rr][n] = (int)((hi < u_hi ? hi : u_hi) - u_lo);
      }
      M = fmaxf(M, p.part_ml[((long long)slot * BM + r) * 2]);
      ++n;
    }
    float L = 0.f;
    for (int e = 0; e < n; ++e) {
      const float* ml = p.part_ml + ((long long)s_slot[rr][k][e] * BM + r) * 2;
      const float w = exp2f(ml[0] - M) * ml[1];
      s_w[rr][k][e] = w;
      L += w;
    }
    const float inv = 1.f / L;
    for (int e = 0; e < n; ++e) s_w[rr][k][e] *= inv;
    s_n[rr][k] = n;
    if (k == 0) { s_M0[rr] = M; s_L0[rr] = L; }
  }
  __syncthreads();
  const int col = tc * 16;
  if (live && col < p.Dv) {
    const int k = col / DVC, cg = (col - k * DVC) >> 4;
    const int n = s_n[rr][k];
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.f;
    for (int e = 0; e < n; ++e) {
      const float w = s_w[rr][k][e];
      const uint4* src = reinterpret_cast<const uint4*>(
          p.part_o + (((long long)s_slot[rr][k][e] * kColGroups + cg) * BM + r) * 16);
      const uint4 u0 = src[0], u1 = src[1];
      const uint32_t uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float2 f = unpack2(uu[c]);
        acc[c * 2] = fmaf(w, f.x, acc[c * 2]);
        acc[c * 2 + 1] = fmaf(w, f.y, acc[c * 2 + 1]);
      }
    }
    if (gate) {
      const uint4* gp = reinterpret_cast<const uint4*>(gate + (long long)i * ldg + col);
      const uint4 g0 = gp[0], g1 = gp[1];
      const uint32_t gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float2 f = unpack2(gg[c]);
        acc[c * 2] *= f.x;
        acc[c * 2 + 1] *= f.y;
      }
    }
    uint4 o0, o1;
    o0.x = pack2(acc[0], acc[1]); o0.y = pack2(acc[2], acc[3]); o0.z = pack2(acc[4], acc[5]); o0.w = pack2(acc[6], acc[7]);
    o1.x = pack2(acc[8], acc[9]); o1.y = pack2(acc[10], acc[11]); o1.z = pack2(acc[12], acc[13]); o1.w = pack2(acc[14], acc[15]);
    uint4* op = reinterpret_cast<uint4*>(out + (long long)i * ldo + col);
    op[0] = o0;
    op[1] = o1;
  }
  if (mass && live && tc < p.T) {
    const int t = tc;
    const int f_lo = t * p.tpf, f_hi = f_lo + p.tpf;
    const float M = s_M0[rr], invL = 1.f / s_L0[rr];
    float a = 0.f;
    for (int e = 0; e < s_n[rr][0]; ++e) {
      if (s_alo[rr][e] < f_hi && f_lo < s_ahi[rr][e]) {
        const float* pc = p.pieces + ((((long long)s_slot[rr][0][e] * p.T + t) * 2) * BM + r) * 2;
        a += exp2f(pc[0] - M) * pc[1] + exp2f(pc[BM * 2] - M) * pc[BM * 2 + 1];
      }
    }
    mass[(long long)i * p.T + t] = a * invL;
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

// RMEM_ATTN_DIRECT=1 enables the one-CTA-per-unit path for single-frame launches.  Off by default: measured at c3 it
// is slower (56 CTAs x 27 tiles leave 92 SMs idle: 1.496 vs 1.463 ms per frame) than stream-K + combine.
bool direct_switch() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("RMEM_ATTN_DIRECT"); v = (e && e[0] == '1') ? 1 : 0; }
  return v != 0;
}

// Static schedule.  Uniform cuts give every CTA the same number of tiles, but a CTA whose range crosses a unit boundary
// pays a second segment epilogue (TMEM read-out of O is 64 B/clk: ~6 tile-times) plus a pipeline restart, and the kernel
// ends with the slowest CTA (per-CTA wall times, profiles/r01_attn_trace_cta95.txt: 99k cycles with one segment, 112k with
// two).  make_bounds() equalises tiles + kSegCost * segments (+ kRestartCost for a second segment) instead: smallest
// budget for which a greedy walk covers all L steps with nCTA CTAs, at most two segments per CTA.
constexpr int kSegCost = 6, kRestartCost = 1;
bool balance_switch() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("RMEM_ATTN_BALANCE"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}
// CTAs [0, m) get `budget`, the rest budget - 1
bool greedy_bounds(long long L, int TPU, int n, long long budget, int m, int* b) {
  long long pos = 0;
  for (int c = 0; c < n; ++c) {
    b[c] = (int)pos;
    long long room = budget - (c >= m ? 1 : 0) - kSegCost;
    bool second = false;
    while (room > 0 && pos < L) {
      const long long unit_end = (pos / TPU + 1) * TPU;
      const long long take = room < unit_end - pos ? room : unit_end - pos;
      pos += take;
      room -= take;
      if (pos == unit_end && pos < L) {
        if (second) break;
        second = true;
        room -= kSegCost + kRestartCost;
      }
    }
  }
  b[n] = (int)L;
  return pos >= L;
}
void make_bounds(long long L, int TPU, int n, int* b) {
  const long long uni = (L + n - 1) / n;
  if (!balance_switch() || uni < 16) {                 // short launches keep the uniform cut (see schedule())
    for (int c = 0; c <= n; ++c) b[c] = (int)((L * c) / n);
    return;
  }
  long long lo = uni, hi = uni + 2 * kSegCost + kRestartCost + 2;
  while (lo < hi) {
    const long long mid = (lo + hi) / 2;
    if (greedy_bounds(L, TPU, n, mid, n, b)) hi = mid; else lo = mid + 1;
  }
  // with the minimal budget the walk usually ends a few CTAs early (idle SMs): give only the first m CTAs the full
  // budget and the others one tile less, m minimal
  int mlo = 0, mhi = n;
  while (mlo < mhi) {
    const int mid = (mlo + mhi) / 2;
    if (greedy_bounds(L, TPU, n, lo, mid, b)) mhi = mid; else mlo = mid + 1;
  }
  greedy_bounds(L, TPU, n, lo, mlo, b);
}

bool short_switch() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("RMEM_ATTN_SHORT"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

void schedule(int HW, int T, int Dv, int* n_units, int* tpf, int* TPU, int* nCTA) {
  const int qtiles = cdiv(HW, BM), n_dv = Dv / DVC;
  *n_units = qtiles * n_dv;
  *tpf = cdiv(HW, BN);
  *TPU = T * *tpf;
  const long long L = (long long)*n_units * *TPU;
  int n = sm_count();
  if (n > kMaxCTA) n = kMaxCTA;
  if ((long long)n > L / 4) n = (int)(L / 4);                    // at least ~4 tiles per CTA
  const int cap = (kMaxSegsPerUnit - 2) * *n_units;             // combine2 resolves <= kMaxSegsPerUnit segments per unit
  if (n > cap) n = cap;
  // Short launches (the T = 1 self-attention: ~10 tiles per SM): a segment epilogue costs about six tiles, so a CTA that
  // straddles two units pays more in epilogues than the last SMs are worth -- use k whole-segment CTAs per unit instead.
  if (L / n < 16 && n / *n_units >= 2 && short_switch()) n = (n / *n_units) * *n_units;
  if (n < *n_units) n = *n_units;                               // a CTA never spans more than two units
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
  o += (size_t)nCTA * 2 * T * 2 * BM * 2 * sizeof(float);
  return o + 256;
}

}  // namespace

// Optional CUDA events recorded right before / after the main kernel (bench.py times the dominant kernel alone with them).
static thread_local cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
void long_attn_tc2_set_events(void* ev0, void* ev1) {
  g_ev0 = reinterpret_cast<cudaEvent_t>(ev0);
  g_ev1 = reinterpret_cast<cudaEvent_t>(ev1);
}

int long_attn_tc2_set_trace(long long* dev_buf) {
  static int cta = -1;
  if (cta < 0) {
    const char* e = getenv("RMEM_TRACE_CTA");
    cta = e ? atoi(e) : 0;
    RMEM_CUDA_CHECK(cudaMemcpyToSymbol(g_trace_cta, &cta, sizeof(cta)));
  }
  RMEM_CUDA_CHECK(cudaMemcpyToSymbol(g_trace, &dev_buf, sizeof(dev_buf)));
  return RMEM_OK;
}

// Host-only view of the static schedule (tests): step range of every CTA for a launch of this shape.
int long_attn_tc2_schedule(int HW, int T, int Dv, int* n_units, int* tiles_per_unit, int* n_cta, int* bounds, int cap) {
  int tpf = 0;
  schedule(HW, T, Dv, n_units, &tpf, tiles_per_unit, n_cta);
  RMEM_REQUIRE(*n_cta + 1 <= cap && *n_cta <= kMaxCTA, "schedule: %d CTAs do not fit the caller's table (%d)", *n_cta, cap);
  make_bounds((long long)*n_units * *tiles_per_unit, *tiles_per_unit, *n_cta, bounds);
  return RMEM_OK;
}

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
  // One frame of keys and no attention mass wanted (the self-attention): a unit is only tpf tiles long, so cutting it into
  // stream-K segments costs more in segment epilogues + the combine launch than it balances.  One CTA per unit, final
  // rows written by the kernel.
  const bool vec_ok = a.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 &&
                      (!a.gate || (a.ldg % 8 == 0 && (reinterpret_cast<uintptr_t>(a.gate) & 15) == 0));
  p.direct = (a.T == 1 && !a.mass && vec_ok && p.n_units <= sm_count() && direct_switch()) ? 1 : 0;
  if (p.direct) p.nCTA = p.n_units;
  p.gate = a.gate; p.ldg = a.ldg; p.out = a.out; p.ldo = a.ldo;
  p.L = (long long)p.n_units * p.TPU;
  RMEM_REQUIRE(p.nCTA <= kMaxCTA, "long_attn_tc2: %d CTAs > %d", p.nCTA, kMaxCTA);
  make_bounds(p.L, p.TPU, p.nCTA, p.bounds);
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

  RMEM_REQUIRE((reinterpret_cast<uintptr_t>(a.qt) & 15) == 0, "long_attn_tc2: q alignment");
  p.q = a.qt;
  const CUtensorMap *mk, *mv;
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
  if (g_ev0) RMEM_CUDA_CHECK(cudaEventRecord(g_ev0, s));
  RMEM_CUDA_CHECK(launch_pdl(long_attn_tc2_kernel, dim3(p.nCTA), dim3(kThreads), SMEM_TOTAL, s, *mk, *mv, p));
  if (g_ev1) RMEM_CUDA_CHECK(cudaEventRecord(g_ev1, s));
  RMEM_LAUNCH_CHECK();
  if (p.direct) return RMEM_OK;
  RMEM_CUDA_CHECK(launch_pdl(combine2_kernel, dim3(cdiv(a.HW, kCombRows)), dim3(256), 0, s, p, a.gate, a.ldg, a.out, a.ldo, a.mass));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

}  // namespace rmem
