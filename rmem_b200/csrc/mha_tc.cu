// Multi-head attention over the restricted bank on tcgen05 for the AOT model (8 heads x 32; K1 / K2b / K3 of SURVEY.md:
// networks/layers/attention.py:28-81 MultiheadAttention, transformer.py:566-571 (self), :632-650 (long-term over the
// bank with the temporal PE as a per-head score bias), :656-662 (short-term, dense over the previous frame)).
// Replaces mha_dense (materialised [H, HW, T*HW] scores on head-batched mma.sync GEMMs) on the engine's hot path.
//
//   out[i, h*32:(h+1)*32] = softmax_j( scale <q_ih, k_jh> + qbias[h, i, t(j)] ) . v_jh      mass[i, t] = mean_h sum_{j in t} P_ijh
//
// Shape of the problem: head dim 32 makes the tensor work small (2 k-steps per score tile, N = 32 per P.V MMA) and the
// exponentials dominant -- 8 heads x 64 keys = 512 ex2 per query row and 64-key sub-tile against 128 for one DeAOT Dv
// chunk: at 16 ex2 / clk / SM the MUFU pipe needs 4096 cycles per (128-query, 64-key) step, the tensor pipe ~1900.  So
// the kernel is organised around keeping MUFU busy, not the tensor pipe:
//   * one CTA = one 128-query tile x a contiguous range of 64-key sub-tiles (stream-K over (query tile, sub-tile) steps,
//     <= 2 segments per CTA, merged by mha_combine_kernel); all 8 heads of a step are processed by the same CTA, so the
//     K tile [64 keys, 256] and the value-major V tile [256, 64 keys] are fetched once for all heads (TMA, 128B swizzle);
//     head h reads its 32 channels as a 64-byte k-offset inside the swizzle atom (scores) / as 32 rows of the V tile (P.V)
//   * a sub-tile is walked in four QUARTERS of two heads; two softmax groups (4 warps each, one warp per TMEM lane
//     quadrant, thread = query row) own the even / odd quarters and with them two 128-column score buffers, so the score
//     MMAs of quarter g+1 and the P.V MMAs of quarter g-1 run under the exponentials of quarter g.  A head always meets
//     the same thread: running maximum / sums live in registers, no hand-over between groups
//   * P is written over its own score columns as packed fp16 and is the TMEM A operand of O_h += P_h . V_h; O for all 8
//     heads is one 256-column accumulator; lazy rescale of a head's 32 columns when its maximum grows by > 2^10
//   TMEM: O[256] | S/P buffer 0 [2 heads x 64] | S/P buffer 1 [2 heads x 64]
//   smem: Q 64 KB | K ring 2 x 32 KB | V^T ring 2 x 32 KB
//   warps 0-7 softmax (+ segment epilogue), 8 TMA producer, 9 MMA issuer + TMEM owner
#include <cstdlib>

#include "attn.cuh"
#include "tcgen05.cuh"

namespace rmem {

namespace {

using namespace tc;

constexpr int BM = 128;        // query rows per CTA
constexpr int BN = 64;         // keys per sub-tile
constexpr int CH = 256;        // channels = heads x head dim
constexpr int NH = 8, DH = 32;
constexpr int KS = 2, VS = 2;  // ring depths (sub-tiles)
constexpr int kSoftmaxWarps = 8, kWarpTma = 8, kWarpMma = 9;
constexpr int kThreads = 10 * 32;

constexpr int SMEM_Q = BM * CH * 2;     // 64 KB: 4 atoms of [128 rows][128 B]
constexpr int SMEM_K = BN * CH * 2;     // 32 KB: 4 atoms of [64 keys][128 B]
constexpr int SMEM_V = CH * BN * 2;     // 32 KB: [256 value rows][64 keys = 128 B]
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + SMEM_Q;
constexpr int OFF_V = OFF_K + KS * SMEM_K;
constexpr int OFF_BAR = OFF_V + VS * SMEM_V;
constexpr int SMEM_TOTAL = OFF_BAR + 256 + 1024;   // barriers + alignment slack
static_assert(SMEM_TOTAL <= 232448, "shared memory budget");

constexpr int TMEM_COLS = 512;
constexpr int TMEM_O = 0;      // 8 heads x 32 fp32 columns
constexpr int TMEM_S = 256;    // buffer b at TMEM_S + b*128, head hh of the quarter at + hh*64; P aliases the first 32

constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 10.0f;     // log2 units: P <= 2^10 before a lazy rescale is forced

constexpr int kMaxCTA = 160;
struct MhaTcParams {
  int HW, HWp, T, tpf, TPU, n_qt, nCTA;        // tpf = sub-tiles per frame, TPU = T * tpf = steps per query tile
  int bounds[kMaxCTA + 1];                     // CTA c owns steps [bounds[c], bounds[c+1]) of the (query tile, sub-tile) sequence
  int slot[kMaxBankFrames];
  float scale_log2;                            // scale * log2(e)
  const float* qbias;                          // [NH][HW][T] (already multiplied by scale) or null
  t16* part_o;                                 // [nCTA][2][BM][CH]      normalised partial O
  float* part_ml;                              // [nCTA][2][BM][NH][2]   (m in log2 units, l)
  float* pieces;                               // [nCTA][2][T][NH][BM][2] per-frame (m, l) or null
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

__device__ __forceinline__ uint32_t pack2_fast(float lo, float hi) {
#ifdef RMEM_OPERAND_BF16
  t162 v = __floats2bfloat162_rn(lo, hi);
#else
  t162 v = __floats2half2_rn(lo, hi);
#endif
  return *reinterpret_cast<uint32_t*>(&v);
}

struct Seg { int unit, lo, hi; };   // sub-tiles [lo, hi) of query tile `unit`

__global__ void __launch_bounds__(kThreads, 1)
mha_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
              const __grid_constant__ CUtensorMap map_v, const MhaTcParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;               // one phase per segment
  uint64_t* q_free = q_full + 1;         // every score MMA of the first segment has read Q
  uint64_t* k_full = q_free + 1;         // [KS]
  uint64_t* k_empty = k_full + KS;       // [KS]
  uint64_t* v_full = k_empty + KS;       // [VS]
  uint64_t* v_empty = v_full + VS;       // [VS]
  uint64_t* s_full = v_empty + VS;       // [2] scores of a quarter are in buffer b
  uint64_t* p_full = s_full + 2;         // [2] P of a quarter stored by the four warps of its group
  uint64_t* seg_done = p_full + 2;       // [2] every P.V MMA of segment s has completed
  uint64_t* o_drained = seg_done + 2;    // all softmax warps have read the first segment's O
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_drained + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x;

  const long long lo = p.bounds[cta], hi = p.bounds[cta + 1];
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

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_free, 1);
    for (int i = 0; i < KS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < VS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&seg_done[i], 1); }
    mbar_init(o_drained, kSoftmaxWarps);
    mbar_fence_init();
  }
  if (warp == kWarpMma) tmem_alloc<TMEM_COLS>(tmem_slot);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_prologue();

  if (warp == kWarpTma) {
    // ================================ TMA producer: Q per segment, K and V^T per sub-tile ================================
    if (ntot > 0) {
      if (elect_one()) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_v);
      }
      __syncwarp();
      int i = 0;
      for (int s = 0; s < nseg; ++s) {
        if (s == 1) mbar_wait(q_free, 0, nullptr, 0);
        if (elect_one()) {
          mbar_expect_tx(q_full, SMEM_Q);
          const int row0 = seg[s].unit * BM;
#pragma unroll
          for (int a = 0; a < 4; ++a) tma_load_2d(smem + OFF_Q + a * (BM * 128), &map_q, q_full, a * 64, row0);
        }
        __syncwarp();
        int t = seg[s].lo / p.tpf, jt = seg[s].lo - t * p.tpf;
        for (int g = seg[s].lo; g < seg[s].hi; ++g, ++i, ++jt) {
          if (jt == p.tpf) { jt = 0; ++t; }
          const int key0 = p.slot[t] * p.HWp + jt * BN;
          const int sk = i % KS, sv = i % VS;
          if (i >= KS) mbar_wait(&k_empty[sk], ((i / KS) - 1) & 1, nullptr, 0);
          if (elect_one()) {
            unsigned char* dk = smem + OFF_K + sk * SMEM_K;
            mbar_expect_tx(&k_full[sk], SMEM_K);
#pragma unroll
            for (int a = 0; a < 4; ++a) tma_load_2d(dk + a * (BN * 128), &map_k, &k_full[sk], a * 64, key0);
          }
          __syncwarp();
          if (i >= VS) mbar_wait(&v_empty[sv], ((i / VS) - 1) & 1, nullptr, 0);
          if (elect_one()) {
            mbar_expect_tx(&v_full[sv], SMEM_V);
            tma_load_2d(smem + OFF_V + sv * SMEM_V, &map_v, &v_full[sv], key0, 0);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ================================ MMA issuer ================================
    // Global quarter index g = sub-tile * 4 + quarter; buffer = g & 1.  Iteration g issues P.V(g-2) (its P sits in the
    // buffer S(g) is about to overwrite; MMAs execute in issue order) and then S(g).
    if (ntot > 0) {
      constexpr uint32_t idesc_s = make_idesc(BM, BN);
      constexpr uint32_t idesc_o = make_idesc(BM, DH);
      const uint32_t smem_base = smem_u32(smem);
      const int G = ntot * 4;
      for (int g = 0; g < G + 2; ++g) {
        if (g >= 2) {
          const int gp = g - 2, i = gp >> 2, q = gp & 3, b = gp & 1;
          if (q == 0) {
            mbar_wait(&v_full[i % VS], (i / VS) & 1, nullptr, 0);
            if (i == n0 && nseg > 1) mbar_wait(o_drained, 0, nullptr, 0);
          }
          mbar_wait(&p_full[b], (gp >> 1) & 1, nullptr, 0);
          fence_after();
          if (elect_one()) {
            const bool first = (i == 0) || (i == n0);
            const uint32_t vbase = smem_base + OFF_V + (i % VS) * SMEM_V;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int h = q * 2 + hh;
              const uint64_t dv = make_desc_sw128(vbase + h * (DH * 128));
              const uint32_t pa = tmem + TMEM_S + b * 128 + hh * 64;
#pragma unroll
              for (int kk = 0; kk < BN / 16; ++kk)
                umma_ts(tmem + TMEM_O + h * DH, pa + kk * 8, dv + (uint64_t)(kk * 2), idesc_o, (first && kk == 0) ? 0u : 1u);
            }
            if (q == 3) {
              commit(&v_empty[i % VS]);
              if (i == n0 - 1) commit(&seg_done[0]);
              else if (i == ntot - 1) commit(&seg_done[1]);
            }
          }
          __syncwarp();
        }
        if (g < G) {
          const int i = g >> 2, q = g & 3, b = g & 1;
          if (q == 0) {
            if (i == 0) mbar_wait(q_full, 0, nullptr, 0);
            if (i == n0 && nseg > 1) mbar_wait(q_full, 1, nullptr, 0);
            mbar_wait(&k_full[i % KS], (i / KS) & 1, nullptr, 0);
          }
          fence_after();
          if (elect_one()) {
            const uint32_t kbase = smem_base + OFF_K + (i % KS) * SMEM_K;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int h = q * 2 + hh;
              const uint64_t dq = make_desc_sw128(smem_base + OFF_Q + (h >> 1) * (BM * 128));
              const uint64_t dk = make_desc_sw128(kbase + (h >> 1) * (BN * 128));
              const uint32_t d = tmem + TMEM_S + b * 128 + hh * 64;
#pragma unroll
              for (int kk = 0; kk < DH / 16; ++kk) {
                const uint64_t off = (uint64_t)(((h & 1) * 64 + kk * 32) >> 4);   // head's 64 B inside the 128 B atom
                umma_ss(d, dq + off, dk + off, idesc_s, kk > 0);
              }
            }
            commit(&s_full[b]);
            if (q == 3) {
              commit(&k_empty[i % KS]);
              if (i == n0 - 1 && nseg > 1) commit(q_free);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (ntot > 0) {
    // ================================ softmax + segment epilogue (warps 0-7) ================================
    const int quad = warp & 3, grp = warp >> 2;
    const int row = quad * 32 + lane;                       // tile row == TMEM lane
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    int i = 0;                                              // sub-tile counter over both segments
    for (int s = 0; s < nseg; ++s) {
      const int qi = seg[s].unit * BM + row;
      const bool row_ok = qi < p.HW;
      // this thread's four heads: k = qq*2 + hh  <->  head 2*(grp + 2*qq) + hh
      float m[4], l[4], lp[4], bias2[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { m[k] = -INFINITY; l[k] = 0.f; lp[k] = 0.f; bias2[k] = 0.f; }
      int cur_t = -1;
      const int i_first = i;
      auto flush_pieces = [&](int t) {
        if (p.pieces) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int h = 2 * (grp + 2 * (k >> 1)) + (k & 1);
            float* d = p.pieces + (((((long long)(cta * 2 + s) * p.T + t) * NH + h) * BM) + row) * 2;
            d[0] = m[k];
            d[1] = lp[k];
          }
        }
      };
      int t = seg[s].lo / p.tpf, jt = seg[s].lo - t * p.tpf;
      for (int gs = seg[s].lo; gs < seg[s].hi; ++gs, ++jt, ++i) {
        if (jt == p.tpf) { jt = 0; ++t; }
        if (t != cur_t) {
          if (cur_t >= 0) flush_pieces(cur_t);
          cur_t = t;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            lp[k] = 0.f;
            const int h = 2 * (grp + 2 * (k >> 1)) + (k & 1);
            bias2[k] = (p.qbias && row_ok) ? p.qbias[((long long)h * p.HW + qi) * p.T + t] * LOG2E : 0.f;
          }
        }
        const int key0 = jt * BN;
        const bool ragged = key0 + BN > p.HW;
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
          const int q = grp + 2 * qq;
          const int g = i * 4 + q;
          mbar_wait(&s_full[grp], (g >> 1) & 1, nullptr, 0);
          fence_after();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int k = qq * 2 + hh;
            const int h = q * 2 + hh;
            const uint32_t sbase = lane_addr + TMEM_S + grp * 128 + hh * 64;
            float sc[64];
            {
              uint32_t r0[32], r1[32];
              tmem_ld32_nowait(sbase, r0);
              tmem_ld32_nowait(sbase + 32, r1);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 32; ++c) { sc[c] = __uint_as_float(r0[c]); sc[32 + c] = __uint_as_float(r1[c]); }
            }
            if (ragged) {
#pragma unroll
              for (int c = 0; c < 64; ++c) sc[c] = (key0 + c < p.HW) ? sc[c] : -INFINITY;
            }
            float mx;
            {
              float a[16];
#pragma unroll
              for (int c = 0; c < 16; ++c) a[c] = fmaxf(fmaxf(sc[c], sc[16 + c]), fmaxf(sc[32 + c], sc[48 + c]));
#pragma unroll
              for (int c = 0; c < 4; ++c) a[c] = fmaxf(fmaxf(a[c], a[4 + c]), fmaxf(a[8 + c], a[12 + c]));
              mx = fmaxf(fmaxf(a[0], a[1]), fmaxf(a[2], a[3]));
            }
            const float mt = fmaf(mx, p.scale_log2, bias2[k]);
            const bool need = mt > m[k] + RESCALE_THRESHOLD;
            if (__any_sync(0xffffffffu, need)) {
              if (i > i_first) {
                // every P.V issued before S(g) has completed (in-order execution, s_full is a commit after S(g)); that
                // includes this head's previous one, so its 32 accumulator columns can be rescaled in place
                const float f = need ? exp2f(m[k] - mt) : 1.f;
                float o[32];
                tmem_ld32(lane_addr + TMEM_O + h * DH, o);
#pragma unroll
                for (int e = 0; e < 32; ++e) o[e] *= f;
                tmem_st32(lane_addr + TMEM_O + h * DH, o);
              }
              if (need) {
                const float f2 = exp2f(m[k] - mt);         // m = -inf on the first tile: f2 = 0, sums are 0 anyway
                l[k] *= f2;
                lp[k] *= f2;
                m[k] = mt;
              }
            }
            const float c0 = bias2[k] - m[k];
            uint32_t pk[32];
            float ls0 = 0.f, ls1 = 0.f, ls2 = 0.f, ls3 = 0.f;
#pragma unroll
            for (int c = 0; c < 64; c += 4) {
              const float e0 = exp2f(fmaf(sc[c], p.scale_log2, c0));
              const float e1 = exp2f(fmaf(sc[c + 1], p.scale_log2, c0));
              const float e2 = exp2f(fmaf(sc[c + 2], p.scale_log2, c0));
              const float e3 = exp2f(fmaf(sc[c + 3], p.scale_log2, c0));
              ls0 += e0; ls1 += e1; ls2 += e2; ls3 += e3;
              pk[c >> 1] = pack2_fast(e0, e1);
              pk[(c >> 1) + 1] = pack2_fast(e2, e3);
            }
            const float lsum = (ls0 + ls1) + (ls2 + ls3);
            l[k] += lsum;
            lp[k] += lsum;
            tmem_st32u(sbase, pk);
          }
          fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[grp]);
        }
      }
      flush_pieces(cur_t);

      // ---- segment epilogue: normalised fp16 partial O + (m, l) per head ----
      mbar_wait(&seg_done[s], 0, nullptr, 0);
      fence_after();
      t16* po = p.part_o + ((long long)(cta * 2 + s) * BM + row) * CH;
      float* ml = p.part_ml + ((long long)(cta * 2 + s) * BM + row) * (NH * 2);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int h = 2 * (grp + 2 * (k >> 1)) + (k & 1);
        const float inv = l[k] > 0.f ? 1.f / l[k] : 0.f;
        float o[32];
        tmem_ld32(lane_addr + TMEM_O + h * DH, o);
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          uint4 u;
          u.x = pack2(o[e] * inv, o[e + 1] * inv);
          u.y = pack2(o[e + 2] * inv, o[e + 3] * inv);
          u.z = pack2(o[e + 4] * inv, o[e + 5] * inv);
          u.w = pack2(o[e + 6] * inv, o[e + 7] * inv);
          *reinterpret_cast<uint4*>(po + h * DH + e) = u;
        }
        ml[h * 2] = m[k];
        ml[h * 2 + 1] = l[k];
      }
      fence_before();
      __syncwarp();
      if (lane == 0 && s + 1 < nseg) mbar_arrive(o_drained);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    fence_after();
    tmem_dealloc<TMEM_COLS>(tmem);
  }
}

// Merge the segments of every query tile: out[i, c] = sum_s w_sh O_s[i, c], w_sh = l_sh 2^(m_sh - M_h) / L_h (h = c / 32);
// mass[i, t] = mean_h sum_{pieces of frame t} l_p 2^(m_p - M_h) / L_h.   One block per query row, thread = channel.
constexpr int kMaxSegs = 40;
__global__ void __launch_bounds__(CH) mha_combine_kernel(const MhaTcParams p, t16* __restrict__ out, long long ldo,
                                                         float* __restrict__ mass) {
  pdl_prologue();
  __shared__ int s_n;
  __shared__ int s_slot[kMaxSegs], s_alo[kMaxSegs], s_ahi[kMaxSegs];
  __shared__ float s_mass[NH][kMaxBankFrames];
  const int i = blockIdx.x;
  const int qt = i / BM, r = i - qt * BM;
  if (threadIdx.x == 0) {
    const long long u_lo = (long long)qt * p.TPU, u_hi = u_lo + p.TPU;
    int c = 0;
    for (int step = 128; step > 0; step >>= 1)
      if (c + step < p.nCTA && p.bounds[c + step] <= u_lo) c += step;
    if (p.bounds[c + 1] <= u_lo) ++c;
    int n = 0;
    for (; c < p.nCTA && n < kMaxSegs; ++c) {
      const long long lo = p.bounds[c], hi = p.bounds[c + 1];
      if (lo >= u_hi) break;
      if (hi <= lo) continue;
      s_slot[n] = c * 2 + (lo < u_lo ? 1 : 0);
      s_alo[n] = (int)((lo > u_lo ? lo : u_lo) - u_lo);
      s_ahi[n] = (int)((hi < u_hi ? hi : u_hi) - u_lo);
      ++n;
    }
    s_n = n;
  }
  __syncthreads();
  const int n = s_n;
  const int col = threadIdx.x, h = col >> 5;
  float M = -INFINITY;
  for (int e = 0; e < n; ++e) {
    const float* ml = p.part_ml + ((long long)s_slot[e] * BM + r) * (NH * 2) + h * 2;
    if (ml[1] > 0.f) M = fmaxf(M, ml[0]);
  }
  float L = 0.f, acc = 0.f;
  for (int e = 0; e < n; ++e) {
    const float* ml = p.part_ml + ((long long)s_slot[e] * BM + r) * (NH * 2) + h * 2;
    const float w = ml[1] > 0.f ? exp2f(ml[0] - M) * ml[1] : 0.f;
    L += w;
    acc = fmaf(w, t2f(p.part_o[((long long)s_slot[e] * BM + r) * CH + col]), acc);
  }
  const float invL = L > 0.f ? 1.f / L : 0.f;
  out[(long long)i * ldo + col] = f2t(acc * invL);
  if (mass) {
    // threads (h, t): lane t of the head's 32-thread group
    const int t = col & 31;
    if (t < p.T) {
      const int f_lo = t * p.tpf, f_hi = f_lo + p.tpf;
      float a = 0.f;
      for (int e = 0; e < n; ++e) {
        if (s_alo[e] < f_hi && f_lo < s_ahi[e]) {
          const float* pc = p.pieces + (((((long long)s_slot[e] * p.T + t) * NH + h) * BM) + r) * 2;
          if (pc[1] > 0.f) a += exp2f(pc[0] - M) * pc[1];
        }
      }
      s_mass[h][t] = a * invL;
    }
    __syncthreads();
    if (threadIdx.x < p.T) {
      float a = 0.f;
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) a += s_mass[hh][threadIdx.x];
      mass[(long long)i * p.T + threadIdx.x] = a * (1.f / NH);
    }
  }
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

size_t part_bytes(int nCTA, int T, size_t* off_ml, size_t* off_pieces) {
  size_t o = (size_t)nCTA * 2 * BM * CH * sizeof(t16);
  o = (o + 255) & ~size_t(255);
  *off_ml = o;
  o += (size_t)nCTA * 2 * BM * NH * 2 * sizeof(float);
  o = (o + 255) & ~size_t(255);
  *off_pieces = o;
  o += (size_t)nCTA * 2 * T * NH * BM * 2 * sizeof(float);
  return o + 256;
}

}  // namespace

size_t mha_tc_workspace(int HW, int HWp, int nslots, int H) {
  (void)HW; (void)HWp; (void)nslots; (void)H;
  size_t a, b;
  return part_bytes(kMaxCTA, kMaxBankFrames, &a, &b);
}

int mha_tc(const MhaArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  RMEM_REQUIRE(a.H == NH && a.dh == DH, "mha_tc: built for 8 heads x 32 (got %d x %d)", a.H, a.dh);
  RMEM_REQUIRE(a.T >= 1 && a.T <= kMaxBankFrames && a.T <= a.nslots, "mha_tc: T=%d nslots=%d", a.T, a.nslots);
  RMEM_REQUIRE(a.HWp % BN == 0 && a.HWp >= a.HW && a.HW >= 1, "mha_tc: HWp=%d must be a multiple of 64 >= HW=%d", a.HWp, a.HW);
  RMEM_REQUIRE(a.ldq % 8 == 0 && a.ldq >= CH && (reinterpret_cast<uintptr_t>(a.q) & 15) == 0, "mha_tc: q alignment");
  RMEM_REQUIRE((reinterpret_cast<uintptr_t>(a.kbank) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.vtbank) & 15) == 0,
               "mha_tc: bank alignment");
  RMEM_REQUIRE(a.ldo >= CH, "mha_tc: ldo");
  MhaTcParams p;
  p.HW = a.HW; p.HWp = a.HWp; p.T = a.T;
  p.tpf = cdiv(a.HW, BN); p.TPU = a.T * p.tpf; p.n_qt = cdiv(a.HW, BM);
  const long long steps = (long long)p.n_qt * p.TPU;
  int nCTA = sm_count() < kMaxCTA ? sm_count() : kMaxCTA;
  if (steps < nCTA) nCTA = (int)steps;
  if (nCTA > p.n_qt * 32) nCTA = p.n_qt * 32;          // <= 34 segments per query tile (mha_combine_kernel: kMaxSegs)
  RMEM_REQUIRE(p.n_qt <= nCTA, "mha_tc: %d query tiles need at least as many CTAs (%d)", p.n_qt, nCTA);
  p.nCTA = nCTA;
  for (int c = 0; c <= nCTA; ++c) p.bounds[c] = (int)(steps * c / nCTA);   // range length <= TPU: at most two segments
  for (int t = 0; t < kMaxBankFrames; ++t) {
    p.slot[t] = t < a.T ? a.slot[t] : 0;
    if (t < a.T) RMEM_REQUIRE(a.slot[t] >= 0 && a.slot[t] < a.nslots, "mha_tc: bad slot");
  }
  p.scale_log2 = a.scale * LOG2E;
  p.qbias = a.qbias;
  size_t off_ml, off_pieces;
  const size_t need = part_bytes(nCTA, a.T, &off_ml, &off_pieces);
  RMEM_REQUIRE(workspace_bytes >= need, "mha_tc: workspace %zu < %zu", workspace_bytes, need);
  RMEM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "mha_tc: workspace alignment");
  char* ws = reinterpret_cast<char*>(workspace);
  p.part_o = reinterpret_cast<t16*>(ws);
  p.part_ml = reinterpret_cast<float*>(ws + off_ml);
  p.pieces = a.mass ? reinterpret_cast<float*>(ws + off_pieces) : nullptr;

  const CUtensorMap *mq, *mk, *mv;
  {
    uint64_t dims[2] = {(uint64_t)CH, (uint64_t)a.HW};
    uint64_t str[1] = {(uint64_t)a.ldq * 2};
    uint32_t box[2] = {64, (uint32_t)BM};
    RMEM_TRY(tma_encode_cached(&mq, a.q, 2, dims, str, box, nullptr));
  }
  {
    uint64_t dims[2] = {(uint64_t)CH, (uint64_t)a.nslots * a.HWp};
    uint64_t str[1] = {(uint64_t)CH * 2};
    uint32_t box[2] = {64, (uint32_t)BN};
    RMEM_TRY(tma_encode_cached(&mk, a.kbank, 2, dims, str, box, nullptr));
  }
  {
    uint64_t dims[2] = {(uint64_t)a.nslots * a.HWp, (uint64_t)CH};
    uint64_t str[1] = {(uint64_t)a.nslots * a.HWp * 2};
    uint32_t box[2] = {(uint32_t)BN, (uint32_t)CH};
    RMEM_TRY(tma_encode_cached(&mv, a.vtbank, 2, dims, str, box, nullptr));
  }
  static bool attr_done = false;
  if (!attr_done) {
    RMEM_CUDA_CHECK(cudaFuncSetAttribute(mha_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    attr_done = true;
  }
  RMEM_CUDA_CHECK(launch_pdl(mha_tc_kernel, dim3(nCTA), dim3(kThreads), SMEM_TOTAL, s, *mq, *mk, *mv, p));
  RMEM_LAUNCH_CHECK();
  RMEM_CUDA_CHECK(launch_pdl(mha_combine_kernel, dim3(a.HW), dim3(CH), 0, s, p, a.out, a.ldo, a.mass));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

}  // namespace rmem
