// Long-term / self attention for sm_100a, CTA-pair version  (K1 + K1b + K8 of SURVEY.md; attention.py:174-193,
// transformer.py:1140-1197).  RMEM_ATTN_TC3.
//
// What changed against attn_tc2.cu, and why (profiles/r01_*):
//   * tc2 was bound by the ~48 B/clk ONE SM can pull from L2: 48 KB of K + V^T per 64-key step.  Here two CTAs of a
//     cluster (one TPC) work on two neighbouring 128-query tiles and share every K / V^T tile:
//     tcgen05.mma.cta_group::2 (M = 256) reads half of the B operand from each CTA's shared memory, so every SM ingests
//     24 KB per 64 keys.  One thread of the leader CTA issues the MMAs for both; tcgen05.commit ... multicast::cluster
//     releases the rings and score buffers of both CTAs; each CTA's TMA signals the leader's `full` barriers.
//   * The query tile moved from TMEM to shared memory (TMA, once per segment) -- with the B traffic halved, shared-memory
//     bandwidth (the reason it lived in TMEM) is no longer the binding resource -- which frees TMEM for separate score and
//     probability buffers.
//   * Scores and probabilities no longer share TMEM columns.  With P written over its own score buffer a score slot was
//     tied up for the whole S -> softmax -> P.V chain (~4300 cycles for two groups in flight, tensor pipe 74 % busy:
//     profiles/r02_attn3_trace_2slots.txt).  Now there is ONE 128-column score slot that goes back to the S issuer as soon
//     as every softmax warp holds its scores in registers (~300 cycles after the MMA), and four 32-column P buffers
//     (two groups in flight) that are recycled when their P.V MMA retires.
//   * The lazy rescale of O (tcgen05.ld/st round trip over 256 columns behind a drained P.V pipeline) was the
//     data-dependent slow path: on peaked scores (sigma ~ 15-20 log2 units in GPM layers 1-2) ~45 % of the tiles hit it.
//     The row maximum is now SEEDED with a lower bound of the true maximum computed by attn_seed_kernel (scores against
//     the 3x3 neighbourhood of the query's own position in every bank frame): the true maximum stays within the fp16
//     head-room (2^15) of the seed, so rescales become rare (tools/attn_rescale_sim.py); the rescale path itself stays as
//     the always-correct fallback.
//
//   cluster = 2 CTAs x 512 threads: warps 0-11 = three softmax groups (64-key sub-tiles round-robin), 12 Q/K producer,
//                                   13 V producer, 14 S issuer (leader) + TMEM owner, 15 P.V issuer (leader);
//                                   setmaxnreg gives the softmax warpgroups 152 registers and the last warpgroup 40
//   TMEM (512 cols per CTA): O[256] | S[128] (one group) | 4 x P[32] (packed fp16, two groups in flight)
//   smem per CTA: Q 32 KB | K ring 6 x 8 KB (32 of a sub-tile's 64 keys) | V^T ring 8 x 16 KB (128 of 256 dv rows x 64 keys)
//   Work: units = (query-tile pair, Dv chunk of 256); a unit is T * ceil(HW / 64) sub-tiles of 64 keys; the (unit, sub-tile)
//   sequence is cut into one contiguous range per cluster (stream-K, cost-aware bounds), <= 2 segments per cluster, each
//   flushed as normalised fp16 rows + fp32 (m, l); combine3_kernel merges, gates and emits the per-frame attention mass.
#include <cstdlib>

#include "attn.cuh"
#include "tcgen05.cuh"

namespace rmem {

namespace {

using namespace tc;

constexpr int BM = 128;        // query rows per CTA (256 per cluster)
constexpr int BNS = 64;        // keys per sub-tile (one softmax step, one P.V MMA group)
constexpr int DK = 128;
constexpr int DVC = 256;       // Dv columns per unit
constexpr int KS = 6;          // K ring depth (sub-tiles)
constexpr int VS = 8;          // V ring depth (sub-tiles)
constexpr int kGroups = 3;                       // softmax groups of four warps (one per TMEM lane quadrant)
constexpr int kSoftmaxWarps = 4 * kGroups;
constexpr int kWarpK = 12, kWarpV = 13, kWarpMmaS = 14, kWarpMmaPV = 15;
constexpr int kThreads = 16 * 32;
// Register budget (64 K per SM): the kernel starts at 128 per thread (512 threads); the producer / issuer warpgroup drops
// to kRegsIssue and the three softmax warpgroups grow to kRegsSoftmax (setmaxnreg): 3*128*152 + 128*56 = 65536.
constexpr int kRegsSoftmax = 152, kRegsIssue = 56;

constexpr int SMEM_Q = BM * DK * 2;              // 32 KB: [2 dk atoms][128 rows][128 B]
constexpr int SMEM_K = (BNS / 2) * DK * 2;       // 8 KB: [2 dk atoms][32 keys][128 B]    (this CTA's half of the sub-tile)
constexpr int SMEM_V = (DVC / 2) * BNS * 2;      // 16 KB: [128 dv rows][64 keys]         (this CTA's half of the chunk)
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + SMEM_Q;
constexpr int OFF_V = OFF_K + KS * SMEM_K;
constexpr int OFF_MSH = OFF_V + VS * SMEM_V;              // float [128]        row-max hand-over
constexpr int OFF_LX = OFF_MSH + BM * 4;                  // float [3][128][2]  (m, l) exchange at segment end
constexpr int OFF_BAR = OFF_LX + kGroups * BM * 2 * 4;
constexpr int SMEM_USED = OFF_BAR + 512;
constexpr int SMEM_TOTAL = SMEM_USED + 1024;              // + alignment slack (the base is rounded up to 1024 B)
static_assert(SMEM_TOTAL <= 232448, "shared memory budget");

constexpr int TMEM_COLS = 512;
constexpr int TMEM_O = 0;
constexpr int TMEM_S = 256;    // ONE 128-key score slot: sub-tile `sub` at TMEM_S + sub*64 (fp32)
constexpr int TMEM_P = 384;    // 4 P buffers of 32 columns (packed fp16 pairs): (group & 1) * 2 + sub

constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 15.0f;     // log2 units: P <= 2^15 (fp16 max 2^16) before a lazy rescale is forced

constexpr int kMaxCL = 96;     // clusters a launch may use (B200: 74)
struct Tc3Params {
  int HW, HWp, T, tpf, TPU, n_units, n_dv, nCL, Dv;   // tpf = 64-key sub-tiles per frame, TPU = T * tpf per unit
  int bounds[kMaxCL + 1];      // cluster c owns steps [bounds[c], bounds[c+1]) of the (unit, sub-tile) sequence
  int slot[kMaxBankFrames];
  float scale_log2;            // scale * log2(e)
  const float* qbias;          // [HW, T] or null (already multiplied by scale)
  const float* mseed;          // [HW] lower bound of the row maximum in log2 units (scale and bias applied) or null
  t16* part_o;                 // [nCTA][2][DVC/16][BM][16]  normalised partial O
  float* part_ml;              // [nCTA][2][BM][2]           (m in log2 units, l)
  float* pieces;               // [nCTA][2][T][3][BM][2]     per-frame (m, l) of each softmax group (Dv chunk 0) or null
  int* rescales;               // debug counter (warp-level rescale events) or null
  // guarded launch: when non-null, the kernel (and combine3 for its parameter set) runs only if *run_if != 0
  const int* run_if;
  // ---- column kernel (long_attn_tc4_kernel): see its header ----
  int column;                  // 1: bounds[] cut the (query pair, sub-tile) sequence, a cluster owns one column for all Dv
                               //    chunks; partial slot (cta * n_dv + chunk), one (m, l) / pieces record per CTA
  int max_sub;                 // sub-tiles a cluster's column may hold (rows of its P scratch)
  t16* pbuf;                   // [nCTA][max_sub][BM][64] probabilities of pass 0, re-read by the later passes
  int* overflow;               // set when a score leaves the fp16 range of the fixed reference: the caller falls back
};

// ---- cluster / pair primitives ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on the LEADER CTA's copy of a barrier (works from either CTA of the pair).  Default semantics (release at CTA
// scope), as CUTLASS's ClusterBarrier::arrive(cta_id): everything these barriers order lives in TMEM / the async proxy
// and is fenced with tcgen05.fence::before/after_thread_sync.  A `.release.cluster` arrive costs the arriving warp
// ~1000 cycles per sub-tile (profiles/r02_attn3_trace_first.txt: exp_done -> p_arrived 1200 cycles in the peer CTA
// against 240 in the leader).
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity, nullptr, 0); }
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity, nullptr, 0); }

// Tried and dropped (profiles/r02_attn3_experiments.txt): a polynomial 2^x on the FMA pipe for 1/8 .. 1/2 of the
// exponentials (slower in every version of this kernel: the added issue slots and registers cost more than the MUFU
// cycles saved), and ex2.approx.f16x2 (two exponentials per MUFU instruction on fp16-rounded exponents: 86 us and 1.6x
// the error).
// TMA tile into THIS CTA's shared memory, transaction bytes counted on the LEADER CTA's barrier (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
// MMA completion -> the same barrier in both CTAs of the pair
__device__ __forceinline__ void commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// D[tmem, 2 x 128 rows] (+)= A[smem of each CTA] . B[smem, half per CTA]^T
__device__ __forceinline__ void umma2_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem of each CTA: packed fp16 pairs, one lane per row] . B[smem, half per CTA]^T
__device__ __forceinline__ void umma2_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(TMEM_COLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(TMEM_COLS));
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

// unsaturated fp32x2 -> t16x2 (values are bounded by 2^15 here)
__device__ __forceinline__ uint32_t pack2_fast(float lo, float hi) {
#ifdef RMEM_OPERAND_BF16
  t162 v = __floats2bfloat162_rn(lo, hi);
#else
  t162 v = __floats2half2_rn(lo, hi);
#endif
  return *reinterpret_cast<uint32_t*>(&v);
}

// Optional event trace (null in production): clock64 per pipeline event of ONE cluster, [rank][256 sub-tiles][16] int64,
// then per-CTA wall times in rows 600.. ([cta][16]: globaltimer start / end, groups, segments, SM id, clock64 start / end).
__device__ long long* g_trace3 = nullptr;
__device__ int g_trace3_cl = 0;
#define TRACE3(j, k)                                                                         \
  do {                                                                                       \
    if (trace && lane == 0 && (j) < 256) trace[(long long)(j) * 16 + (k)] = clock64();        \
  } while (0)

struct Seg { int unit, lo, hi; };   // sub-tiles [lo, hi) of the unit

__global__ void __launch_bounds__(kThreads, 1)
long_attn_tc3_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const Tc3Params p) {
  extern __shared__ unsigned char smem_raw[];
  // 128B-swizzled TMA / UMMA tiles need 1024B alignment; the offset is the same in both CTAs of the pair
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* m_sh = reinterpret_cast<float*>(smem + OFF_MSH);
  float* lx = reinterpret_cast<float*>(smem + OFF_LX);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;                      // leader: both CTAs' query tiles have landed (one phase per load)
  uint64_t* q_free = q_full + 1;                // both:   every score MMA of the first segment has read Q
  uint64_t* k_full = q_free + 1;                // [KS] leader
  uint64_t* k_empty = k_full + KS;              // [KS] both
  uint64_t* v_full = k_empty + KS;              // [VS] leader
  uint64_t* v_empty = v_full + VS;              // [VS] both
  // S(sub-tile j) complete, both CTAs: six barriers in rotation (j % 6).  A softmax group owns every third sub-tile, so
  // all uses of one barrier belong to the same group and no waiter ever skips a phase.
  uint64_t* s_full = v_empty + VS;              // [6]
  uint64_t* s_free = s_full + 6;                // [2] leader: the owner (both CTAs) of score half-slot `sub` holds its scores
  uint64_t* p_full = s_free + 2;                // [4]  leader: P(sub-tile) stored by its softmax group in BOTH CTAs
  uint64_t* sp_free = p_full + 4;               // [4]  both:   P.V(sub-tile) complete: P buffer (and all before) retired
  uint64_t* o_drained = sp_free + 4;            // leader: both CTAs' segment epilogues have read O
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_drained + 1);
  static_assert((2 + 2 * KS + 2 * VS + 8 + 4 + 4 + 1) * 8 + 4 <= 512, "barrier area");

  if (p.run_if) {                               // fallback launch behind the column kernel: nothing to do unless it bailed out
    pdl_wait();
    if (*p.run_if == 0) return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cl = (int)cluster_id_x();
  const int cta = cl * 2 + (int)rank;           // partial-result slot owner
  long long* const trace = (g_trace3 && cl == g_trace3_cl) ? g_trace3 + (long long)rank * 256 * 16 : nullptr;
  long long* const cta_times = (g_trace3 && cta < 148) ? g_trace3 + (600 + cta) * 16 : nullptr;
  if (cta_times && threadIdx.x == 0) {
    unsigned long long gt; unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    cta_times[0] = (long long)gt; cta_times[4] = smid; cta_times[5] = clock64();
  }

  // ---- this cluster's work: a contiguous range of (unit, sub-tile) steps -> at most two segments ----
  const long long lo = p.bounds[cl], hi = p.bounds[cl + 1];
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
  const int n0 = nseg > 0 ? seg[0].hi - seg[0].lo : 0;                       // sub-tiles of the first segment
  const int ntot = n0 + (nseg > 1 ? seg[1].hi - seg[1].lo : 0);
  // units are ordered (query pair major, Dv chunk minor): the Q tiles change only when unit / n_dv changes
  const bool q_reload1 = nseg > 1 && (seg[1].unit / p.n_dv != seg[0].unit / p.n_dv);

  if (threadIdx.x == 0) {
    mbar_init(q_full, 2);
    mbar_init(q_free, 1);
    for (int i = 0; i < KS; ++i) { mbar_init(&k_full[i], 2); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < VS; ++i) { mbar_init(&v_full[i], 2); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 6; ++i) mbar_init(&s_full[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&s_free[i], 2 * 4);   // the four warps of the owning group, both CTAs
    for (int i = 0; i < 4; ++i) { mbar_init(&p_full[i], 8); mbar_init(&sp_free[i], 1); }
    mbar_init(o_drained, 2 * kSoftmaxWarps);           // every softmax warp of both CTAs
    mbar_fence_init();
  }
  if (warp == kWarpMmaS) tmem_alloc_pair(tmem_slot);
  fence_before();
  __syncthreads();
  cluster_sync_all();           // the peer's barriers are initialised and both TMEM allocations are done
  fence_after();
  const uint32_t tmem = *tmem_slot;
  if (cta_times && threadIdx.x == 0) cta_times[7] = clock64();     // set-up done (barriers, TMEM, cluster sync)
  pdl_prologue();   // barriers, TMEM and descriptors are set up; global memory is touched only from here on
  if (cta_times && threadIdx.x == 0) cta_times[8] = clock64();     // previous kernel complete

  if (warp >= kSoftmaxWarps) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsIssue));
  if (warp == kWarpK) {
    // ================================ Q + K producer (both CTAs, own halves) ================================
    if (ntot > 0) {
      if (elect_one()) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_k);
      }
      __syncwarp();
      int i = 0;
      for (int s = 0; s < nseg; ++s) {
        if (s == 0 || q_reload1) {
          if (s == 1) mbar_wait_b(q_free, 0);                // the first segment's score MMAs are done with Q
          if (elect_one()) {
            const int row0 = ((seg[s].unit / p.n_dv) * 2 + (int)rank) * BM;
            tma_load_2d_pair(smem + OFF_Q, &map_q, q_full, 0, row0);
            tma_load_2d_pair(smem + OFF_Q + BM * 128, &map_q, q_full, 64, row0);
            if (leader) mbar_expect_tx(q_full, 2 * SMEM_Q); else mbar_arrive_leader(q_full);
          }
          __syncwarp();
        }
        int t = seg[s].lo / p.tpf, jt = seg[s].lo - t * p.tpf;
        for (int g = seg[s].lo; g < seg[s].hi; ++g, ++i, ++jt) {
          if (jt == p.tpf) { jt = 0; ++t; }
          const int st = i % KS;
          if (i >= KS) mbar_wait_b(&k_empty[st], ((i / KS) - 1) & 1);
          if (elect_one()) {
            const int key0 = p.slot[t] * p.HWp + jt * BNS + (int)rank * (BNS / 2);
            unsigned char* sk = smem + OFF_K + st * SMEM_K;
            tma_load_2d_pair(sk, &map_k, &k_full[st], 0, key0);
            tma_load_2d_pair(sk + (BNS / 2) * 128, &map_k, &k_full[st], 64, key0);
            if (leader) mbar_expect_tx(&k_full[st], 2 * SMEM_K); else mbar_arrive_leader(&k_full[st]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kWarpV) {
    // ================================ V^T producer (both CTAs, own dv half) ================================
    if (ntot > 0) {
      if (elect_one()) tma_prefetch_desc(&map_v);
      __syncwarp();
      int i = 0;
      for (int s = 0; s < nseg; ++s) {
        const int dv0 = (seg[s].unit % p.n_dv) * DVC + (int)rank * (DVC / 2);
        int t = seg[s].lo / p.tpf, jt = seg[s].lo - t * p.tpf;
        for (int g = seg[s].lo; g < seg[s].hi; ++g, ++i, ++jt) {
          if (jt == p.tpf) { jt = 0; ++t; }
          const int st = i % VS;
          if (i >= VS) mbar_wait_b(&v_empty[st], ((i / VS) - 1) & 1);
          if (elect_one()) {
            const int key0 = p.slot[t] * p.HWp + jt * BNS;
            tma_load_2d_pair(smem + OFF_V + st * SMEM_V, &map_v, &v_full[st], key0, dv0);
            if (leader) mbar_expect_tx(&v_full[st], 2 * SMEM_V); else mbar_arrive_leader(&v_full[st]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kWarpMmaS) {
    // ================================ S = Q.K^T issuer (leader only) ================================
    if (ntot > 0 && leader) {
      constexpr uint32_t idesc_s = make_idesc(2 * BM, BNS);
      const uint32_t smem_base = smem_u32(smem);
      // One 64-key score MMA group per sub-tile into half-slot (j & 1).  The two half-slots are recycled independently, each
      // as soon as its single owner group has the scores in registers (a 128-key MMA into one slot had to wait for BOTH
      // owners: 2180 cycles per key group, profiles/r02_attn3_trace_3groups_1slot.txt).
      for (int i = 0; i < ntot; ++i) {
        const int st = i % KS, sub = i & 1;
        if (i == 0) mbar_wait_cl(q_full, 0);
        if (i == n0 && q_reload1) mbar_wait_cl(q_full, 1);
        mbar_wait_cl(&k_full[st], (i / KS) & 1);
        TRACE3(i, 10);
        if (i >= 2) mbar_wait_cl(&s_free[sub], ((i >> 1) - 1) & 1);   // the scores of sub-tile i-2 are in registers
        fence_after();
        TRACE3(i, 2);
        if (elect_one()) {
          const uint64_t dq = make_desc_sw128(smem_base + OFF_Q);
          const uint64_t dk = make_desc_sw128(smem_base + OFF_K + st * SMEM_K);
          const uint32_t d = tmem + TMEM_S + sub * BNS;
#pragma unroll
          for (int kk = 0; kk < DK / 16; ++kk) {
            // 32 B per k-step inside the 128 B swizzle atom; the second 64-wide dk atom starts one tile-half later
            const uint64_t oa = (uint64_t)(((kk >> 2) * (BM * 128) + (kk & 3) * 32) >> 4);
            const uint64_t ob = (uint64_t)(((kk >> 2) * ((BNS / 2) * 128) + (kk & 3) * 32) >> 4);
            umma2_ss(d, dq + oa, dk + ob, idesc_s, kk > 0);
          }
          commit_pair(&k_empty[st]);
          commit_pair(&s_full[i % 6]);
          if (i == n0 - 1 && q_reload1) commit_pair(q_free);
        }
        __syncwarp();
        TRACE3(i, 3);
      }
    }
  } else {
    // ================================ O += P.V issuer (leader only) ================================
    if (ntot > 0 && leader) {
      constexpr uint32_t idesc_o = make_idesc(2 * BM, DVC);
      const uint32_t smem_base = smem_u32(smem);
      for (int j = 0; j < ntot; ++j) {
        const int b = j & 3, sv = j % VS;
        mbar_wait_cl(&v_full[sv], (j / VS) & 1);             // long since complete: its latency hides behind the P wait
        TRACE3(j, 13);
        const bool first = (j == 0) || (j == n0);
        if (j == n0 && nseg > 1) mbar_wait_cl(o_drained, 0);
        mbar_wait_cl(&p_full[b], (j >> 2) & 1);
        fence_after();
        TRACE3(j, 0);
        if (cta_times && lane == 0) { if (j == 0) cta_times[9] = clock64(); if (j == ntot - 1) cta_times[10] = clock64(); }
        if (elect_one()) {
          const uint64_t dv = make_desc_sw128(smem_base + OFF_V + sv * SMEM_V);
          const uint32_t pa = tmem + TMEM_P + b * 32;
#pragma unroll
          for (int kk = 0; kk < BNS / 16; ++kk)
            umma2_ts(tmem + TMEM_O, pa + kk * 8, dv + (uint64_t)(kk * 2), idesc_o, (first && kk == 0) ? 0u : 1u);
          commit_pair(&v_empty[sv]);
          commit_pair(&sp_free[b]);
        }
        __syncwarp();
        TRACE3(j, 1);
      }
    }
  }
  } else {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsSoftmax));
  if (ntot > 0) {
    // ================================ softmax + epilogue (warps 0-11, both CTAs) ================================
    // Three softmax groups of four warps (one warp per TMEM lane quadrant) take the 64-key sub-tiles round-robin.  Two
    // groups could not keep up with the tensor pipe: a sub-tile costs a softmax warp ~2300 cycles (MUFU: 64 ex2 per thread
    // at 8 cycles per warp instruction, shared by the warps of an SM sub-core, plus ~1100 cycles of barrier / TMEM
    // latencies) against 1580 cycles of tensor work per 128-key group (profiles/r02_attn3_trace_1slot.txt).
    const int quad = warp & 3, grp = warp >> 2;
    const int row = quad * 32 + lane;                       // tile row == TMEM lane
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    // named barriers: hand-over of the running maximum from group a to group (a+1)%3 on id 1 + a*4 + quad (64 threads),
    // segment-end exchange on id 13 (all softmax warps)
    const int id_in = 1 + ((grp + kGroups - 1) % kGroups) * 4 + quad;
    const int id_out = 1 + grp * 4 + quad;
    constexpr int id_ex = 13;
    int j = 0;                                              // sub-tile counter over both segments
    int own = grp;                                          // next sub-tile this group handles (j % 3 == grp)

    for (int s = 0; s < nseg; ++s) {
      const int unit = seg[s].unit;
      const int qt = (unit / p.n_dv) * 2 + (int)rank, dvc = unit % p.n_dv;
      const int qi = qt * BM + row;
      const bool row_ok = qi < p.HW;
      const float seed = (p.mseed && row_ok) ? p.mseed[qi] : -INFINITY;
      // m_ref: the running maximum this group's l / l_piece are relative to
      float m_ref = -INFINITY, l_tot = 0.f, l_piece = 0.f, bias2 = 0.f;
      int cur_t = -1;
      const int j_first = j, j_last = j + (seg[s].hi - seg[s].lo) - 1;
      float* piece_base = (p.pieces && dvc == 0)
                              ? p.pieces + (((long long)(cta * 2 + s) * p.T) * kGroups + grp) * (BM * 2) + row * 2
                              : nullptr;
      auto flush_piece = [&](int t) {
        if (piece_base) {
          float* d = piece_base + (long long)t * (kGroups * BM * 2);
          d[0] = m_ref;
          d[1] = l_piece;
        }
      };
      int t = seg[s].lo / p.tpf, jt = seg[s].lo - t * p.tpf;
      for (int g = seg[s].lo; g < seg[s].hi; ++g, ++jt, ++j) {
        if (jt == p.tpf) { jt = 0; ++t; }
        if (t != cur_t) {                                   // every group walks every sub-tile's frame index
          if (cur_t >= 0) flush_piece(cur_t);
          cur_t = t;
          l_piece = 0.f;
          bias2 = (p.qbias && row_ok) ? p.qbias[(long long)qi * p.T + t] * LOG2E : 0.f;
        }
        {
          const int jj = j;
          if (jj != own) continue;
          own += kGroups;
          const int sub = jj & 1, b = jj & 3;               // score half-slot, P buffer
          long long* const trace_s = quad == 0 ? trace : nullptr;
#define TRACE3_S(k) do { if (trace_s && lane == 0 && jj < 256) trace_s[(long long)jj * 16 + (k)] = clock64(); } while (0)
          TRACE3_S(4);
          mbar_wait_b(&s_full[jj % 6], (jj / 6) & 1);
          fence_after();
          TRACE3_S(5);
          float sc[64];
          {
            uint32_t r0[32], r1[32];
            tmem_ld32_nowait(lane_addr + TMEM_S + sub * BNS, r0);
            tmem_ld32_nowait(lane_addr + TMEM_S + sub * BNS + 32, r1);
            // P buffer b was last read by P.V of sub-tile jj-4 (long retired in the steady state): the wait's own latency
            // overlaps the TMEM load
            if (jj >= 4) {
              mbar_wait_b(&sp_free[b], ((jj >> 2) - 1) & 1);
              fence_after();
            }
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) { sc[c] = __uint_as_float(r0[c]); sc[32 + c] = __uint_as_float(r1[c]); }
          }
          // the scores are in registers: hand the slot back to the S issuer (the next group's scores are computed while
          // this one goes through max / exp / P store)
          fence_before();
          __syncwarp();
          if (lane == 0) { if (leader) mbar_arrive(&s_free[sub]); else mbar_arrive_leader(&s_free[sub]); }
          const int key0 = jt * BNS;
          if (key0 + BNS > p.HW) {                             // ragged / padding sub-tile at the end of the frame
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
          TRACE3_S(6);
          // ---- hand-over of the lazily updated row maximum from the previous group's sub-tile jj-1 ----
          float m_prev = seed;
          if (jj > j_first) {
            named_bar_sync(id_in, 64);
            m_prev = m_sh[row];
          }
          const bool need = mt > m_prev + RESCALE_THRESHOLD;
          float m_new = m_prev;
          if (__any_sync(0xffffffffu, need)) {
            if (jj > j_first) {
              const int jp = jj - 1;
              mbar_wait_b(&sp_free[jp & 3], (jp >> 2) & 1);    // P.V(<= jj-1) retired
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
              if (p.rescales && lane == 0) atomicAdd(p.rescales, 1);
            }
            if (need) m_new = mt;
          }
          if (jj < j_last) {
            m_sh[row] = m_new;
            named_bar_arrive(id_out, 64);
          }
          TRACE3_S(7);
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
            const float e3 = exp2f(fmaf(sc[c + 3], p.scale_log2, c0));
            ls0 += e0; ls1 += e1; ls2 += e2; ls3 += e3;
            pk[c >> 1] = pack2_fast(e0, e1);
            pk[(c >> 1) + 1] = pack2_fast(e2, e3);
          }
          const float lsum = (ls0 + ls1) + (ls2 + ls3);
          l_tot += lsum;
          l_piece += lsum;
          TRACE3_S(8);
          tmem_st32u(lane_addr + TMEM_P + b * 32, pk);
          fence_before();
          __syncwarp();
          if (lane == 0) { if (leader) mbar_arrive(&p_full[b]); else mbar_arrive_leader(&p_full[b]); }
          TRACE3_S(9);
          if (trace && quad > 0 && lane == 0 && jj < 256) trace[(long long)jj * 16 + (quad == 1 ? 11 : quad == 2 ? 12 : 14)] = clock64();
        }
      }
      flush_piece(cur_t);

      // ---- segment epilogue: normalised fp16 partial O + (m, l) ----
      lx[(grp * BM + row) * 2 + 0] = m_ref;
      lx[(grp * BM + row) * 2 + 1] = l_tot;
      named_bar_sync(id_ex, kSoftmaxWarps * 32);
      float M = -INFINITY;                                  // == the maximum O is relative to (m is monotone)
#pragma unroll
      for (int a = 0; a < kGroups; ++a) M = fmaxf(M, lx[(a * BM + row) * 2]);
      float l_row = 0.f;
#pragma unroll
      for (int a = 0; a < kGroups; ++a) {
        const float la = lx[(a * BM + row) * 2 + 1];
        if (la > 0.f) l_row += la * exp2f(lx[(a * BM + row) * 2] - M);
      }
      const float inv = l_row > 0.f ? 1.f / l_row : 0.f;    // a seed far above this segment's scores can leave l = 0
      mbar_wait_b(&sp_free[j_last & 3], (j_last >> 2) & 1);
      fence_after();
      // part_o: [slot][16-column group][row][16] -- the 32 rows of a warp are contiguous per group, so every store
      // instruction covers whole lines.  32-column chunks 0-2 / 3-5 / 6-7 go to softmax groups 0 / 1 / 2.
      t16* po = p.part_o + (((long long)(cta * 2 + s) * (DVC / 16)) * BM + row) * 16;
      const int c_lo = grp * 96, c_hi = grp == 2 ? DVC : c_lo + 96;
#pragma unroll 1
      for (int c = c_lo; c < c_hi; c += 32) {
        float o[32];
        tmem_ld32(lane_addr + TMEM_O + c, o);
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
      if (grp == 0) {
        float* ml = p.part_ml + ((long long)(cta * 2 + s) * BM + row) * 2;
        ml[0] = M;
        ml[1] = l_row;
      }
      fence_before();
      __syncwarp();
      if (cta_times && warp == 0 && lane == 0) cta_times[11 + s] = clock64();   // segment epilogue written
      if (lane == 0 && s + 1 < nseg) mbar_arrive_leader(o_drained);
      // the exchange area is reused by the next segment: every group must have read it before anyone writes again
      if (s + 1 < nseg) named_bar_sync(id_ex, kSoftmaxWarps * 32);
    }
  }
  }
  fence_before();
  __syncthreads();
  cluster_sync_all();           // the leader's MMAs read the peer's shared memory and write its TMEM until here
  if (cta_times && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    cta_times[1] = (long long)gt; cta_times[2] = ntot; cta_times[3] = nseg; cta_times[6] = clock64();
  }
  if (warp == kWarpMmaS) {
    fence_after();
    tmem_dealloc_pair(tmem);
  }
}

// Merge the segments of every unit: out = (sum_s w_s O_s) * gate with w_s = l_s 2^(m_s - M) / L;
// mass[i,t] = sum_{pieces of frame t} l_p 2^(m_p - M) / L  (from the Dv-chunk-0 units).
// Four query rows per block, 64 threads x 16 columns per row (see attn_tc2.cu's combine2_kernel; here a unit covers a
// PAIR of query tiles and the partial slot is (cluster * 2 + rank) * 2 + segment).
constexpr int kMaxSegsPerUnit = 24;
constexpr int kCombRows = 4;
constexpr int kColGroups = DVC / 16;
static_assert(BM % kCombRows == 0, "the rows of a combine block share a query tile");
__device__ __forceinline__ void combine3_body(const Tc3Params& p, const t16* __restrict__ gate, long long ldg,
                                              t16* __restrict__ out, long long ldo, float* __restrict__ mass);
__global__ void __launch_bounds__(256) combine3_kernel(const Tc3Params p, const t16* __restrict__ gate, long long ldg,
                                                       t16* __restrict__ out, long long ldo,
                                                       float* __restrict__ mass) {
  pdl_prologue();
  combine3_body(p, gate, ldg, out, ldo, mass);
}
// Column layout (long_attn_tc4_kernel): every cluster of a query pair contributes a segment to ALL four Dv chunks, so a row
// merges 10-11 segments instead of 3-4 and its (m, l) are shared by the chunks.  The serial walk of combine3_body (two
// dependent L2 reads per segment before the first partial row is touched) would cost more than the column kernel saves:
// here lane e of the row's first warp fetches segment e's (m, l), the weights come from two warp reductions, and the partial
// rows are read four segments at a time.
__device__ __forceinline__ void combine4_body(const Tc3Params& p, const t16* __restrict__ gate, long long ldg,
                                              t16* __restrict__ out, long long ldo, float* __restrict__ mass) {
  __shared__ float s_w[kCombRows][32];           // l_s 2^(m_s - M) / L per segment
  __shared__ float s_M[kCombRows], s_invL[kCombRows];
  __shared__ int s_c0[kCombRows], s_n[kCombRows];
  const int rr = threadIdx.x >> 6, tc = threadIdx.x & 63;
  const int i = blockIdx.x * kCombRows + rr;
  const bool live = i < p.HW;
  const int qt = (blockIdx.x * kCombRows) / BM, r = i - qt * BM;
  const int qp = qt >> 1, rank = qt & 1;
  if (tc < 32) {
    const int lo = qp * p.TPU, hi = lo + p.TPU;
    int c0 = 0;                                      // first cluster of the query pair (bounds ascend, cut at pair boundaries)
    for (int step = 64; step > 0; step >>= 1)
      if (c0 + step < p.nCL && p.bounds[c0 + step] <= lo) c0 += step;
    int n = 0;
    while (c0 + n < p.nCL && p.bounds[c0 + n] < hi && n < 32) ++n;
    float m = -INFINITY, l = 0.f;
    if (live && tc < n) {
      const float2 ml = *reinterpret_cast<const float2*>(p.part_ml + ((long long)((c0 + tc) * 2 + rank) * BM + r) * 2);
      if (ml.y > 0.f) { m = ml.x; l = ml.y; }
    }
    const float M = warp_max(m);
    const float w = l > 0.f ? exp2f(m - M) * l : 0.f;
    const float L = warp_sum(w);
    const float invL = L > 0.f ? 1.f / L : 0.f;
    s_w[rr][tc] = w * invL;
    if (tc == 0) { s_M[rr] = M; s_invL[rr] = invL; s_c0[rr] = c0; s_n[rr] = n; }
  }
  __syncthreads();
  const int c0 = s_c0[rr], n = s_n[rr];
  const int col = tc * 16;
  if (live && col < p.Dv) {
    const int k = col / DVC, cg = (col - k * DVC) >> 4;
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.f;
    const t16* base = p.part_o + ((long long)cg * BM + r) * 16;
    const long long slot_stride = (long long)kColGroups * BM * 16;
    for (int e0 = 0; e0 < n; e0 += 4) {
      uint4 u0[4], u1[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (e0 + q < n) {
          const uint4* src = reinterpret_cast<const uint4*>(base + (long long)(((c0 + e0 + q) * 2 + rank) * p.n_dv + k) * slot_stride);
          u0[q] = src[0];
          u1[q] = src[1];
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (e0 + q < n) {
          const float w = s_w[rr][e0 + q];
          const uint32_t uu[8] = {u0[q].x, u0[q].y, u0[q].z, u0[q].w, u1[q].x, u1[q].y, u1[q].z, u1[q].w};
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float2 f = unpack2(uu[c]);
            acc[c * 2] = fmaf(w, f.x, acc[c * 2]);
            acc[c * 2 + 1] = fmaf(w, f.y, acc[c * 2 + 1]);
          }
        }
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
    const float M = s_M[rr];
    float a = 0.f;
    for (int e = 0; e < n; ++e) {
      const int lo = p.bounds[c0 + e] - qp * p.TPU, hi = p.bounds[c0 + e + 1] - qp * p.TPU;
      if (lo < f_hi && f_lo < hi) {
        const float* pc = p.pieces + ((((long long)((c0 + e) * 2 + rank) * p.T + t) * kGroups) * BM + r) * 2;
        const float2 p0 = *reinterpret_cast<const float2*>(pc);
        const float2 p1 = *reinterpret_cast<const float2*>(pc + BM * 2);
        const float2 p2 = *reinterpret_cast<const float2*>(pc + 2 * BM * 2);
        if (p0.y > 0.f) a += exp2f(p0.x - M) * p0.y;
        if (p1.y > 0.f) a += exp2f(p1.x - M) * p1.y;
        if (p2.y > 0.f) a += exp2f(p2.x - M) * p2.y;
      }
    }
    mass[(long long)i * p.T + t] = a * s_invL[rr];
  }
}
// Column kernel + its guarded fallback: merges whichever of the two partial sets is valid (pb when *flag is set).
__global__ void __launch_bounds__(256) combine34_kernel(const Tc3Params pa, const Tc3Params pb, const int* __restrict__ flag,
                                                        const t16* __restrict__ gate, long long ldg,
                                                        t16* __restrict__ out, long long ldo, float* __restrict__ mass) {
  pdl_prologue();
  if (*flag) combine3_body(pb, gate, ldg, out, ldo, mass);
  else combine4_body(pa, gate, ldg, out, ldo, mass);
}
__device__ __forceinline__ void combine3_body(const Tc3Params& p, const t16* __restrict__ gate, long long ldg,
                                              t16* __restrict__ out, long long ldo, float* __restrict__ mass) {
  __shared__ int s_n[kCombRows][4];
  __shared__ int s_ml[kCombRows][4][kMaxSegsPerUnit];          // (m, l) / pieces record of the segment
  __shared__ int s_slot[kCombRows][4][kMaxSegsPerUnit];        // (cta * 2 + seg)
  __shared__ float s_w[kCombRows][4][kMaxSegsPerUnit];         // l_s 2^(m_s - M) / L
  __shared__ float s_M0[kCombRows], s_L0[kCombRows];
  __shared__ int s_alo[kCombRows][kMaxSegsPerUnit], s_ahi[kCombRows][kMaxSegsPerUnit];   // unit-0 group ranges
  const int rr = threadIdx.x >> 6, tc = threadIdx.x & 63;
  const int i = blockIdx.x * kCombRows + rr;
  const bool live = i < p.HW;
  const int qt = (blockIdx.x * kCombRows) / BM, r = i - qt * BM;
  const int qp = qt >> 1, rank = qt & 1;
  if (live && tc < p.n_dv) {
    const int k = tc;
    const int unit = p.column ? qp : qp * p.n_dv + k;
    const long long u_lo = (long long)unit * p.TPU, u_hi = u_lo + p.TPU;
    int c = 0;                                       // first cluster whose range reaches past u_lo (bounds ascend)
    for (int step = 64; step > 0; step >>= 1)
      if (c + step < p.nCL && p.bounds[c + step] <= u_lo) c += step;
    if (p.bounds[c + 1] <= u_lo) ++c;
    int n = 0;
    float M = -INFINITY;
    for (; c < p.nCL && n < kMaxSegsPerUnit; ++c) {
      const long long lo = p.bounds[c], hi = p.bounds[c + 1];
      if (lo >= u_hi) break;
      if (hi <= lo) continue;
      const int slot = p.column ? (c * 2 + rank) * p.n_dv + k : (c * 2 + rank) * 2 + (lo < u_lo ? 1 : 0);
      const int mls = p.column ? (c * 2 + rank) : slot;
      s_slot[rr][k][n] = slot;
      s_ml[rr][k][n] = mls;
      if (k == 0) {
        s_alo[rr][n] = (int)((lo > u_lo ? lo : u_lo) - u_lo);
        s_ahi[rr][n] = (int)((hi < u_hi ? hi : u_hi) - u_lo);
      }
      const float* ml = p.part_ml + ((long long)mls * BM + r) * 2;
      if (ml[1] > 0.f) M = fmaxf(M, ml[0]);
      ++n;
    }
    float L = 0.f;
    for (int e = 0; e < n; ++e) {
      const float* ml = p.part_ml + ((long long)s_ml[rr][k][e] * BM + r) * 2;
      const float w = ml[1] > 0.f ? exp2f(ml[0] - M) * ml[1] : 0.f;
      s_w[rr][k][e] = w;
      L += w;
    }
    const float inv = L > 0.f ? 1.f / L : 0.f;
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
    const float M = s_M0[rr], invL = s_L0[rr] > 0.f ? 1.f / s_L0[rr] : 0.f;
    float a = 0.f;
    for (int e = 0; e < s_n[rr][0]; ++e) {
      if (s_alo[rr][e] < f_hi && f_lo < s_ahi[rr][e]) {
        const float* pc = p.pieces + ((((long long)s_ml[rr][0][e] * p.T + t) * kGroups) * BM + r) * 2;
        // a piece whose sum is zero may carry m = -inf (no sub-tile of that frame seen by the group): skip it
#pragma unroll
        for (int gq = 0; gq < kGroups; ++gq)
          if (pc[gq * BM * 2 + 1] > 0.f) a += exp2f(pc[gq * BM * 2] - M) * pc[gq * BM * 2 + 1];
      }
    }
    mass[(long long)i * p.T + t] = a * invL;
  }
}

// =================================================================================================================
// Column kernel (RMEM_ATTN_TC4): the scores and their exponentials are computed ONCE per (query pair, sub-tile) instead
// of once per Dv chunk.  long_attn_tc3_kernel recomputes S = Q.K^T and the softmax for each of the four 256-column value
// chunks because a CTA pair's TMEM holds 128 x 1024 fp32 accumulators with no column left for S: a third
// of its MMA time and three quarters of its MUFU work are repeats.  Here a cluster owns a COLUMN -- one query pair, a
// contiguous range of n <= max_sub sub-tiles, all four chunks -- and walks it in four passes:
//   pass 0   exactly the pipeline above for chunk 0 (S -> softmax -> P -> P.V into O_a = TMEM 0..255), and every
//            softmax thread also stores its 64 probabilities (128 B) to the CTA's scratch rows in global memory (L2);
//   pass c   (c = 1, 2, 3) the Q/K producer streams those P tiles back by TMA (16 KB per sub-tile, 128B swizzle, into a
//            5-stage ring over the dead Q + K ring), the V producer streams chunk c, and the P.V issuer runs
//            O += P.V with A = P from SHARED memory: no score MMAs, no exponentials, no TMEM round trip.  Passes alternate
//            between O_a and O_b = TMEM 256..511 (the dead S / P columns), so the softmax warps -- idle otherwise -- read
//            out and store pass c-1 while pass c accumulates.
// What makes the passes independent is a FIXED softmax reference: P = 2^(s - m_ref) with m_ref = seed + kRefShift per row
// (seed = the 3x3-neighbourhood lower bound of the row maximum, qprep_seed_kernel).  No running maximum, no hand-over
// between the softmax groups, no lazy rescale -- so nothing of pass 0 has to be replayed later.  The price is range: fp16
// P covers s in [m_ref - 24, m_ref + 15.5] = [seed - 18, seed + 21.5].  Measured on the engine's c3 operands the true row
// maximum sits 1-5 (at most 15.2) above the seed and what falls below the range carries < 2e-3 of a row's mass (mean
// 5e-5).  A score above the range sets *overflow: the launcher has long_attn_tc3_kernel queued right behind, which runs
// only then (always-correct fallback), and combine34_kernel merges whichever partial set is valid.
constexpr int PS = 5;                                    // P ring depth (16 KB stages over the Q tile + K ring)
static_assert(PS * (BM * BNS * 2) <= SMEM_Q + KS * SMEM_K, "P ring must fit over the Q tile and the K ring");
constexpr uint32_t kLboP = BM * 16, kSboP = 128;     // P scratch tile: see the store in pass 0
constexpr float kRefShift = 6.0f;
constexpr float kOverflowAt = 15.5f;
constexpr int OFF_BAR4 = OFF_BAR + 512;                  // the column kernel's extra barriers
constexpr int SMEM_TOTAL4 = SMEM_USED + 512 + 1024;
static_assert(SMEM_TOTAL4 <= 232448, "shared memory budget");

__global__ void __launch_bounds__(kThreads, 1)
long_attn_tc4_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_p,
                     const Tc3Params p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* lx = reinterpret_cast<float*>(smem + OFF_LX);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;                      // leader
  uint64_t* k_full = q_full + 2;                // [KS] leader
  uint64_t* k_empty = k_full + KS;              // [KS] both
  uint64_t* v_full = k_empty + KS;              // [VS] leader
  uint64_t* v_empty = v_full + VS;              // [VS] both
  uint64_t* s_full = v_empty + VS;              // [6] both
  uint64_t* s_free = s_full + 6;                // [2] leader
  uint64_t* p_full = s_free + 2;                // [4] leader
  uint64_t* sp_free = p_full + 4;               // [4] both
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sp_free + 4 + 1);
  uint64_t* bars4 = reinterpret_cast<uint64_t*>(smem + OFF_BAR4);
  uint64_t* p_ready = bars4;                    // own CTA: every softmax warp has stored (and fenced) its pass-0 probabilities
  uint64_t* pk_full = p_ready + 1;              // [PS] leader: both CTAs' P tiles of a sub-tile have landed
  uint64_t* pk_empty = pk_full + PS;            // [PS] both:   the P.V MMAs that read the stage have completed
  uint64_t* pass_done = pk_empty + PS;          // [4]  both:   every P.V MMA of pass c has completed
  uint64_t* o_drained = pass_done + 4;          // [2]  leader: both CTAs have read out pass c (c = 0, 1)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cl = (int)cluster_id_x();
  const int cta = cl * 2 + (int)rank;
  // optional per-CTA phase times (tools/trace_attn4.py): rows 600 + cta of the trace buffer
  long long* const cta_times = (g_trace3 && cta < 148) ? g_trace3 + (600 + cta) * 16 : nullptr;
  if (cta_times && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    cta_times[0] = (long long)gt; cta_times[5] = clock64();
  }

  // this cluster's column: query pair `qp`, sub-tiles [lo, lo + n) of its T * tpf
  const int g_lo = p.bounds[cl], g_hi = p.bounds[cl + 1];
  const int qp = g_lo / p.TPU;
  const int lo = g_lo - qp * p.TPU;
  const int n = g_hi - g_lo;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 2);
    for (int i = 0; i < KS; ++i) { mbar_init(&k_full[i], 2); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < VS; ++i) { mbar_init(&v_full[i], 2); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 6; ++i) mbar_init(&s_full[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&s_free[i], 2 * 4);
    for (int i = 0; i < 4; ++i) { mbar_init(&p_full[i], 8); mbar_init(&sp_free[i], 1); }
    mbar_init(p_ready, kSoftmaxWarps);
    for (int i = 0; i < PS; ++i) { mbar_init(&pk_full[i], 2); mbar_init(&pk_empty[i], 1); }
    for (int i = 0; i < 4; ++i) mbar_init(&pass_done[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&o_drained[i], 2 * kSoftmaxWarps);
    mbar_fence_init();
  }
  if (warp == kWarpMmaS) tmem_alloc_pair(tmem_slot);
  fence_before();
  __syncthreads();
  cluster_sync_all();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  if (cta_times && threadIdx.x == 0) cta_times[7] = clock64();
  pdl_prologue();
  if (cta_times && threadIdx.x == 0) cta_times[8] = clock64();

  if (warp >= kSoftmaxWarps) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsIssue));
  if (warp == kWarpK) {
    // ================================ Q + K producer (pass 0), P producer (passes 1-3) ================================
    if (n > 0) {
      if (elect_one()) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_p);
      }
      __syncwarp();
      if (elect_one()) {
        const int row0 = (qp * 2 + (int)rank) * BM;
        tma_load_2d_pair(smem + OFF_Q, &map_q, q_full, 0, row0);
        tma_load_2d_pair(smem + OFF_Q + BM * 128, &map_q, q_full, 64, row0);
        if (leader) mbar_expect_tx(q_full, 2 * SMEM_Q); else mbar_arrive_leader(q_full);
      }
      __syncwarp();
      {
        int t = lo / p.tpf, jt = lo - t * p.tpf;
        for (int i = 0; i < n; ++i, ++jt) {
          if (jt == p.tpf) { jt = 0; ++t; }
          const int st = i % KS;
          if (i >= KS) mbar_wait_b(&k_empty[st], ((i / KS) - 1) & 1);
          if (elect_one()) {
            const int key0 = p.slot[t] * p.HWp + jt * BNS + (int)rank * (BNS / 2);
            unsigned char* sk = smem + OFF_K + st * SMEM_K;
            tma_load_2d_pair(sk, &map_k, &k_full[st], 0, key0);
            tma_load_2d_pair(sk + (BNS / 2) * 128, &map_k, &k_full[st], 64, key0);
            if (leader) mbar_expect_tx(&k_full[st], 2 * SMEM_K); else mbar_arrive_leader(&k_full[st]);
          }
          __syncwarp();
        }
      }
      // passes 1-3: this CTA's own probabilities, written by its softmax warps in pass 0 (generic proxy, fenced for the
      // async proxy before p_ready); the Q tile and the K ring are dead by then (every score MMA has completed)
      mbar_wait_b(p_ready, 0);
      int x = 0;
      for (int c = 1; c < p.n_dv; ++c)
        for (int i = 0; i < n; ++i, ++x) {
          const int st = x % PS;
          if (x >= PS) mbar_wait_b(&pk_empty[st], ((x / PS) - 1) & 1);
          if (elect_one()) {
            tma_load_2d_pair(smem + OFF_Q + st * (BM * BNS * 2), &map_p, &pk_full[st], 0, (cta * p.max_sub + i) * BM);
            if (leader) mbar_expect_tx(&pk_full[st], 2 * BM * BNS * 2); else mbar_arrive_leader(&pk_full[st]);
          }
          __syncwarp();
        }
    }
  } else if (warp == kWarpV) {
    // ================================ V^T producer: chunk c of pass c ================================
    if (n > 0) {
      if (elect_one()) tma_prefetch_desc(&map_v);
      __syncwarp();
      int x = 0;
      for (int c = 0; c < p.n_dv; ++c) {
        const int dv0 = c * DVC + (int)rank * (DVC / 2);
        int t = lo / p.tpf, jt = lo - t * p.tpf;
        for (int i = 0; i < n; ++i, ++x, ++jt) {
          if (jt == p.tpf) { jt = 0; ++t; }
          const int st = x % VS;
          if (x >= VS) mbar_wait_b(&v_empty[st], ((x / VS) - 1) & 1);
          if (elect_one()) {
            const int key0 = p.slot[t] * p.HWp + jt * BNS;
            tma_load_2d_pair(smem + OFF_V + st * SMEM_V, &map_v, &v_full[st], key0, dv0);
            if (leader) mbar_expect_tx(&v_full[st], 2 * SMEM_V); else mbar_arrive_leader(&v_full[st]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kWarpMmaS) {
    // ================================ S = Q.K^T issuer (leader, pass 0 only) ================================
    if (n > 0 && leader) {
      constexpr uint32_t idesc_s = make_idesc(2 * BM, BNS);
      const uint32_t smem_base = smem_u32(smem);
      for (int i = 0; i < n; ++i) {
        const int st = i % KS, sub = i & 1;
        if (i == 0) mbar_wait_cl(q_full, 0);
        mbar_wait_cl(&k_full[st], (i / KS) & 1);
        if (i >= 2) mbar_wait_cl(&s_free[sub], ((i >> 1) - 1) & 1);
        fence_after();
        if (elect_one()) {
          const uint64_t dq = make_desc_sw128(smem_base + OFF_Q);
          const uint64_t dk = make_desc_sw128(smem_base + OFF_K + st * SMEM_K);
          const uint32_t d = tmem + TMEM_S + sub * BNS;
#pragma unroll
          for (int kk = 0; kk < DK / 16; ++kk) {
            const uint64_t oa = (uint64_t)(((kk >> 2) * (BM * 128) + (kk & 3) * 32) >> 4);
            const uint64_t ob = (uint64_t)(((kk >> 2) * ((BNS / 2) * 128) + (kk & 3) * 32) >> 4);
            umma2_ss(d, dq + oa, dk + ob, idesc_s, kk > 0);
          }
          commit_pair(&k_empty[st]);
          commit_pair(&s_full[i % 6]);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================ O += P.V issuer (leader): P from TMEM in pass 0, from shared memory after ================================
    if (n > 0 && leader) {
      constexpr uint32_t idesc_o = make_idesc(2 * BM, DVC);
      const uint32_t smem_base = smem_u32(smem);
      int xv = 0;                                           // V ring position over all passes
      for (int j = 0; j < n; ++j, ++xv) {
        const int b = j & 3, sv = xv % VS;
        mbar_wait_cl(&v_full[sv], (xv / VS) & 1);
        mbar_wait_cl(&p_full[b], (j >> 2) & 1);
        fence_after();
        if (elect_one()) {
          const uint64_t dv = make_desc_sw128(smem_base + OFF_V + sv * SMEM_V);
          const uint32_t pa = tmem + TMEM_P + b * 32;
#pragma unroll
          for (int kk = 0; kk < BNS / 16; ++kk)
            umma2_ts(tmem + TMEM_O, pa + kk * 8, dv + (uint64_t)(kk * 2), idesc_o, (j == 0 && kk == 0) ? 0u : 1u);
          commit_pair(&v_empty[sv]);
          commit_pair(&sp_free[b]);
          if (j == n - 1) commit_pair(&pass_done[0]);
        }
        __syncwarp();
      }
      int xp = 0;
      for (int c = 1; c < p.n_dv; ++c) {
        const uint32_t od = tmem + ((c & 1) ? TMEM_S : TMEM_O);     // O_b = the 256 columns S and P occupied in pass 0
        if (c >= 2) mbar_wait_cl(&o_drained[c - 2], 0);
        for (int j = 0; j < n; ++j, ++xv, ++xp) {
          const int sv = xv % VS, sp = xp % PS;
          mbar_wait_cl(&v_full[sv], (xv / VS) & 1);
          mbar_wait_cl(&pk_full[sp], (xp / PS) & 1);
          fence_after();
          if (elect_one()) {
            const uint64_t dv = make_desc_sw128(smem_base + OFF_V + sv * SMEM_V);
            // P tile = [8 chunks][128 rows][16 B]: core matrices 2048 B apart along K, 128 B apart along M; a k-step of 16
            // keys = two chunks = 4096 B
            const uint64_t dp = make_desc_noswz(smem_base + OFF_Q + sp * (BM * BNS * 2), kLboP, kSboP);
#pragma unroll
            for (int kk = 0; kk < BNS / 16; ++kk)
              umma2_ss(od, dp + (uint64_t)(kk * (4096 >> 4)), dv + (uint64_t)(kk * 2), idesc_o, (j == 0 && kk == 0) ? 0u : 1u);
            commit_pair(&v_empty[sv]);
            commit_pair(&pk_empty[sp]);
            if (j == n - 1) commit_pair(&pass_done[c]);
          }
          __syncwarp();
        }
      }
    }
  }
  } else {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsSoftmax));
  if (n > 0) {
    // ================================ softmax (pass 0) + read-out of every pass (warps 0-11, both CTAs) ================================
    const int quad = warp & 3, grp = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    const int qt = qp * 2 + (int)rank;
    const int qi = qt * BM + row;
    const bool row_ok = qi < p.HW;
    // fixed reference; rows past HW get a reference no score reaches (their probabilities become 0)
    const float m_ref = row_ok ? p.mseed[qi] + kRefShift : 3.0e38f;
    float l_tot = 0.f, l_piece = 0.f, bias2 = 0.f;
    bool over = false;
    int cur_t = -1;
    float* piece_base = p.pieces ? p.pieces + (((long long)cta * p.T) * kGroups + grp) * (BM * 2) + row * 2 : nullptr;
    auto flush_piece = [&](int t) {
      if (piece_base) {
        float* d = piece_base + (long long)t * (kGroups * BM * 2);
        d[0] = m_ref;
        d[1] = l_piece;
      }
    };
    t16* prow = p.pbuf + (long long)cta * p.max_sub * (BM * BNS) + row * 8;     // + j * tile + chunk * (BM * 8) elements
    int t = lo / p.tpf, jt = lo - t * p.tpf;
    for (int j = 0; j < n; ++j, ++jt) {
      if (jt == p.tpf) { jt = 0; ++t; }
      if (t != cur_t) {
        if (cur_t >= 0) flush_piece(cur_t);
        cur_t = t;
        l_piece = 0.f;
        bias2 = (p.qbias && row_ok) ? p.qbias[(long long)qi * p.T + t] * LOG2E : 0.f;
      }
      if (j % kGroups != grp) continue;
      const int sub = j & 1, b = j & 3;
      mbar_wait_b(&s_full[j % 6], (j / 6) & 1);
      fence_after();
      float sc[64];
      {
        uint32_t r0[32], r1[32];
        tmem_ld32_nowait(lane_addr + TMEM_S + sub * BNS, r0);
        tmem_ld32_nowait(lane_addr + TMEM_S + sub * BNS + 32, r1);
        if (j >= 4) {
          mbar_wait_b(&sp_free[b], ((j >> 2) - 1) & 1);
          fence_after();
        }
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) { sc[c] = __uint_as_float(r0[c]); sc[32 + c] = __uint_as_float(r1[c]); }
      }
      fence_before();
      __syncwarp();
      if (lane == 0) { if (leader) mbar_arrive(&s_free[sub]); else mbar_arrive_leader(&s_free[sub]); }
      const int key0 = jt * BNS;
      if (key0 + BNS > p.HW) {
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
      const float c0 = bias2 - m_ref;
      over = over || (row_ok && fmaf(mx, p.scale_log2, c0) > kOverflowAt);
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
      l_tot += lsum;
      l_piece += lsum;
      tmem_st32u(lane_addr + TMEM_P + b * 32, pk);
      fence_before();
      __syncwarp();
      if (lane == 0) { if (leader) mbar_arrive(&p_full[b]); else mbar_arrive_leader(&p_full[b]); }
      // the same probabilities for passes 1-3.  Scratch tile layout = [8 key chunks][128 rows][16 B] (the UMMA no-swizzle
      // K-major core-matrix order): store c of a warp covers 32 consecutive rows = 512 contiguous bytes.  (Row-major rows
      // -- 32 scattered 16-byte pieces per store instruction -- doubled the pass-0 time per sub-tile: 2200 cycles.)
      {
        uint4* dst = reinterpret_cast<uint4*>(prow + (long long)j * (BM * BNS));
#pragma unroll
        for (int c = 0; c < 8; ++c) dst[c * BM] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      }
    }
    flush_piece(cur_t);
    if (__any_sync(0xffffffffu, over) && lane == 0) atomicExch(p.overflow, 1);
    // the stores above become visible to the TMA (async proxy) loads of this CTA's producer warp
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(p_ready);
    if (cta_times && warp == 0 && lane == 0) cta_times[13] = clock64();      // pass-0 softmax done

    // ---- row sums of the three groups -> 1 / l (the same for every pass) ----
    lx[grp * BM + row] = l_tot;
    named_bar_sync(13, kSoftmaxWarps * 32);
    const float l_row = (lx[row] + lx[BM + row]) + lx[2 * BM + row];
    const float inv = l_row > 0.f ? 1.f / l_row : 0.f;
    if (grp == 0) {
      float* ml = p.part_ml + ((long long)cta * BM + row) * 2;
      ml[0] = m_ref;
      ml[1] = row_ok ? l_row : 0.f;
    }
    // ---- read-out of pass c while pass c + 1 accumulates into the other half of TMEM ----
    const int c_lo = grp * 96, c_hi = grp == 2 ? DVC : c_lo + 96;
    for (int c = 0; c < p.n_dv; ++c) {
      mbar_wait_b(&pass_done[c], 0);
      fence_after();
      if (cta_times && warp == 0 && lane == 0) cta_times[9 + c] = clock64();   // pass c complete
      const uint32_t ob = lane_addr + ((c & 1) ? TMEM_S : TMEM_O);
      t16* po = p.part_o + (((long long)(cta * p.n_dv + c) * (DVC / 16)) * BM + row) * 16;
#pragma unroll 1
      for (int cc = c_lo; cc < c_hi; cc += 32) {
        float o[32];
        tmem_ld32(ob + cc, o);
        if (row_ok) {
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            uint4 u;
            u.x = pack2(o[e] * inv, o[e + 1] * inv);
            u.y = pack2(o[e + 2] * inv, o[e + 3] * inv);
            u.z = pack2(o[e + 4] * inv, o[e + 5] * inv);
            u.w = pack2(o[e + 6] * inv, o[e + 7] * inv);
            *reinterpret_cast<uint4*>(po + (long long)((cc + e) >> 4) * (BM * 16) + ((cc + e) & 8)) = u;
          }
        }
      }
      fence_before();
      __syncwarp();
      if (lane == 0 && c + 2 < p.n_dv) mbar_arrive_leader(&o_drained[c]);
    }
  }
  }
  fence_before();
  __syncthreads();
  if (cta_times && threadIdx.x == 0) cta_times[14] = clock64();               // last read-out stored
  cluster_sync_all();
  if (cta_times && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    cta_times[1] = (long long)gt; cta_times[2] = n; cta_times[6] = clock64();
  }
  if (warp == kWarpMmaS) {
    fence_after();
    tmem_dealloc_pair(tmem);
  }
}

// Seed of the running row maximum: the best score of query i against the keys at its own position and the 8 neighbours
// in every bank frame (T x 9 dot products of 128) -- a LOWER bound of the true maximum that is close to it whenever the
// match is local or the score distribution is broad.  One warp per query; eight lanes share a key (16 channels = two
// 16-byte loads each, 3 shuffles to reduce), so a warp scores four keys per step.
__global__ void __launch_bounds__(256) attn_seed_kernel(const t16* __restrict__ qt, const float* __restrict__ qbias,
                                                        const t16* __restrict__ kbank, Tc3Params p, int h, int w,
                                                        float* __restrict__ mseed, int* __restrict__ zero_flag) {
  pdl_prologue();
  if (zero_flag && blockIdx.x == 0 && threadIdx.x == 0) *zero_flag = 0;   // arms the column kernel's overflow flag
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= p.HW) return;
  const int sub = lane >> 3, part = lane & 7;            // key within the step, 16-channel block
  float q[16];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(qt + (long long)i * DK + part * 16);
    const uint4 a = qp[0], b = qp[1];
    const uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int c = 0; c < 8; ++c) { const float2 f = unpack2(u[c]); q[2 * c] = f.x; q[2 * c + 1] = f.y; }
  }
  const int y = i / w, x = i - y * w;
  float best = -INFINITY;
  const int n = p.T * 9;
  for (int k0 = 0; k0 < n; k0 += 4) {
    const int k = k0 + sub;
    const int t = k / 9, nb = k - t * 9;
    const int yy = y + nb / 3 - 1, xx = x + nb % 3 - 1;
    const bool ok = k < n && yy >= 0 && yy < h && xx >= 0 && xx < w;
    float s = 0.f;
    if (ok) {
      const uint4* kp = reinterpret_cast<const uint4*>(kbank + ((long long)p.slot[t] * p.HWp + yy * w + xx) * DK + part * 16);
      const uint4 a = kp[0], b = kp[1];
      const uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int c = 0; c < 8; ++c) { const float2 f = unpack2(u[c]); s = fmaf(q[2 * c], f.x, s); s = fmaf(q[2 * c + 1], f.y, s); }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (ok) {
      const float b2 = qbias ? qbias[(long long)i * p.T + t] * LOG2E : 0.f;
      best = fmaxf(best, fmaf(s, p.scale_log2, b2));
    }
  }
  best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 8));
  best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 16));
  if (lane == 0) mseed[i] = best;
}

// qprep + seed in one launch (the engine's long-term path): Qt = t16(Q + cur_pos_emb), qbias[i,t] = scale <Qt_i, pe_mem[slot_T(t)]>
// (transformer.py:1140-1175 as a score bias, see ops.cu qprep_kernel) and the row-maximum seed of attn_seed_kernel above.
// One warp per query; eight lanes share a 128-channel row (16 channels each); the four 8-lane groups take the bank frames
// round-robin and issue the loads of all nine neighbour keys of a frame before the first dot product, so a query costs
// ceil(T / 4) L2 round trips instead of T * 9 / 4 (attn_seed_kernel: 16 us at T = 8).
struct PeSlots3 { int s[kMaxBankFrames]; };
__global__ void __launch_bounds__(256) qprep_seed_kernel(const t16* __restrict__ q, long long ldq,
                                                         const float* __restrict__ pe_cur, const float* __restrict__ pe_mem,
                                                         PeSlots3 ps, float scale, const t16* __restrict__ kbank, Tc3Params p,
                                                         int h, int w, t16* __restrict__ qt, float* __restrict__ qbias,
                                                         float* __restrict__ mseed, int* __restrict__ zero_flag) {
  pdl_prologue();
  if (zero_flag && blockIdx.x == 0 && threadIdx.x == 0) *zero_flag = 0;   // arms the column kernel's overflow flag
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= p.HW) return;
  const int sub = lane >> 3, part = lane & 7;
  float qv[16];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(q + (long long)i * ldq + part * 16);
    const uint4 a = qp[0], b = qp[1];
    const uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t r[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float2 f = unpack2(u[c]);
      const float x0 = f.x + (pe_cur ? pe_cur[part * 16 + 2 * c] : 0.f);
      const float x1 = f.y + (pe_cur ? pe_cur[part * 16 + 2 * c + 1] : 0.f);
      r[c] = pack2(x0, x1);
      const float2 g = unpack2(r[c]);
      qv[2 * c] = g.x; qv[2 * c + 1] = g.y;
    }
    if (sub == 0) {
      uint4* op = reinterpret_cast<uint4*>(qt + (long long)i * DK + part * 16);
      op[0] = make_uint4(r[0], r[1], r[2], r[3]);
      op[1] = make_uint4(r[4], r[5], r[6], r[7]);
    }
  }
  const int y = i / w, x = i - y * w;
  float best = -INFINITY;
  for (int r0 = 0; r0 < p.T; r0 += 4) {                   // uniform trip count: the shuffles below need the whole warp
    const int t = r0 + sub;
    const bool live = t < p.T;
    float bs = 0.f;
    if (live) {
      const float4* pm = reinterpret_cast<const float4*>(pe_mem + (long long)ps.s[t] * DK + part * 16);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 e = pm[c];
        bs = fmaf(qv[4 * c], e.x, bs); bs = fmaf(qv[4 * c + 1], e.y, bs);
        bs = fmaf(qv[4 * c + 2], e.z, bs); bs = fmaf(qv[4 * c + 3], e.w, bs);
      }
    }
    bs += __shfl_xor_sync(0xffffffffu, bs, 1);
    bs += __shfl_xor_sync(0xffffffffu, bs, 2);
    bs += __shfl_xor_sync(0xffffffffu, bs, 4);
    const float bias = bs * scale;
    if (live && part == 0) qbias[(long long)i * p.T + t] = bias;
    uint4 ka[9], kb[9];
    bool ok[9];
    const t16* kf = kbank + ((long long)(live ? p.slot[t] : 0) * p.HWp) * DK + part * 16;
#pragma unroll
    for (int nb = 0; nb < 9; ++nb) {
      const int yy = y + nb / 3 - 1, xx = x + nb % 3 - 1;
      ok[nb] = live && yy >= 0 && yy < h && xx >= 0 && xx < w;
      if (ok[nb]) {
        const uint4* kp = reinterpret_cast<const uint4*>(kf + (long long)(yy * w + xx) * DK);
        ka[nb] = kp[0];
        kb[nb] = kp[1];
      }
    }
#pragma unroll
    for (int nb = 0; nb < 9; ++nb) {
      float sc = 0.f;
      if (ok[nb]) {
        const uint32_t u[8] = {ka[nb].x, ka[nb].y, ka[nb].z, ka[nb].w, kb[nb].x, kb[nb].y, kb[nb].z, kb[nb].w};
#pragma unroll
        for (int c = 0; c < 8; ++c) { const float2 f = unpack2(u[c]); sc = fmaf(qv[2 * c], f.x, sc); sc = fmaf(qv[2 * c + 1], f.y, sc); }
      }
      sc += __shfl_xor_sync(0xffffffffu, sc, 1);
      sc += __shfl_xor_sync(0xffffffffu, sc, 2);
      sc += __shfl_xor_sync(0xffffffffu, sc, 4);
      if (ok[nb]) best = fmaxf(best, fmaf(sc, p.scale_log2, bias * LOG2E));
    }
  }
  best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 8));
  best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 16));
  if (lane == 0) mseed[i] = best;
}

int sm_count3() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// Static schedule in 64-key sub-tiles (see attn_tc2.cu make_bounds): equalise sub-tiles + kSegCost * segments over the
// clusters.  Measured on the per-CTA wall times of profiles/r02_attn3_trace_3groups.txt: a cluster with 40 groups and
// two segments ends 4 us BEFORE one with 43 groups and one segment, i.e. a segment epilogue is nearly free here (three
// softmax groups share the TMEM read-out and the score MMAs of the next segment run meanwhile): two sub-tiles.
constexpr int kSegCost = 2, kRestartCost = 0;
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
  if (uni < 16) {                                        // short launches keep the uniform cut
    for (int c = 0; c <= n; ++c) b[c] = (int)((L * c) / n);
    return;
  }
  long long lo = uni, hi = uni + 2 * kSegCost + kRestartCost + 2;
  while (lo < hi) {
    const long long mid = (lo + hi) / 2;
    if (greedy_bounds(L, TPU, n, mid, n, b)) hi = mid; else lo = mid + 1;
  }
  int mlo = 0, mhi = n;
  while (mlo < mhi) {
    const int mid = (mlo + mhi) / 2;
    if (greedy_bounds(L, TPU, n, lo, mid, b)) mhi = mid; else mlo = mid + 1;
  }
  greedy_bounds(L, TPU, n, lo, mlo, b);
}

void schedule3(int HW, int T, int Dv, int* n_units, int* tpf, int* TPU, int* nCL) {
  const int qpairs = cdiv(cdiv(HW, BM), 2), n_dv = Dv / DVC;
  *n_units = qpairs * n_dv;
  *tpf = cdiv(HW, BNS);
  *TPU = T * *tpf;
  const long long L = (long long)*n_units * *TPU;
  int n = sm_count3() / 2;
  if (n > kMaxCL) n = kMaxCL;
  if ((long long)n > L / 4) n = (int)(L / 4);                   // at least ~4 sub-tiles per cluster
  const int cap = (kMaxSegsPerUnit - 2) * *n_units;             // combine3 resolves <= kMaxSegsPerUnit segments per unit
  if (n > cap) n = cap;
  // Short launches (the T = 1 self-attention: ~10 sub-tiles per cluster): a cluster that straddles two units pays a
  // pipeline restart for a handful of tiles -- k whole-segment clusters per unit instead.
  if (L / n < 16 && n / *n_units >= 2) n = (n / *n_units) * *n_units;
  if (n < *n_units) n = *n_units;                               // a cluster never spans more than two units
  if ((long long)n > L) n = (int)L;
  if (n < 1) n = 1;
  *nCL = n;
}

size_t part_bytes3(int nCL, int T, size_t* off_ml, size_t* off_pieces, size_t* off_seed, int HW) {
  const size_t nCTA = (size_t)nCL * 2;
  size_t o = nCTA * 2 * BM * DVC * sizeof(t16);
  o = (o + 255) & ~size_t(255);
  *off_ml = o;
  o += nCTA * 2 * BM * 2 * sizeof(float);
  o = (o + 255) & ~size_t(255);
  *off_pieces = o;
  o += nCTA * 2 * T * kGroups * BM * 2 * sizeof(float);
  o = (o + 255) & ~size_t(255);
  *off_seed = o;
  o += (size_t)HW * sizeof(float);
  return o + 256;
}

}  // namespace

static thread_local cudaEvent_t g3_ev0 = nullptr, g3_ev1 = nullptr;
void long_attn_tc3_set_events(void* ev0, void* ev1) {
  g3_ev0 = reinterpret_cast<cudaEvent_t>(ev0);
  g3_ev1 = reinterpret_cast<cudaEvent_t>(ev1);
}
static thread_local int* g3_rescales = nullptr;
void long_attn_tc3_set_rescale_counter(int* dev_counter) { g3_rescales = dev_counter; }

int long_attn_tc3_set_trace(long long* dev_buf) {
  static int cl = -1;
  if (cl < 0) {
    const char* e = getenv("RMEM_TRACE_CTA");
    cl = e ? atoi(e) : 0;
    RMEM_CUDA_CHECK(cudaMemcpyToSymbol(g_trace3_cl, &cl, sizeof(cl)));
  }
  RMEM_CUDA_CHECK(cudaMemcpyToSymbol(g_trace3, &dev_buf, sizeof(dev_buf)));
  return RMEM_OK;
}

// Host-only view of the static schedule (tests): step range of every cluster for a launch of this shape.
int long_attn_tc3_schedule(int HW, int T, int Dv, int* n_units, int* groups_per_unit, int* n_clusters, int* bounds,
                           int cap) {
  int tpf = 0;
  schedule3(HW, T, Dv, n_units, &tpf, groups_per_unit, n_clusters);
  RMEM_REQUIRE(*n_clusters + 1 <= cap && *n_clusters <= kMaxCL, "schedule: %d clusters do not fit the caller's table (%d)",
               *n_clusters, cap);
  make_bounds((long long)*n_units * *groups_per_unit, *groups_per_unit, *n_clusters, bounds);
  return RMEM_OK;
}

size_t long_attn_tc3_workspace(int HW, int HWp, int nslots, int Dv) {
  (void)HWp;
  size_t best = 0;
  for (int T = 1; T <= nslots && T <= kMaxBankFrames; ++T) {
    int n_units, tpf, TPU, nCL;
    schedule3(HW, T, Dv, &n_units, &tpf, &TPU, &nCL);
    size_t a, b, c;
    const size_t n = part_bytes3(nCL, T, &a, &b, &c, HW);
    if (n > best) best = n;
  }
  return best;
}

int qprep_seed_tc3(const t16* q, long long ldq, const float* pe_cur, const float* pe_mem, const int* pe_slot, int T,
                   float scale, const t16* kbank, const int* slot, int HW, int HWp, int h, int w, t16* qt, float* qbias,
                   float* mseed, cudaStream_t s, int* zero_flag) {
  RMEM_REQUIRE(T >= 1 && T <= kMaxBankFrames && h * w == HW, "qprep_seed: T=%d grid %dx%d HW=%d", T, h, w, HW);
  RMEM_REQUIRE(ldq % 8 == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(qt) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(pe_mem) & 15) == 0, "qprep_seed: alignment");
  Tc3Params p = {};
  p.HW = HW; p.HWp = HWp; p.T = T;
  p.scale_log2 = scale * LOG2E;
  PeSlots3 ps;
  for (int t = 0; t < kMaxBankFrames; ++t) { ps.s[t] = t < T ? pe_slot[t] : 0; p.slot[t] = t < T ? slot[t] : 0; }
  RMEM_CUDA_CHECK(launch_pdl(qprep_seed_kernel, dim3(cdiv(HW, 8)), dim3(256), 0, s, q, ldq, pe_cur, pe_mem, ps, scale, kbank,
                             p, h, w, qt, qbias, mseed, zero_flag));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

namespace {
int tc3_launch(const LongAttnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s, const int* run_if,
               int* zero_flag, bool do_combine, Tc3Params* p_out);
}
int long_attn_tc3(const LongAttnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  return tc3_launch(a, workspace, workspace_bytes, s, nullptr, nullptr, true, nullptr);
}

namespace {
int tc3_launch(const LongAttnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s, const int* run_if,
               int* zero_flag, bool do_combine, Tc3Params* p_out) {
  RMEM_REQUIRE(a.Dk == DK, "long_attn_tc3: Dk=%d (built for 128)", a.Dk);
  RMEM_REQUIRE(a.Dv % DVC == 0 && a.Dv <= 1024, "long_attn_tc3: Dv=%d must be a multiple of 256, <= 1024", a.Dv);
  RMEM_REQUIRE(a.HWp % BNS == 0 && a.HWp >= a.HW, "long_attn_tc3: HWp=%d must be a multiple of 64", a.HWp);
  RMEM_REQUIRE(a.T >= 1 && a.T <= kMaxBankFrames && a.T <= a.nslots, "long_attn_tc3: T=%d nslots=%d", a.T, a.nslots);
  RMEM_REQUIRE(a.ldo % 8 == 0 && (!a.gate || a.ldg % 8 == 0), "long_attn_tc3: ldo/ldg alignment");
  RMEM_REQUIRE(a.seed_h * a.seed_w == a.HW || a.seed_h == 0, "long_attn_tc3: seed grid %dx%d != HW=%d", a.seed_h,
               a.seed_w, a.HW);
  Tc3Params p = {};
  p.run_if = run_if;
  p.HW = a.HW; p.HWp = a.HWp; p.T = a.T; p.Dv = a.Dv; p.n_dv = a.Dv / DVC;
  schedule3(a.HW, a.T, a.Dv, &p.n_units, &p.tpf, &p.TPU, &p.nCL);
  RMEM_REQUIRE(p.nCL <= kMaxCL, "long_attn_tc3: %d clusters > %d", p.nCL, kMaxCL);
  make_bounds((long long)p.n_units * p.TPU, p.TPU, p.nCL, p.bounds);
  size_t off_ml, off_pieces, off_seed;
  const size_t need = part_bytes3(p.nCL, a.T, &off_ml, &off_pieces, &off_seed, a.HW);
  RMEM_REQUIRE(workspace_bytes >= need, "long_attn_tc3: workspace %zu < %zu", workspace_bytes, need);
  RMEM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "long_attn_tc3: workspace alignment");
  for (int t = 0; t < kMaxBankFrames; ++t) p.slot[t] = t < a.T ? a.slot[t] : 0;
  p.scale_log2 = a.scale * LOG2E;
  p.qbias = a.qbias;
  char* ws = reinterpret_cast<char*>(workspace);
  p.part_o = reinterpret_cast<t16*>(ws);
  p.part_ml = reinterpret_cast<float*>(ws + off_ml);
  p.pieces = a.mass ? reinterpret_cast<float*>(ws + off_pieces) : nullptr;
  float* mseed = reinterpret_cast<float*>(ws + off_seed);
  p.mseed = a.mseed ? a.mseed : (a.seed_h > 0 ? mseed : nullptr);   // a.mseed: the caller ran qprep_seed_tc3
  p.rescales = g3_rescales;

  RMEM_REQUIRE((reinterpret_cast<uintptr_t>(a.qt) & 15) == 0, "long_attn_tc3: q alignment");
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
    uint32_t box[2] = {64, (uint32_t)(BNS / 2)};
    RMEM_TRY(tma_encode_cached(&mk, a.kbank, 2, dims, str, box, nullptr));
  }
  {
    uint64_t dims[2] = {(uint64_t)a.nslots * a.HWp, (uint64_t)a.Dv};
    uint64_t str[1] = {(uint64_t)a.nslots * a.HWp * 2};
    uint32_t box[2] = {(uint32_t)BNS, (uint32_t)(DVC / 2)};
    RMEM_TRY(tma_encode_cached(&mv, a.vtbank, 2, dims, str, box, nullptr));
  }
  static bool attr_done = false;
  if (!attr_done) {
    RMEM_CUDA_CHECK(cudaFuncSetAttribute(long_attn_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    attr_done = true;
  }
  if (p.mseed && !a.mseed) {
    RMEM_CUDA_CHECK(launch_pdl(attn_seed_kernel, dim3(cdiv(a.HW, 8)), dim3(256), 0, s, a.qt, a.qbias, a.kbank, p,
                               a.seed_h, a.seed_w, mseed, zero_flag));
    RMEM_LAUNCH_CHECK();
  }
  if (g3_ev0 && !run_if) RMEM_CUDA_CHECK(cudaEventRecord(g3_ev0, s));
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.nCL * 2);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = SMEM_TOTAL;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    RMEM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, long_attn_tc3_kernel, *mq, *mk, *mv, static_cast<const Tc3Params&>(p)));
  }
  if (g3_ev1 && !run_if) RMEM_CUDA_CHECK(cudaEventRecord(g3_ev1, s));
  RMEM_LAUNCH_CHECK();
  if (p_out) *p_out = p;
  if (!do_combine) return RMEM_OK;
  RMEM_CUDA_CHECK(launch_pdl(combine3_kernel, dim3(cdiv(a.HW, kCombRows)), dim3(256), 0, s, p, a.gate, a.ldg, a.out,
                             a.ldo, a.mass));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

// ---- column kernel: schedule, workspace, launch ----
struct Sched4 { int n_qp, tpf, TPU, nCL, max_sub; };
bool schedule4(int HW, int T, Sched4* sc, int* bounds) {
  sc->n_qp = cdiv(cdiv(HW, BM), 2);
  sc->tpf = cdiv(HW, BNS);
  sc->TPU = T * sc->tpf;
  int n = sm_count3() / 2;
  if (n > kMaxCL) n = kMaxCL;
  // clusters per query pair: at most kMaxSegsPerUnit - 2 segments for the combine, and never more than sub-tiles
  const int per = sc->TPU < kMaxSegsPerUnit - 2 ? sc->TPU : kMaxSegsPerUnit - 2;
  if (n > per * sc->n_qp) n = per * sc->n_qp;
  if (n < sc->n_qp) return false;                               // more query pairs than clusters: not this kernel's case
  sc->nCL = n;
  const int base = n / sc->n_qp, extra = n % sc->n_qp;
  sc->max_sub = cdiv(sc->TPU, base);
  if (bounds) {
    int c = 0;
    for (int qp = 0; qp < sc->n_qp; ++qp) {
      const int cnt = base + (qp < extra ? 1 : 0);
      for (int k = 0; k < cnt; ++k) bounds[c++] = qp * sc->TPU + (int)((long long)sc->TPU * k / cnt);
    }
    bounds[c] = sc->n_qp * sc->TPU;
  }
  return true;
}
constexpr size_t kFlagBytes = 256;
size_t part_bytes4(const Sched4& sc, int T, int n_dv, size_t* off_ml, size_t* off_pieces, size_t* off_pbuf) {
  const size_t nCTA = (size_t)sc.nCL * 2;
  size_t o = nCTA * n_dv * BM * DVC * sizeof(t16);
  o = (o + 255) & ~size_t(255);
  *off_ml = o;
  o += nCTA * BM * 2 * sizeof(float);
  o = (o + 255) & ~size_t(255);
  *off_pieces = o;
  o += nCTA * T * kGroups * BM * 2 * sizeof(float);
  o = (o + 1023) & ~size_t(1023);
  *off_pbuf = o;
  o += nCTA * sc.max_sub * BM * BNS * sizeof(t16);
  return o + 256;
}
bool tc4_switch() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("RMEM_ATTN_TC4"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}
constexpr int kMinT4 = 3;      // below this a column is a handful of sub-tiles: the four pass transitions cost more than the repeats
}  // namespace

size_t long_attn_tc4_workspace(int HW, int HWp, int nslots, int Dv) {
  const size_t w3 = (long_attn_tc3_workspace(HW, HWp, nslots, Dv) + 1023) & ~size_t(1023);
  size_t best = 0;
  for (int T = kMinT4; T <= nslots && T <= kMaxBankFrames; ++T) {
    Sched4 sc;
    if (!schedule4(HW, T, &sc, nullptr)) continue;
    size_t a, b, c;
    const size_t n = part_bytes4(sc, T, Dv / DVC, &a, &b, &c);
    if (n > best) best = n;
  }
  return kFlagBytes + w3 + best;
}

// Workspace: [overflow flag, 256 B][tc3 workspace (fallback)][column partials | (m, l) | pieces | P scratch].
// Needs the seed (a.mseed, from qprep_seed_tc3 with zero_flag = the workspace's first int, or the token grid): anything else
// -- unseeded calls, short banks, more query pairs than clusters -- is the pair kernel's case.
int long_attn_tc4(const LongAttnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  RMEM_REQUIRE(workspace_bytes >= long_attn_tc4_workspace(a.HW, a.HWp, a.nslots, a.Dv), "long_attn_tc4: workspace too small");
  RMEM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "long_attn_tc4: workspace must be 256-byte aligned");
  char* ws = reinterpret_cast<char*>(workspace);
  int* flag = reinterpret_cast<int*>(ws);
  char* ws3 = ws + kFlagBytes;
  const size_t w3 = (long_attn_tc3_workspace(a.HW, a.HWp, a.nslots, a.Dv) + 1023) & ~size_t(1023);
  Sched4 sc;
  const bool seeded = a.mseed != nullptr || a.seed_h > 0;
  Tc3Params p = {};
  if (!tc4_switch() || !seeded || a.T < kMinT4 || a.Dk != DK || a.Dv % DVC != 0 || a.Dv > 1024 ||
      !schedule4(a.HW, a.T, &sc, p.bounds))
    return tc3_launch(a, ws3, w3, s, nullptr, nullptr, true, nullptr);
  RMEM_REQUIRE(a.HWp % BNS == 0 && a.HWp >= a.HW, "long_attn_tc4: HWp=%d must be a multiple of 64", a.HWp);
  RMEM_REQUIRE(a.T <= kMaxBankFrames && a.T <= a.nslots, "long_attn_tc4: T=%d nslots=%d", a.T, a.nslots);
  RMEM_REQUIRE(a.ldo % 8 == 0 && (!a.gate || a.ldg % 8 == 0), "long_attn_tc4: ldo/ldg alignment");
  RMEM_REQUIRE((reinterpret_cast<uintptr_t>(a.qt) & 15) == 0, "long_attn_tc4: q alignment");
  p.HW = a.HW; p.HWp = a.HWp; p.T = a.T; p.Dv = a.Dv; p.n_dv = a.Dv / DVC;
  p.tpf = sc.tpf; p.TPU = sc.TPU; p.n_units = sc.n_qp; p.nCL = sc.nCL;
  p.column = 1; p.max_sub = sc.max_sub;
  for (int t = 0; t < kMaxBankFrames; ++t) p.slot[t] = t < a.T ? a.slot[t] : 0;
  p.scale_log2 = a.scale * LOG2E;
  p.qbias = a.qbias;
  size_t off_ml, off_pieces, off_pbuf;
  part_bytes4(sc, a.T, p.n_dv, &off_ml, &off_pieces, &off_pbuf);
  char* ws4 = ws3 + w3;
  p.part_o = reinterpret_cast<t16*>(ws4);
  p.part_ml = reinterpret_cast<float*>(ws4 + off_ml);
  p.pieces = a.mass ? reinterpret_cast<float*>(ws4 + off_pieces) : nullptr;
  p.pbuf = reinterpret_cast<t16*>(ws4 + off_pbuf);
  p.overflow = flag;

  // the seed: the caller's (qprep_seed_tc3 zeroed the flag) or one pass of attn_seed_kernel into the tc3 workspace's slot
  LongAttnArgs a3 = a;
  if (!a.mseed) {
    size_t o_ml, o_pc, o_seed;
    int nu, tpf3, TPU3, nCL3;
    schedule3(a.HW, a.T, a.Dv, &nu, &tpf3, &TPU3, &nCL3);
    part_bytes3(nCL3, a.T, &o_ml, &o_pc, &o_seed, a.HW);
    float* mseed = reinterpret_cast<float*>(ws3 + o_seed);
    RMEM_CUDA_CHECK(launch_pdl(attn_seed_kernel, dim3(cdiv(a.HW, 8)), dim3(256), 0, s, a.qt, a.qbias, a.kbank, p, a.seed_h,
                               a.seed_w, mseed, flag));
    RMEM_LAUNCH_CHECK();
    a3.mseed = mseed;
  }
  p.mseed = a3.mseed;

  const CUtensorMap *mq, *mk, *mv, *mp;
  {
    uint64_t dims[2] = {(uint64_t)DK, (uint64_t)a.HW};
    uint64_t str[1] = {(uint64_t)DK * 2};
    uint32_t box[2] = {64, (uint32_t)BM};
    RMEM_TRY(tma_encode_cached(&mq, a.qt, 2, dims, str, box, nullptr));
  }
  {
    uint64_t dims[2] = {(uint64_t)DK, (uint64_t)a.nslots * a.HWp};
    uint64_t str[1] = {(uint64_t)DK * 2};
    uint32_t box[2] = {64, (uint32_t)(BNS / 2)};
    RMEM_TRY(tma_encode_cached(&mk, a.kbank, 2, dims, str, box, nullptr));
  }
  {
    uint64_t dims[2] = {(uint64_t)a.nslots * a.HWp, (uint64_t)a.Dv};
    uint64_t str[1] = {(uint64_t)a.nslots * a.HWp * 2};
    uint32_t box[2] = {(uint32_t)BNS, (uint32_t)(DVC / 2)};
    RMEM_TRY(tma_encode_cached(&mv, a.vtbank, 2, dims, str, box, nullptr));
  }
  {
    // the scratch tiles are copied verbatim (no swizzle): 128 lines of 128 B per tile
    uint64_t dims[2] = {(uint64_t)BNS, (uint64_t)sc.nCL * 2 * sc.max_sub * BM};
    uint64_t str[1] = {(uint64_t)BNS * 2};
    uint32_t box[2] = {(uint32_t)BNS, (uint32_t)BM};
    RMEM_TRY(tma_encode_cached(&mp, p.pbuf, 2, dims, str, box, nullptr, /*swizzle128=*/0));
  }
  static bool attr_done = false;
  if (!attr_done) {
    RMEM_CUDA_CHECK(cudaFuncSetAttribute(long_attn_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL4));
    attr_done = true;
  }
  if (g3_ev0) RMEM_CUDA_CHECK(cudaEventRecord(g3_ev0, s));
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.nCL * 2);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = SMEM_TOTAL4;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    RMEM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, long_attn_tc4_kernel, *mq, *mk, *mv, *mp, static_cast<const Tc3Params&>(p)));
  }
  RMEM_LAUNCH_CHECK();
  // always-correct fallback queued right behind: runs only if the column kernel met a score outside its fp16 range
  Tc3Params p3;
  a3.seed_h = a3.seed_w = 0;                       // a3.mseed is set
  RMEM_TRY(tc3_launch(a3, ws3, w3, s, flag, nullptr, false, &p3));
  // measurement events bracket the column kernel AND the guarded fallback: what the op costs on these operands either way
  if (g3_ev1) RMEM_CUDA_CHECK(cudaEventRecord(g3_ev1, s));
  RMEM_CUDA_CHECK(launch_pdl(combine34_kernel, dim3(cdiv(a.HW, kCombRows)), dim3(256), 0, s, p, p3,
                             static_cast<const int*>(flag), a.gate, a.ldg, a.out, a.ldo, a.mass));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

}  // namespace rmem
