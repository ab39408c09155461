// Fused long-term / self attention for sm_100a:  tcgen05.mma (UMMA) with TMEM accumulators, operands staged
// by TMA (128B-swizzled, K-major on both sides), online softmax fused in registers, per-frame attention mass
// and the temporal-PE score bias for free.  Flash-decoding style split over memory frames so 148 SMs are busy
// with 14 query tiles (SURVEY.md section 7 "hard parts").
//
//   grid = (query tiles of 128) x (Dv chunks of 256) x (frame, sub-split)        block = 192 threads
//   warps 0-3  softmax + epilogue (thread <-> one query row <-> one TMEM lane)
//   warp  4    TMA producer (Q once; K tile [64 keys x 128] + V^T tile [256 x 64 keys] per stage, 3 stages)
//   warp  5    MMA issuer + TMEM owner:  S = Q.K^T (M128 N64 K128)  ->  TMEM cols 256+64b
//                                        O += P.V   (M128 N256 K64) ->  TMEM cols 0..255, P via swizzled smem
//   TMEM: 512 columns = O[256] | S0[64] | S1[64] | spare.
//   Each CTA writes un-normalised (O, m, l) partials; combine_kernel merges the splits, applies the gate and
//   emits mass[i,t] = sum_{splits of frame t} l_s 2^(m_s - M) / L.
#include "attn.cuh"
#include "tcgen05.cuh"

namespace rmem {

namespace {

using namespace tc;

constexpr int BM = 128;        // query rows per CTA
constexpr int BN = 64;         // keys per KV tile
constexpr int DK = 128;
constexpr int DVC = 256;       // Dv columns per CTA
constexpr int STAGES = 3;
constexpr int kThreads = 192;

constexpr int SMEM_Q = BM * DK * 2;            // 32 KB (two 64-col swizzle atoms)
constexpr int SMEM_K = BN * DK * 2;            // 16 KB per stage
constexpr int SMEM_V = DVC * BN * 2;           // 32 KB per stage
constexpr int SMEM_P = BM * BN * 2;            // 16 KB per buffer
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + SMEM_Q;
constexpr int OFF_V = OFF_K + STAGES * SMEM_K;
constexpr int OFF_P = OFF_V + STAGES * SMEM_V;
constexpr int OFF_BAR = OFF_P + 2 * SMEM_P;
constexpr int SMEM_TOTAL = OFF_BAR + 256 + 1024;   // + barriers + 1024B alignment slack

constexpr int TMEM_COLS = 512;
constexpr int TMEM_O = 0;
constexpr int TMEM_S = 256;

constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 8.0f;      // log2 units: P <= 2^8 before a lazy rescale is forced

struct TcParams {
  int HW, HWp, T, nsub, tiles_per_frame, Dv;
  int slot[kMaxBankFrames];
  float scale_log2;           // scale * log2(e)
  const float* qbias;         // [HW, T] or null (already multiplied by scale)
  float* part_o;              // [nsplit][HW][Dv] fp32, un-normalised
  float* part_ml;             // [nsplit][HW][2]  (m in log2 units, l)
  int* err;                   // device error flag (deadlock watchdog)
};

// ---------------------------------------------------------------------------------------------- kernel
__global__ void __launch_bounds__(kThreads, 1)
long_attn_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const TcParams p) {
  extern __shared__ unsigned char smem_raw[];
  // 128B swizzle needs 1024B-aligned tiles
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;               // [STAGES]
  uint64_t* kv_empty = bars + 1 + STAGES;     // [STAGES]
  uint64_t* s_full = bars + 1 + 2 * STAGES;   // [2]
  uint64_t* s_free = s_full + 2;              // [2]
  uint64_t* p_full = s_free + 2;              // [2]
  uint64_t* p_free = p_full + 2;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BM;
  const int dv0 = blockIdx.y * DVC;
  const int split = blockIdx.z;
  const int t = split / p.nsub, sub = split - t * p.nsub;
  // this CTA's tiles of frame t
  const int tile_lo = (int)(((long long)p.tiles_per_frame * sub) / p.nsub);
  const int tile_hi = (int)(((long long)p.tiles_per_frame * (sub + 1)) / p.nsub);
  const int n_tiles = tile_hi - tile_lo;
  const int key_base = p.slot[t] * p.HWp;     // column / row offset of the frame inside the bank

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 128);
      mbar_init(&p_full[i], 128); mbar_init(&p_free[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      mbar_expect_tx(q_full, SMEM_Q);
      tma_load_2d(smem + OFF_Q, &map_q, q_full, 0, q0);
      tma_load_2d(smem + OFF_Q + BM * 128, &map_q, q_full, 64, q0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % STAGES;
        if (j >= STAGES) mbar_wait(&kv_empty[st], ((j / STAGES) - 1) & 1, p.err, 1);
        const int key0 = key_base + (tile_lo + j) * BN;
        mbar_expect_tx(&kv_full[st], SMEM_K + SMEM_V);
        unsigned char* sk = smem + OFF_K + st * SMEM_K;
        tma_load_2d(sk, &map_k, &kv_full[st], 0, key0);
        tma_load_2d(sk + BN * 128, &map_k, &kv_full[st], 64, key0);
        tma_load_2d(smem + OFF_V + st * SMEM_V, &map_v, &kv_full[st], key0, dv0);
      }
    }
  } else if (warp == 5) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(BM, BN);
      constexpr uint32_t idesc_o = make_idesc(BM, DVC);
      const uint32_t q_addr = smem_u32(smem + OFF_Q);
      auto issue_s = [&](int j) {
        const int st = j % STAGES, b = j & 1;
        const uint32_t k_addr = smem_u32(smem + OFF_K + st * SMEM_K);
#pragma unroll
        for (int kk = 0; kk < DK / 16; ++kk) {
          const uint32_t a = q_addr + (kk >> 2) * (BM * 128) + (kk & 3) * 32;
          const uint32_t bb = k_addr + (kk >> 2) * (BN * 128) + (kk & 3) * 32;
          umma_ss(tmem + TMEM_S + b * BN, make_desc_sw128(a), make_desc_sw128(bb), idesc_s, kk > 0);
        }
        commit(&s_full[b]);
      };
      mbar_wait(q_full, 0, p.err, 2);
      mbar_wait(&kv_full[0], 0, p.err, 3);
      fence_after();
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) {
          const int jn = j + 1;
          mbar_wait(&kv_full[jn % STAGES], (jn / STAGES) & 1, p.err, 4);
          if (jn >= 2) mbar_wait(&s_free[jn & 1], ((jn - 2) >> 1) & 1, p.err, 5);
          fence_after();
          issue_s(jn);
        }
        const int st = j % STAGES, b = j & 1;
        mbar_wait(&p_full[b], (j >> 1) & 1, p.err, 6);
        fence_after();
        const uint32_t p_addr = smem_u32(smem + OFF_P + b * SMEM_P);
        const uint32_t v_addr = smem_u32(smem + OFF_V + st * SMEM_V);
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
          umma_ss(tmem + TMEM_O, make_desc_sw128(p_addr + kk * 32), make_desc_sw128(v_addr + kk * 32), idesc_o,
                  (j > 0 || kk > 0) ? 1u : 0u);
        }
        commit(&kv_empty[st]);
        commit(&p_free[b]);
      }
    }
  } else {
    // ================================ softmax + epilogue (warps 0-3) ================================
    const int row = warp * 32 + lane;                 // tile row == TMEM lane
    const int qi = q0 + row;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    const float bias2 = (p.qbias && qi < p.HW) ? p.qbias[(long long)qi * p.T + t] * LOG2E : 0.f;
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int b = j & 1;
      mbar_wait(&s_full[b], (j >> 1) & 1, p.err, 7);
      fence_after();
      float s[BN];
      tmem_ld32(lane_addr + TMEM_S + b * BN, s);
      tmem_ld32(lane_addr + TMEM_S + b * BN + 32, s + 32);
      fence_before();
      mbar_arrive(&s_free[b]);
      const int key0 = (tile_lo + j) * BN;            // key index inside the frame
      float mt = -INFINITY;
#pragma unroll
      for (int c = 0; c < BN; ++c) {
        float x = fmaf(s[c], p.scale_log2, bias2);
        x = (key0 + c < p.HW) ? x : -INFINITY;
        s[c] = x;
        mt = fmaxf(mt, x);
      }
      // lazy rescale (warp-uniform decision; each warp owns its 32 TMEM lanes)
      const bool need = mt > m_used + RESCALE_THRESHOLD;
      if (__any_sync(0xffffffffu, need)) {
        if (j > 0) {
          mbar_wait(&p_free[(j - 1) & 1], ((j - 1) >> 1) & 1, p.err, 8);   // PV(j-1) retired
          fence_after();
          const float f = need ? exp2f(m_used - mt) : 1.f;
          l *= f;
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
        if (need) m_used = mt;
      }
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < BN; ++c) {
        s[c] = exp2f(s[c] - m_used);
        lsum += s[c];
      }
      l += lsum;
      if (j >= 2) mbar_wait(&p_free[b], ((j - 2) >> 1) & 1, p.err, 9);     // PV(j-2) done reading P[b]
      unsigned char* prow = smem + OFF_P + b * SMEM_P + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
      for (int c = 0; c < BN / 8; ++c) {
        uint4 u;
        u.x = pack2(s[c * 8 + 0], s[c * 8 + 1]);
        u.y = pack2(s[c * 8 + 2], s[c * 8 + 3]);
        u.z = pack2(s[c * 8 + 4], s[c * 8 + 5]);
        u.w = pack2(s[c * 8 + 6], s[c * 8 + 7]);
        *reinterpret_cast<uint4*>(prow + ((c ^ (row & 7)) << 4)) = u;
      }
      fence_async_smem();
      mbar_arrive(&p_full[b]);
    }
    // epilogue: un-normalised partial O + (m, l)
    const int last = n_tiles - 1;
    mbar_wait(&p_free[last & 1], (last >> 1) & 1, p.err, 10);
    fence_after();
    float* po = p.part_o + ((long long)split * p.HW + qi) * p.Dv + dv0;
#pragma unroll 1
    for (int c = 0; c < DVC; c += 32) {
      float o[32];
      tmem_ld32(lane_addr + TMEM_O + c, o);
      if (qi < p.HW) {
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          *reinterpret_cast<float4*>(po + c + e) = make_float4(o[e], o[e + 1], o[e + 2], o[e + 3]);
      }
    }
    if (qi < p.HW && blockIdx.y == 0) {
      float* ml = p.part_ml + ((long long)split * p.HW + qi) * 2;
      ml[0] = m_used;
      ml[1] = l;
    }
    fence_before();
  }
  __syncthreads();
  if (warp == 5) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
  }
}

// Merge the split partials: out = (sum_s 2^(m_s-M) O_s) / L * gate ; mass[i,t] = sum_{s in t} l_s 2^(m_s-M) / L.
__global__ void __launch_bounds__(256) combine_kernel(const float* __restrict__ part_o,
                                                      const float* __restrict__ part_ml, int nsplit, int nsub, int T,
                                                      int HW, int Dv, const t16* __restrict__ gate, long long ldg,
                                                      t16* __restrict__ out, long long ldo,
                                                      float* __restrict__ mass) {
  __shared__ float w[kMaxBankFrames * 8];
  __shared__ float s_inv;
  const int i = blockIdx.x;
  if (threadIdx.x == 0) {
    float M = -INFINITY;
    for (int s = 0; s < nsplit; ++s) M = fmaxf(M, part_ml[((long long)s * HW + i) * 2]);
    float L = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const float* ml = part_ml + ((long long)s * HW + i) * 2;
      float f = exp2f(ml[0] - M);
      w[s] = f;
      L += f * ml[1];
    }
    s_inv = 1.f / L;
    if (mass) {
      for (int t = 0; t < T; ++t) {
        float a = 0.f;
        for (int u = 0; u < nsub; ++u) {
          int s = t * nsub + u;
          a += w[s] * part_ml[((long long)s * HW + i) * 2 + 1];
        }
        mass[(long long)i * T + t] = a / L;
      }
    }
  }
  __syncthreads();
  const float inv = s_inv;
  for (int c = threadIdx.x * 4; c < Dv; c += blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < nsplit; ++s) {
      const float4 o = *reinterpret_cast<const float4*>(part_o + ((long long)s * HW + i) * Dv + c);
      const float f = w[s];
      acc.x += f * o.x; acc.y += f * o.y; acc.z += f * o.z; acc.w += f * o.w;
    }
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    if (gate) {
      const uint2 g = *reinterpret_cast<const uint2*>(gate + (long long)i * ldg + c);
      const float2 g0 = unpack2(g.x), g1 = unpack2(g.y);
      acc.x *= g0.x; acc.y *= g0.y; acc.z *= g1.x; acc.w *= g1.y;
    }
    uint2 o;
    o.x = pack2(acc.x, acc.y);
    o.y = pack2(acc.z, acc.w);
    *reinterpret_cast<uint2*>(out + (long long)i * ldo + c) = o;
  }
}

int encode_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
              uint32_t box_inner, uint32_t box_outer) {
  uint64_t dims[2] = {inner, outer};
  uint64_t strides[1] = {row_stride_bytes};
  uint32_t box[2] = {box_inner, box_outer};
  return tma_encode(map, base, 2, dims, strides, box, nullptr);
}

int pick_nsub(int qtiles, int dvchunks, int T, int tiles_per_frame) {
  // aim for >= 2 waves of 148 CTAs without exceeding the tile count of a frame or the combine capacity
  int base = qtiles * dvchunks * T;
  int nsub = 1;
  while (base * nsub < 296 && nsub * 2 <= tiles_per_frame && T * nsub * 2 <= kMaxBankFrames * 8) nsub *= 2;
  return nsub;
}

}  // namespace

size_t long_attn_tc_workspace(int HW, int HWp, int nslots, int Dv) {
  (void)HWp;
  const int qtiles = cdiv(HW, BM), dvchunks = Dv / DVC > 0 ? Dv / DVC : 1, tiles_per_frame = cdiv(HW, BN);
  int max_split = 1;
  for (int T = 1; T <= nslots && T <= kMaxBankFrames; ++T) {
    int ns = T * pick_nsub(qtiles, dvchunks, T, tiles_per_frame);
    if (ns > max_split) max_split = ns;
  }
  return (size_t)max_split * (size_t)HW * ((size_t)Dv + 2) * sizeof(float) + 2048;
}

int long_attn_tc(const LongAttnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  RMEM_REQUIRE(a.Dk == DK, "long_attn_tc: Dk=%d (built for 128)", a.Dk);
  RMEM_REQUIRE(a.Dv % DVC == 0, "long_attn_tc: Dv=%d must be a multiple of 256", a.Dv);
  RMEM_REQUIRE(a.HWp % BN == 0 && a.HWp >= a.HW, "long_attn_tc: HWp=%d must be a multiple of 64", a.HWp);
  RMEM_REQUIRE(a.T >= 1 && a.T <= kMaxBankFrames && a.T <= a.nslots, "long_attn_tc: T=%d nslots=%d", a.T, a.nslots);
  RMEM_REQUIRE(a.ldo % 4 == 0 && (!a.gate || a.ldg % 4 == 0), "long_attn_tc: ldo/ldg alignment");
  const int qtiles = cdiv(a.HW, BM), dvchunks = a.Dv / DVC;
  const int tiles_per_frame = cdiv(a.HW, BN);
  const int nsub = pick_nsub(qtiles, dvchunks, a.T, tiles_per_frame);
  const int nsplit = a.T * nsub;
  const size_t need = (size_t)nsplit * a.HW * ((size_t)a.Dv + 2) * sizeof(float) + 2048;
  RMEM_REQUIRE(workspace_bytes >= need, "long_attn_tc: workspace %zu < %zu", workspace_bytes, need);

  CUtensorMap mq, mk, mv;
  RMEM_TRY(encode_2d(&mq, a.qt, DK, a.HW, DK * 2, 64, BM));
  RMEM_TRY(encode_2d(&mk, a.kbank, DK, (uint64_t)a.nslots * a.HWp, DK * 2, 64, BN));
  RMEM_TRY(encode_2d(&mv, a.vtbank, (uint64_t)a.nslots * a.HWp, a.Dv, (uint64_t)a.nslots * a.HWp * 2, BN, DVC));

  TcParams p;
  p.HW = a.HW; p.HWp = a.HWp; p.T = a.T; p.nsub = nsub; p.tiles_per_frame = tiles_per_frame; p.Dv = a.Dv;
  for (int t = 0; t < kMaxBankFrames; ++t) p.slot[t] = t < a.T ? a.slot[t] : 0;
  p.scale_log2 = a.scale * LOG2E;
  p.qbias = a.qbias;
  char* ws = reinterpret_cast<char*>(workspace);
  p.err = reinterpret_cast<int*>(ws);
  p.part_ml = reinterpret_cast<float*>(ws + 256);
  p.part_o = reinterpret_cast<float*>(ws + 1024 + (size_t)nsplit * a.HW * 2 * sizeof(float));
  p.part_o = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(p.part_o) + 15) & ~uintptr_t(15));

  static bool attr_done = false;
  if (!attr_done) {
    RMEM_CUDA_CHECK(cudaFuncSetAttribute(long_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    attr_done = true;
  }
  dim3 grid(qtiles, dvchunks, nsplit);
  long_attn_tc_kernel<<<grid, kThreads, SMEM_TOTAL, s>>>(mq, mk, mv, p);
  RMEM_LAUNCH_CHECK();
  combine_kernel<<<a.HW, 256, 0, s>>>(p.part_o, p.part_ml, nsplit, nsub, a.T, a.HW, a.Dv, a.gate, a.ldg, a.out, a.ldo,
                                      a.mass);
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

}  // namespace rmem
