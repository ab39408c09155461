// Materialised-score long-term attention (generic GEMMs + fused row softmax / per-frame mass) and the
// windowed local attention.  The dense path is the simple, always-correct implementation of K1/K3;
// the fused tcgen05 kernel in attn_tc.cu is the fast one and is checked against it on the GPU.
#include "attn.cuh"
#include "gemm.cuh"
#include "ops.cuh"

namespace rmem {

namespace {

struct SlotMap { int slot[kMaxBankFrames]; };

// One block per query row.  S row layout: [nslots][HWp] fp32 (only live slots / first HW columns valid).
__global__ void softmax_mass_kernel(const float* __restrict__ S, t16* __restrict__ P, long long ld, int HW, int HWp,
                                    int nslots, int T, SlotMap sm, const float* __restrict__ qbias,
                                    float* __restrict__ mass) {
  pdl_prologue();
  __shared__ float red[32];
  __shared__ float s_mass[kMaxBankFrames];
  const int row = blockIdx.x;
  // multi-head (AOT): blockIdx.y = head; S / P / qbias / mass are laid out [head][row][...]
  const long long hrow = (long long)blockIdx.y * HW + row;
  const float* Sr = S + hrow * ld;
  t16* Pr = P + hrow * ld;
  float bias[kMaxBankFrames];
#pragma unroll
  for (int t = 0; t < kMaxBankFrames; ++t) bias[t] = (t < T && qbias) ? qbias[hrow * T + t] : 0.f;

  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < kMaxBankFrames; ++t) {
    if (t >= T) break;
    const float* St = Sr + (long long)sm.slot[t] * HWp;
    for (int j = threadIdx.x; j < HW; j += blockDim.x) mx = fmaxf(mx, St[j] + bias[t]);
  }
  mx = block_max(mx, red);
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < kMaxBankFrames; ++t) {
    if (t >= T) break;
    const float* St = Sr + (long long)sm.slot[t] * HWp;
    for (int j = threadIdx.x; j < HW; j += blockDim.x) sum += __expf(St[j] + bias[t] - mx);
  }
  sum = block_sum(sum, red);
  const float inv = 1.f / sum;

  // zero everything first (dead slots + pad columns), then fill live frames
  for (int j = threadIdx.x; j < nslots * HWp; j += blockDim.x) {
    int s = j / HWp, c = j - s * HWp;
    bool live = false;
#pragma unroll
    for (int t = 0; t < kMaxBankFrames; ++t) live = live || (t < T && sm.slot[t] == s);
    if (!live || c >= HW) Pr[j] = f2t(0.f);
  }
#pragma unroll
  for (int t = 0; t < kMaxBankFrames; ++t) {
    if (t >= T) break;
    const float* St = Sr + (long long)sm.slot[t] * HWp;
    t16* Pt = Pr + (long long)sm.slot[t] * HWp;
    float m = 0.f;
    for (int j = threadIdx.x; j < HW; j += blockDim.x) {
      float p = __expf(St[j] + bias[t] - mx) * inv;
      Pt[j] = f2t(p);
      m += p;
    }
    if (mass) {
      m = block_sum(m, red);
      if (threadIdx.x == 0) s_mass[t] = m;
    }
  }
  if (mass) {
    __syncthreads();
    if (threadIdx.x < T) mass[hrow * T + threadIdx.x] = s_mass[threadIdx.x];
  }
}

// ------------------------------------------------------------------------------------------------
// Local attention, one warp per query (CUDA cores; 0.87 GFLOP at 480p).
template <int NCH>  // Dv = NCH * 256
__global__ void __launch_bounds__(128) local_attn_kernel(const t16* __restrict__ q, long long ldq,
                                                         const t16* __restrict__ k, long long ldk,
                                                         const t16* __restrict__ v, long long ldv,
                                                         const float* __restrict__ rel, long long ldrel,
                                                         const t16* __restrict__ gate, long long ldg,
                                                         t16* __restrict__ out, long long ldo, int h, int w,
                                                         float scale) {
  pdl_prologue();
  __shared__ float s_q[4][128];
  __shared__ float s_p[4][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + warp;
  if (i >= h * w) return;
  const int py = i / w, px = i - py * w;
  {
    uint2 u = *reinterpret_cast<const uint2*>(q + (long long)i * ldq + lane * 4);
    float2 a = unpack2(u.x), b = unpack2(u.y);
    s_q[warp][lane * 4 + 0] = a.x; s_q[warp][lane * 4 + 1] = a.y;
    s_q[warp][lane * 4 + 2] = b.x; s_q[warp][lane * 4 + 3] = b.y;
  }
  __syncwarp();
  float sc[8];
  float mx = -INFINITY;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    int d = r * 32 + lane;
    sc[r] = -INFINITY;
    if (d < 225) {
      int dy = d / 15 - 7, dx = d % 15 - 7;
      int ny = py + dy, nx = px + dx;
      if ((unsigned)ny < (unsigned)h && (unsigned)nx < (unsigned)w) {
        const t16* kr = k + (long long)(ny * w + nx) * ldk;
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          uint4 u = *reinterpret_cast<const uint4*>(kr + c * 8);
          float2 a = unpack2(u.x), b = unpack2(u.y), e = unpack2(u.z), f = unpack2(u.w);
          const float* qq = &s_q[warp][c * 8];
          dot += a.x * qq[0] + a.y * qq[1] + b.x * qq[2] + b.y * qq[3] + e.x * qq[4] + e.y * qq[5] + f.x * qq[6] +
                 f.y * qq[7];
        }
        sc[r] = dot * scale + rel[(long long)i * ldrel + d];
      }
    }
    mx = fmaxf(mx, sc[r]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    sc[r] = (sc[r] == -INFINITY) ? 0.f : __expf(sc[r] - mx);
    sum += sc[r];
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
#pragma unroll
  for (int r = 0; r < 8; ++r) s_p[warp][r * 32 + lane] = sc[r] * inv;
  __syncwarp();

  float acc[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[c][e] = 0.f;
  const int y_lo = max(py - 7, 0), y_hi = min(py + 7, h - 1);
  const int x_lo = max(px - 7, 0), x_hi = min(px + 7, w - 1);
  for (int ny = y_lo; ny <= y_hi; ++ny) {
    for (int nx = x_lo; nx <= x_hi; ++nx) {
      const float p = s_p[warp][(ny - py + 7) * 15 + (nx - px + 7)];
      const t16* vr = v + (long long)(ny * w + nx) * ldv + lane * 8;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint4 u = *reinterpret_cast<const uint4*>(vr + c * 256);
        float2 a = unpack2(u.x), b = unpack2(u.y), e = unpack2(u.z), f = unpack2(u.w);
        acc[c][0] += p * a.x; acc[c][1] += p * a.y; acc[c][2] += p * b.x; acc[c][3] += p * b.y;
        acc[c][4] += p * e.x; acc[c][5] += p * e.y; acc[c][6] += p * f.x; acc[c][7] += p * f.y;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = c * 256 + lane * 8;
    if (gate) {
      uint4 u = *reinterpret_cast<const uint4*>(gate + (long long)i * ldg + col);
      float2 a = unpack2(u.x), b = unpack2(u.y), e = unpack2(u.z), f = unpack2(u.w);
      acc[c][0] *= a.x; acc[c][1] *= a.y; acc[c][2] *= b.x; acc[c][3] *= b.y;
      acc[c][4] *= e.x; acc[c][5] *= e.y; acc[c][6] *= f.x; acc[c][7] *= f.y;
    }
    uint4 o;
    o.x = pack2(acc[c][0], acc[c][1]); o.y = pack2(acc[c][2], acc[c][3]);
    o.z = pack2(acc[c][4], acc[c][5]); o.w = pack2(acc[c][6], acc[c][7]);
    *reinterpret_cast<uint4*>(out + (long long)i * ldo + col) = o;
  }
}

}  // namespace

size_t long_attn_dense_workspace(int HW, int HWp, int nslots) {
  size_t cols = (size_t)nslots * HWp;
  return (size_t)HW * cols * (sizeof(float) + sizeof(t16)) + 256;
}

int long_attn_dense(const LongAttnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  RMEM_REQUIRE(a.T >= 1 && a.T <= kMaxBankFrames && a.T <= a.nslots, "long_attn: T=%d nslots=%d", a.T, a.nslots);
  RMEM_REQUIRE(a.HWp % 8 == 0 && a.HWp >= a.HW, "long_attn: HWp=%d HW=%d", a.HWp, a.HW);
  RMEM_REQUIRE(workspace_bytes >= long_attn_dense_workspace(a.HW, a.HWp, a.nslots), "long_attn: workspace too small");
  const long long ld = (long long)a.nslots * a.HWp;
  float* S = reinterpret_cast<float*>(workspace);
  t16* P = reinterpret_cast<t16*>(S + (size_t)a.HW * ld);
  for (int t = 0; t < a.T; ++t) {
    RMEM_REQUIRE(a.slot[t] >= 0 && a.slot[t] < a.nslots, "long_attn: bad slot");
    GemmParams g;
    g.A = a.qt; g.lda = a.Dk;
    g.B = a.kbank + (size_t)a.slot[t] * a.HWp * a.Dk; g.ldb = a.Dk;
    g.M = a.HW; g.N = a.HW; g.K = a.Dk;
    g.alpha = a.scale;
    g.C = S + (size_t)a.slot[t] * a.HWp; g.ldc = ld; g.c_fp32 = 1;
    RMEM_TRY(gemm_launch(g, s));
  }
  SlotMap sm;
  for (int t = 0; t < kMaxBankFrames; ++t) sm.slot[t] = t < a.T ? a.slot[t] : -1;
  RMEM_CUDA_CHECK(launch_pdl(softmax_mass_kernel, dim3(a.HW), dim3(256), 0, s, S, P, ld, a.HW, a.HWp, a.nslots, a.T, sm, a.qbias, a.mass));
  RMEM_LAUNCH_CHECK();
  GemmParams g;
  g.A = P; g.lda = ld;
  g.B = a.vtbank; g.ldb = ld;
  g.M = a.HW; g.N = a.Dv; g.K = (int)ld;
  g.gate = a.gate; g.ldg = a.ldg;
  g.C = a.out; g.ldc = a.ldo; g.c_fp32 = 0;
  return gemm_launch(g, s);
}

// ------------------------------------------------------------------------------------------------
// Multi-head attention over the bank, materialised scores (AOT MultiheadAttention, attention.py:28-81 and the mass
// record of transformer.py:636-643): head-batched legacy GEMMs (head dim 32 is below the tcgen05 kernel's K = 64 atom).
size_t mha_dense_workspace(int HW, int HWp, int nslots, int H) {
  size_t cols = (size_t)nslots * HWp;
  return (size_t)H * HW * cols * (sizeof(float) + sizeof(t16)) + (size_t)H * HW * kMaxBankFrames * sizeof(float) + 512;
}

int mha_dense(const MhaArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  RMEM_REQUIRE(a.T >= 1 && a.T <= kMaxBankFrames && a.T <= a.nslots, "mha: T=%d nslots=%d", a.T, a.nslots);
  RMEM_REQUIRE(a.HWp % 8 == 0 && a.HWp >= a.HW && a.dh % 8 == 0, "mha: HWp=%d HW=%d dh=%d", a.HWp, a.HW, a.dh);
  RMEM_REQUIRE(workspace_bytes >= mha_dense_workspace(a.HW, a.HWp, a.nslots, a.H), "mha: workspace too small");
  const int C = a.H * a.dh;
  const long long ld = (long long)a.nslots * a.HWp;
  float* S = reinterpret_cast<float*>(workspace);
  t16* P = reinterpret_cast<t16*>(S + (size_t)a.H * a.HW * ld);
  float* mass_h = reinterpret_cast<float*>(reinterpret_cast<char*>(P) + (((size_t)a.H * a.HW * ld * sizeof(t16) + 255) & ~size_t(255)));
  for (int t = 0; t < a.T; ++t) {
    RMEM_REQUIRE(a.slot[t] >= 0 && a.slot[t] < a.nslots, "mha: bad slot");
    GemmParams g;
    g.A = a.q; g.lda = a.ldq;
    g.B = a.kbank + (size_t)a.slot[t] * a.HWp * C; g.ldb = C;
    g.M = a.HW; g.N = a.HW; g.K = a.dh;
    g.alpha = a.scale;
    g.C = S + (size_t)a.slot[t] * a.HWp; g.ldc = ld; g.c_fp32 = 1;
    g.batch = a.H; g.sA = a.dh; g.sB = a.dh; g.sC = (long long)a.HW * ld;
    RMEM_TRY(gemm_launch(g, s));
  }
  SlotMap sm;
  for (int t = 0; t < kMaxBankFrames; ++t) sm.slot[t] = t < a.T ? a.slot[t] : -1;
  RMEM_CUDA_CHECK(launch_pdl(softmax_mass_kernel, dim3(dim3(a.HW, a.H)), dim3(256), 0, s, S, P, ld, a.HW, a.HWp, a.nslots, a.T, sm, a.qbias, a.mass ? mass_h : nullptr));
  RMEM_LAUNCH_CHECK();
  if (a.mass) RMEM_TRY(mean_heads(mass_h, a.mass, a.H, (long long)a.HW * a.T, s));
  GemmParams g;
  g.A = P; g.lda = ld;
  g.B = a.vtbank; g.ldb = ld;
  g.M = a.HW; g.N = a.dh; g.K = (int)ld;
  g.C = a.out; g.ldc = a.ldo; g.c_fp32 = 0;
  g.batch = a.H; g.sA = (long long)a.HW * ld; g.sB = (long long)a.dh * ld; g.sC = a.dh;
  return gemm_launch(g, s);
}

int local_attn(const t16* q, long long ldq, const t16* k, long long ldk, const t16* v, long long ldv,
               const float* rel, long long ldrel, const t16* gate, long long ldg, t16* out, long long ldo, int h,
               int w, int Dv, float scale, cudaStream_t s) {
  RMEM_REQUIRE(Dv == 1024 || Dv == 512 || Dv == 256, "local_attn: Dv=%d unsupported", Dv);
  RMEM_REQUIRE(ldq % 4 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && (!gate || ldg % 8 == 0),
               "local_attn: row strides must keep 16B alignment");
  const int grid = cdiv(h * w, 4);
  if (Dv == 1024)
    RMEM_CUDA_CHECK(launch_pdl(local_attn_kernel<4>, dim3(grid), dim3(128), 0, s, q, ldq, k, ldk, v, ldv, rel, ldrel, gate, ldg, out, ldo, h, w, scale));
  else if (Dv == 512)
    RMEM_CUDA_CHECK(launch_pdl(local_attn_kernel<2>, dim3(grid), dim3(128), 0, s, q, ldq, k, ldk, v, ldv, rel, ldrel, gate, ldg, out, ldo, h, w, scale));
  else
    RMEM_CUDA_CHECK(launch_pdl(local_attn_kernel<1>, dim3(grid), dim3(128), 0, s, q, ldq, k, ldk, v, ldv, rel, ldrel, gate, ldg, out, ldo, h, w, scale));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

}  // namespace rmem
