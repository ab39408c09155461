// HBM-bound kernels of the RMem propagation path: packing, pooling, norms, depthwise conv, resize,
// ID-bank gather, mask head, evict relevance.  Coalesced along the channel (innermost) dimension,
// 16-byte vector accesses where the layout allows.  See ops.cuh for the contracts.
#include "ops.cuh"

namespace rmem {

namespace {

__device__ __forceinline__ void load8(const t16* p, float* v) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack2(u.x), b = unpack2(u.y), c = unpack2(u.z), d = unpack2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void store8(t16* p, const float* v) {
  uint4 u;
  u.x = pack2(v[0], v[1]); u.y = pack2(v[2], v[3]);
  u.z = pack2(v[4], v[5]); u.w = pack2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ------------------------------------------------------------------------------------------------
__global__ void pack_image_kernel(const float* __restrict__ img, t16* __restrict__ out, int HW) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  float v[8] = {img[i], img[HW + i], img[2 * HW + i], 0.f, 0.f, 0.f, 0.f, 0.f};
  store8(out + (size_t)i * 8, v);
}

// Same into the zero-padded layout the stem convolution reads (gemm.cuh, conv = 2): [H + 6][W + 8][8], pixel (y, x) at
// padded (y + 3, x + 3).  Only the interior is written: the caller zeroes the buffer once.
__global__ void pack_image_padded_kernel(const float* __restrict__ img, t16* __restrict__ out, int H, int W) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;
  if (i >= HW) return;
  const int y = i / W, x = i - y * W;
  float v[8] = {img[i], img[HW + i], img[2 * HW + i], 0.f, 0.f, 0.f, 0.f, 0.f};
  store8(out + ((size_t)(y + 3) * (W + 8) + x + 3) * 8, v);
}

__global__ void maxpool_kernel(const t16* __restrict__ x, t16* __restrict__ y, int Hin, int Win, int C, int Hout,
                               int Wout) {
  pdl_prologue();
  const int cv = C / 8;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Hout * Wout * cv) return;
  int c8 = (int)(i % cv);
  int ox = (int)((i / cv) % Wout), oy = (int)(i / ((long long)cv * Wout));
  float m[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
  for (int ky = 0; ky < 3; ++ky) {
    int iy = oy * 2 - 1 + ky;
    if ((unsigned)iy >= (unsigned)Hin) continue;
    for (int kx = 0; kx < 3; ++kx) {
      int ix = ox * 2 - 1 + kx;
      if ((unsigned)ix >= (unsigned)Win) continue;
      float v[8];
      load8(x + ((size_t)iy * Win + ix) * C + c8 * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], v[k]);
    }
  }
  store8(y + ((size_t)oy * Wout + ox) * C + c8 * 8, m);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, two-pass statistics in registers.
constexpr int LN_MAXV = 16;  // C <= 512
__global__ void layernorm_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, t16* __restrict__ y, long long ldy,
                                 t16* __restrict__ y2, long long ldy2, const float* __restrict__ add2, int P, int C,
                                 float* __restrict__ init_res, long long ld_init) {
  pdl_prologue();
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= P) return;
  const float* xr = x + (long long)row * ldx;
  float v[LN_MAXV];
  const int nv = C / 32;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAXV; ++j)
    if (j < nv) { v[j] = xr[j * 32 + lane]; s += v[j]; }
  float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAXV; ++j)
    if (j < nv) { float d = v[j] - mean; q += d * d; }
  float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
#pragma unroll
  for (int j = 0; j < LN_MAXV; ++j)
    if (j < nv) {
      int c = j * 32 + lane;
      const float of = (v[j] - mean) * rstd * gamma[c] + beta[c];
      t16 o = f2t(of);
      y[(long long)row * ldy + c] = o;
      if (y2) y2[(long long)row * ldy2 + c] = add2 ? f2t(of + add2[(long long)row * C + c]) : o;
      if (init_res) {                       // start of a residual stream [x | 0] (DualBranchGPM: tgt = x, tgt_id = 0)
        init_res[(long long)row * ld_init + c] = v[j];
        init_res[(long long)row * ld_init + C + c] = 0.f;
      }
    }
}

// Two LayerNorms over the two C-wide halves of one [P, 2C] row (norm2 on tgt, id_norm2 on tgt_id, transformer.py:1222)
// in one launch: warp -> (row, half).
__global__ void layernorm_pair_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ g0,
                                      const float* __restrict__ b0, const float* __restrict__ g1,
                                      const float* __restrict__ b1, t16* __restrict__ y, long long ldy,
                                      t16* __restrict__ y1, long long ldy1, int P, int C) {
  pdl_prologue();
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int row = w >> 1, half = w & 1;
  if (row >= P) return;
  const float* xr = x + (long long)row * ldx + half * C;
  const float* gamma = half ? g1 : g0;
  const float* beta = half ? b1 : b0;
  float v[LN_MAXV];
  const int nv = C / 32;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAXV; ++j)
    if (j < nv) { v[j] = xr[j * 32 + lane]; s += v[j]; }
  float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAXV; ++j)
    if (j < nv) { float d = v[j] - mean; q += d * d; }
  float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
#pragma unroll
  for (int j = 0; j < LN_MAXV; ++j)
    if (j < nv) {
      int c = j * 32 + lane;
      const float o = (v[j] - mean) * rstd * gamma[c] + beta[c];
      if (half && y1) y1[(long long)row * ldy1 + c] = f2t(o);     // second half to its own destination
      else y[(long long)row * ldy + half * C + c] = f2t(o);
    }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm: stats (double atomics) + apply.
template <typename T>
__device__ __forceinline__ void gn_load8(const T* p, float* v);
template <>
__device__ __forceinline__ void gn_load8<t16>(const t16* p, float* v) { load8(p, v); }
template <>
__device__ __forceinline__ void gn_load8<float>(const float* p, float* v) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// Deterministic two-level reduction: fixed-order per-block partials, the last block to finish folds them in
// block order (no floating-point atomics, so run-to-run results are bit-identical).
template <typename T>
__global__ void gn_stats_kernel(const T* __restrict__ x, int P, int C, int G, double* __restrict__ stats,
                                double* __restrict__ partials, unsigned int* __restrict__ counter) {
  pdl_prologue();
  // thread -> fixed 8-channel vector column; rows strided over the grid.
  const int cv = C / 8;
  const int col = threadIdx.x % cv;
  const int rows_per_block = blockDim.x / cv;
  const int r0 = threadIdx.x / cv;
  float s = 0.f, q = 0.f;
  for (int r = blockIdx.x * rows_per_block + r0; r < P; r += gridDim.x * rows_per_block) {
    float v[8];
    gn_load8<T>(x + (size_t)r * C + col * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) { s += v[k]; q += v[k] * v[k]; }
  }
  __shared__ float sh_s[256], sh_q[256];
  __shared__ bool is_last;
  sh_s[threadIdx.x] = s;
  sh_q[threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.x < 2 * G) {
    const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
    const int cpg = cv / G;              // vector columns per group
    const float* src = which ? sh_q : sh_s;
    double a = 0.0;
    for (int r = 0; r < rows_per_block; ++r)
      for (int c = 0; c < cpg; ++c) a += (double)src[r * cv + g * cpg + c];
    partials[(size_t)blockIdx.x * 2 * G + threadIdx.x] = a;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    __threadfence();
    // fold the per-block partials with the whole block: thread -> (output o, slice j); slice j adds blocks j, j+J, ...
    // in order, then the J slices are added in order -- still a fixed summation order, ~16x shorter serial chain
    __shared__ double sh_d[256];
    const int n_out = 2 * G, J = blockDim.x / n_out;
    const int o = threadIdx.x % n_out, j = threadIdx.x / n_out;
    double a = 0.0;
    if (j < J)
      for (unsigned b = j; b < gridDim.x; b += J) a += partials[(size_t)b * n_out + o];
    sh_d[threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.x < n_out) {
      double t = 0.0;
      for (int jj = 0; jj < J; ++jj) t += sh_d[jj * n_out + threadIdx.x];
      stats[threadIdx.x] = t;
    }
    if (threadIdx.x == 0) *counter = 0u;   // re-arm for the next call on this stream
  }
}

template <typename T>
__global__ void gn_apply_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                const float* __restrict__ beta, t16* __restrict__ y, int P, int C, int G, int relu,
                                const double* __restrict__ stats, const t16* __restrict__ add) {
  pdl_prologue();
  const int cv = C / 8;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)P * cv) return;
  int col = (int)(i % cv);
  long long r = i / cv;
  int g = (col * 8) / (C / G);
  double n = (double)P * (C / G);
  double mean = stats[g * 2] / n;
  double var = stats[g * 2 + 1] / n - mean * mean;
  float fm = (float)mean, rstd = rsqrtf(fmaxf((float)var, 0.f) + 1e-5f);
  float v[8];
  gn_load8<T>(x + (size_t)r * C + col * 8, v);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int c = col * 8 + k;
    float o = (v[k] - fm) * rstd * gamma[c] + beta[c];
    v[k] = relu == 1 ? fmaxf(o, 0.f) : (relu == 2 ? 0.5f * o * (1.f + erff(o * 0.70710678118654752f)) : o);
  }
  if (add) {                                   // e.g. the FPN adapter branch, added to the normalised + activated map
    float a[8];
    load8(add + (size_t)r * C + col * 8, a);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] += a[k];
  }
  store8(y + (size_t)r * C + col * 8, v);
}

constexpr int kGnMaxBlocks = 148 * 4;

template <typename T>
int groupnorm_impl(const T* x, const float* gamma, const float* beta, t16* y, int P, int C, int G, int relu,
                   double* stats, cudaStream_t s, const t16* add = nullptr, bool stats_ready = false) {
  RMEM_REQUIRE(C % 8 == 0 && G <= 32 && C % G == 0 && (C / G) % 8 == 0 && 256 % (C / 8) == 0,
               "groupnorm: unsupported C=%d G=%d", C, G);
  // scratch layout (doubles): [0,64) stats | [64] counter (zero-initialised once, self re-arming) | [72,..) partials
  if (!stats_ready) {     // (stats_ready: the producing GEMM's epilogue left them there, GemmParams::gn_stats)
    int rows_per_block = 256 / (C / 8);
    int grid = min(cdiv(P, rows_per_block), kGnMaxBlocks);     // one row group per block until the grid cap: parallelism first
    RMEM_CUDA_CHECK(launch_pdl(gn_stats_kernel<T>, dim3(grid), dim3(256), 0, s, x, P, C, G, stats, stats + 72, reinterpret_cast<unsigned int*>(stats + 64)));
    RMEM_LAUNCH_CHECK();
  }
  long long nvec = (long long)P * (C / 8);
  RMEM_CUDA_CHECK(launch_pdl(gn_apply_kernel<T>, dim3((unsigned)((nvec + 255) / 256)), dim3(256), 0, s, x, gamma, beta, y, P, C, G, relu, static_cast<const double*>(stats), add));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

// ------------------------------------------------------------------------------------------------
// 5x5 depthwise convolution, NHWC.  Block = 64 channels (lane = channel pair) x an 8 x 18 output tile: the 12 x 22 input
// halo is staged once in shared memory (coalesced 128-byte pixel rows, all loads in flight before the single barrier),
// warp w then slides a 5 x 5 register window along row w of the tile: 5 shared-memory reads per output instead of 25
// global ones, the 25 weight pairs in registers (fetched before the PDL wait: not produced by the preceding kernel).
// fp32 accumulation in (ky, kx) order.  The register-only versions before it sat at ~13 us for 6.6 MB of traffic
// (148 registers -> 2.4 waves of L2-latency-bound threads).
constexpr int DW_TH = 8;
template <int DW_TW>
__global__ void __launch_bounds__(256) dwconv5_kernel(const t16* __restrict__ x, int ldx, const float* __restrict__ w,
                                                      t16* __restrict__ y, int ldy, int h, int wd, int C) {
  __shared__ uint32_t tile[(DW_TH + 4) * (DW_TW + 4)][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cp = ldx / 2, cpy = ldy / 2;                 // pixel pitch of input / output in channel pairs
  const int c2 = blockIdx.x * 32 + lane;                 // channel pair (C % 64 == 0)
  const int x0 = blockIdx.y * DW_TW, y0 = blockIdx.z * DW_TH;
  float2 wt[25];
#pragma unroll
  for (int k = 0; k < 25; ++k) wt[k] = *reinterpret_cast<const float2*>(w + (size_t)k * C + c2 * 2);
  pdl_prologue();
  const uint32_t* xin = reinterpret_cast<const uint32_t*>(x) + c2;
  constexpr int NPIX = (DW_TH + 4) * (DW_TW + 4);        // 264 halo pixels, 33 per warp
  uint32_t v[(NPIX + 7) / 8];
#pragma unroll
  for (int i = 0; i < (NPIX + 7) / 8; ++i) {
    const int p = i * 8 + warp;
    const int iy = y0 - 2 + p / (DW_TW + 4), ix = x0 - 2 + p % (DW_TW + 4);
    v[i] = (p < NPIX && (unsigned)iy < (unsigned)h && (unsigned)ix < (unsigned)wd) ? xin[((size_t)iy * wd + ix) * cp] : 0u;
  }
#pragma unroll
  for (int i = 0; i < (NPIX + 7) / 8; ++i) {
    const int p = i * 8 + warp;
    if (p < NPIX) tile[p][lane] = v[i];
  }
  __syncthreads();
  const int oy = y0 + warp;
  if (oy >= h) return;
  uint32_t* yout = reinterpret_cast<uint32_t*>(y) + c2;
  float2 win[5][5];                                        // [kx][ky]
#pragma unroll
  for (int kx = 1; kx < 5; ++kx)
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) win[kx][ky] = unpack2(tile[(warp + ky) * (DW_TW + 4) + kx - 1][lane]);
#pragma unroll
  for (int px = 0; px < DW_TW; ++px) {
    if (x0 + px >= wd) break;
#pragma unroll
    for (int kx = 0; kx < 4; ++kx)
#pragma unroll
      for (int ky = 0; ky < 5; ++ky) win[kx][ky] = win[kx + 1][ky];
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) win[4][ky] = unpack2(tile[(warp + ky) * (DW_TW + 4) + px + 4][lane]);
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int ky = 0; ky < 5; ++ky)
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        a0 = fmaf(win[kx][ky].x, wt[ky * 5 + kx].x, a0);
        a1 = fmaf(win[kx][ky].y, wt[ky * 5 + kx].y, a1);
      }
    yout[((size_t)oy * wd + x0 + px) * cpy] = pack2(a0, a1);
  }
}

// ------------------------------------------------------------------------------------------------
// align_corners=True source coordinate, as ATen's area_pixel_compute_source_index.
__device__ __forceinline__ void src_index(int dst, int in_size, int out_size, int& i0, int& i1, float& l0, float& l1) {
  if (in_size == out_size) { i0 = i1 = dst; l0 = 1.f; l1 = 0.f; return; }
  float scale = out_size > 1 ? __fdiv_rn((float)(in_size - 1), (float)(out_size - 1)) : 0.f;  // exact: --use_fast_math
  float real = __fmul_rn(scale, (float)dst);
  i0 = min((int)real, in_size - 1);
  float lam = fminf(fmaxf(__fsub_rn(real, (float)i0), 0.f), 1.f);
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = lam;
  l0 = __fsub_rn(1.f, lam);
}
__device__ __forceinline__ float bilerp(float v00, float v01, float v10, float v11, float wy0, float wy1, float wx0,
                                        float wx1) {
  float top = __fadd_rn(__fmul_rn(wx0, v00), __fmul_rn(wx1, v01));
  float bot = __fadd_rn(__fmul_rn(wx0, v10), __fmul_rn(wx1, v11));
  return __fadd_rn(__fmul_rn(wy0, top), __fmul_rn(wy1, bot));
}

// gn_stats != nullptr: x is the un-normalised convolution output; relu(GroupNorm(x)) (statistics given, G groups) is
// applied to each of the four taps on load, so the normalised map is never written (fpn.py:46-60: up(relu(gn(conv)))).
__global__ void upsample_t16_kernel(const t16* __restrict__ x, t16* __restrict__ y, int hin, int win, int hout,
                                     int wout, int C, const t16* __restrict__ add, const double* __restrict__ gn_stats,
                                     const float* __restrict__ gamma, const float* __restrict__ beta, int G) {
  pdl_prologue();
  const int cv = C / 8;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)hout * wout * cv) return;
  int c8 = (int)(i % cv);
  int ox = (int)((i / cv) % wout), oy = (int)(i / ((long long)cv * wout));
  int y0, y1, x0, x1;
  float wy0, wy1, wx0, wx1;
  src_index(oy, hin, hout, y0, y1, wy0, wy1);
  src_index(ox, win, wout, x0, x1, wx0, wx1);
  float a[8], b[8], c[8], d[8], o[8];
  load8(x + ((size_t)y0 * win + x0) * C + c8 * 8, a);
  load8(x + ((size_t)y0 * win + x1) * C + c8 * 8, b);
  load8(x + ((size_t)y1 * win + x0) * C + c8 * 8, c);
  load8(x + ((size_t)y1 * win + x1) * C + c8 * 8, d);
  if (gn_stats) {                              // same arithmetic as gn_apply_kernel (8 channels never straddle a group)
    const int g = (c8 * 8) / (C / G);
    const double n = (double)hin * win * (C / G);
    const double mean = gn_stats[g * 2] / n;
    const double var = gn_stats[g * 2 + 1] / n - mean * mean;
    const float fm = (float)mean, rstd = rsqrtf(fmaxf((float)var, 0.f) + 1e-5f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float ga = gamma[c8 * 8 + k], be = beta[c8 * 8 + k];
      a[k] = fmaxf((a[k] - fm) * rstd * ga + be, 0.f);
      b[k] = fmaxf((b[k] - fm) * rstd * ga + be, 0.f);
      c[k] = fmaxf((c[k] - fm) * rstd * ga + be, 0.f);
      d[k] = fmaxf((d[k] - fm) * rstd * ga + be, 0.f);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = bilerp(a[k], b[k], c[k], d[k], wy0, wy1, wx0, wx1);
  if (add) {                                   // FPN: adapter(feature) + upsample(x), fpn.py:50-60
    float e[8];
    load8(add + ((size_t)oy * wout + ox) * C + c8 * 8, e);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] += e[k];
  }
  store8(y + ((size_t)oy * wout + ox) * C + c8 * 8, o);
}

// ------------------------------------------------------------------------------------------------
// conv_out (1x1, Cin <= 128 -> Cout <= 16, planar fp32 output): one thread per pixel.  A block stages its 128 pixel rows
// in shared memory with coalesced 16-byte loads (row pitch + 16 B: conflict-free row reads), the fp32 weights sit next to
// them and are read as broadcasts, the Cout planes are written with lanes = consecutive pixels.  (One warp per pixel with
// a shuffle tree per output channel and single-lane stores took 23 us for 25680 pixels; row-per-thread global loads 18.)
__global__ void __launch_bounds__(128) conv_out_kernel(const t16* __restrict__ x, const t16* __restrict__ w,
                                                       const float* __restrict__ b, float* __restrict__ out, int P,
                                                       int Cin, int Cout) {
  extern __shared__ __align__(16) unsigned char co_smem[];
  float* sw = reinterpret_cast<float*>(co_smem);                 // [Cin / 8][16][8]: 8 input channels per output channel
  const int nch = Cin / 8, pitch = Cin * 2 + 16;
  unsigned char* tile = co_smem + (size_t)nch * 16 * 8 * sizeof(float);
  for (int i = threadIdx.x; i < Cout * Cin; i += blockDim.x) {
    const int o = i / Cin, c = i - o * Cin;
    sw[((c >> 3) * 16 + o) * 8 + (c & 7)] = t2f(w[i]);          // weights are not produced by the preceding kernel
  }
  pdl_prologue();
  const int p0 = blockIdx.x * 128;
  for (int i = threadIdx.x; i < 128 * nch; i += blockDim.x) {
    const int pix = i / nch, ch = i - pix * nch;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (p0 + pix < P) v = *reinterpret_cast<const uint4*>(x + (size_t)(p0 + pix) * Cin + ch * 8);
    *reinterpret_cast<uint4*>(tile + (size_t)pix * pitch + ch * 16) = v;
  }
  __syncthreads();
  const int p = p0 + threadIdx.x;
  if (p >= P) return;
  float acc[16];
#pragma unroll
  for (int o = 0; o < 16; ++o) acc[o] = 0.f;
  const unsigned char* row = tile + (size_t)threadIdx.x * pitch;
  for (int c8 = 0; c8 < nch; ++c8) {
    const uint4 r = *reinterpret_cast<const uint4*>(row + c8 * 16);
    const float2 a = unpack2(r.x), bb = unpack2(r.y), cc = unpack2(r.z), dd = unpack2(r.w);
#pragma unroll
    for (int o = 0; o < 16; ++o) {
      if (o < Cout) {
        const float4 w0 = *reinterpret_cast<const float4*>(&sw[(c8 * 16 + o) * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&sw[(c8 * 16 + o) * 8 + 4]);
        acc[o] += a.x * w0.x + a.y * w0.y + bb.x * w0.z + bb.y * w0.w + cc.x * w1.x + cc.y * w1.y + dd.x * w1.z +
                  dd.y * w1.w;
      }
    }
  }
#pragma unroll
  for (int o = 0; o < 16; ++o)
    if (o < Cout) out[(size_t)o * P + p] = acc[o] + b[o];
}

// Decoder tail in one launch: y = conv_out( relu( GroupNorm(x) ) ) -- fpn.py:62-67.  The GroupNorm statistics come from
// gn_stats_kernel (`stats`, per-group sum / sum of squares in double); the normalised activation is rounded to t16 when
// it is staged in shared memory, exactly what gn_apply_kernel would have stored; the logits equal the two-kernel path's
// up to the fp32 summation order (two half-sums over the input channels instead of one chain) -- minus one 6.6 MB round
// trip and one launch.  64 pixels per block, two threads per pixel (each takes half of the input channels, one shuffle to
// add): 403 blocks at P4 = 25.8 k pixels instead of 202 single-warp-per-scheduler blocks.
__global__ void __launch_bounds__(128) conv_out_gn_kernel(const t16* __restrict__ x, const t16* __restrict__ w,
                                                          const float* __restrict__ b, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const double* __restrict__ stats,
                                                          int G, float* __restrict__ out, int P, int Cin, int Cout) {
  extern __shared__ __align__(16) unsigned char co_smem[];
  float* sw = reinterpret_cast<float*>(co_smem);                 // [Cin / 8][16][8]
  const int nch = Cin / 8, pitch = Cin * 2 + 16;
  float* s_a = sw + (size_t)nch * 16 * 8;                        // [Cin] scale, [Cin] shift of the normalisation
  float* s_b = s_a + Cin;
  unsigned char* tile = reinterpret_cast<unsigned char*>(s_b + Cin);
  for (int i = threadIdx.x; i < Cout * Cin; i += blockDim.x) {
    const int o = i / Cin, c = i - o * Cin;
    sw[((c >> 3) * 16 + o) * 8 + (c & 7)] = t2f(w[i]);
  }
  pdl_prologue();
  for (int c = threadIdx.x; c < Cin; c += blockDim.x) {
    const int g = c / (Cin / G);
    const double n = (double)P * (Cin / G);
    const double mean = stats[g * 2] / n;
    const double var = stats[g * 2 + 1] / n - mean * mean;
    s_a[c] = (float)mean;                                        // same arithmetic as gn_apply_kernel
    s_b[c] = rsqrtf(fmaxf((float)var, 0.f) + 1e-5f);
  }
  __syncthreads();
  const int p0 = blockIdx.x * 64;
  for (int i = threadIdx.x; i < 64 * nch; i += blockDim.x) {
    const int pix = i / nch, ch = i - pix * nch;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = 0.f;
    if (p0 + pix < P) {
      load8(x + (size_t)(p0 + pix) * Cin + ch * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = ch * 8 + k;
        v[k] = fmaxf((v[k] - s_a[c]) * s_b[c] * gamma[c] + beta[c], 0.f);
      }
    }
    uint4 u;
    u.x = pack2(v[0], v[1]); u.y = pack2(v[2], v[3]); u.z = pack2(v[4], v[5]); u.w = pack2(v[6], v[7]);
    *reinterpret_cast<uint4*>(tile + (size_t)pix * pitch + ch * 16) = u;
  }
  __syncthreads();
  const int pix = threadIdx.x >> 1, half = threadIdx.x & 1;
  const int p = p0 + pix;
  float acc[16];
#pragma unroll
  for (int o = 0; o < 16; ++o) acc[o] = 0.f;
  const unsigned char* row = tile + (size_t)pix * pitch;
  const int c_lo = half * (nch / 2), c_hi = c_lo + nch / 2;
  for (int c8 = c_lo; c8 < c_hi; ++c8) {
    const uint4 r = *reinterpret_cast<const uint4*>(row + c8 * 16);
    const float2 a = unpack2(r.x), bb = unpack2(r.y), cc = unpack2(r.z), dd = unpack2(r.w);
#pragma unroll
    for (int o = 0; o < 16; ++o) {
      if (o < Cout) {
        const float4 w0 = *reinterpret_cast<const float4*>(&sw[(c8 * 16 + o) * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&sw[(c8 * 16 + o) * 8 + 4]);
        acc[o] += a.x * w0.x + a.y * w0.y + bb.x * w0.z + bb.y * w0.w + cc.x * w1.x + cc.y * w1.y + dd.x * w1.z +
                  dd.y * w1.w;
      }
    }
  }
#pragma unroll
  for (int o = 0; o < 16; ++o) {
    const float other = __shfl_xor_sync(0xffffffffu, acc[o], 1);
    acc[o] = half == 0 ? acc[o] + other : other + acc[o];          // low half of the channels first in both threads
  }
  if (half == 0 && p < P) {
#pragma unroll
    for (int o = 0; o < 16; ++o)
      if (o < Cout) out[(size_t)o * P + p] = acc[o] + b[o];
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void transpose_kernel(const t16* __restrict__ x, long long ldx, t16* __restrict__ y, long long ldy, int P,
                                 int C) {
  pdl_prologue();
  __shared__ t16 tile[64][66];
  int p0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  for (int i = threadIdx.y; i < 64; i += blockDim.y) {
    int p = p0 + i, c = c0 + threadIdx.x * 2;
    t16 a = f2t(0.f), b = a;
    if (p < P && c < C) {
      a = x[(long long)p * ldx + c];
      if (c + 1 < C) b = x[(long long)p * ldx + c + 1];
    }
    tile[i][threadIdx.x * 2] = a;
    tile[i][threadIdx.x * 2 + 1] = b;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 64; i += blockDim.y) {
    int c = c0 + i;
    if (c >= C) continue;
    int p = p0 + threadIdx.x * 2;
    if (p < P) y[(long long)c * ldy + p] = tile[threadIdx.x * 2][i];
    if (p + 1 < P) y[(long long)c * ldy + p + 1] = tile[threadIdx.x * 2 + 1][i];
  }
}

__global__ void copy2d_kernel(const t16* __restrict__ src, long long lds, t16* __restrict__ dst, long long ldd, int P,
                              int C) {
  pdl_prologue();
  const int cv = C / 8;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)P * cv) return;
  int c8 = (int)(i % cv);
  long long r = i / cv;
  *reinterpret_cast<uint4*>(dst + r * ldd + c8 * 8) = *reinterpret_cast<const uint4*>(src + r * lds + c8 * 8);
}
__global__ void fill_kernel(t16* __restrict__ dst, long long ldd, int P, int C, float v) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)P * C) return;
  dst[(i / C) * ldd + (i % C)] = f2t(v);
}

__global__ void separate_label_kernel(const void* __restrict__ label, int is_f32, uint8_t* __restrict__ out, int n,
                                      int engine, int n_engines) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int v = is_f32 ? (int)reinterpret_cast<const float*>(label)[i] : (int)reinterpret_cast<const uint8_t*>(label)[i];
  if (n_engines > 1) {
    int lo = engine * 10 + 1, hi = (engine + 1) * 10;
    v = (v >= lo && v <= hi) ? (v - lo + 1) : 0;
  }
  out[i] = (uint8_t)v;
}

// ------------------------------------------------------------------------------------------------
// ID bank: one block per token, one thread per output channel.  Summing one conv-weight row per pixel of the 17x17 label
// patch is a 289 x 1 KB gather per token (484 MB of L2 reads per 480p frame: 47-72 us).  Three cheaper decompositions of
// the same sum, picked per patch by their read count:
//   RECT  the in-bounds pixels share one class: one rectangle of that class' 2-D prefix table (4 reads);
//   RUNS  every image row of the patch is cut into runs of one class, each run is a difference of two entries of the
//         per-(row, class) 1-D prefix table (2 reads per run; label maps are piecewise constant, ~2 runs per row);
//   MINOR rectangle of the dominant class + (w[pixel, class] - w[pixel, dominant]) for the other pixels (salt-and-pepper);
//   TAPS  the plain gather (also taken when no prefix table is given).
// Steady-state c3 labels (argmax maps of the network): 59 reads per patch on average instead of 289.
__global__ void idbank_kernel(const uint8_t* __restrict__ label, int H, int W, int use_ignore,
                              const float* __restrict__ wp, const float* __restrict__ prefix,
                              const float* __restrict__ prefix_rows, const float* __restrict__ bias,
                              const float* __restrict__ ln_g, const float* __restrict__ ln_b, t16* __restrict__ out,
                              long long ldo, float* __restrict__ out_f32, int w, int C, t16* __restrict__ out2,
                              t16* __restrict__ out3, long long ldo23) {
  pdl_prologue();
  __shared__ int8_t ch[17 * 17];
  __shared__ float red[32];
  __shared__ int s_hist[13];                     // in-bounds pixels per class (12 = in no one-hot channel)
  __shared__ int s_nruns, s_nlist;
  __shared__ int s_rowruns[17];                  // valid-class runs per patch row
  __shared__ unsigned short list[17 * 17];       // RUNS: ky | x0 << 5 | x1 << 10 (class from ch);  MINOR: tap index
  // (both lists are built in a fixed order -- row-major -- so the fp32 sum, and with it the whole engine, stays
  // bit-reproducible from run to run)
  const int tok = blockIdx.x;
  const int py = tok / w, px = tok - py * w;
  if (threadIdx.x < 13) s_hist[threadIdx.x] = 0;
  if (threadIdx.x == 13) { s_nruns = 0; s_nlist = 0; }
  __syncthreads();
  for (int t = threadIdx.x; t < 289; t += blockDim.x) {
    int ky = t / 17, kx = t - ky * 17;
    int iy = py * 16 - 8 + ky, ix = px * 16 - 8 + kx;
    int c = -1;
    if ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W) {
      int lab = label[(size_t)iy * W + ix];
      if (lab <= 10) c = lab;
      else if (lab == 255 && use_ignore) c = 11;
      else c = 12;                      // in bounds but in no one-hot channel: contributes nothing
      atomicAdd(&s_hist[c], 1);
    }
    ch[t] = (int8_t)c;
  }
  __syncthreads();
  // in-bounds taps form the rectangle [ky0, ky1) x [kx0, kx1)
  const int ky0 = max(0, 8 - py * 16), ky1 = min(17, H + 8 - py * 16);
  const int kx0 = max(0, 8 - px * 16), kx1 = min(17, W + 8 - px * 16);
  enum { TAPS = 0, RECT = 1, RUNS = 2, MINOR = 3 };
  int mode = TAPS, dom = 0;
  if (prefix) {
    int tot = 0, best = -1;
    for (int k = 0; k < 13; ++k) { tot += s_hist[k]; if (k < 12 && s_hist[k] > best) { best = s_hist[k]; dom = k; } }
    const int minor = tot - best;                            // pixels that are not of the dominant one-hot class
    if (minor == 0) mode = RECT;
    else {
      int cost = 289;
      if (4 + 2 * minor < cost) { cost = 4 + 2 * minor; mode = MINOR; }
      if (prefix_rows) {
        // valid-class runs of this patch (counted by 17 threads, one per row)
        if (threadIdx.x < 17) {
          const int ky = threadIdx.x;
          int n = 0;
          if (ky >= ky0 && ky < ky1)
            for (int kx = kx0; kx < kx1; ++kx) {
              const int k = ch[ky * 17 + kx];
              if (k < 12 && (kx == kx0 || ch[ky * 17 + kx - 1] != k)) ++n;
            }
          s_rowruns[ky] = n;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
          int n = 0;
          for (int ky = 0; ky < 17; ++ky) n += s_rowruns[ky];
          s_nruns = n;
        }
        __syncthreads();
        if (2 * s_nruns < cost) { cost = 2 * s_nruns; mode = RUNS; }
      }
    }
  }
  if (mode == RUNS) {
    if (threadIdx.x < 17) {
      const int ky = threadIdx.x;
      int pos = 0;
      for (int r = 0; r < ky; ++r) pos += s_rowruns[r];
      if (ky >= ky0 && ky < ky1) {
        int x0 = kx0;
        for (int kx = kx0 + 1; kx <= kx1; ++kx)
          if (kx == kx1 || ch[ky * 17 + kx] != ch[ky * 17 + x0]) {
            if (ch[ky * 17 + x0] < 12) list[pos++] = (unsigned short)(ky | (x0 << 5) | (kx << 10));
            x0 = kx;
          }
      }
      if (ky == 16) s_nlist = pos;
    }
    __syncthreads();
  } else if (mode == MINOR) {
    if (threadIdx.x < 32) {                                    // ordered compaction by one warp
      int pos = 0;
      for (int t0 = 0; t0 < 289; t0 += 32) {
        const int t = t0 + threadIdx.x;
        const bool m = t < 289 && ch[t] >= 0 && ch[t] != dom;
        const unsigned b = __ballot_sync(0xffffffffu, m);
        if (m) list[pos + __popc(b & ((1u << threadIdx.x) - 1u))] = (unsigned short)t;
        pos += __popc(b);
      }
      if (threadIdx.x == 0) s_nlist = pos;
    }
    __syncthreads();
  }
  const int c = threadIdx.x;
  float acc = 0.f;
  if (c < C) {
    if (mode == RECT || mode == MINOR) {
      const float* P = prefix + (size_t)dom * 18 * 18 * C + c;      // P[ky][kx] = sum over taps [0,ky) x [0,kx)
      acc = (P[((size_t)ky1 * 18 + kx1) * C] - P[((size_t)ky0 * 18 + kx1) * C]) -
            (P[((size_t)ky1 * 18 + kx0) * C] - P[((size_t)ky0 * 18 + kx0) * C]);
      if (mode == MINOR) {
        const int n = s_nlist;
        for (int i = 0; i < n; ++i) {
          const int t = list[i], k = ch[t];
          const float wd = wp[((size_t)t * 12 + dom) * C + c];
          const float wk = k < 12 ? wp[((size_t)t * 12 + k) * C + c] : 0.f;
          acc += wk - wd;
        }
      }
    } else if (mode == RUNS) {
      const int n = s_nlist;
      int i = 0;
      for (; i + 4 <= n; i += 4) {                               // four runs (eight reads) in flight
        float v[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = list[i + u], ky = r & 31, x0 = (r >> 5) & 31, x1 = r >> 10;
          const float* P = prefix_rows + ((size_t)(ky * 12 + ch[ky * 17 + x0]) * 18) * C + c;
          v[2 * u] = P[(size_t)x1 * C];
          v[2 * u + 1] = P[(size_t)x0 * C];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += v[2 * u] - v[2 * u + 1];
      }
      for (; i < n; ++i) {
        const int r = list[i], ky = r & 31, x0 = (r >> 5) & 31, x1 = r >> 10;
        const float* P = prefix_rows + ((size_t)(ky * 12 + ch[ky * 17 + x0]) * 18) * C + c;
        acc += P[(size_t)x1 * C] - P[(size_t)x0 * C];
      }
    } else {
      // predicated loads, eight taps in flight (the tap loop is L2-latency bound otherwise); summed in tap order
      int t = 0;
      for (; t + 8 <= 289; t += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int k = ch[t + u];
          v[u] = (k >= 0 && k < 12) ? wp[((size_t)(t + u) * 12 + k) * C + c] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
      }
      for (; t < 289; ++t) {
        const int k = ch[t];
        if (k >= 0 && k < 12) acc += wp[((size_t)t * 12 + k) * C + c];
      }
    }
    acc += bias[c];
  }
  float o = acc;
  if (ln_g) {
    float mean = block_sum(c < C ? acc : 0.f, red) / (float)C;
    float d = c < C ? acc - mean : 0.f;
    float var = block_sum(d * d, red) / (float)C;
    o = d * rsqrtf(var + 1e-5f) * (c < C ? ln_g[c] : 0.f) + (c < C ? ln_b[c] : 0.f);
  }
  if (c < C) {
    if (out) out[(long long)tok * ldo + c] = f2t(o);
    if (out2) out2[(long long)tok * ldo23 + c] = f2t(o);       // the same embedding, written where later layers read it
    if (out3) out3[(long long)tok * ldo23 + c] = f2t(o);
    if (out_f32) out_f32[(long long)tok * C + c] = o;
  }
}

// GRU_MEMORY ablation (ConvGRUCell.forward, transformer.py:84-100), elementwise halves around the two convolutions:
//   reset:  comb[:, C:2C] = t16( sigmoid(gates[:, 0:C]) * h )            (input of conv_can next to the unchanged x half)
//   blend:  h = (1 - u) h + u tanh(cand),  u = sigmoid(gates[:, C:2C]);  h16 = t16(h)
__global__ void gru_reset_kernel(const float* __restrict__ gates, long long ldg, const float* __restrict__ h,
                                 t16* __restrict__ comb_h, long long ldc, int P, int C) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)P * C) return;
  const long long r = i / C;
  const int c = (int)(i - r * C);
  const float g = gates[r * ldg + c];
  const float reset = 1.f / (1.f + expf(-g));
  comb_h[r * ldc + c] = f2t(reset * h[r * C + c]);
}
__global__ void gru_blend_kernel(const float* __restrict__ gates_u, long long ldg, const float* __restrict__ cand,
                                 float* __restrict__ h, t16* __restrict__ h16, int P, int C) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)P * C) return;
  const long long r = i / C;
  const int c = (int)(i - r * C);
  const float u = 1.f / (1.f + expf(-gates_u[r * ldg + c]));
  const float hn = (1.f - u) * h[i] + u * tanhf(cand[i]);
  h[i] = hn;
  h16[i] = f2t(hn);
}

// ------------------------------------------------------------------------------------------------
struct LogitPtrs { const float* p[4]; };

__device__ __forceinline__ void softmax11(const float* x, float* p) {
  float m = x[0];
#pragma unroll
  for (int c = 1; c < 11; ++c) m = fmaxf(m, x[c]);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 11; ++c) { p[c] = expf(__fsub_rn(x[c], m)); s = __fadd_rn(s, p[c]); }
#pragma unroll
  for (int c = 0; c < 11; ++c) p[c] = __fdiv_rn(p[c], s);
}

__global__ void mask_head_kernel(LogitPtrs lp, int k, int h4, int w4, int Ho, int Wo, float* __restrict__ out_logits,
                                 uint8_t* __restrict__ out_label) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Ho * Wo) return;
  int oy = i / Wo, ox = i - oy * Wo;
  int y0, y1, x0, x1;
  float wy0, wy1, wx0, wx1;
  src_index(oy, h4, Ho, y0, y1, wy0, wy1);
  src_index(ox, w4, Wo, x0, x1, wx0, wx1);
  const size_t plane = (size_t)h4 * w4, oplane = (size_t)Ho * Wo;
  const size_t i00 = (size_t)y0 * w4 + x0, i01 = (size_t)y0 * w4 + x1, i10 = (size_t)y1 * w4 + x0,
               i11 = (size_t)y1 * w4 + x1;
  if (k == 1) {
    float lg[11], p[11];
    const float* L = lp.p[0];
#pragma unroll
    for (int c = 0; c < 11; ++c) {
      const float* pl = L + c * plane;
      lg[c] = bilerp(pl[i00], pl[i01], pl[i10], pl[i11], wy0, wy1, wx0, wx1);
      if (out_logits) out_logits[c * oplane + i] = lg[c];
    }
    softmax11(lg, p);
    int best = 0;
#pragma unroll
    for (int c = 1; c < 11; ++c)
      if (p[c] > p[best]) best = c;
    if (out_label) out_label[i] = (uint8_t)best;
    return;
  }
  // soft aggregation over k engines (aot_engine.py:650-673)
  float merged[41];
  float bg = 1.f;
  for (int e = 0; e < k; ++e) {
    float lg[11], p[11];
    const float* L = lp.p[e];
#pragma unroll
    for (int c = 0; c < 11; ++c) {
      const float* pl = L + c * plane;
      lg[c] = bilerp(pl[i00], pl[i01], pl[i10], pl[i11], wy0, wy1, wx0, wx1);
    }
    softmax11(lg, p);
    bg = (e == 0) ? p[0] : __fmul_rn(bg, p[0]);
    for (int c = 1; c < 11; ++c) merged[1 + e * 10 + (c - 1)] = p[c];
  }
  merged[0] = bg;
  const int nc = 1 + 10 * k;
  float mx = -INFINITY;
  for (int c = 0; c < nc; ++c) {
    float m = fminf(fmaxf(merged[c], 1e-5f), 1.f - 1e-5f);
    float lg = logf(__fdiv_rn(m, __fsub_rn(1.f, m)));
    merged[c] = lg;
    if (out_logits) out_logits[c * oplane + i] = lg;
    mx = fmaxf(mx, lg);
  }
  float s = 0.f;
  for (int c = 0; c < nc; ++c) { merged[c] = expf(__fsub_rn(merged[c], mx)); s = __fadd_rn(s, merged[c]); }
  int best = 0;
  float bp = __fdiv_rn(merged[0], s);
  for (int c = 1; c < nc; ++c) {
    float p = __fdiv_rn(merged[c], s);
    if (p > bp) { bp = p; best = c; }
  }
  if (out_label) out_label[i] = (uint8_t)best;
}

// ------------------------------------------------------------------------------------------------
// Test-time augmentation head (evaluator.py:338-441): every augmentation a (scale x flip) has its own engine and its
// own 1/4-res logits; per output pixel each one is upsampled (bilinear, align_corners), soft-aggregated over its k object
// groups, soft-maxed, mirrored back if the augmentation was flipped, and the PROBABILITIES are averaged before the argmax.
struct TtaPtrs { const float* p[8][4]; int h4[8], w4[8], flip[8]; };
__global__ void tta_head_kernel(TtaPtrs tp, int n_aug, int k, int Ho, int Wo, float* __restrict__ out_prob,
                                uint8_t* __restrict__ out_label) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Ho * Wo) return;
  const int oy = i / Wo, ox = i - oy * Wo;
  const int nc = 1 + 10 * k;
  float mean[41];
  for (int c = 0; c < nc; ++c) mean[c] = 0.f;
  for (int a = 0; a < n_aug; ++a) {
    const int h4 = tp.h4[a], w4 = tp.w4[a];
    const int sx = tp.flip[a] ? (Wo - 1 - ox) : ox;        // flip_tensor(pred_logit, 3) on the full-size logits
    int y0, y1, x0, x1;
    float wy0, wy1, wx0, wx1;
    src_index(oy, h4, Ho, y0, y1, wy0, wy1);
    src_index(sx, w4, Wo, x0, x1, wx0, wx1);
    const size_t plane = (size_t)h4 * w4;
    const size_t i00 = (size_t)y0 * w4 + x0, i01 = (size_t)y0 * w4 + x1, i10 = (size_t)y1 * w4 + x0,
                 i11 = (size_t)y1 * w4 + x1;
    float merged[41];
    if (k == 1) {
      float lg[11];
#pragma unroll
      for (int c = 0; c < 11; ++c) {
        const float* pl = tp.p[a][0] + c * plane;
        lg[c] = bilerp(pl[i00], pl[i01], pl[i10], pl[i11], wy0, wy1, wx0, wx1);
      }
      softmax11(lg, merged);
    } else {
      float bg = 1.f;
      for (int e = 0; e < k; ++e) {
        float lg[11], p[11];
#pragma unroll
        for (int c = 0; c < 11; ++c) {
          const float* pl = tp.p[a][e] + c * plane;
          lg[c] = bilerp(pl[i00], pl[i01], pl[i10], pl[i11], wy0, wy1, wx0, wx1);
        }
        softmax11(lg, p);
        bg = (e == 0) ? p[0] : __fmul_rn(bg, p[0]);
        for (int c = 1; c < 11; ++c) merged[1 + e * 10 + (c - 1)] = p[c];
      }
      merged[0] = bg;
      float mx = -INFINITY;
      for (int c = 0; c < nc; ++c) {
        float m = fminf(fmaxf(merged[c], 1e-5f), 1.f - 1e-5f);
        merged[c] = logf(__fdiv_rn(m, __fsub_rn(1.f, m)));
        mx = fmaxf(mx, merged[c]);
      }
      float s = 0.f;
      for (int c = 0; c < nc; ++c) { merged[c] = expf(__fsub_rn(merged[c], mx)); s = __fadd_rn(s, merged[c]); }
      for (int c = 0; c < nc; ++c) merged[c] = __fdiv_rn(merged[c], s);
    }
    for (int c = 0; c < nc; ++c) mean[c] = __fadd_rn(mean[c], merged[c]);
  }
  const float inv = __fdiv_rn(1.f, (float)n_aug);
  int best = 0;
  float bp = -1.f;
  for (int c = 0; c < nc; ++c) {
    const float p = __fmul_rn(mean[c], inv);
    if (out_prob) out_prob[(size_t)c * Ho * Wo + i] = p;
    if (p > bp) { bp = p; best = c; }
  }
  if (out_label) out_label[i] = (uint8_t)best;
}

// Frame preprocessing of the evaluation loader on the GPU (MultiRestrictSize + MultiToTensor,
// dataloaders/video_transforms.py:559-682): uint8 HWC frame -> fp32 NCHW at the network size, cubic resize with
// OpenCV's INTER_CUBIC rule (a = -0.75, half-pixel centres, no antialiasing, taps clamped to the image), optional
// horizontal flip, x / 255, ImageNet mean / std.  `bgr`: the frame is in cv2.imread order and the network wants RGB.
__device__ __forceinline__ void cubic_w(float t, float* w) {
  const float A = -0.75f;
  w[0] = ((A * (t + 1.f) - 5.f * A) * (t + 1.f) + 8.f * A) * (t + 1.f) - 4.f * A;
  w[1] = ((A + 2.f) * t - (A + 3.f)) * t * t + 1.f;
  w[2] = ((A + 2.f) * (1.f - t) - (A + 3.f)) * (1.f - t) * (1.f - t) + 1.f;
  w[3] = 1.f - w[0] - w[1] - w[2];
}
__global__ void preprocess_kernel(const uint8_t* __restrict__ img, int H, int W, int bgr, int nh, int nw, int flip,
                                  float* __restrict__ out) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nh * nw) return;
  const int oy = i / nw, ox0 = i - oy * nw;
  const int ox = flip ? (nw - 1 - ox0) : ox0;                // tmp[:, ::-1] after the resize
  float v[3];
  if (nh == H && nw == W) {
    const uint8_t* p = img + ((size_t)oy * W + ox) * 3;
    v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
  } else {
    const float fy = ((float)oy + 0.5f) * ((float)H / (float)nh) - 0.5f;
    const float fx = ((float)ox + 0.5f) * ((float)W / (float)nw) - 0.5f;
    const int iy = (int)floorf(fy), ix = (int)floorf(fx);
    float wy[4], wx[4];
    cubic_w(fy - (float)iy, wy);
    cubic_w(fx - (float)ix, wx);
    v[0] = v[1] = v[2] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), H - 1);
      float r[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int xx = min(max(ix - 1 + b, 0), W - 1);
        const uint8_t* p = img + ((size_t)yy * W + xx) * 3;
        r[0] += wx[b] * (float)p[0]; r[1] += wx[b] * (float)p[1]; r[2] += wx[b] * (float)p[2];
      }
      v[0] += wy[a] * r[0]; v[1] += wy[a] * r[1]; v[2] += wy[a] * r[2];
    }
  }
  const float mean[3] = {0.485f, 0.456f, 0.406f}, sd[3] = {0.229f, 0.224f, 0.225f};
  const size_t plane = (size_t)nh * nw;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float x = v[bgr ? 2 - c : c];
    out[c * plane + (size_t)oy * nw + ox0] = (x / 255.f - mean[c]) / sd[c];
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void evict_rel_kernel(const float* __restrict__ mass, int T, const float* __restrict__ logits4, int h4,
                                 int w4, int h, int w, float* __restrict__ rel) {
  pdl_prologue();
  __shared__ float red[32];
  float acc[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) acc[t] = 0.f;
  const size_t plane = (size_t)h4 * w4;
  for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
    int oy = i / w, ox = i - oy * w;
    int y0, y1, x0, x1;
    float wy0, wy1, wx0, wx1;
    src_index(oy, h4, h, y0, y1, wy0, wy1);
    src_index(ox, w4, w, x0, x1, wx0, wx1);
    float lg[11], p[11];
#pragma unroll
    for (int c = 0; c < 11; ++c) {
      const float* pl = logits4 + c * plane;
      lg[c] = bilerp(pl[(size_t)y0 * w4 + x0], pl[(size_t)y0 * w4 + x1], pl[(size_t)y1 * w4 + x0],
                     pl[(size_t)y1 * w4 + x1], wy0, wy1, wx0, wx1);
    }
    softmax11(lg, p);
    float fg = 1.f - p[0];
#pragma unroll
    for (int t = 0; t < 16; ++t)
      if (t < T) acc[t] += mass[(size_t)i * T + t] * fg;
  }
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    if (t < T) {
      float v = block_sum(acc[t], red);
      if (threadIdx.x == 0) rel[t] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// ---- small elementwise kernels of the AOT block (transformer.py:553-692) ----
__global__ void add_t16_kernel(const t16* __restrict__ a, long long lda, const t16* __restrict__ b, long long ldb,
                               t16* __restrict__ y, long long ldy, int P, int C) {
  pdl_prologue();
  const int cv = C / 8;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)P * cv) return;
  int c8 = (int)(i % cv);
  long long r = i / cv;
  float u[8], v[8];
  load8(a + r * lda + c8 * 8, u);
  load8(b + r * ldb + c8 * 8, v);
#pragma unroll
  for (int k = 0; k < 8; ++k) u[k] += v[k];
  store8(y + r * ldy + c8 * 8, u);
}

// y = LayerNorm(a + b): one warp per row, C <= 512
__global__ void add_layernorm_kernel(const t16* __restrict__ a, long long lda, const t16* __restrict__ b,
                                     long long ldb, const float* __restrict__ gamma, const float* __restrict__ beta,
                                     t16* __restrict__ y, long long ldy, int P, int C) {
  pdl_prologue();
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= P) return;
  float v[LN_MAXV];
  const int nv = C / 32;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAXV; ++j)
    if (j < nv) {
      v[j] = t2f(a[(long long)row * lda + j * 32 + lane]) + t2f(b[(long long)row * ldb + j * 32 + lane]);
      s += v[j];
    }
  float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAXV; ++j)
    if (j < nv) { float d = v[j] - mean; q += d * d; }
  float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
#pragma unroll
  for (int j = 0; j < LN_MAXV; ++j)
    if (j < nv) {
      int c = j * 32 + lane;
      y[(long long)row * ldy + c] = f2t((v[j] - mean) * rstd * gamma[c] + beta[c]);
    }
}

__global__ void accum_t16_kernel(const t16* __restrict__ x, long long ldx, float* __restrict__ y, long long ldy, int P,
                                 int C) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)P * C) return;
  long long r = i / C;
  int c = (int)(i - r * C);
  y[r * ldy + c] += t2f(x[r * ldx + c]);
}

__global__ void cvt_f32_t16_kernel(const float* __restrict__ x, long long ldx, t16* __restrict__ y, long long ldy,
                                   int P, int C) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)P * C) return;
  long long r = i / C;
  int c = (int)(i - r * C);
  y[r * ldy + c] = f2t(x[r * ldx + c]);
}

// PositionEmbeddingSine(num_pos_feats = C/2, normalize = True)  (position.py:35-77) -> fp32 [h*w, C]
__global__ void sine_pe_kernel(float* __restrict__ out, int h, int w, int C) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)h * w * C) return;
  const int c = (int)(i % C);
  const int tok = (int)(i / C);
  const int py = tok / w, px = tok - py * w;
  const int npf = C / 2;
  const bool is_x = c >= npf;
  const int k = is_x ? c - npf : c;
  const float pos = is_x ? (float)px / ((float)(w - 1) + 1e-6f) : (float)py / ((float)(h - 1) + 1e-6f);
  const float e = pos * 6.283185307179586f;
  const float dim_t = powf(10000.f, (float)(2 * (k / 2)) / (float)npf);
  const float a = e / dim_t;
  out[i] = (k & 1) ? cosf(a) : sinf(a);
}

// mean over heads of the per-head attention mass: in [H][P][T] -> out [P][T]
__global__ void mean_heads_kernel(const float* __restrict__ in, float* __restrict__ out, int H, long long n) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int h = 0; h < H; ++h) s += in[(long long)h * n + i];
  out[i] = s / (float)H;
}

struct PeSlots { int s[16]; };
// AOT: qt = t16(q + pe_cur); qbias[h][i][t] = scale * <qt_i[h*dh .. (h+1)*dh), pe_mem[slot(t)][h*dh ..]>
// one warp per row, C = 256, H = 8 heads of 32: lane j of chunk h covers channel h*32 + lane
__global__ void qprep_heads_kernel(const t16* __restrict__ q, long long ldq, const float* __restrict__ pe_cur,
                                   const float* __restrict__ pe_mem, PeSlots ps, int T, float scale,
                                   t16* __restrict__ qt, float* __restrict__ qbias, int P, int H) {
  pdl_prologue();
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= P) return;
  const int C = H * 32;
  for (int h = 0; h < H; ++h) {
    const int c = h * 32 + lane;
    t16 r = f2t(t2f(q[(long long)row * ldq + c]) + pe_cur[c]);
    qt[(long long)row * C + c] = r;
    const float v = t2f(r);
    for (int t = 0; t < T; ++t) {
      float s = warp_sum(v * pe_mem[ps.s[t] * C + c]);
      if (lane == 0) qbias[((long long)h * P + row) * T + t] = s * scale;
    }
  }
}

__global__ void qprep_kernel(const t16* __restrict__ q, long long ldq, const float* __restrict__ pe_cur,
                             const float* __restrict__ pe_mem, PeSlots ps, int T, float scale, t16* __restrict__ qt,
                             float* __restrict__ qbias, int P, int C) {
  pdl_prologue();
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= P) return;
  float v[8];
  const int nv = C / 32;  // <= 8
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) {
      int c = j * 32 + lane;
      float x = t2f(q[(long long)row * ldq + c]) + (pe_cur ? pe_cur[c] : 0.f);
      t16 r = f2t(x);
      qt[(long long)row * C + c] = r;
      v[j] = t2f(r);
    }
  for (int t = 0; t < T; ++t) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < nv) s += v[j] * pe_mem[ps.s[t] * C + j * 32 + lane];
    s = warp_sum(s);
    if (lane == 0) qbias[(long long)row * T + t] = s * scale;
  }
}

}  // namespace

// ================================================================================================
int pack_image(const float* img, t16* out, int H, int W, cudaStream_t s) {
  RMEM_CUDA_CHECK(launch_pdl(pack_image_kernel, dim3(cdiv(H * W, 256)), dim3(256), 0, s, img, out, H * W));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int pack_image_padded(const float* img, t16* out, int H, int W, cudaStream_t s) {
  RMEM_CUDA_CHECK(launch_pdl(pack_image_padded_kernel, dim3(cdiv(H * W, 256)), dim3(256), 0, s, img, out, H, W));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int maxpool3x3s2(const t16* x, t16* y, int Hin, int Win, int C, int Hout, int Wout, cudaStream_t s) {
  RMEM_REQUIRE(C % 8 == 0, "maxpool: C %% 8");
  long long n = (long long)Hout * Wout * (C / 8);
  RMEM_CUDA_CHECK(launch_pdl(maxpool_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, x, y, Hin, Win, C, Hout, Wout));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int layernorm(const float* x, long long ldx, const float* gamma, const float* beta, t16* y, long long ldy, t16* y2,
              long long ldy2, int P, int C, cudaStream_t s, const float* add2, float* init_res, long long ld_init) {
  RMEM_REQUIRE(C % 32 == 0 && C <= 32 * LN_MAXV, "layernorm: unsupported C=%d", C);
  RMEM_REQUIRE(!init_res || ld_init >= 2 * C, "layernorm: init_res needs a row stride >= 2C");
  RMEM_CUDA_CHECK(launch_pdl(layernorm_kernel, dim3(cdiv(P, 8)), dim3(256), 0, s, x, ldx, gamma, beta, y, ldy, y2, ldy2, add2, P, C,
                             init_res, ld_init));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int layernorm_pair(const float* x, long long ldx, const float* g0, const float* b0, const float* g1, const float* b1,
                   t16* y, long long ldy, int P, int C, cudaStream_t s, t16* y1, long long ldy1) {
  RMEM_REQUIRE(C % 32 == 0 && C <= 32 * LN_MAXV, "layernorm_pair: unsupported C=%d", C);
  RMEM_CUDA_CHECK(launch_pdl(layernorm_pair_kernel, dim3(cdiv(2 * P, 8)), dim3(256), 0, s, x, ldx, g0, b0, g1, b1, y, ldy,
                             y1, ldy1, P, C));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int groupnorm_t16(const t16* x, const float* gamma, const float* beta, t16* y, int P, int C, int G, int relu,
                   double* stats, cudaStream_t s, const t16* add, bool stats_ready) {
  return groupnorm_impl<t16>(x, gamma, beta, y, P, C, G, relu, stats, s, add, stats_ready);
}
int groupnorm_f32(const float* x, const float* gamma, const float* beta, t16* y, int P, int C, int G, int relu,
                  double* stats, cudaStream_t s) {
  return groupnorm_impl<float>(x, gamma, beta, y, P, C, G, relu, stats, s);
}

int dwconv5x5(const t16* x, const float* w, t16* y, int h, int wd, int C, cudaStream_t s, int ldx, int ldy) {
  RMEM_REQUIRE(C % 64 == 0, "dwconv: C %% 64");
  if (ldx <= 0) ldx = C;
  if (ldy <= 0) ldy = C;
  RMEM_REQUIRE(ldx % 2 == 0 && ldy % 2 == 0 && ldx >= C && ldy >= C, "dwconv: ldx=%d ldy=%d", ldx, ldy);
  // tile width: wider tiles re-read less halo, narrower ones give more blocks and a shorter serial chain per warp
  // (RMEM_DW_TW = 6 | 9 | 18 | 27 for A/B; c3's 31 x 54 token grid: 576 / 384 / 192 / 128 blocks)
  static const int tw = [] { const char* e = getenv("RMEM_DW_TW"); const int v = e ? atoi(e) : 18;
                             return (v == 6 || v == 9 || v == 18 || v == 27) ? v : 18; }();
  const dim3 grid(C / 64, cdiv(wd, tw), cdiv(h, DW_TH));
  if (tw == 6) RMEM_CUDA_CHECK(launch_pdl(dwconv5_kernel<6>, grid, dim3(256), 0, s, x, ldx, w, y, ldy, h, wd, C));
  else if (tw == 9) RMEM_CUDA_CHECK(launch_pdl(dwconv5_kernel<9>, grid, dim3(256), 0, s, x, ldx, w, y, ldy, h, wd, C));
  else if (tw == 27) RMEM_CUDA_CHECK(launch_pdl(dwconv5_kernel<27>, grid, dim3(256), 0, s, x, ldx, w, y, ldy, h, wd, C));
  else RMEM_CUDA_CHECK(launch_pdl(dwconv5_kernel<18>, grid, dim3(256), 0, s, x, ldx, w, y, ldy, h, wd, C));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int upsample_bilinear_t16(const t16* x, t16* y, int hin, int win, int hout, int wout, int C, cudaStream_t s,
                          const t16* add, const double* gn_stats, const float* gamma, const float* beta, int G) {
  RMEM_REQUIRE(C % 8 == 0, "upsample: C %% 8");
  RMEM_REQUIRE(!gn_stats || (gamma && beta && G > 0 && C % G == 0 && (C / G) % 8 == 0), "upsample: GroupNorm arguments (C=%d G=%d)", C, G);
  long long n = (long long)hout * wout * (C / 8);
  RMEM_CUDA_CHECK(launch_pdl(upsample_t16_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, x, y, hin, win, hout, wout, C, add,
                             gn_stats, gamma, beta, G));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int conv_out_logits(const t16* x, const t16* w, const float* b, float* out, int P, int Cin, int Cout,
                    cudaStream_t s) {
  RMEM_REQUIRE(Cin % 32 == 0 && Cin <= 128 && Cout <= 16, "conv_out: unsupported Cin=%d Cout=%d", Cin, Cout);
  const size_t smem = (size_t)(Cin / 8) * 16 * 8 * sizeof(float) + (size_t)128 * (Cin * 2 + 16);
  RMEM_CUDA_CHECK(launch_pdl(conv_out_kernel, dim3(cdiv(P, 128)), dim3(128), smem, s, x, w, b, out, P, Cin, Cout));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int conv_out_gn_logits(const t16* x, const float* gamma, const float* beta, int G, double* stats, const t16* w,
                       const float* b, float* out, int P, int Cin, int Cout, cudaStream_t s, bool stats_ready) {
  RMEM_REQUIRE(Cin % 32 == 0 && Cin <= 128 && Cout <= 16 && (Cin / 8) % 2 == 0, "conv_out_gn: unsupported Cin=%d Cout=%d", Cin, Cout);
  RMEM_REQUIRE(G <= 32 && Cin % G == 0 && (Cin / G) % 8 == 0 && 256 % (Cin / 8) == 0, "conv_out_gn: unsupported C=%d G=%d", Cin, G);
  if (!stats_ready) {
    const int rows_per_block = 256 / (Cin / 8);
    const int grid = min(cdiv(P, rows_per_block), kGnMaxBlocks);
    RMEM_CUDA_CHECK(launch_pdl(gn_stats_kernel<t16>, dim3(grid), dim3(256), 0, s, x, P, Cin, G, stats, stats + 72, reinterpret_cast<unsigned int*>(stats + 64)));
    RMEM_LAUNCH_CHECK();
  }
  const size_t smem = (size_t)(Cin / 8) * 16 * 8 * sizeof(float) + (size_t)2 * Cin * sizeof(float) + (size_t)64 * (Cin * 2 + 16);
  RMEM_CUDA_CHECK(launch_pdl(conv_out_gn_kernel, dim3(cdiv(P, 64)), dim3(128), smem, s, x, w, b, gamma, beta,
                             static_cast<const double*>(stats), G, out, P, Cin, Cout));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int gru_reset(const float* gates, long long ldg, const float* h, t16* comb_h, long long ldc, int P, int C, cudaStream_t s) {
  const long long n = (long long)P * C;
  RMEM_CUDA_CHECK(launch_pdl(gru_reset_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, gates, ldg, h, comb_h, ldc, P, C));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}
int gru_blend(const float* gates_u, long long ldg, const float* cand, float* h, t16* h16, int P, int C, cudaStream_t s) {
  const long long n = (long long)P * C;
  RMEM_CUDA_CHECK(launch_pdl(gru_blend_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, gates_u, ldg, cand, h, h16, P, C));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int transpose_t16(const t16* x, long long ldx, t16* y, long long ldy, int P, int C, cudaStream_t s) {
  dim3 grid(cdiv(P, 64), cdiv(C, 64)), block(32, 8);
  RMEM_CUDA_CHECK(launch_pdl(transpose_kernel, dim3(grid), dim3(block), 0, s, x, ldx, y, ldy, P, C));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int copy2d_t16(const t16* src, long long lds, t16* dst, long long ldd, int P, int C, cudaStream_t s) {
  RMEM_REQUIRE(C % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0, "copy2d: alignment");
  long long n = (long long)P * (C / 8);
  RMEM_CUDA_CHECK(launch_pdl(copy2d_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, src, lds, dst, ldd, P, C));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}
int fill_t16(t16* dst, long long ldd, int P, int C, float v, cudaStream_t s) {
  long long n = (long long)P * C;
  RMEM_CUDA_CHECK(launch_pdl(fill_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, dst, ldd, P, C, v));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int separate_label(const void* label, int label_is_f32, uint8_t* out, int H, int W, int engine, int n_engines,
                   cudaStream_t s) {
  RMEM_CUDA_CHECK(launch_pdl(separate_label_kernel, dim3(cdiv(H * W, 256)), dim3(256), 0, s, label, label_is_f32, out, H * W, engine, n_engines));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int idbank_embed(const uint8_t* label, int H, int W, int use_ignore, const float* w_packed, const float* bias,
                 const float* ln_g, const float* ln_b, t16* out, long long ldo, float* out_f32, int h, int w, int C,
                 cudaStream_t s, const float* prefix, const float* prefix_rows, t16* out2, t16* out3, long long ldo23) {
  RMEM_REQUIRE(C <= 256 && C % 32 == 0, "idbank: unsupported C=%d", C);
  RMEM_REQUIRE(prefix || !prefix_rows, "idbank: the row-prefix table needs the 2-D prefix table as well");
  RMEM_CUDA_CHECK(launch_pdl(idbank_kernel, dim3(h * w), dim3(256), 0, s, label, H, W, use_ignore, w_packed, prefix, prefix_rows, bias, ln_g, ln_b, out, ldo, out_f32, w, C, out2, out3, ldo23));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int mask_head(const float* const* logits4, int k, int h4, int w4, int Ho, int Wo, float* out_logits,
              uint8_t* out_label, cudaStream_t s) {
  RMEM_REQUIRE(k >= 1 && k <= 4, "mask_head: 1..4 engines supported, got %d", k);
  LogitPtrs lp;
  for (int e = 0; e < 4; ++e) lp.p[e] = e < k ? logits4[e] : nullptr;
  RMEM_CUDA_CHECK(launch_pdl(mask_head_kernel, dim3(cdiv(Ho * Wo, 128)), dim3(128), 0, s, lp, k, h4, w4, Ho, Wo, out_logits, out_label));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int tta_head(const float* const* logits4, int n_aug, int k, const int* h4, const int* w4, const int* flip, int Ho, int Wo,
             float* out_prob, uint8_t* out_label, cudaStream_t s) {
  RMEM_REQUIRE(n_aug >= 1 && n_aug <= 8 && k >= 1 && k <= 4, "tta_head: %d augmentations x %d object groups (max 8 x 4)", n_aug, k);
  TtaPtrs tp;
  for (int a = 0; a < 8; ++a) {
    tp.h4[a] = a < n_aug ? h4[a] : 0; tp.w4[a] = a < n_aug ? w4[a] : 0; tp.flip[a] = a < n_aug ? flip[a] : 0;
    for (int e = 0; e < 4; ++e) tp.p[a][e] = (a < n_aug && e < k) ? logits4[a * k + e] : nullptr;
  }
  RMEM_CUDA_CHECK(launch_pdl(tta_head_kernel, dim3(cdiv(Ho * Wo, 128)), dim3(128), 0, s, tp, n_aug, k, Ho, Wo, out_prob, out_label));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int preprocess_frame(const uint8_t* img, int H, int W, int bgr, int nh, int nw, int flip, float* out, cudaStream_t s) {
  RMEM_REQUIRE(H > 0 && W > 0 && nh > 0 && nw > 0, "preprocess: empty frame");
  RMEM_CUDA_CHECK(launch_pdl(preprocess_kernel, dim3(cdiv(nh * nw, 256)), dim3(256), 0, s, img, H, W, bgr, nh, nw, flip, out));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int evict_relevance(const float* mass, int T, const float* logits4, int h4, int w4, int h, int w, float* rel,
                    cudaStream_t s) {
  RMEM_REQUIRE(T >= 1 && T <= 16, "evict_relevance: T=%d out of range", T);
  RMEM_CUDA_CHECK(launch_pdl(evict_rel_kernel, dim3(1), dim3(1024), 0, s, mass, T, logits4, h4, w4, h, w, rel));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int add_t16(const t16* a, long long lda, const t16* b, long long ldb, t16* y, long long ldy, int P, int C,
            cudaStream_t s) {
  RMEM_REQUIRE(C % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldy % 8 == 0, "add_t16: alignment");
  long long n = (long long)P * (C / 8);
  RMEM_CUDA_CHECK(launch_pdl(add_t16_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, a, lda, b, ldb, y, ldy, P, C));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int add_layernorm_t16(const t16* a, long long lda, const t16* b, long long ldb, const float* gamma, const float* beta,
                      t16* y, long long ldy, int P, int C, cudaStream_t s) {
  RMEM_REQUIRE(C % 32 == 0 && C <= 32 * LN_MAXV, "add_layernorm: unsupported C=%d", C);
  RMEM_CUDA_CHECK(launch_pdl(add_layernorm_kernel, dim3(cdiv(P, 8)), dim3(256), 0, s, a, lda, b, ldb, gamma, beta, y, ldy, P, C));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int accum_t16_into_f32(const t16* x, long long ldx, float* y, long long ldy, int P, int C, cudaStream_t s) {
  long long n = (long long)P * C;
  RMEM_CUDA_CHECK(launch_pdl(accum_t16_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, x, ldx, y, ldy, P, C));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int cvt_f32_t16(const float* x, long long ldx, t16* y, long long ldy, int P, int C, cudaStream_t s) {
  long long n = (long long)P * C;
  RMEM_CUDA_CHECK(launch_pdl(cvt_f32_t16_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, x, ldx, y, ldy, P, C));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int sine_pos_emb(float* out, int h, int w, int C, cudaStream_t s) {
  RMEM_REQUIRE(C % 4 == 0 && h > 1 && w > 1, "sine_pos_emb: unsupported geometry");
  long long n = (long long)h * w * C;
  RMEM_CUDA_CHECK(launch_pdl(sine_pe_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, out, h, w, C));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int mean_heads(const float* in, float* out, int H, long long n, cudaStream_t s) {
  RMEM_CUDA_CHECK(launch_pdl(mean_heads_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, in, out, H, n));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

int qprep_heads(const t16* q, long long ldq, const float* pe_cur, const float* pe_mem, const int* pe_slot, int T,
                float scale, t16* qt, float* qbias, int P, int H, cudaStream_t s) {
  RMEM_REQUIRE(T >= 1 && T <= 16 && H >= 1 && H <= 16, "qprep_heads: T=%d H=%d", T, H);
  PeSlots ps;
  for (int t = 0; t < 16; ++t) ps.s[t] = (t < T && pe_slot) ? pe_slot[t] : 0;
  RMEM_CUDA_CHECK(launch_pdl(qprep_heads_kernel, dim3(cdiv(P, 8)), dim3(256), 0, s, q, ldq, pe_cur, pe_mem, ps, T, scale, qt, qbias, P, H));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

void temporal_pe_slots(int T, int n_slots, int* out) {
  if (T <= n_slots) {
    for (int t = 0; t < T; ++t) out[t] = t;
    return;
  }
  const float scale = (float)n_slots / (float)T;  // ATen nearest: src = min(floor(dst * scale), in - 1)
  for (int t = 0; t < T; ++t) {
    int i = T - 1 - t;
    int src = (int)floorf((float)i * scale);
    if (src > n_slots - 1) src = n_slots - 1;
    out[t] = n_slots - 1 - src;
  }
}

int qprep(const t16* q, long long ldq, const float* pe_cur, const float* pe_mem, const int* pe_slot, int T,
          float scale, t16* qt, float* qbias, int P, int C, cudaStream_t s) {
  RMEM_REQUIRE(C % 32 == 0 && C <= 256, "qprep: unsupported C=%d", C);
  RMEM_REQUIRE(T >= 0 && T <= 16, "qprep: T=%d", T);
  PeSlots ps;
  for (int t = 0; t < 16; ++t) ps.s[t] = (t < T && pe_slot) ? pe_slot[t] : 0;
  RMEM_CUDA_CHECK(launch_pdl(qprep_kernel, dim3(cdiv(P, 8)), dim3(256), 0, s, q, ldq, pe_cur, pe_mem, ps, T, scale, qt, qbias, P, C));
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
}

}  // namespace rmem
