// Training-side loss head (first slice): value AND gradient of the reference's per-frame training
// loss with respect to the 1/4-resolution ID logits the engine's decoder leaves behind.
//
//   AOTEngine.calculate_current_loss        networks/engines/aot_engine.py:484-511
//     bilinear (align_corners) upsampling of pred_id_logits to the label size, channels 0 .. obj_num
//   CrossEntropyLoss.forward, top-k branch  networks/layers/loss.py:163-211
//     per-pixel cross entropy (ignore_index 255 -> 0), mean of the top_k largest of ALL H*W values
//   SoftJaccordLoss.forward / tversky_loss  networks/layers/loss.py:30-74, 136-160   (alpha = beta = 1, eps = 1e-6)
//     over the pixels that are not 255, for every class that owns at least one pixel: 1 - I / (I + A + B + eps), mean
//   loss = 0.5 * ce + 0.5 * jaccard         aot_engine.py:141-142
//
// HBM-bound pixel work, fp32, IEEE expf / logf / division (this file is compiled WITHOUT --use_fast_math).  Every
// reduction runs in a fixed order (warp shuffles -> per-block partials -> one finalising block, integer histograms for the
// top-k selection), so the loss and the gradient are bit-reproducible from run to run.
//
//   tl_pixel_kernel     per output pixel: 4-tap upsample, softmax, cross entropy -> ce[P]; per-block class sums
//                       (I_c = sum p_c [g = c], S_c = sum p_c, N_c = #[g = c] over the valid pixels); histogram of the
//                       leading 11 bits of ce.  Last block: class sums folded, first digit of the threshold picked.
//   tl_hist_kernel x 2  the other two passes of an exact radix select of the k-th largest ce on the fp32 bit patterns
//                       (11 + 11 + 10 bits; ce >= 0, so the patterns order like the values); warp-aggregated
//                       shared-memory histograms, integer atomics only.  Last block: next digit (scan_select).
//   tl_topk_sum_kernel  sum of the values above the threshold (ties at the threshold enter as k_rem * threshold).
//                       Last block: losses, Jaccard coefficients per class, tie weight.
//   tl_grad_pixel       d loss / d upsampled logits per pixel (softmax recomputed), channel-major scratch
//   tl_grad_gather      transpose of the upsampling as a GATHER per 1/4-res logit (fixed order; no float atomics)
// "Last block": the block that takes the last ticket of a per-launch counter folds what all blocks left in global memory
// (__threadfence + atomic ticket, as gn_stats_kernel in ops.cu): 4 launches for the value, 6 with the gradient, where
// the first version needed 9 / 11 (a one-block scan after each histogram, a finalising launch).
#include "../../include/rmem_b200.h"
#include "common.cuh"

namespace rmem {
namespace {

constexpr int kMaxCh = 11;        // background + MODEL_MAX_OBJ_NUM
constexpr int kThreads = 256;
constexpr int kBins = 2048;
constexpr int kPartStride = 3 * kMaxCh;

struct SelectState {
  unsigned int prefix;            // leading bits of the k-th largest pattern resolved so far (right-aligned)
  unsigned int k_rem;             // rank, counted from the top, still to resolve inside that prefix bucket (>= 1)
  unsigned int n_ties;            // after the last pass: how many values equal the threshold
  unsigned int pad;
};
struct Coef {                     // written by tl_finalize_kernel, read by tl_grad_pixel_kernel
  float a[kMaxCh];                // d jaccard / d p_c at a pixel of class c
  float b[kMaxCh];                // d jaccard / d p_c at a valid pixel of another class
  float inv_k, tie_w;
  unsigned int thr_bits, pad;
};

struct Layout {
  size_t ce, gup, part, tpart, hist, state, counters, coef, cls, total;   // [hist, coef) is zeroed at the start of a call
  int P, nb;
};
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
inline Layout make_layout(int H, int W) {
  Layout L;
  L.P = H * W;
  L.nb = cdiv(L.P, kThreads);
  size_t o = 0;
  L.ce = o; o = align256(o + (size_t)L.P * 4);
  L.gup = o; o = align256(o + (size_t)L.P * kMaxCh * 4);
  L.part = o; o = align256(o + (size_t)L.nb * kPartStride * 4);
  L.tpart = o; o = align256(o + (size_t)L.nb * 8);
  L.hist = o; o = align256(o + (size_t)3 * kBins * 4);
  L.state = o; o = align256(o + sizeof(SelectState));
  L.counters = o; o = align256(o + 4 * sizeof(unsigned int));
  L.coef = o; o = align256(o + sizeof(Coef));
  L.cls = o; o = align256(o + kPartStride * sizeof(double));
  L.total = o;
  return L;
}

// align_corners=True source taps, as ATen's area_pixel_compute_source_index (the mask head uses the same arithmetic)
struct Axis { float scale; int in_size; };
__host__ __device__ inline Axis make_axis(int in_size, int out_size) {
  Axis a;
  a.in_size = in_size;
  a.scale = out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
  return a;
}
__device__ __forceinline__ void taps(const Axis& a, int dst, int& i0, int& i1, float& l0, float& l1) {
  const float real = __fmul_rn(a.scale, (float)dst);
  i0 = min((int)real, a.in_size - 1);
  const float lam = fminf(fmaxf(__fsub_rn(real, (float)i0), 0.f), 1.f);
  i1 = i0 + ((i0 < a.in_size - 1) ? 1 : 0);
  l1 = lam;
  l0 = __fsub_rn(1.f, lam);
}

// upsampled logits of one output pixel and their softmax; returns log(sum exp(x - m)) and m through the references
__device__ __forceinline__ void pixel_softmax(const float* __restrict__ lg, int h4, int w4, Axis ay, Axis ax, int oy,
                                              int ox, int n_ch, float* x, float* p, float& m, float& lse) {
  int y0, y1, x0, x1;
  float wy0, wy1, wx0, wx1;
  taps(ay, oy, y0, y1, wy0, wy1);
  taps(ax, ox, x0, x1, wx0, wx1);
  const size_t plane = (size_t)h4 * w4;
  m = -INFINITY;
#pragma unroll
  for (int c = 0; c < kMaxCh; ++c) {
    if (c < n_ch) {
      const float* L = lg + c * plane;
      const float top = __fadd_rn(__fmul_rn(wx0, L[y0 * w4 + x0]), __fmul_rn(wx1, L[y0 * w4 + x1]));
      const float bot = __fadd_rn(__fmul_rn(wx0, L[y1 * w4 + x0]), __fmul_rn(wx1, L[y1 * w4 + x1]));
      x[c] = __fadd_rn(__fmul_rn(wy0, top), __fmul_rn(wy1, bot));
      m = fmaxf(m, x[c]);
    } else {
      x[c] = 0.f;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxCh; ++c) {
    p[c] = c < n_ch ? expf(x[c] - m) : 0.f;
    s += p[c];
  }
  lse = logf(s);
  const float inv = 1.f / s;
#pragma unroll
  for (int c = 0; c < kMaxCh; ++c) p[c] *= inv;
}

// ---- pieces shared by the selection kernels ----
// Ticket of the calling block on `counter`; true for the block that arrives last, after which everything the other
// blocks wrote to global memory before their ticket is visible to it (read it with __ldcg).
__device__ __forceinline__ bool last_block_done(unsigned int* counter) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// One histogram vote per lane, aggregated over the lanes of the warp that vote for the same bin (ce values cluster in a
// few exponent bins: unaggregated shared-memory atomics would serialise 32 deep).  Every lane of the warp must call it;
// vote = false keeps a lane out.
__device__ __forceinline__ void hist_vote(unsigned int* sh, bool vote, unsigned int digit) {
  const unsigned int key = vote ? digit : 0xffffffffu;
  const unsigned int peers = __match_any_sync(0xffffffffu, key);
  if (vote && (int)(threadIdx.x & 31) == __ffs((int)peers) - 1) atomicAdd(&sh[digit], (unsigned int)__popc(peers));
}
__device__ __forceinline__ void hist_flush(const unsigned int* sh, unsigned int* hist) {
  for (int i = threadIdx.x; i < kBins; i += kThreads)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// The bin, counted from the top, in which rank k_rem falls -- by the whole (last) block: thread t owns bins
// [8t, 8t + 8), a suffix sum over the threads (warp shuffles + 8 warp totals) gives the count above each thread's bins,
// and the one thread whose bins straddle the rank walks its eight bins.
__device__ __forceinline__ void scan_select(const unsigned int* hist, int pass, SelectState* st, unsigned int k) {
  __shared__ unsigned int wtot[kThreads / 32];
  constexpr int per = kBins / kThreads;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  unsigned int h[per], s = 0;
#pragma unroll
  for (int j = 0; j < per; ++j) { h[j] = __ldcg(&hist[t * per + j]); s += h[j]; }
  const unsigned int k_rem = pass ? st->k_rem : k;
  const unsigned int prev = pass ? st->prefix : 0u;
  unsigned int v = s;                                      // -> sum over the lanes >= this one
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int n = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += n;
  }
  if (lane == 0) wtot[warp] = v;
  __syncthreads();                                         // also: every thread has read st before one of them writes it
  unsigned int above = v - s;
  for (int w = warp + 1; w < kThreads / 32; ++w) above += wtot[w];
  if (above < k_rem && k_rem <= above + s) {
    unsigned int cum = above, hit = h[0];
    int bin = 0;
    bool found = false;
#pragma unroll
    for (int j = per - 1; j >= 0; --j) {
      if (!found) {
        if (j == 0 || cum + h[j] >= k_rem) { found = true; bin = j; hit = h[j]; }
        else cum += h[j];
      }
    }
    const unsigned int b = (unsigned int)(t * per + bin);
    st->prefix = pass == 0 ? b : (prev << (pass == 1 ? 11 : 10)) | b;
    st->k_rem = k_rem - cum;
    st->n_ties = hit;
  }
}

__global__ void __launch_bounds__(kThreads) tl_pixel_kernel(const float* __restrict__ lg, int h4, int w4,
                                                            const uint8_t* __restrict__ gt, int H, int W, int n_ch,
                                                            float* __restrict__ ce, float* part, unsigned int* hist,
                                                            unsigned int* counter, double* cls, SelectState* st,
                                                            unsigned int k) {
  __shared__ float sm[kThreads / 32][kPartStride];
  __shared__ unsigned int sh[kBins];
  __shared__ double fold[kThreads];
  for (int i = threadIdx.x; i < kBins; i += kThreads) sh[i] = 0u;
  const int P = H * W;
  const int pix = blockIdx.x * kThreads + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float p[kMaxCh], x[kMaxCh];
  int g = 255;
  bool in = pix < P;
  float loss = 0.f;
  if (in) {
    g = gt[pix];
    float m, lse;
    pixel_softmax(lg, h4, w4, make_axis(h4, H), make_axis(w4, W), pix / W, pix % W, n_ch, x, p, m, lse);
    if (g < n_ch) {
      float xg = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxCh; ++c) xg = (c == g) ? x[c] : xg;
      loss = lse - (xg - m);                       // -log_softmax[g] >= 0
      loss = loss > 0.f ? loss : 0.f;              // also folds -0.0 (the radix select orders bit patterns)
    }
    ce[pix] = loss;                                // 255 (ignored) and ids above obj_num: 0
  }
  __syncthreads();                                 // sh is zeroed
  hist_vote(sh, in, __float_as_uint(loss) >> 21);
  const bool valid = in && g != 255;               // flatten_probas keeps everything but 255
#pragma unroll
  for (int c = 0; c < kMaxCh; ++c) {
    const float pc = valid ? p[c] : 0.f;
    const bool mine = valid && g == c;
    const float i_c = warp_sum(mine ? pc : 0.f);
    const float s_c = warp_sum(pc);
    const unsigned n_c = __popc(__ballot_sync(0xffffffffu, mine));
    if (lane == 0) {
      sm[warp][c] = i_c;
      sm[warp][kMaxCh + c] = s_c;
      sm[warp][2 * kMaxCh + c] = (float)n_c;
    }
  }
  __syncthreads();
  if (threadIdx.x < kPartStride) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) acc += sm[w][threadIdx.x];
    part[(size_t)blockIdx.x * kPartStride + threadIdx.x] = acc;
  }
  hist_flush(sh, hist);
  if (!last_block_done(counter)) return;
  // class sums: thread -> (column o, slice j); slice j adds blocks j, j + J, ... in order, then the J slices in order
  constexpr int J = kThreads / kPartStride;
  const int o = threadIdx.x % kPartStride, j = threadIdx.x / kPartStride;
  double a = 0.0;
  if (j < J)
    for (unsigned int b = j; b < gridDim.x; b += J) a += (double)__ldcg(&part[(size_t)b * kPartStride + o]);
  fold[threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.x < kPartStride) {
    double tsum = 0.0;
    for (int jj = 0; jj < J; ++jj) tsum += fold[jj * kPartStride + threadIdx.x];
    cls[threadIdx.x] = tsum;
  }
  scan_select(hist, 0, st, k);
}

// pass 1: bits 20..10 of the values whose bits 31..21 equal the prefix, pass 2: bits 9..0 of those whose bits 31..10 do
__global__ void __launch_bounds__(kThreads) tl_hist_kernel(const float* __restrict__ ce, int P, int pass,
                                                           SelectState* st, unsigned int* hist, unsigned int* counter,
                                                           unsigned int k) {
  __shared__ unsigned int sh[kBins];
  for (int i = threadIdx.x; i < kBins; i += kThreads) sh[i] = 0u;
  __syncthreads();
  const unsigned int prefix = st->prefix;
  for (int base = blockIdx.x * kThreads; base < P; base += gridDim.x * kThreads) {   // warp-uniform trip count
    const int i = base + threadIdx.x;
    const unsigned int b = i < P ? __float_as_uint(ce[i]) : 0u;
    if (pass == 1) hist_vote(sh, i < P && (b >> 21) == prefix, (b >> 10) & 2047u);
    else hist_vote(sh, i < P && (b >> 10) == prefix, b & 1023u);
  }
  __syncthreads();
  hist_flush(sh, hist);
  if (!last_block_done(counter)) return;
  scan_select(hist, pass, st, k);
}

__global__ void __launch_bounds__(kThreads) tl_topk_sum_kernel(const float* __restrict__ ce, int P,
                                                               const SelectState* __restrict__ st, double* tpart,
                                                               unsigned int* counter, const double* __restrict__ cls,
                                                               int n_ch, unsigned int k, Coef* __restrict__ coef,
                                                               float* __restrict__ losses) {
  __shared__ double red[kThreads];
  const int i = blockIdx.x * kThreads + threadIdx.x;
  const unsigned int thr = st->prefix;
  double v = 0.0;
  if (i < P) {
    const float c = ce[i];
    if (__float_as_uint(c) > thr) v = (double)c;
  }
  v = warp_sum_d(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) acc += red[w];
    tpart[blockIdx.x] = acc;
  }
  if (!last_block_done(counter)) return;
  v = 0.0;
  for (unsigned int b = threadIdx.x; b < gridDim.x; b += kThreads) v += __ldcg(&tpart[b]);
  red[threadIdx.x] = v;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int present = 0;
    for (int c = 0; c < n_ch; ++c) present += cls[2 * kMaxCh + c] > 0.0 ? 1 : 0;
    double jac = 0.0;
    for (int c = 0; c < kMaxCh; ++c) {
      float a = 0.f, b = 0.f;
      if (c < n_ch && cls[2 * kMaxCh + c] > 0.0) {
        const double I = cls[c], S = cls[kMaxCh + c], N = cls[2 * kMaxCh + c];
        const double D = S + N - I + 1e-6;          // I + (S - I) + (N - I) + eps
        jac += 1.0 - I / D;
        a = (float)(-1.0 / (D * present));
        b = (float)(I / (D * D * present));
      }
      coef->a[c] = a;
      coef->b[c] = b;
    }
    if (present) jac /= present;
    const float thr_val = __uint_as_float(thr);
    const double ce_loss = (red[0] + (double)st->k_rem * (double)thr_val) / (double)k;
    coef->inv_k = (float)(1.0 / (double)k);
    coef->tie_w = (float)((double)st->k_rem / (double)st->n_ties);
    coef->thr_bits = thr;
    losses[0] = (float)(0.5 * ce_loss + 0.5 * jac);
    losses[1] = (float)ce_loss;
    losses[2] = (float)jac;
  }
}

__global__ void __launch_bounds__(kThreads) tl_grad_pixel_kernel(const float* __restrict__ lg, int h4, int w4,
                                                                 const uint8_t* __restrict__ gt, int H, int W, int n_ch,
                                                                 const float* __restrict__ ce,
                                                                 const Coef* __restrict__ coef, float scale,
                                                                 float* __restrict__ gup) {
  const int P = H * W;
  const int pix = blockIdx.x * kThreads + threadIdx.x;
  if (pix >= P) return;
  float p[kMaxCh], x[kMaxCh], m, lse;
  pixel_softmax(lg, h4, w4, make_axis(h4, H), make_axis(w4, W), pix / W, pix % W, n_ch, x, p, m, lse);
  const int g = gt[pix];
  const bool valid = g != 255;
  const unsigned int bits = __float_as_uint(ce[pix]);
  const unsigned int thr = coef->thr_bits;
  const float w_ce = (valid && g < n_ch) ? (bits > thr ? 1.f : (bits == thr ? coef->tie_w : 0.f)) * coef->inv_k : 0.f;
  float q[kMaxCh], dot = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxCh; ++c) {
    q[c] = valid ? (g == c ? coef->a[c] : coef->b[c]) : 0.f;
    dot += p[c] * q[c];
  }
#pragma unroll
  for (int c = 0; c < kMaxCh; ++c) {
    if (c < n_ch) {
      const float d_ce = w_ce * (p[c] - (g == c ? 1.f : 0.f));
      const float d_j = p[c] * (q[c] - dot);
      gup[(size_t)c * P + pix] = scale * 0.5f * (d_ce + d_j);
    }
  }
}

// grad[c][y4][x4] = sum over the output pixels whose taps touch (y4, x4) of wy * wx * gup[c][oy][ox]
__global__ void __launch_bounds__(kThreads) tl_grad_gather_kernel(const float* __restrict__ gup, int H, int W, int h4,
                                                                  int w4, int n_ch, int n_logit_ch,
                                                                  float* __restrict__ grad) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  const int plane = h4 * w4;
  if (i >= n_logit_ch * plane) return;
  const int c = i / plane, r = i - c * plane, y4 = r / w4, x4 = r - y4 * w4;
  if (c >= n_ch) { grad[i] = 0.f; return; }
  const Axis ay = make_axis(h4, H), ax = make_axis(w4, W);
  int ylo = 0, yhi = H - 1, xlo = 0, xhi = W - 1;
  if (ay.scale > 0.f) {
    ylo = max(0, (int)floorf((float)(y4 - 1) / ay.scale) - 1);
    yhi = min(H - 1, (int)ceilf((float)(y4 + 1) / ay.scale) + 1);
  }
  if (ax.scale > 0.f) {
    xlo = max(0, (int)floorf((float)(x4 - 1) / ax.scale) - 1);
    xhi = min(W - 1, (int)ceilf((float)(x4 + 1) / ax.scale) + 1);
  }
  const float* G = gup + (size_t)c * H * W;
  float acc = 0.f;
  for (int oy = ylo; oy <= yhi; ++oy) {
    int i0, i1;
    float l0, l1;
    taps(ay, oy, i0, i1, l0, l1);
    const float wy = (i0 == y4 ? l0 : 0.f) + (i1 == y4 ? l1 : 0.f);
    if (wy == 0.f) continue;
    float row = 0.f;
    for (int ox = xlo; ox <= xhi; ++ox) {
      int j0, j1;
      float m0, m1;
      taps(ax, ox, j0, j1, m0, m1);
      const float wx = (j0 == x4 ? m0 : 0.f) + (j1 == x4 ? m1 : 0.f);
      if (wx != 0.f) row += wx * G[(size_t)oy * W + ox];
    }
    acc += wy * row;
  }
  grad[i] = acc;
}

// predict_current_mask of the TRAINING engine (aot_engine.py:467-483 after decode_current_logits has pushed the channels
// above obj_num to -1e10, :449-452): argmax over channels 0 .. obj_num of the upsampled logits, first maximum on ties.
__global__ void __launch_bounds__(kThreads) tl_predict_mask_kernel(const float* __restrict__ lg, int h4, int w4, int H,
                                                                   int W, int n_ch, uint8_t* __restrict__ label) {
  const int pix = blockIdx.x * kThreads + threadIdx.x;
  if (pix >= H * W) return;
  float p[kMaxCh], x[kMaxCh], m, lse;
  pixel_softmax(lg, h4, w4, make_axis(h4, H), make_axis(w4, W), pix / W, pix % W, n_ch, x, p, m, lse);
  int best = 0;
#pragma unroll
  for (int c = 1; c < kMaxCh; ++c)
    if (c < n_ch && x[c] > x[best]) best = c;
  label[pix] = (uint8_t)best;
}

}  // namespace
}  // namespace rmem

using namespace rmem;

extern "C" {

int rmem_train_loss_workspace_bytes(int H, int W, size_t* bytes) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(bytes && H > 0 && W > 0 && (long long)H * W < (1ll << 30), "train_loss: bad size %d x %d", H, W);
  *bytes = make_layout(H, W).total;
  return RMEM_OK;
  RMEM_API_END
}

int rmem_train_loss_fwd_bwd(const float* logits4, int n_logit_ch, int h4, int w4, const uint8_t* gt, int H, int W,
                            int obj_num, long long top_k_pixels, float grad_scale, float* losses, float* grad_logits4,
                            void* workspace, size_t workspace_bytes, void* stream) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(logits4 && gt && losses && workspace, "train_loss: null argument");
  RMEM_REQUIRE(H > 0 && W > 0 && h4 > 0 && w4 > 0 && (long long)H * W < (1ll << 30), "train_loss: bad size");
  RMEM_REQUIRE(obj_num >= 0 && obj_num + 1 <= kMaxCh && obj_num + 1 <= n_logit_ch,
               "train_loss: obj_num=%d needs 1..%d logit channels, got %d", obj_num, kMaxCh, n_logit_ch);
  const Layout L = make_layout(H, W);
  RMEM_REQUIRE(top_k_pixels >= 1 && top_k_pixels <= L.P, "train_loss: top_k_pixels=%lld outside 1..%d", top_k_pixels, L.P);
  RMEM_REQUIRE(workspace_bytes >= L.total, "train_loss: workspace %zu < %zu bytes", workspace_bytes, L.total);
  RMEM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "train_loss: workspace must be 256-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  float* ce = reinterpret_cast<float*>(ws + L.ce);
  float* gup = reinterpret_cast<float*>(ws + L.gup);
  float* part = reinterpret_cast<float*>(ws + L.part);
  double* tpart = reinterpret_cast<double*>(ws + L.tpart);
  unsigned int* hist = reinterpret_cast<unsigned int*>(ws + L.hist);
  SelectState* st = reinterpret_cast<SelectState*>(ws + L.state);
  Coef* coef = reinterpret_cast<Coef*>(ws + L.coef);
  const int n_ch = obj_num + 1;
  const unsigned int k = (unsigned int)top_k_pixels;

  unsigned int* counters = reinterpret_cast<unsigned int*>(ws + L.counters);
  double* cls = reinterpret_cast<double*>(ws + L.cls);
  RMEM_CUDA_CHECK(cudaMemsetAsync(ws + L.hist, 0, L.coef - L.hist, s));      // histograms, select state, tickets
  tl_pixel_kernel<<<L.nb, kThreads, 0, s>>>(logits4, h4, w4, gt, H, W, n_ch, ce, part, hist, counters, cls, st, k);
  RMEM_LAUNCH_CHECK();
  const int hist_grid = min(L.nb, 148 * 4);
  for (int pass = 1; pass < 3; ++pass) {
    tl_hist_kernel<<<hist_grid, kThreads, 0, s>>>(ce, L.P, pass, st, hist + pass * kBins, counters + pass, k);
    RMEM_LAUNCH_CHECK();
  }
  tl_topk_sum_kernel<<<L.nb, kThreads, 0, s>>>(ce, L.P, st, tpart, counters + 3, cls, n_ch, k, coef, losses);
  RMEM_LAUNCH_CHECK();
  if (grad_logits4) {
    tl_grad_pixel_kernel<<<L.nb, kThreads, 0, s>>>(logits4, h4, w4, gt, H, W, n_ch, ce, coef, grad_scale, gup);
    RMEM_LAUNCH_CHECK();
    const int gather_grid = cdiv(n_logit_ch * h4 * w4, kThreads);
    tl_grad_gather_kernel<<<gather_grid, kThreads, 0, s>>>(gup, H, W, h4, w4, n_ch, n_logit_ch, grad_logits4);
    RMEM_LAUNCH_CHECK();
  }
  return RMEM_OK;
  RMEM_API_END
}

int rmem_train_predict_mask(const float* logits4, int n_logit_ch, int h4, int w4, int H, int W, int obj_num,
                            uint8_t* label, void* stream) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(logits4 && label, "train_predict_mask: null argument");
  RMEM_REQUIRE(H > 0 && W > 0 && h4 > 0 && w4 > 0 && (long long)H * W < (1ll << 30), "train_predict_mask: bad size");
  RMEM_REQUIRE(obj_num >= 0 && obj_num + 1 <= kMaxCh && obj_num + 1 <= n_logit_ch,
               "train_predict_mask: obj_num=%d needs 1..%d logit channels, got %d", obj_num, kMaxCh, n_logit_ch);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = cdiv(H * W, kThreads);
  tl_predict_mask_kernel<<<grid, kThreads, 0, s>>>(logits4, h4, w4, H, W, obj_num + 1, label);
  RMEM_LAUNCH_CHECK();
  return RMEM_OK;
  RMEM_API_END
}

}  // extern "C"
