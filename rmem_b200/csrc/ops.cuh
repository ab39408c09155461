// Host launchers of the HBM-bound kernels on the RMem propagation path.
// All tensors are token-major / NHWC ("[pixels, channels]"), device pointers, explicit stream.
#pragma once
#include "common.cuh"

namespace rmem {

// img NCHW fp32 [3,H,W] -> NHWC t16 [H,W,8] (channels 3..7 zero).              (encoder input)
int pack_image(const float* img, t16* out, int H, int W, cudaStream_t s);
// zero-padded [H+6][W+8][8] layout for the stem convolution (gemm.cuh conv = 2); writes the interior only
int pack_image_padded(const float* img, t16* out, int H, int W, cudaStream_t s);

// 3x3 stride-2 pad-1 max pooling on NHWC t16 (resnet.py:186).
int maxpool3x3s2(const t16* x, t16* y, int Hin, int Win, int C, int Hout, int Wout, cudaStream_t s);

// LayerNorm over C (eps 1e-5): x fp32 [P, ldx] -> y t16 [P, ldy] (+ optional second copy y2).
// `add2` (fp32 [P, C], optional): y2 = t16(LN(x) + add2) instead of a plain copy (AOT sine PE on q, k).
// layernorm_pair: the two C-wide halves of a [P, 2C] row normalised in one launch; the second half goes to
// y[:, C:2C] or, when y1 is given, to y1[:, 0:C] (row stride ldy1).
int layernorm_pair(const float* x, long long ldx, const float* g0, const float* b0, const float* g1, const float* b1,
                   t16* y, long long ldy, int P, int C, cudaStream_t s, t16* y1 = nullptr, long long ldy1 = 0);
int layernorm(const float* x, long long ldx, const float* gamma, const float* beta, t16* y, long long ldy,
              t16* y2, long long ldy2, int P, int C, cudaStream_t s, const float* add2 = nullptr,
              // optional: also start a residual stream from x: init_res[:, 0:C] = x, init_res[:, C:2C] = 0 (fp32, row stride ld_init)
              float* init_res = nullptr, long long ld_init = 0);

// GroupNorm over (pixels x C/G) per group, eps 1e-5, optional ReLU.  `stats` = kGnScratchDoubles doubles of
// scratch whose element [64] (the block counter) must be zero before the first call; it re-arms itself.
constexpr int kGnScratchDoubles = 72 + 148 * 4 * 64;
// `add` (optional, [P,C] t16): added to the normalised + activated map before the store
int groupnorm_t16(const t16* x, const float* gamma, const float* beta, t16* y, int P, int C, int G, int relu,
                   double* stats, cudaStream_t s, const t16* add = nullptr, bool stats_ready = false);
int groupnorm_f32(const float* x, const float* gamma, const float* beta, t16* y, int P, int C, int G, int relu,
                  double* stats, cudaStream_t s);

// AOT block helpers (transformer.py:553-692, basic.py:15-35, position.py:35-77).  groupnorm_*'s `relu` argument is an
// activation code: 0 none, 1 ReLU, 2 exact GELU.
int add_t16(const t16* a, long long lda, const t16* b, long long ldb, t16* y, long long ldy, int P, int C,
            cudaStream_t s);
int add_layernorm_t16(const t16* a, long long lda, const t16* b, long long ldb, const float* gamma, const float* beta,
                      t16* y, long long ldy, int P, int C, cudaStream_t s);
int accum_t16_into_f32(const t16* x, long long ldx, float* y, long long ldy, int P, int C, cudaStream_t s);
int cvt_f32_t16(const float* x, long long ldx, t16* y, long long ldy, int P, int C, cudaStream_t s);
int sine_pos_emb(float* out, int h, int w, int C, cudaStream_t s);
int mean_heads(const float* in, float* out, int H, long long n, cudaStream_t s);
int qprep_heads(const t16* q, long long ldq, const float* pe_cur, const float* pe_mem, const int* pe_slot, int T,
                float scale, t16* qt, float* qbias, int P, int H, cudaStream_t s);

// Depthwise 5x5 (pad 2) on a token-major map: x t16 [h*w, C], w fp32 [25, C] -> y t16.   (basic.py:38-59)
// ldx / ldy: pixel pitch of input / output in elements (<= 0: C)
int dwconv5x5(const t16* x, const float* w, t16* y, int h, int wd, int C, cudaStream_t s, int ldx = 0, int ldy = 0);

// Bilinear resize, align_corners=True, NHWC t16.                                         (fpn.py:50,58)
int upsample_bilinear_t16(const t16* x, t16* y, int hin, int win, int hout, int wout, int C, cudaStream_t s,
                          const t16* add = nullptr,    // add: optional [hout*wout, C] t16 added to the interpolated map
                          // optional: x is an un-normalised conv output, relu(GroupNorm(x)) with these statistics (the
                          // stats block of groupnorm_t16 / GemmParams::gn_stats) is applied to the taps on load
                          const double* gn_stats = nullptr, const float* gamma = nullptr, const float* beta = nullptr, int G = 0);

// 1x1 conv to the 11 ID logits, planar fp32 output [11, P].                                (fpn.py:66)
int conv_out_logits(const t16* x, const t16* w, const float* b, float* out, int P, int Cin, int Cout, cudaStream_t s);
// GRU_MEMORY ablation: the elementwise halves of ConvGRUCell.forward (transformer.py:84-100), fp32 hidden state
int gru_reset(const float* gates, long long ldg, const float* h, t16* comb_h, long long ldc, int P, int C, cudaStream_t s);
int gru_blend(const float* gates_u, long long ldg, const float* cand, float* h, t16* h16, int P, int C, cudaStream_t s);
// out = conv_out(relu(GroupNorm_G(x))) in two launches (statistics, fused normalise + 1x1 conv): the decoder tail, fpn.py:62-67
int conv_out_gn_logits(const t16* x, const float* gamma, const float* beta, int G, double* stats, const t16* w,
                       const float* b, float* out, int P, int Cin, int Cout, cudaStream_t s, bool stats_ready = false);

// [P, C] (row stride ldx) -> [C, ldy] transposed copy.
int transpose_t16(const t16* x, long long ldx, t16* y, long long ldy, int P, int C, cudaStream_t s);

// dst[p, 0:C] = src[p, 0:C] with independent row strides (t16).
int copy2d_t16(const t16* src, long long lds, t16* dst, long long ldd, int P, int C, cudaStream_t s);
int fill_t16(t16* dst, long long ldd, int P, int C, float v, cudaStream_t s);

// Per-engine label map (aot_engine.py:604-618): label fp32/uint8 [H,W] -> uint8 [H,W].
int separate_label(const void* label, int label_is_f32, uint8_t* out, int H, int W, int engine, int n_engines,
                   cudaStream_t s);

// ID bank (aot.py:111-114, deaot.py:65-69): gather-sum of Conv2d(12->C,k17,s16,p8) weight slices indexed by the
// label, + bias, optional LayerNorm.  w_packed fp32 [17*17][12][C].  out t16 [h*w, ldo].
// `prefix` (optional) fp32 [12][18][18][C]: per-class inclusive 2-D prefix sums of the weight over (ky, kx), lets a patch
// of uniform class be a 4-read rectangle sum.
int idbank_embed(const uint8_t* label, int H, int W, int use_ignore, const float* w_packed, const float* bias,
                 const float* ln_g, const float* ln_b, t16* out, long long ldo, float* out_f32, int h, int w, int C,
                 cudaStream_t s, const float* prefix = nullptr, const float* prefix_rows = nullptr,
                 t16* out2 = nullptr, t16* out3 = nullptr, long long ldo23 = 0);

// Mask head (aot_engine.py:457-463, 650-673; evaluator.py:430-441): k engines' planar logits [11,h4,w4] ->
// bilinear(align_corners=True) -> soft aggregation -> out_logits [1+10k, Ho, Wo] (optional) and uint8 label.
int mask_head(const float* const* logits4, int k, int h4, int w4, int Ho, int Wo, float* out_logits,
              uint8_t* out_label, cudaStream_t s);
// Test-time-augmentation head: mean over augmentations of the soft-maxed, (un)flipped, upsampled logits -> argmax.
// logits4: HOST array [n_aug * k] of device pointers (augmentation major), h4 / w4 / flip: HOST [n_aug].
int tta_head(const float* const* logits4, int n_aug, int k, const int* h4, const int* w4, const int* flip, int Ho, int Wo,
             float* out_prob, uint8_t* out_label, cudaStream_t s);
// uint8 HWC frame -> normalised fp32 NCHW at nh x nw (OpenCV INTER_CUBIC rule), optional horizontal flip.
int preprocess_frame(const uint8_t* img, int H, int W, int bgr, int nh, int nw, int flip, float* out, cudaStream_t s);

// Relevance part of the evict score (aot_engine.py:355-362, transformer.py:891-906):
//   fg = 1 - softmax(bilinear(logits4 -> h x w))[0];  rel[t] = sum_i mass[i,t] * fg[i]   (un-normalised)
int evict_relevance(const float* mass, int T, const float* logits4, int h4, int w4, int h, int w, float* rel,
                    cudaStream_t s);

// Qt = t16(Q + pe_cur);  qbias[i,t] = scale * <Qt_i, pe_mem[t]>        (temporal PE as a score bias, K8)
// pe_mem = mem_pos_emb [n_slots, C]; pe_slot[t] = slot used by memory frame t (temporal_pe_slots).
int qprep(const t16* q, long long ldq, const float* pe_cur, const float* pe_mem, const int* pe_slot, int T,
          float scale, t16* qt, float* qbias, int P, int C, cudaStream_t s);
// transformer.py:1140-1170: identity for T <= n_slots, flip -> nearest -> flip above.
void temporal_pe_slots(int T, int n_slots, int* out);

}  // namespace rmem
