/* rmem_b200 -- C ABI of the B200-native RMem (restricted-memory VOS) propagation engine.
 *
 * The reference (Restricted-Memory/RMem, pure Python/PyTorch) has no FFI layer: its boundary for this
 * path is the Python method surface of AOTInferEngine / AOTEngine plus the LSTT/GPM module forwards
 * (SURVEY.md section 8b).  This header is what a ctypes binding under those Python classes calls.
 * Conventions: plain `extern "C"`, every entry returns int (0 = ok, <0 = error; message via
 * rmem_last_error(), thread-local), never throws, never calls exit(), never allocates device memory
 * (the caller passes outputs and a workspace/arena it sized with the matching *_bytes query), takes an
 * explicit cudaStream_t (as void*), holds no global mutable state.  Device pointers only unless a
 * parameter is marked HOST.  All activations are token-major ("[pixels, channels]", i.e. NHWC); "t16" below is
 * the 16-bit tensor-core operand type (fp16 by default, see rmem_operand_dtype), fp32 the residual / statistics /
 * logits type.
 *
 * Paths are relative to /root/reference/aot_plus/.
 */
#ifndef RMEM_B200_H
#define RMEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMEM_MAX_BANK_FRAMES 16
#define RMEM_GN_SCRATCH_DOUBLES (72 + 148 * 4 * 64)
#define RMEM_ATTN_DENSE 0 /* materialised scores: generic GEMMs + row softmax */
#define RMEM_ATTN_TC 1    /* (removed: the first tcgen05 kernel of round 1; selecting it is an error) */
#define RMEM_ATTN_TC2 2   /* stream-K schedule over single CTAs, 8 softmax warps, P through TMEM, fp16 partials */
#define RMEM_ATTN_TC3 3   /* CTA pairs (tcgen05.mma.cta_group::2) sharing every K / V^T tile, 128-key score MMAs, seeded
                             row maximum (default) */
#define RMEM_ATTN_TC4 4   /* TC3 + the column kernel for seeded banks of >= 3 frames: scores and exponentials once per (query
                             pair, sub-tile), probabilities re-read from L2 for the other Dv chunks, fixed softmax reference
                             from the seed, guarded TC3 fallback; everything else (self-attention, unseeded calls) runs TC3 */

int rmem_version(void);
const char* rmem_last_error(void);
/* "fp16" (default build) or "t16": the 16-bit tensor-core operand / activation type every `void*` half tensor uses. */
const char* rmem_operand_dtype(void);

/* ---------------------------------------------------------------- op level (module forwards) ---- */

/* Generic t16 GEMM / implicit-GEMM conv with fused epilogue: nn.Linear / nn.Conv2d(+FrozenBatchNorm2d+ReLU)
 * call sites of networks/encoders/resnet.py:48-68,178-195, networks/decoders/fpn.py:36-68,
 * networks/layers/transformer.py:1104-1123, attention.py:151-172.  See rmem_b200/csrc/gemm.cuh. */
typedef struct rmem_gemm_desc {
  const void* A; long long lda;      /* t16 [M,K] row-major, or NHWC map when conv != 0 */
  const void* B; long long ldb;      /* t16 [N,K] row-major weight (K ordered ky,kx,ci for conv) */
  int M, N, K;
  int conv, Hin, Win, Cin, Wout, kw, stride, pad;
  float alpha;
  const float* bias; int bias_along_m;
  int act; int act_from_col;         /* 0 none, 1 relu, 2 silu; applied to columns >= act_from_col */
  const void* residual; long long ldr;  /* t16, added before the activation */
  const void* gate; long long ldg;      /* t16, multiplied after the activation */
  int accumulate;                    /* C += result (fp32 destinations only) */
  void* C; long long ldc; int c_is_f32;
  void* C2; long long ldc2; int c2_is_f32; int n_split; /* columns >= n_split go to C2 */
  int pad_n_ok;                      /* N % 32 != 0: columns N..round_up(N,32)-1 of C are writable padding */
  int n_images;                      /* conv != 0 only: images stacked in A ([n][Hin][Win][Cin]) and in C / residual
                                      * ([n][Hout*Wout][N]), M = n * Hout * Wout; 0 or 1 = a single image */
} rmem_gemm_desc;
int rmem_gemm_fwd(const rmem_gemm_desc* d, void* stream);
/* 0 = auto: the tcgen05/TMA kernel whenever its alignment rules hold (K % 64 == 0, 16-byte aligned operands), else the
 * legacy mma.sync kernel; 1 = force legacy (parity tests compare the two).  Thread-local.  Returns the previous value. */
int rmem_set_gemm_impl(int impl);

/* Long-term / self attention over the restricted bank with per-frame attention mass and the temporal
 * positional embedding applied as a per-(query,frame) score bias.
 * Replaces GatedPropagation.forward's QK^T -> softmax -> .V (networks/layers/attention.py:174-193), the
 * mass record of transformer.py:1185-1192 and the temporal-PE add of transformer.py:1140-1175. */
int rmem_long_attn_workspace_bytes(int impl, int HW, int HWp, int nslots, int Dv, size_t* bytes);
int rmem_long_attn_fwd(int impl, const void* qt, const float* qbias, const void* kbank, const void* vtbank,
                       int nslots, int T, const int* slots /*HOST [T]*/, int HW, int HWp, int Dk, int Dv,
                       float scale, const void* gate, long long ldg, void* out, long long ldo, float* mass,
                       void* workspace, size_t workspace_bytes, void* stream);
/* Same op with the token grid of the queries / bank frames (grid_h * grid_w == HW).  RMEM_ATTN_TC3 uses it to seed the
 * running row maximum of the online softmax with the scores against the 3x3 neighbourhood of the query's own position in
 * every bank frame (a lower bound of the true maximum), which keeps the lazy-rescale path of the kernel rare on peaked
 * score distributions; results are the same softmax either way.  grid_h = grid_w = 0: no seed (= rmem_long_attn_fwd). */
int rmem_long_attn_grid_fwd(int impl, const void* qt, const float* qbias, const void* kbank, const void* vtbank,
                            int nslots, int T, const int* slots /*HOST [T]*/, int HW, int HWp, int Dk, int Dv,
                            float scale, const void* gate, long long ldg, void* out, long long ldo, float* mass,
                            int grid_h, int grid_w, void* workspace, size_t workspace_bytes, void* stream);
/* Multi-head attention over the bank (AOT model, 8 heads x 32): MultiheadAttention.forward of
 * networks/layers/attention.py:28-81 as used by SimplifiedTransformerBlock (transformer.py:566-571 self-attention, :632-650
 * long-term attention over the restricted bank with the temporal PE as a per-head score bias, :656-662 short-term).
 * q [HW, H*dh] (row stride ldq), kbank [nslots][HWp][H*dh], vtbank [H*dh][nslots*HWp] (value-major), qbias [H][HW][T]
 * (already multiplied by scale) or NULL, out [HW, H*dh] (row stride ldo), mass [HW, T] = head mean of the per-frame
 * probability mass or NULL.  impl RMEM_ATTN_DENSE = materialised scores; anything else = the fused tcgen05 kernel. */
int rmem_mha_workspace_bytes(int impl, int HW, int HWp, int nslots, int H, size_t* bytes);
int rmem_mha_fwd(int impl, const void* q, long long ldq, const void* kbank, const void* vtbank, int nslots, int T,
                 const int* slots /*HOST [T]*/, int HW, int HWp, int H, int dh, float scale, const float* qbias,
                 void* out, long long ldo, float* mass, void* workspace, size_t workspace_bytes, void* stream);
/* Debug aid (RMEM_ATTN_TC3): device int the kernel increments once per warp-level lazy-rescale event; NULL disables.
 * Thread-local. */
int rmem_debug_attn_rescale_counter(void* dev_int);

/* Debug aid: per-event clock64 trace of CTA 0 of the RMEM_ATTN_TC2 kernel into dev_buf ([tiles][16] int64); NULL disables. */
int rmem_debug_attn_trace(void* dev_buf);
/* Host-only: the static (unit, tile)-step schedule the fused attention kernel uses for a launch of this shape: CTA c owns
 * steps [bounds[c], bounds[c+1]) of n_units * tiles_per_unit (impl RMEM_ATTN_TC2: CTAs and 64-key tiles; RMEM_ATTN_TC3:
 * clusters of two CTAs and 128-key groups, a unit being a PAIR of query tiles x one Dv chunk).  `bounds` has room for `cap` ints (>= n_cta + 1). */
int rmem_debug_attn_schedule(int impl, int HW, int T, int Dv, int* n_units, int* tiles_per_unit, int* n_cta, int* bounds, int cap);
/* Measurement aid: cudaEvent_t handles recorded immediately before / after the main attention kernel launch (TC2 / TC3) inside
 * rmem_long_attn_fwd (NULL, NULL clears).  Thread-local. */
int rmem_debug_attn_events(void* ev0, void* ev1);
/* Same for the tcgen05 GEMM: first 64 CTAs of every launch, [cta][8] int64. */
int rmem_debug_gemm_trace(void* dev_buf);
/* Tuning aids for the tcgen05 GEMM (tools/tune_gemm.py), thread-local.  force: tile width bn (0 = auto | 64 | 128 | 256),
 * split-K factor (0 = auto, 1..4; bn = 64 only), TMA ring depth (0 = auto, 1..6) for every following launch.  log: the
 * shape of every following tcgen05 GEMM launch is appended to a HOST buffer of 16-int records (M N K conv Hin Win Cin Wout
 * kw stride pad act has_res has_gate flags n_split); NULL stops logging; log_count = records written so far. */
int rmem_debug_gemm_force(int bn, int splitk, int stages);
int rmem_debug_gemm_log(int* host_buf, int cap_records);
int rmem_debug_gemm_log_count(void);

/* Qt = t16(Q + cur_pos_emb); qbias[i,t] = scale * <Qt_i, pe_mem[t]>          (transformer.py:1140-1175) */
/* pe_mem = mem_pos_emb [n_slots, C]; pe_slot HOST [T] = slot of each memory frame (rmem_temporal_pe_slots). */
int rmem_qprep_fwd(const void* q, long long ldq, const float* pe_cur, const float* pe_mem, const int* pe_slot, int T,
                   float scale, void* qt, float* qbias, int P, int C, void* stream);
/* Slot map of transformer.py:1140-1170 (identity for T<=4, flip-nearest-flip above): HOST out [T]. */
int rmem_temporal_pe_slots(int T, int n_slots, int* out);

/* Windowed short-term attention: LocalGatedPropagation.forward core (attention.py:289-353, 363-413). */
int rmem_local_attn_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                        const float* rel, long long ldrel, const void* gate, long long ldg, void* out,
                        long long ldo, int h, int w, int Dv, float scale, void* stream);

/* Same op on tensor cores (tcgen05 + 3-D TMA boxes over the key halo of a 4x32 query patch).  rel_pitch = floats per
 * window row of `rel`: 15 = the reference's relative_emb_k order (225 columns), 16 = one aligned 64-byte line per window
 * row (240 columns, what the engine's packed weights produce; ~3x faster bias fetch). */
int rmem_local_attn_tc_workspace_bytes(int h, int w, int Dv, size_t* bytes);
int rmem_local_attn_tc_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                           const float* rel, long long ldrel, int rel_pitch, const void* gate, long long ldg,
                           void* out, long long ldo, int h, int w, int Dv, float scale, void* workspace,
                           size_t workspace_bytes, void* stream);

/* nn.LayerNorm (transformer.py:1104,1119,1222), GroupNorm (basic.py:6-12, 62-70), DWConv2d (basic.py:38-59). */
int rmem_layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, void* y, long long ldy,
                       int P, int C, void* stream);
int rmem_groupnorm_fwd(const void* x, int x_is_f32, const float* gamma, const float* beta, void* y, int P, int C,
                       int G, int relu, double* stats /* RMEM_GN_SCRATCH_DOUBLES doubles, zero-initialised once */,
                       void* stream);
int rmem_dwconv5x5_fwd(const void* x, const float* w /* [25,C] */, void* y, int h, int w_, int C, void* stream);

/* F.interpolate(bilinear, align_corners=True) on NHWC t16 (fpn.py:50,58). */
int rmem_upsample_bilinear_fwd(const void* x, void* y, int hin, int win, int hout, int wout, int C, void* stream);
int rmem_transpose_fwd(const void* x, long long ldx, void* y, long long ldy, int P, int C, void* stream);
int rmem_maxpool3x3s2_fwd(const void* x, void* y, int Hin, int Win, int C, int Hout, int Wout, void* stream);
int rmem_pack_image_fwd(const float* img_nchw, void* out_nhwc8, int H, int W, void* stream);
/* Same into the zero-padded layout [H+6][W+8][8] (pixel (y,x) at (y+3, x+3)) that the stem convolution reads
 * (rmem_gemm_desc.conv = 2: conv1 7x7 stride 2 pad 3 of networks/encoders/resnet.py:178-181 as a tcgen05 GEMM with one
 * k-block per window row; B = [Cout][7][8 pixels][8 channels], zero weights for the 8th pixel).  Writes the interior
 * only: the caller zeroes the buffer once. */
int rmem_pack_image_padded_fwd(const float* img_nchw, void* out, int H, int W, void* stream);

/* ID bank: one_hot_mask (utils/image.py:69-74) + assign_identity (networks/engines/aot_engine.py:208-232) +
 * patch_wise_id_bank Conv2d(12->256,k17,s16,p8) (networks/models/aot.py:63-74,111-114) + id_norm
 * (networks/models/deaot.py:65-69), as a gather-sum indexed by the uint8 label. */
/* prefix (nullable): fp32 [12][18][18][C] inclusive 2-D prefix sums of the weight over (ky, kx) per class; a patch whose
 * in-bounds pixels share one class then costs four reads instead of up to 289 weight rows, a patch with few pixels off
 * its dominant class the rectangle plus two reads per such pixel.  prefix_rows (nullable, needs prefix): fp32
 * [17][12][18][C] 1-D prefix sums over kx per (ky, class): every run of equal labels in a patch row costs two reads.
 * The kernel picks the cheapest decomposition per patch; all of them compute the same sum (fp32 order differs). */
int rmem_idbank_fwd(const uint8_t* label, int H, int W, int use_ignore, const float* w_packed /* [289,12,C] */,
                    const float* prefix, const float* prefix_rows, const float* bias, const float* ln_gamma,
                    const float* ln_beta, void* out_t16, long long ldo, float* out_f32, int h, int w, int C,
                    void* stream);

/* Mask-ID assignment: bilinear(align_corners=True) upsample of the 1/4-res logits (aot_engine.py:457-463),
 * soft_logit_aggregation over k object groups (aot_engine.py:650-673), softmax -> argmax
 * (networks/managers/evaluator.py:430-441).  logits4: HOST array of k device pointers, each planar [11,h4,w4]. */
int rmem_mask_head_fwd(const float* const* logits4, int k, int h4, int w4, int Ho, int Wo, float* out_logits,
                       uint8_t* out_label, void* stream);

/* Test-time augmentation (networks/managers/evaluator.py:338-441): every augmentation (scale x flip) runs its own engine;
 * per output pixel each one's 1/4-res logits are upsampled (bilinear, align_corners), soft-aggregated over its k object
 * groups, soft-maxed, mirrored back when that augmentation was flipped (flip_tensor(pred_logit, 3)), the probabilities
 * are averaged over the augmentations and the argmax is taken.  logits4: HOST array [n_aug * k] of device pointers
 * (augmentation major), each planar [11, h4[a], w4[a]]; h4 / w4 / flip: HOST [n_aug].  out_prob [1+10k, Ho, Wo] nullable. */
int rmem_tta_head_fwd(const float* const* logits4, int n_aug, int k, const int* h4, const int* w4, const int* flip,
                      int Ho, int Wo, float* out_prob, uint8_t* out_label, void* stream);
/* Loader-side preprocessing on the GPU: MultiRestrictSize + MultiToTensor (dataloaders/video_transforms.py:559-682).
 * img: device uint8 [H, W, 3] as decoded (bgr != 0: cv2.imread order, swapped to RGB); out: fp32 [3, nh, nw] =
 * ((resize(img) / 255) - mean) / std with OpenCV's INTER_CUBIC rule (a = -0.75, half-pixel centres, clamped taps; no
 * resize when nh x nw == H x W), mirrored horizontally when flip != 0. */
int rmem_preprocess_fwd(const uint8_t* img, int H, int W, int bgr, int nh, int nw, int flip, float* out, void* stream);

/* Relevance term of the evict score: fg-prob from the low-res logits (aot_engine.py:355-362) times the layer-0
 * attention mass, summed over tokens (transformer.py:891-906).  rel[T] un-normalised. */
int rmem_evict_relevance_fwd(const float* mass, int T, const float* logits4, int h4, int w4, int h, int w,
                             float* rel, void* stream);
/* EMA + UCB freshness + argmin (transformer.py:907-964).  HOST arithmetic on T_old floats.  idx = long_memories_indexes
 * after the append (T_old+1 entries).  ema_keys/ema_vals/times_keys/times_vals are the engine's dictionaries
 * (capacity RMEM_MAX_BANK_FRAMES+1, counts in/out).  Returns the logical index to drop in *drop. */
int rmem_evict_pick(const float* rel_host, int T_old, const int* idx, int former, int* ema_keys, float* ema_vals,
                    int* n_ema, int* times_keys, int* times_vals, int* n_times, int* drop);

/* ---------------------------------------------------------------- engine level ---- */
/* Per-clip state machine: AOTInferEngine.{add_reference_frame, match_propogate_one_frame, update_memory,
 * restart_engine} (networks/engines/aot_engine.py:571-725, deaot_engine.py:20-56) over DeAOT
 * (networks/models/deaot.py) = ResNet-50 encoder + DualBranchGPM (networks/layers/transformer.py:700-1249) +
 * FPN decoder, with the restricted long-term bank (transformer.py:880-991) held in a caller-provided arena. */
typedef struct rmem_engine_config {
  int model;            /* 0 = r50_deaotl */
  int H, W;             /* snapped input size (16k+1), dataloaders/video_transforms.py:607-615 */
  int former_mem_len;   /* FORMER_MEM_LEN */
  int latter_mem_len;   /* LATTER_MEM_LEN */
  int max_engines;      /* ceil(max objects / 10) object groups */
  int attn_impl;        /* RMEM_ATTN_DENSE | RMEM_ATTN_TC2 | RMEM_ATTN_TC3 | RMEM_ATTN_TC4 */
  int long_term_mem_gap;
  /* Ablation knobs of configs/models/r50_deaotl.py:9-28 (all off in the shipped configs):
   *   no_long_memory  NO_LONG_MEMORY (aot_engine.py:339): never append to the long-term bank (reference frame only).
   *   reverse_infer   REVERSE_INFER (aot_engine.py:371-396) and
   *   time_encode     TIME_ENCODE[_NORM] (aot_engine.py:293-303, 413-421): accepted and ignored -- at inference both leave
   *                   every output of the unmodified reference bit-identical (the reverse pass only feeds a training loss,
   *                   the sin/cos encoding is stored and never read); checked by oracle/make_golden.py, tests/golden/knobs.json.
   *   gru_memory      GRU_MEMORY (transformer.py:35-119, 337-338, 395-396, 406-430; model 1 = R50_AOTL only -- DualBranchGPM
   *                   hard-codes gru_memory = False, :728): on every eviction the dropped frame's K / V memories of each layer
   *                   go through that layer's ConvGRU cells (K: 2x2, V: 1x1, padding "same"; fp32 hidden state, zeroed by
   *                   add_reference_frame) and the cells' outputs replace bank position 1, which like position 0 is never
   *                   dropped.  Needs the `lstt.<l>.gru.<i>.{gates,can,out}` weights (LSTT.layers.<l>.memory_grus.<i>.* of a
   *                   checkpoint trained with GRU_MEMORY); refused for model 0.  Golden: tests/golden/aot_gru_memory.npz. */
  int no_long_memory, reverse_infer, time_encode, gru_memory;
} rmem_engine_config;

typedef struct rmem_weight_entry {
  const char* name;     /* packed tensor name (see rmem_b200/weights.py) */
  size_t offset;        /* byte offset into the weight blob */
  size_t nbytes;
} rmem_weight_entry;

typedef struct rmem_engine rmem_engine; /* opaque, host-side */

int rmem_engine_arena_bytes(const rmem_engine_config* cfg, size_t* bytes);
int rmem_engine_create(const rmem_engine_config* cfg, const void* weight_blob, const rmem_weight_entry* entries,
                       int n_entries, void* arena, size_t arena_bytes, rmem_engine** out);
void rmem_engine_destroy(rmem_engine* e);
int rmem_engine_restart(rmem_engine* e);
int rmem_engine_set_gap(rmem_engine* e, int long_term_mem_gap);
/* label: device, fp32 (label_is_f32) or uint8, [H,W] integer object ids (255 = ignore). */
int rmem_engine_add_reference_frame(rmem_engine* e, const float* img, const void* label, int label_is_f32,
                                    int n_objects, int frame_step, void* stream);
/* out_logits [1+10k, Ho, Wo] fp32 (nullable), out_label uint8 [Ho, Wo] (nullable). */
int rmem_engine_propagate(rmem_engine* e, const float* img, int Ho, int Wo, float* out_logits, uint8_t* out_label,
                          void* stream);
int rmem_engine_update_memory(rmem_engine* e, const void* label, int label_is_f32, void* stream);
/* introspection (parity tests): */
int rmem_engine_num_groups(const rmem_engine* e);
int rmem_engine_long_indexes(const rmem_engine* e, int group, int* idx /*HOST, cap 17*/, int* n);
int rmem_engine_pred_logits(const rmem_engine* e, int group, const float** logits4, int* h4, int* w4);
int rmem_engine_last_evict(const rmem_engine* e, int group, float* rel /*HOST cap 16*/, int* n, int* drop);
/* One GPM layer's memories of an object group (DeAOT), valid after update_memory: the restricted long-term bank
 * (LongShortTermTransformer long_term_memories, transformer.py:993-1007) as kbank t16 [nslots][HWp][128] and vtbank t16
 * [1024][nslots*HWp] (V || ID_V, value-major), the logical->physical slot table (HOST out, cap 16, T entries), and the
 * short-term memory = last propagated frame's K (= Q) [HW,128] and V || ID_V [HW,1024].  Device pointers into the arena;
 * any out parameter may be NULL.  Used by the layer-level parity tests and by bench.py's per-layer roofline. */
int rmem_engine_layer_memory(const rmem_engine* e, int group, int layer, const void** kbank, const void** vtbank,
                             const void** q_last, const void** vid_last, int* nslots, int* HWp, int* T,
                             int* slots /*HOST cap 16*/);
long long rmem_engine_launch_count(const rmem_engine* e);
/* Software pipelining across frames: encode the NEXT frame (aot.py:116-134, no dependency on the memory bank) on the
 * engine's side stream while the current frame is propagated.  `img` must be ready on `stream`; the following
 * rmem_engine_propagate must be given the same pointer (otherwise the inline encoder runs).  Bit-identical results. */
int rmem_engine_prefetch(rmem_engine* e, const float* img, void* stream);
/* Pair prefetch: frames i+2 and i+3 encoded in ONE pass of the image encoder (every GEMM / conv launch covers both
 * images: the launches are latency-bound at one 480p image, a pair costs ~1.2x one frame).  Issue it every second frame,
 * two frames ahead (before rmem_engine_propagate of frame i); each image is consumed by the propagate call that passes the
 * same pointer and must stay untouched until then.  Same math as the single-frame encoder, differently tiled: results agree to
 * fp16 rounding, not bit for bit. */
int rmem_engine_prefetch2(rmem_engine* e, const float* img_a, const float* img_b, void* stream);
/* The general form: n = 1, 2 or 4 coming frames (HOST array of n device pointers) in one encoder pass.  A group of n is
 * meant to be issued n frames ahead -- before rmem_engine_propagate of frame i for frames i+n .. i+2n-1, every n-th frame. */
int rmem_engine_prefetch_n(rmem_engine* e, const float* const* imgs, int n, void* stream);
/* Profiling aid: CUDA events between pipeline stages (adds a stream sync per call while on).  get_timing writes
 * "stage total_ms count" lines into buf. */
int rmem_engine_set_timing(rmem_engine* e, int on);
int rmem_engine_get_timing(rmem_engine* e, char* buf, size_t cap);

/* ---------------------------------------------------------------- training side (first slice) ---- */
/* Value and gradient of the reference's per-frame training loss with respect to the 1/4-resolution ID logits
 * (rmem_engine_pred_logits): AOTEngine.calculate_current_loss (networks/engines/aot_engine.py:484-511) =
 * bilinear align_corners upsampling to the label size, channels 0..obj_num, then
 *   0.5 * CrossEntropyLoss (networks/layers/loss.py:163-211: per-pixel CE with ignore_index 255, mean of the
 *         top_k_pixels largest of all H*W values -- the caller computes top_k_pixels from the training step, :190-198)
 * + 0.5 * SoftJaccordLoss  (loss.py:30-74, 136-160: tversky alpha = beta = 1, eps 1e-6, over the non-255 pixels, mean over
 *         the classes that own a pixel).
 * logits4 fp32 [n_logit_ch, h4, w4]; gt uint8 [H, W] (255 = ignore; other ids above obj_num contribute no cross
 * entropy -- the reference raises there); losses = device float[3] {total, ce, jaccard}; grad_logits4 (nullable) =
 * grad_scale * d total / d logits4, same shape as logits4 (channels above obj_num: 0).  fp32, IEEE exp / log, every
 * reduction in a fixed order: bit-reproducible.  workspace: device, 256-byte aligned, rmem_train_loss_workspace_bytes(H, W).
 * Backward of the layers UNDER the logits (decoder, GPM, encoder) is not built. */
int rmem_train_loss_workspace_bytes(int H, int W, size_t* bytes);
int rmem_train_loss_fwd_bwd(const float* logits4, int n_logit_ch, int h4, int w4, const uint8_t* gt, int H, int W,
                            int obj_num, long long top_k_pixels, float grad_scale, float* losses, float* grad_logits4,
                            void* workspace, size_t workspace_bytes, void* stream);
/* predict_current_mask of the training engine (aot_engine.py:467-483; decode_current_logits has pushed the channels above
 * obj_num to -1e10 there, :449-452): label uint8 [H, W] = argmax over channels 0..obj_num of the upsampled logits. */
int rmem_train_predict_mask(const float* logits4, int n_logit_ch, int h4, int w4, int H, int W, int obj_num,
                            uint8_t* label, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RMEM_B200_H */
