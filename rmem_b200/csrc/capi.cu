// C-ABI glue: error state, op-level entry points declared in include/rmem_b200.h.
#include <cstdarg>
#include <map>

#include "../../include/rmem_b200.h"
#include "attn.cuh"
#include "common.cuh"
#include "gemm.cuh"
#include "ops.cuh"

namespace rmem {

static thread_local char g_err[1024] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }
long long& launch_counter() { return g_launches; }

int evict_pick_host(const float* rel_raw, int T_old, const int* idx, int former, std::map<int, float>& ema,
                    std::map<int, int>& times, int* drop, float* rel_norm_out, bool gru = false);

}  // namespace rmem

using namespace rmem;

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int rmem_version(void) { return 100; }
const char* rmem_operand_dtype(void) { return RMEM_OPERAND_NAME; }
const char* rmem_last_error(void) { return get_error(); }

int rmem_gemm_fwd(const rmem_gemm_desc* d, void* stream) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(d, "null desc");
  GemmParams p;
  p.A = (const t16*)d->A; p.lda = d->lda; p.B = (const t16*)d->B; p.ldb = d->ldb;
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.conv = d->conv; p.Hin = d->Hin; p.Win = d->Win; p.Cin = d->Cin; p.Wout = d->Wout; p.kw = d->kw;
  p.stride = d->stride; p.pad = d->pad;
  p.alpha = d->alpha; p.bias = d->bias; p.bias_m = d->bias_along_m; p.act = d->act; p.act_from = d->act_from_col;
  p.res = (const t16*)d->residual; p.ldr = d->ldr; p.gate = (const t16*)d->gate; p.ldg = d->ldg;
  p.accumulate = d->accumulate;
  p.C = d->C; p.ldc = d->ldc; p.c_fp32 = d->c_is_f32;
  p.C2 = d->C2; p.ldc2 = d->ldc2; p.c2_fp32 = d->c2_is_f32;
  p.n_split = d->C2 ? d->n_split : (1 << 30);
  p.pad_n_ok = d->pad_n_ok;
  p.nimg = d->n_images > 1 ? d->n_images : 1;
  return gemm_launch(p, STREAM(stream));
  RMEM_API_END
}

int rmem_set_gemm_impl(int impl) {
  RMEM_API_BEGIN
  int prev = gemm_impl_switch();
  gemm_impl_switch() = impl;
  return prev;
  RMEM_API_END
}

int rmem_long_attn_workspace_bytes(int impl, int HW, int HWp, int nslots, int Dv, size_t* bytes) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(bytes, "null bytes");
  RMEM_REQUIRE(impl == RMEM_ATTN_DENSE || impl == RMEM_ATTN_TC2 || impl == RMEM_ATTN_TC3 || impl == RMEM_ATTN_TC4,
               "attention impl %d", impl);
  *bytes = impl == RMEM_ATTN_TC4 ? long_attn_tc4_workspace(HW, HWp, nslots, Dv)
           : impl == RMEM_ATTN_TC3 ? long_attn_tc3_workspace(HW, HWp, nslots, Dv)
           : impl == RMEM_ATTN_TC2 ? long_attn_tc2_workspace(HW, HWp, nslots, Dv)
                                   : long_attn_dense_workspace(HW, HWp, nslots);
  return RMEM_OK;
  RMEM_API_END
}

int rmem_long_attn_fwd(int impl, const void* qt, const float* qbias, const void* kbank, const void* vtbank, int nslots,
                       int T, const int* slots, int HW, int HWp, int Dk, int Dv, float scale, const void* gate,
                       long long ldg, void* out, long long ldo, float* mass, void* workspace, size_t workspace_bytes,
                       void* stream) {
  RMEM_API_BEGIN
  return rmem_long_attn_grid_fwd(impl, qt, qbias, kbank, vtbank, nslots, T, slots, HW, HWp, Dk, Dv, scale, gate, ldg, out,
                                 ldo, mass, 0, 0, workspace, workspace_bytes, stream);
  RMEM_API_END
}

int rmem_long_attn_grid_fwd(int impl, const void* qt, const float* qbias, const void* kbank, const void* vtbank,
                            int nslots, int T, const int* slots, int HW, int HWp, int Dk, int Dv, float scale,
                            const void* gate, long long ldg, void* out, long long ldo, float* mass, int grid_h,
                            int grid_w, void* workspace, size_t workspace_bytes, void* stream) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(qt && kbank && vtbank && out && slots && workspace, "null argument");
  RMEM_REQUIRE(T >= 1 && T <= kMaxBankFrames, "T=%d out of range", T);
  RMEM_REQUIRE((grid_h == 0 && grid_w == 0) || (grid_h > 0 && grid_w > 0 && grid_h * grid_w == HW),
               "token grid %dx%d does not match HW=%d", grid_h, grid_w, HW);
  LongAttnArgs a;
  a.qt = (const t16*)qt; a.qbias = qbias; a.kbank = (const t16*)kbank; a.vtbank = (const t16*)vtbank;
  a.nslots = nslots; a.T = T;
  for (int t = 0; t < T; ++t) a.slot[t] = slots[t];
  a.HW = HW; a.HWp = HWp; a.Dk = Dk; a.Dv = Dv; a.scale = scale;
  a.gate = (const t16*)gate; a.ldg = ldg; a.out = (t16*)out; a.ldo = ldo; a.mass = mass;
  a.seed_h = grid_h; a.seed_w = grid_w;
  if (impl == RMEM_ATTN_TC4) return long_attn_tc4(a, workspace, workspace_bytes, STREAM(stream));
  if (impl == RMEM_ATTN_TC3) return long_attn_tc3(a, workspace, workspace_bytes, STREAM(stream));
  if (impl == RMEM_ATTN_TC2) return long_attn_tc2(a, workspace, workspace_bytes, STREAM(stream));
  RMEM_REQUIRE(impl == RMEM_ATTN_DENSE, "attention impl %d (0 dense, 2 tc2, 3 tc3)", impl);
  return long_attn_dense(a, workspace, workspace_bytes, STREAM(stream));
  RMEM_API_END
}

int rmem_mha_workspace_bytes(int impl, int HW, int HWp, int nslots, int H, size_t* bytes) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(bytes, "null argument");
  *bytes = impl == RMEM_ATTN_DENSE ? mha_dense_workspace(HW, HWp, nslots, H) : mha_tc_workspace(HW, HWp, nslots, H);
  return RMEM_OK;
  RMEM_API_END
}

int rmem_mha_fwd(int impl, const void* q, long long ldq, const void* kbank, const void* vtbank, int nslots, int T,
                 const int* slots, int HW, int HWp, int H, int dh, float scale, const float* qbias, void* out,
                 long long ldo, float* mass, void* workspace, size_t workspace_bytes, void* stream) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(q && kbank && vtbank && out && slots && workspace, "null argument");
  RMEM_REQUIRE(T >= 1 && T <= kMaxBankFrames, "T=%d out of range", T);
  MhaArgs a;
  a.q = (const t16*)q; a.ldq = ldq; a.kbank = (const t16*)kbank; a.vtbank = (const t16*)vtbank;
  a.nslots = nslots; a.T = T;
  for (int t = 0; t < T; ++t) a.slot[t] = slots[t];
  a.HW = HW; a.HWp = HWp; a.H = H; a.dh = dh; a.scale = scale; a.qbias = qbias;
  a.out = (t16*)out; a.ldo = ldo; a.mass = mass;
  if (impl == RMEM_ATTN_DENSE) return mha_dense(a, workspace, workspace_bytes, STREAM(stream));
  return mha_tc(a, workspace, workspace_bytes, STREAM(stream));
  RMEM_API_END
}

int rmem_debug_gemm_trace(void* dev_buf) { return gemm_tc_set_trace(reinterpret_cast<long long*>(dev_buf)); }
int rmem_debug_gemm_force(int bn, int splitk, int stages) {
  RMEM_API_BEGIN
  return gemm_tc_set_force(bn, splitk, stages);
  RMEM_API_END
}
int rmem_debug_gemm_log(int* host_buf, int cap_records) {
  RMEM_API_BEGIN
  return gemm_tc_set_log(host_buf, cap_records);
  RMEM_API_END
}
int rmem_debug_gemm_log_count(void) { return gemm_tc_log_count(); }
int rmem_debug_attn_events(void* ev0, void* ev1) {
  RMEM_API_BEGIN
  long_attn_tc2_set_events(ev0, ev1);
  long_attn_tc3_set_events(ev0, ev1);
  return RMEM_OK;
  RMEM_API_END
}
int rmem_debug_attn_rescale_counter(void* dev_int) {
  RMEM_API_BEGIN
  long_attn_tc3_set_rescale_counter(reinterpret_cast<int*>(dev_int));
  return RMEM_OK;
  RMEM_API_END
}
int rmem_debug_attn_schedule(int impl, int HW, int T, int Dv, int* n_units, int* tiles_per_unit, int* n_cta, int* bounds,
                             int cap) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(n_units && tiles_per_unit && n_cta && bounds, "null argument");
  if (impl == RMEM_ATTN_TC3) return long_attn_tc3_schedule(HW, T, Dv, n_units, tiles_per_unit, n_cta, bounds, cap);
  RMEM_REQUIRE(impl == RMEM_ATTN_TC2, "attention impl %d has no static schedule", impl);
  return long_attn_tc2_schedule(HW, T, Dv, n_units, tiles_per_unit, n_cta, bounds, cap);
  RMEM_API_END
}

int rmem_debug_attn_trace(void* dev_buf) {
  RMEM_API_BEGIN
  RMEM_TRY(local_attn_tc_set_trace(reinterpret_cast<long long*>(dev_buf)));
  RMEM_TRY(long_attn_tc3_set_trace(reinterpret_cast<long long*>(dev_buf)));
  return long_attn_tc2_set_trace(reinterpret_cast<long long*>(dev_buf));
  RMEM_API_END
}

int rmem_qprep_fwd(const void* q, long long ldq, const float* pe_cur, const float* pe_mem, const int* pe_slot, int T,
                   float scale, void* qt, float* qbias, int P, int C, void* stream) {
  RMEM_API_BEGIN
  return qprep((const t16*)q, ldq, pe_cur, pe_mem, pe_slot, T, scale, (t16*)qt, qbias, P, C, STREAM(stream));
  RMEM_API_END
}

int rmem_temporal_pe_slots(int T, int n_slots, int* out) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(out && T >= 1 && T <= kMaxBankFrames && n_slots >= 1, "bad argument");
  temporal_pe_slots(T, n_slots, out);
  return RMEM_OK;
  RMEM_API_END
}

int rmem_local_attn_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                        const float* rel, long long ldrel, const void* gate, long long ldg, void* out, long long ldo,
                        int h, int w, int Dv, float scale, void* stream) {
  RMEM_API_BEGIN
  return local_attn((const t16*)q, ldq, (const t16*)k, ldk, (const t16*)v, ldv, rel, ldrel, (const t16*)gate, ldg,
                    (t16*)out, ldo, h, w, Dv, scale, STREAM(stream));
  RMEM_API_END
}

int rmem_local_attn_tc_workspace_bytes(int h, int w, int Dv, size_t* bytes) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(bytes, "null bytes");
  *bytes = local_attn_tc_workspace(h, w, Dv);
  return RMEM_OK;
  RMEM_API_END
}

int rmem_local_attn_tc_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                           const float* rel, long long ldrel, int rel_pitch, const void* gate, long long ldg,
                           void* out, long long ldo, int h, int w, int Dv, float scale, void* workspace,
                           size_t workspace_bytes, void* stream) {
  RMEM_API_BEGIN
  return local_attn_tc((const t16*)q, ldq, (const t16*)k, ldk, (const t16*)v, ldv, rel, ldrel, rel_pitch,
                       (const t16*)gate, ldg,
                       (t16*)out, ldo, h, w, Dv, scale, workspace, workspace_bytes, STREAM(stream));
  RMEM_API_END
}

int rmem_layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, void* y, long long ldy,
                       int P, int C, void* stream) {
  RMEM_API_BEGIN
  return layernorm(x, ldx, gamma, beta, (t16*)y, ldy, nullptr, 0, P, C, STREAM(stream));
  RMEM_API_END
}

int rmem_groupnorm_fwd(const void* x, int x_is_f32, const float* gamma, const float* beta, void* y, int P, int C, int G,
                       int relu, double* stats, void* stream) {
  RMEM_API_BEGIN
  if (x_is_f32) return groupnorm_f32((const float*)x, gamma, beta, (t16*)y, P, C, G, relu, stats, STREAM(stream));
  return groupnorm_t16((const t16*)x, gamma, beta, (t16*)y, P, C, G, relu, stats, STREAM(stream));
  RMEM_API_END
}

int rmem_dwconv5x5_fwd(const void* x, const float* w, void* y, int h, int w_, int C, void* stream) {
  RMEM_API_BEGIN
  return dwconv5x5((const t16*)x, w, (t16*)y, h, w_, C, STREAM(stream));
  RMEM_API_END
}

int rmem_upsample_bilinear_fwd(const void* x, void* y, int hin, int win, int hout, int wout, int C, void* stream) {
  RMEM_API_BEGIN
  return upsample_bilinear_t16((const t16*)x, (t16*)y, hin, win, hout, wout, C, STREAM(stream));
  RMEM_API_END
}

int rmem_transpose_fwd(const void* x, long long ldx, void* y, long long ldy, int P, int C, void* stream) {
  RMEM_API_BEGIN
  return transpose_t16((const t16*)x, ldx, (t16*)y, ldy, P, C, STREAM(stream));
  RMEM_API_END
}

int rmem_maxpool3x3s2_fwd(const void* x, void* y, int Hin, int Win, int C, int Hout, int Wout, void* stream) {
  RMEM_API_BEGIN
  return maxpool3x3s2((const t16*)x, (t16*)y, Hin, Win, C, Hout, Wout, STREAM(stream));
  RMEM_API_END
}

int rmem_pack_image_fwd(const float* img, void* out, int H, int W, void* stream) {
  RMEM_API_BEGIN
  return pack_image(img, (t16*)out, H, W, STREAM(stream));
  RMEM_API_END
}

int rmem_pack_image_padded_fwd(const float* img, void* out, int H, int W, void* stream) {
  return pack_image_padded(img, (t16*)out, H, W, STREAM(stream));
}

int rmem_idbank_fwd(const uint8_t* label, int H, int W, int use_ignore, const float* w_packed, const float* prefix,
                    const float* prefix_rows,
                    const float* bias, const float* ln_gamma, const float* ln_beta, void* out_t16, long long ldo,
                    float* out_f32, int h, int w, int C, void* stream) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(label && w_packed && bias && (out_t16 || out_f32), "null argument");
  return idbank_embed(label, H, W, use_ignore, w_packed, bias, ln_gamma, ln_beta, (t16*)out_t16, ldo, out_f32, h, w,
                      C, STREAM(stream), prefix, prefix_rows);
  RMEM_API_END
}

int rmem_mask_head_fwd(const float* const* logits4, int k, int h4, int w4, int Ho, int Wo, float* out_logits,
                       uint8_t* out_label, void* stream) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(logits4, "null argument");
  return mask_head(logits4, k, h4, w4, Ho, Wo, out_logits, out_label, STREAM(stream));
  RMEM_API_END
}

int rmem_tta_head_fwd(const float* const* logits4, int n_aug, int k, const int* h4, const int* w4, const int* flip,
                      int Ho, int Wo, float* out_prob, uint8_t* out_label, void* stream) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(logits4 && h4 && w4 && flip && (out_prob || out_label), "null argument");
  return tta_head(logits4, n_aug, k, h4, w4, flip, Ho, Wo, out_prob, out_label, STREAM(stream));
  RMEM_API_END
}

int rmem_preprocess_fwd(const uint8_t* img, int H, int W, int bgr, int nh, int nw, int flip, float* out, void* stream) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(img && out, "null argument");
  return preprocess_frame(img, H, W, bgr, nh, nw, flip, out, STREAM(stream));
  RMEM_API_END
}

int rmem_evict_relevance_fwd(const float* mass, int T, const float* logits4, int h4, int w4, int h, int w, float* rel,
                             void* stream) {
  RMEM_API_BEGIN
  return evict_relevance(mass, T, logits4, h4, w4, h, w, rel, STREAM(stream));
  RMEM_API_END
}

int rmem_evict_pick(const float* rel_host, int T_old, const int* idx, int former, int* ema_keys, float* ema_vals,
                    int* n_ema, int* times_keys, int* times_vals, int* n_times, int* drop) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(rel_host && idx && ema_keys && ema_vals && n_ema && times_keys && times_vals && n_times && drop,
               "null argument");
  RMEM_REQUIRE(T_old >= 1 && T_old <= kMaxBankFrames, "T_old=%d out of range", T_old);
  std::map<int, float> ema;
  std::map<int, int> times;
  for (int i = 0; i < *n_ema; ++i) ema[ema_keys[i]] = ema_vals[i];
  for (int i = 0; i < *n_times; ++i) times[times_keys[i]] = times_vals[i];
  RMEM_TRY(evict_pick_host(rel_host, T_old, idx, former, ema, times, drop, nullptr));
  int i = 0;
  for (auto& kv : ema) { ema_keys[i] = kv.first; ema_vals[i] = kv.second; ++i; }
  *n_ema = i;
  i = 0;
  for (auto& kv : times) { times_keys[i] = kv.first; times_vals[i] = kv.second; ++i; }
  *n_times = i;
  return RMEM_OK;
  RMEM_API_END
}

}  // extern "C"
