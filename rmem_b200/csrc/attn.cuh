// Attention contractions of the RMem path.
#pragma once
#include "common.cuh"

namespace rmem {

constexpr int kMaxBankFrames = 16;

// Long-term (and self) attention over the restricted bank:
//   P = softmax_j( scale * <qt_i, K_j> + qbias[i, t(j)] )   over all live frames' tokens
//   out[i,:] = (sum_j P_ij * V_j) * gate[i,:] ;   mass[i,t] = sum_{j in frame t} P_ij
// Bank layout (ours, ring of physical slots):  kbank [nslots][HWp][Dk] t16 (token-major),
// vtbank [Dv][nslots*HWp] t16 (value-major = K-major for the PV contraction, frame `s` at columns s*HWp..).
struct LongAttnArgs {
  const t16* qt = nullptr;      // [HW, Dk]
  const float* qbias = nullptr;  // [HW, T] or null
  const t16* kbank = nullptr;
  const t16* vtbank = nullptr;
  int nslots = 1;
  int T = 1;
  int slot[kMaxBankFrames] = {0};  // logical frame t -> physical slot
  int HW = 0, HWp = 0, Dk = 128, Dv = 1024;
  float scale = 1.f;
  const t16* gate = nullptr;    // [HW, Dv] or null
  long long ldg = 0;
  t16* out = nullptr;           // [HW, Dv]
  long long ldo = 0;
  float* mass = nullptr;         // [HW, T] or null
  // RMEM_ATTN_TC3 only: token grid (seed_h * seed_w == HW) for the row-maximum seed (scores against the 3x3 neighbourhood
  // of the query's own position in every bank frame); 0 = no seed (the first key tile seeds the maximum)
  int seed_h = 0, seed_w = 0;
  // RMEM_ATTN_TC3: seed already computed by qprep_seed_tc3 ([HW] fp32, log2 units); the kernel then launches no seed pass
  const float* mseed = nullptr;
};

// Materialised-score implementation (generic GEMM + row softmax).  Workspace: S fp32 + P t16.
size_t long_attn_dense_workspace(int HW, int HWp, int nslots);
int long_attn_dense(const LongAttnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s);

// Multi-head attention over a bank (AOT: 8 heads x 32): q [HW, H*dh] (row stride ldq), kbank [nslots][HWp][H*dh],
// vtbank [H*dh][nslots*HWp], qbias [H][HW][T] or null, out [HW, H*dh] (row stride ldo), mass [HW, T] = head mean or null.
struct MhaArgs {
  const t16* q = nullptr;
  long long ldq = 0;
  const t16* kbank = nullptr;
  const t16* vtbank = nullptr;
  int nslots = 1, T = 1;
  int slot[kMaxBankFrames] = {0};
  int HW = 0, HWp = 0, H = 8, dh = 32;
  float scale = 1.f;
  const float* qbias = nullptr;
  t16* out = nullptr;
  long long ldo = 0;
  float* mass = nullptr;
};
size_t mha_dense_workspace(int HW, int HWp, int nslots, int H);
int mha_dense(const MhaArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s);
// Fused tcgen05 version (mha_tc.cu): scores, online softmax and P.V of all 8 heads of a (128-query, 64-key) step in one
// CTA, P through TMEM, stream-K over the SMs; the engine's AOT path (attn_impl != dense).
size_t mha_tc_workspace(int HW, int HWp, int nslots, int H);
int mha_tc(const MhaArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s);

// v2 (attn_tc2.cu): stream-K schedule over the SMs, 8 softmax warps, P through TMEM, fp16 partials.
size_t long_attn_tc2_workspace(int HW, int HWp, int nslots, int Dv);
int long_attn_tc2_schedule(int HW, int T, int Dv, int* n_units, int* tiles_per_unit, int* n_cta, int* bounds, int cap);
int long_attn_tc2(const LongAttnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s);
// v3 (attn_tc3.cu): CTA pairs (tcgen05.mma.cta_group::2) share every K / V^T tile, 128-key score MMAs, seeded row maximum.
size_t long_attn_tc3_workspace(int HW, int HWp, int nslots, int Dv);
int long_attn_tc3_schedule(int HW, int T, int Dv, int* n_units, int* groups_per_unit, int* n_clusters, int* bounds, int cap);
int long_attn_tc3(const LongAttnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s);
// qprep (Qt, qbias: ops.cuh) and the row-maximum seed of the long-term attention in ONE launch (Dk = 128).
int qprep_seed_tc3(const t16* q, long long ldq, const float* pe_cur, const float* pe_mem, const int* pe_slot, int T,
                   float scale, const t16* kbank, const int* slot, int HW, int HWp, int h, int w, t16* qt, float* qbias,
                   float* mseed, cudaStream_t s, int* zero_flag = nullptr);
// v4 (attn_tc3.cu, long_attn_tc4_kernel): column kernel -- scores and exponentials once per (query pair, sub-tile), the
// probabilities re-read from L2 for the other Dv chunks; fixed softmax reference from the seed, guarded tc3 fallback.
// Falls through to long_attn_tc3 for unseeded calls and banks of fewer than 3 frames.  The first int of the workspace is
// the overflow flag: qprep_seed_tc3(..., zero_flag = workspace) arms it when the caller provides a.mseed.
size_t long_attn_tc4_workspace(int HW, int HWp, int nslots, int Dv);
int long_attn_tc4(const LongAttnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s);
int long_attn_tc3_set_trace(long long* dev_buf);
void long_attn_tc3_set_events(void* ev0, void* ev1);
// Debug: device int incremented once per warp-level lazy-rescale event (null disables).  Thread-local.
void long_attn_tc3_set_rescale_counter(int* dev_counter);
// Debug: clock64 event trace of CTA 0 ([tile][16] long long, see attn_tc2.cu); nullptr disables.
int long_attn_tc2_set_trace(long long* dev_buf);
// Measurement aid: record these CUDA events around the main kernel launch only (null clears).  Thread-local.
void long_attn_tc2_set_events(void* ev0, void* ev1);

// Windowed short-term attention (LocalGatedPropagation core, attention.py:289-353), 15x15 window:
//   s[i,d] = scale*<q_i, k_{i+d}> + rel[i,d] ; p = softmax_d ; out_i = (sum_d p[i,d] v_{i+d}) * gate_i
// q,k [HW,128]; v [HW,Dv]; rel fp32 [HW, ldrel] (first 225 columns); all token-major with row strides.
int local_attn(const t16* q, long long ldq, const t16* k, long long ldk, const t16* v, long long ldv,
               const float* rel, long long ldrel, const t16* gate, long long ldg, t16* out, long long ldo, int h,
               int w, int Dv, float scale, cudaStream_t s);

// Tensor-core version (local_attn_tc.cu): 8x16 query patches x 256-column Dv chunks, key halo walked as 2x32 tiles via
// 3-D TMA boxes; needs a value-major padded copy of v in `workspace` (local_attn_tc_workspace bytes, 256B aligned).
size_t local_attn_tc_workspace(int h, int w, int Dv);
int local_attn_tc_set_trace(long long* dev_buf);
int local_attn_tc(const t16* q, long long ldq, const t16* k, long long ldk, const t16* v, long long ldv,
                  const float* rel, long long ldrel, int rel_pitch, const t16* gate, long long ldg, t16* out, long long ldo, int h,
                  int w, int Dv, float scale, void* workspace, size_t workspace_bytes, cudaStream_t s,
                  bool v_prepared = false);
// The value-major copy alone (same stream order as the attention that reads it): lets a caller whose v is ready early
// take the transpose off the path between q and the output; pass v_prepared = true to local_attn_tc afterwards.
int local_attn_tc_prepare_v(const t16* v, long long ldv, int h, int w, int Dv, void* workspace, size_t workspace_bytes,
                            cudaStream_t s);

}  // namespace rmem
