// Host-side per-clip state machine + frame orchestration of the DeAOT+RMem propagation path.
//
// Mirrors (behaviour, not code) of the reference, paths relative to /root/reference/aot_plus/:
//   AOTInferEngine / DeAOTInferEngine   networks/engines/aot_engine.py:571-725, deaot_engine.py:20-56
//   AOTEngine.{add_reference_frame, match_propogate_one_frame, update_short_term_memory}
//                                        networks/engines/aot_engine.py:241-436
//   DeAOT.{encode_image, get_id_emb, LSTT_forward, decode_id_logits}   networks/models/{aot,deaot}.py
//   DualBranchGPM / GatedPropagationModule                             networks/layers/transformer.py:700-1249
//   restrict_long_memories                                             networks/layers/transformer.py:880-991
//
// B200-first differences from the reference's structure: the bank is a fixed ring of physical slots in
// a caller-provided HBM arena (no torch.cat re-allocation), V / ID_V are stored value-major so the
// P.V contraction is K-major on both operands, the temporal positional embedding is a score bias
// instead of a re-materialised K bank, the encoder runs once per frame for all object groups, and the
// only host<->device sync is the T-float relevance read-back on frames that append to the bank.
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <cmath>
#include <cstring>

#include "../../include/rmem_b200.h"
#include "attn.cuh"
#include "common.cuh"
#include "gemm.cuh"
#include "ops.cuh"

namespace rmem {

int evict_pick_host(const float* rel_raw, int T_old, const int* idx, int former, std::map<int, float>& ema,
                    std::map<int, int>& times, int* drop, float* rel_norm_out, bool gru = false);

namespace {

constexpr int kLayers = 3;
constexpr int kEncGroup = 4;    // most frames one encoder pass covers (rmem_engine_prefetch_n)
constexpr int kFeatSets = 8;    // encoder output sets: two groups of kEncGroup contiguous members
constexpr int kD = 256;      // d_model
constexpr int kDk = 128;     // att dim
constexpr int kDv = 1024;    // V (512) || ID_V (512)
constexpr int kMaxObj = 10;

struct Geo {
  int H, W, H1, W1, H4, W4, H8, W8, h, w;
  int P1, P4, P8, HW, HWp;
};

Geo make_geo(int H, int W) {
  Geo g;
  g.H = H; g.W = W;
  g.H1 = (H - 1) / 2 + 1; g.W1 = (W - 1) / 2 + 1;      // conv1 k7 s2 p3
  g.H4 = (g.H1 - 1) / 2 + 1; g.W4 = (g.W1 - 1) / 2 + 1;  // maxpool k3 s2 p1
  g.H8 = (g.H4 - 1) / 2 + 1; g.W8 = (g.W4 - 1) / 2 + 1;  // layer2 stride 2
  g.h = (g.H8 - 1) / 2 + 1; g.w = (g.W8 - 1) / 2 + 1;    // layer3 stride 2
  g.P1 = g.H1 * g.W1; g.P4 = g.H4 * g.W4; g.P8 = g.H8 * g.W8;
  g.HW = g.h * g.w;
  g.HWp = round_up(g.HW, 128);
  return g;
}

struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  bool dry = false;
  template <typename T>
  T* take(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    size_t o = off;
    off += bytes;
    if (dry) return nullptr;
    return (o + bytes <= cap) ? reinterpret_cast<T*>(base + o) : nullptr;
  }
};

struct LayerState {
  t16* kc[2];       // [HW,128]   current / previous frame K (= Q)
  t16* vid[2];      // [HW,1024]  V || ID_V, token-major (short-term memory + source of the bank append)
  t16* cat;         // [HW,512]   curr_ID_V || id_emb  (linear_ID_V input; layer 0 uses only the id half)
  t16* kbank;       // [nslots][HWp][128]
  t16* vtbank;      // [1024][nslots*HWp]
};

struct LayerStateA {  // AOT (SimplifiedTransformerBlock) per-layer memories, transformer.py:438-441, 689-692
  t16* k_cur;       // [HWp,256]  Q = K of the current frame
  t16* v_cur;       // [HW,256]   LN2(tgt) = V of the current frame
  t16* v_lin;       // [HW,256]   linear_V(V + id): what the long bank stores
  t16* sk[2];       // [HWp,256]  short-term K = linear_QMem(o3)
  t16* sv[2];       // [HW,256]   short-term V = linear_VMem(o3 + id)
  t16* kbank;       // [nslots][HWp][256]
  t16* vtbank;      // [256][nslots*HWp]
  float* gru_h[2];  // GRU_MEMORY: fp32 hidden state of the K / V ConvGRU [HW,256] (zeroed per clip with the state region)
};

struct Group {       // one AOTEngine: <= 10 objects, own bank (reference + per-engine deepcopy)
  LayerState L[kLayers];
  LayerStateA A[kLayers];
  float* mass0;      // [HW,16] layer-0 attention mass of the last propagate (logical frame order)
  int mass_T = 0;
  float* logits4;    // planar [11, P4] = pred_id_logits
  int parity = 0;    // which of kc/vid is "current"
  std::vector<int> slots;       // logical frame -> physical slot
  std::vector<int> free_slots;
  std::vector<int> long_idx;    // long_memories_indexes
  std::map<int, float> ema;     // stored_attn_weight_dict
  std::map<int, int> times;     // stored_frame_times
  int frame_step = 0, last_mem_step = -1;
  bool has_ref = false;
  std::vector<float> last_rel;
  int last_drop = -1;
  // deferred eviction: update_memory leaves the T relevance floats in flight to pinned host memory and returns; the
  // table edit happens at the next call that reads the tables (rmem_engine::resolve_evict) -- no host sync per append
  bool evict_pending = false;
  int evict_T_old = 0;
};

}  // namespace
}  // namespace rmem

using namespace rmem;

struct rmem_engine {
  rmem_engine_config cfg;
  Geo g;
  int nslots = 0;
  std::map<std::string, std::pair<const char*, size_t>> weights;
  int n_groups = 0;
  std::vector<Group> groups;
  long long launches0 = 0;

  // shared scratch
  t16 *img8, *c1, *x0, *x1, *t1, *t2, *ds, *feat4, *feat8, *feat16;
  size_t img8_elems = 0;
  float* enc_tgt;     // [HW,256] projector output
  // Encoder outputs are double-buffered: the image encoder does not depend on the memory state, so the NEXT frame can be
  // encoded on a side stream (rmem_engine_prefetch) while this frame's propagation / decoder / memory update run.
  // feat4/feat8/feat16/enc_tgt above always point at the set the current frame reads.
  // Eight feature sets = two GROUPS of four contiguous members ([4][pixels][channels]): rmem_engine_prefetch_n encodes
  // two or four frames in one pass (every encoder GEMM / conv runs once over all images, M multiplied) into an aligned
  // pair / a whole group, a single prefetch or the inline encoder uses one member.
  t16 *feat4s[kFeatSets], *feat8s[kFeatSets], *feat16s[kFeatSets];
  float* enc_tgts[kFeatSets];
  // The decoder's adapter convolutions (fpn.py:44,50,56: 1x1 convs of the ENCODER features) do not depend on the memory
  // state either: they are computed with the encoder (prefetch stream) and added by the decoder's GroupNorm / upsample
  // kernels instead of sitting on the per-frame critical path as three more GEMM launches per object group.
  t16 *ad16s[kFeatSets], *ad8s[kFeatSets], *ad4s[kFeatSets], *ad16, *ad8, *ad4;
  int fslot = 0;
  // Independent GEMMs of one layer (U / ID_U next to QV, the self-attention projections, the three linear_ID_V of the
  // memory update) are forked onto a second engine-owned stream: each is a ~3 us kernel behind ~5 us of launch latency.
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool use_aux() const { return aux_stream != nullptr && !timing && aux_on; }
  bool aux_on = true, branch_par = true;
  int fork(cudaStream_t s) {
    RMEM_CUDA_CHECK(cudaEventRecord(ev_fork, s));
    RMEM_CUDA_CHECK(cudaStreamWaitEvent(aux_stream, ev_fork, 0));
    return RMEM_OK;
  }
  int join(cudaStream_t s) {
    RMEM_CUDA_CHECK(cudaEventRecord(ev_join, aux_stream));
    RMEM_CUDA_CHECK(cudaStreamWaitEvent(s, ev_join, 0));
    return RMEM_OK;
  }
  float* rel_pinned = nullptr;             // [max_engines][kMaxBankFrames] pinned host landing zone of the relevance sums
  cudaEvent_t ev_rel[4] = {nullptr, nullptr, nullptr, nullptr};
  // transformer.py:891-964 on the host, run lazily: waits for the T floats of the last append (normally long complete),
  // then EMA / UCB / argmin and the slot-table edit exactly as before.
  int resolve_evict(int gi, cudaStream_t s = nullptr) {
    Group& gr = groups[gi];
    if (!gr.evict_pending) return RMEM_OK;
    gr.evict_pending = false;
    RMEM_CUDA_CHECK(cudaEventSynchronize(ev_rel[gi]));
    const int T_old = gr.evict_T_old;
    const int cap = cfg.former_mem_len + cfg.latter_mem_len;
    int drop = cfg.former_mem_len;
    gr.last_rel.assign(T_old, 0.f);
    RMEM_TRY(evict_pick_host(rel_pinned + (size_t)gi * kMaxBankFrames, T_old, gr.long_idx.data(), cfg.former_mem_len, gr.ema,
                             gr.times, &drop, gr.last_rel.data(), cfg.gru_memory != 0));
    gr.last_drop = drop;
    if ((int)gr.slots.size() > cap) {
      if (cfg.gru_memory) RMEM_TRY(gru_condense(gr, drop, s));   // GRU_MEMORY resolves inside update_memory (stream known)
      gr.free_slots.push_back(gr.slots[drop]);
      gr.slots.erase(gr.slots.begin() + drop);
      gr.long_idx.erase(gr.long_idx.begin() + drop);
    }
    return RMEM_OK;
  }
  int resolve_all() {
    for (int gi = 0; gi < (int)groups.size(); ++gi) RMEM_TRY(resolve_evict(gi));
    return RMEM_OK;
  }
  cudaStream_t enc_stream = nullptr;
  cudaEvent_t ev_img_ready = nullptr, ev_inline = nullptr, ev_done[kFeatSets] = {}, ev_feat_free[kFeatSets] = {};
  bool feat_free_valid[kFeatSets] = {}, inline_valid = false;
  bool pending[kFeatSets] = {};            // features of pf_img[sl] are (being) produced on enc_stream, not consumed yet
  const float* pf_img[kFeatSets] = {};
  long long pf_seq[kFeatSets] = {}, pf_counter = 0;
  int pf_age[kFeatSets] = {};              // features() calls since the prefetch was issued
  int pf_ttl[kFeatSets] = {};              // ... at which an unconsumed entry is stale (2: single prefetch, 4: pair)
  // slot for a single frame: not pending, preferably not the current one; all pending -> the oldest (stale) entry
  int pick_slot() const {
    for (int k = 1; k <= kFeatSets; ++k) {
      const int sl = (fslot + k) % kFeatSets;          // k = kFeatSets: the current slot itself (its readers are all issued)
      if (!pending[sl]) return sl;
    }
    int o = 0;
    for (int sl = 1; sl < kFeatSets; ++sl)
      if (pf_seq[sl] < pf_seq[o]) o = sl;
    return o;
  }
  // first slot of an aligned run of n (2 | 4) slots with no unconsumed prefetch -- preferably outside the current frame's
  // group -- else of the run whose newest entry is the oldest (stale)
  int pick_run(int n) const {
    const int groups = kFeatSets / kEncGroup, cur_group = fslot / kEncGroup;
    for (int k = 1; k <= groups; ++k) {
      const int base = ((cur_group + k) % groups) * kEncGroup;
      for (int r = 0; r + n <= kEncGroup; r += n) {
        bool free_run = true;
        for (int j = 0; j < n; ++j) free_run = free_run && !pending[base + r + j];
        if (free_run) return base + r;
      }
    }
    long long best = -1;
    int o = 0;
    for (int r = 0; r + n <= kFeatSets; r += n) {
      long long newest = 0;
      for (int j = 0; j < n; ++j) newest = pf_seq[r + j] > newest ? pf_seq[r + j] : newest;
      if (best < 0 || newest < best) { best = newest; o = r; }
    }
    return o;
  }
  int last_enc_slot = -1;                  // last slot encoded on enc_stream (its event orders the encoder temporaries)
  void use_slot(int sl) {
    fslot = sl; feat4 = feat4s[sl]; feat8 = feat8s[sl]; feat16 = feat16s[sl]; enc_tgt = enc_tgts[sl];
    ad16 = ad16s[sl]; ad8 = ad8s[sl]; ad4 = ad4s[sl];
  }
  int init_streams() {
    // Stream priorities were measured and do not help (c3, 1.335 ms per frame with equal priorities): a low-priority
    // prefetch stream under a high-priority caller stream loses the overlap altogether (1.52 ms), a high-priority second
    // stream alone costs 0.05 ms.  RMEM_ENC_PRIO = 1 / -1 gives the prefetch stream the highest / lowest priority.
    int prio_least = 0, prio_greatest = 0;
    RMEM_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    int enc_prio = prio_least;
    { const char* e = getenv("RMEM_ENC_PRIO"); if (e && e[0] == '1') enc_prio = prio_greatest; }
    RMEM_CUDA_CHECK(cudaStreamCreateWithPriority(&enc_stream, cudaStreamNonBlocking, enc_prio));
    RMEM_CUDA_CHECK(cudaStreamCreateWithFlags(&aux_stream, cudaStreamNonBlocking));
    RMEM_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    RMEM_CUDA_CHECK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    { const char* e = getenv("RMEM_AUX_STREAM"); aux_on = !(e && e[0] == '0'); }
    { const char* e = getenv("RMEM_BRANCH_PAR"); branch_par = !(e && e[0] == '0'); }
    { const char* e = getenv("RMEM_FUSED_SEED"); fused_seed = !(e && e[0] == '0'); }
    { const char* e = getenv("RMEM_SELF_SEED"); self_seed = !(e && e[0] == '0'); }
    RMEM_CUDA_CHECK(cudaEventCreateWithFlags(&ev_img_ready, cudaEventDisableTiming));
    RMEM_CUDA_CHECK(cudaEventCreateWithFlags(&ev_inline, cudaEventDisableTiming));
    for (int i = 0; i < kFeatSets; ++i) {
      RMEM_CUDA_CHECK(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming));
      RMEM_CUDA_CHECK(cudaEventCreateWithFlags(&ev_feat_free[i], cudaEventDisableTiming));
    }
    RMEM_CUDA_CHECK(cudaMallocHost(&rel_pinned, sizeof(float) * 4 * kMaxBankFrames));
    for (int i = 0; i < 4; ++i) RMEM_CUDA_CHECK(cudaEventCreateWithFlags(&ev_rel[i], cudaEventDisableTiming));
    return RMEM_OK;
  }
  ~rmem_engine() {
    if (enc_stream) {
      cudaStreamSynchronize(enc_stream);
      cudaStreamDestroy(enc_stream);
    }
    if (aux_stream) {
      cudaStreamSynchronize(aux_stream);
      cudaStreamDestroy(aux_stream);
    }
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    for (int i = 0; i < 4; ++i)
      if (ev_rel[i]) cudaEventDestroy(ev_rel[i]);
    if (rel_pinned) cudaFreeHost(rel_pinned);
    if (ev_img_ready) cudaEventDestroy(ev_img_ready);
    if (ev_inline) cudaEventDestroy(ev_inline);
    for (int i = 0; i < kFeatSets; ++i) {
      if (ev_done[i]) cudaEventDestroy(ev_done[i]);
      if (ev_feat_free[i]) cudaEventDestroy(ev_feat_free[i]);
    }
  }
  float* res;         // [HW,512] tgt || tgt_id residual stream
  t16 *attn_b, *dwo2;
  t16 *t_ln, *qt, *cu, *cu0, *attn_a, *dwo, *z, *qk, *vt_self, *u_self, *gpm_out, *idemb;
  float *qbias, *rel, *rel_dev, *mseed;
  bool self_seed = true;                   // RMEM_SELF_SEED=0: no seed pass for the T = 1 self-attention (A/B)
  bool fused_seed = true;                  // RMEM_FUSED_SEED=0: separate qprep + attn_seed launches (A/B)
  t16 *d0, *d1, *d2;
  double* stats;
  uint8_t* label8;
  void* attn_ws;
  size_t attn_ws_bytes;
  void* local_ws = nullptr;
  size_t local_ws_bytes = 0;
  char* arena_base = nullptr;
  size_t state_begin = 0, state_bytes = 0;   // region zeroed per clip (banks, short-term memory)
  bool ones_ready = false;
  // AOT scratch (model 1)
  float *a_pos = nullptr, *a_tgt = nullptr, *a_qbias = nullptr;
  t16 *a_tln, *a_qkpos, *a_q, *a_k, *a_vt, *a_att, *a_o3, *a_sum, *a_n4k, *a_n4v, *a_n4vt, *a_ff1, *a_ff2, *a_decin, *a_qt;
  void* mha_ws = nullptr;
  size_t mha_ws_bytes = 0;
  // GRU_MEMORY scratch: [x | h] (then [x | reset*h]) t16, gates fp32 [HW,512], candidate fp32, token-major x / h / output
  t16 *g_comb = nullptr, *g_x = nullptr, *g_h16 = nullptr, *g_out = nullptr;
  float *g_gates = nullptr, *g_cand = nullptr;
  bool pos_ready = false;
  // optional stage timing (debug / profiling aid): CUDA events between pipeline stages, accumulated per name
  bool timing = false;
  std::vector<std::pair<std::string, cudaEvent_t>> marks;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  std::map<std::string, std::pair<double, long long>> stage_ms;
  void mark(const char* name, cudaStream_t s) {
    if (!timing) return;
    if (ev_used == ev_pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ev_pool.push_back(e);
    }
    cudaEvent_t e = ev_pool[ev_used++];
    cudaEventRecord(e, s);
    marks.emplace_back(name, e);
  }
  void flush_marks(cudaStream_t s) {
    if (!timing || marks.empty()) return;
    cudaStreamSynchronize(s);
    for (size_t i = 1; i < marks.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
      auto& acc = stage_ms[marks[i].first];
      acc.first += ms;
      acc.second += 1;
    }
    marks.clear();
    ev_used = 0;
  }

  // ---------------------------------------------------------------------------------------------
  template <typename T>
  const T* Wt(const std::string& name, size_t elems, int* rc) {
    auto it = weights.find(name);
    if (it == weights.end()) {
      set_error("missing weight '%s'", name.c_str());
      *rc = RMEM_ERR_WEIGHT;
      return nullptr;
    }
    if (it->second.second != elems * sizeof(T)) {
      set_error("weight '%s': expected %zu bytes, blob has %zu", name.c_str(), elems * sizeof(T), it->second.second);
      *rc = RMEM_ERR_WEIGHT;
      return nullptr;
    }
    return reinterpret_cast<const T*>(it->second.first);
  }

  int layout(Arena& a) {
    const Geo& G = g;
    // encoder temporaries hold kEncGroup images; every activation is [n][pixels][channels]
    constexpr size_t NG = kEncGroup;
    img8_elems = (size_t)(G.H + 6) * (G.W + 8) * 8;
    img8 = a.take<t16>(NG * img8_elems);                      // zero-padded stem inputs (zeroed once at create)
    c1 = a.take<t16>(NG * (size_t)G.P1 * 64);
    x0 = a.take<t16>(NG * (size_t)G.P4 * 256);
    x1 = a.take<t16>(NG * (size_t)G.P4 * 256);
    t1 = a.take<t16>(NG * (size_t)G.P4 * 128);
    t2 = a.take<t16>(NG * (size_t)G.P4 * 64);
    ds = a.take<t16>(NG * (size_t)G.P4 * 256);
    for (int sl = 0; sl < kFeatSets; sl += kEncGroup) {
      feat4s[sl] = a.take<t16>(NG * (size_t)G.P4 * 256);
      feat8s[sl] = a.take<t16>(NG * (size_t)G.P8 * 512);
      feat16s[sl] = a.take<t16>(NG * (size_t)G.HW * 1024);
      enc_tgts[sl] = a.take<float>(NG * (size_t)G.HW * kD);
      ad16s[sl] = a.take<t16>(NG * (size_t)G.HW * 256);
      ad8s[sl] = a.take<t16>(NG * (size_t)G.P8 * 256);
      ad4s[sl] = a.take<t16>(NG * (size_t)G.P4 * 128);
      for (int j = 1; j < kEncGroup; ++j) {
        feat4s[sl + j] = feat4s[sl] + (size_t)j * G.P4 * 256;
        feat8s[sl + j] = feat8s[sl] + (size_t)j * G.P8 * 512;
        feat16s[sl + j] = feat16s[sl] + (size_t)j * G.HW * 1024;
        enc_tgts[sl + j] = enc_tgts[sl] + (size_t)j * G.HW * kD;
        ad16s[sl + j] = ad16s[sl] + (size_t)j * G.HW * 256;
        ad8s[sl + j] = ad8s[sl] + (size_t)j * G.P8 * 256;
        ad4s[sl + j] = ad4s[sl] + (size_t)j * G.P4 * 128;
      }
    }
    use_slot(0);
    res = a.take<float>((size_t)G.HW * 2 * kD);
    t_ln = a.take<t16>((size_t)G.HW * kD);
    qt = a.take<t16>((size_t)G.HW * kDk);
    cu = a.take<t16>((size_t)G.HW * kDv);
    cu0 = a.take<t16>((size_t)G.HW * kDv);
    attn_a = a.take<t16>((size_t)G.HW * kDv);
    dwo = a.take<t16>((size_t)G.HW * kDv);
    attn_b = a.take<t16>((size_t)G.HW * kDv);     // short-term branch's own input (it runs beside the long-term branch)
    dwo2 = a.take<t16>((size_t)G.HW * 2 * kDv);   // [HW, 2048]: depthwise-conv outputs of the long | short branch = the
                                                  // A operand of the merged tail projection
    z = a.take<t16>((size_t)G.HW * 2 * kD);
    qk = a.take<t16>((size_t)G.HWp * kDk);
    vt_self = a.take<t16>((size_t)kDv * G.HWp);
    u_self = a.take<t16>((size_t)G.HW * kDv);
    gpm_out = a.take<t16>((size_t)G.HW * 2 * kD);
    idemb = a.take<t16>((size_t)G.HW * kD);
    qbias = a.take<float>((size_t)G.HW * kMaxBankFrames);
    rel = a.take<float>((size_t)G.HW * 256);
    rel_dev = a.take<float>(64);
    mseed = a.take<float>((size_t)G.HW);
    d0 = a.take<t16>((size_t)G.P4 * 128);
    d1 = a.take<t16>((size_t)G.P4 * 128);
    d2 = a.take<t16>((size_t)G.P4 * 128);
    stats = a.take<double>(kGnScratchDoubles);
    label8 = a.take<uint8_t>((size_t)G.H * G.W);
    if (cfg.model == 1) {
      a_pos = a.take<float>((size_t)G.HW * kD);
      a_tgt = a.take<float>((size_t)G.HW * kD);
      a_qbias = a.take<float>((size_t)8 * G.HW * kMaxBankFrames);
      a_tln = a.take<t16>((size_t)G.HW * kD);
      a_qkpos = a.take<t16>((size_t)G.HW * kD);
      a_q = a.take<t16>((size_t)G.HWp * kD);
      a_k = a.take<t16>((size_t)G.HWp * kD);
      a_vt = a.take<t16>((size_t)kD * G.HWp);
      a_att = a.take<t16>((size_t)G.HW * kD);
      a_o3 = a.take<t16>((size_t)G.HW * kD);
      a_sum = a.take<t16>((size_t)G.HW * kD);
      a_n4k = a.take<t16>((size_t)G.HWp * kD);
      a_n4v = a.take<t16>((size_t)G.HW * kD);
      a_n4vt = a.take<t16>((size_t)kD * G.HWp);
      a_ff1 = a.take<t16>((size_t)G.HW * 4 * kD);
      a_ff2 = a.take<t16>((size_t)G.HW * 4 * kD);
      a_decin = a.take<t16>((size_t)G.HW * 4 * kD);
      a_qt = a.take<t16>((size_t)G.HW * kD);
      mha_ws_bytes = cfg.attn_impl == RMEM_ATTN_DENSE ? mha_dense_workspace(G.HW, G.HWp, nslots, 8)
                                                      : mha_tc_workspace(G.HW, G.HWp, nslots, 8);
      mha_ws = a.take<char>(mha_ws_bytes);
      if (cfg.gru_memory) {
        g_comb = a.take<t16>((size_t)G.HW * 2 * kD);
        g_gates = a.take<float>((size_t)G.HW * 2 * kD);
        g_cand = a.take<float>((size_t)G.HW * kD);
        g_x = a.take<t16>((size_t)G.HW * kD);
        g_h16 = a.take<t16>((size_t)G.HW * kD);
        g_out = a.take<t16>((size_t)G.HW * kD);
      }
    }
    size_t ws_dense = long_attn_dense_workspace(G.HW, G.HWp, nslots);
    size_t ws_tc2 = long_attn_tc2_workspace(G.HW, G.HWp, nslots, kDv);
    size_t ws_tc3 = long_attn_tc3_workspace(G.HW, G.HWp, nslots, kDv);
    const size_t ws_tc4 = cfg.attn_impl == RMEM_ATTN_TC4 ? long_attn_tc4_workspace(G.HW, G.HWp, nslots, kDv) : 0;
    attn_ws_bytes = cfg.attn_impl == RMEM_ATTN_TC4 ? ws_tc4
                    : cfg.attn_impl == RMEM_ATTN_TC3 ? ws_tc3 : (cfg.attn_impl == RMEM_ATTN_TC2 ? ws_tc2 : ws_dense);
    attn_ws = a.take<char>(attn_ws_bytes);
    local_ws_bytes = local_attn_tc_workspace(G.h, G.w, kDv);
    local_ws = a.take<char>(local_ws_bytes);
    state_begin = a.off;
    groups.assign(cfg.max_engines, Group());
    for (auto& gr : groups) {
      for (int l = 0; l < kLayers && cfg.model == 1; ++l) {
        LayerStateA& L = gr.A[l];
        L.k_cur = a.take<t16>((size_t)G.HWp * kD);
        L.v_cur = a.take<t16>((size_t)G.HW * kD);
        L.v_lin = a.take<t16>((size_t)G.HW * kD);
        for (int p = 0; p < 2; ++p) {
          L.sk[p] = a.take<t16>((size_t)G.HWp * kD);
          L.sv[p] = a.take<t16>((size_t)G.HW * kD);
        }
        L.kbank = a.take<t16>((size_t)nslots * G.HWp * kD);
        L.vtbank = a.take<t16>((size_t)kD * nslots * G.HWp);
        for (int i = 0; i < 2; ++i) L.gru_h[i] = cfg.gru_memory ? a.take<float>((size_t)G.HW * kD) : nullptr;
      }
      for (int l = 0; l < kLayers && cfg.model == 0; ++l) {
        LayerState& L = gr.L[l];
        for (int p = 0; p < 2; ++p) {
          L.kc[p] = a.take<t16>((size_t)G.HWp * kDk);
          L.vid[p] = a.take<t16>((size_t)G.HW * kDv);
        }
        L.cat = a.take<t16>((size_t)G.HW * 2 * kD);
        L.kbank = a.take<t16>((size_t)nslots * G.HWp * kDk);
        L.vtbank = a.take<t16>((size_t)kDv * nslots * G.HWp);
      }
      gr.mass0 = a.take<float>((size_t)G.HW * kMaxBankFrames);
      gr.logits4 = a.take<float>((size_t)11 * G.P4);
    }
    state_bytes = a.off - state_begin;
    if (!a.dry && a.off > a.cap) {
      set_error("arena too small: need %zu bytes, have %zu", a.off, a.cap);
      return RMEM_ERR_ARENA;
    }
    return RMEM_OK;
  }

  // ---------------------------------------------------------------------------------------------
  // conv / linear helpers over the generic GEMM
  int conv(const t16* x, int Hin, int Win, int Cin, const std::string& wname, int Cout, int k, int stride, int pad,
           int act, const t16* resid, t16* out, cudaStream_t s, int nimg = 1, double* gn_stats = nullptr, int gn_groups = 0) {
    int rc = RMEM_OK;
    const t16* w = Wt<t16>(wname + ".w", (size_t)Cout * k * k * Cin, &rc);
    const float* b = Wt<float>(wname + ".b", Cout, &rc);
    if (rc) return rc;
    int Hout = (Hin + 2 * pad - k) / stride + 1, Wout = (Win + 2 * pad - k) / stride + 1;
    GemmParams p;
    p.A = x; p.B = w; p.ldb = (long long)k * k * Cin;
    p.M = nimg * Hout * Wout; p.N = Cout; p.K = k * k * Cin;     // images are stacked [n][pixels][channels]
    if (k == 1 && stride == 1) {
      p.lda = Cin;
    } else {
      p.conv = 1; p.Hin = Hin; p.Win = Win; p.Cin = Cin; p.Wout = Wout; p.kw = k; p.stride = stride; p.pad = pad;
      p.nimg = nimg;
    }
    p.bias = b; p.act = act;
    p.res = resid; p.ldr = Cout;
    p.C = out; p.ldc = Cout;
    p.gn_stats = gn_stats; p.gn_groups = gn_groups;
    return gemm_launch(p, s);
  }

  struct Lin {
    const t16* A; long long lda; int M; int K; int N;
    std::string w;
    int act = ACT_NONE; int act_from = 0;
    void* C = nullptr; long long ldc = 0; int c_fp32 = 0;
    void* C2 = nullptr; long long ldc2 = 0; int c2_fp32 = 0; int n_split = 1 << 30;
    int accumulate = 0;
    int n_weight_rows = -1;   // rows stored in the blob (>= N when padded)
  };
  int linear(const Lin& L, cudaStream_t s) {
    int rc = RMEM_OK;
    int rows = L.n_weight_rows > 0 ? L.n_weight_rows : L.N;
    const t16* w = Wt<t16>(L.w + ".w", (size_t)rows * L.K, &rc);
    const float* b = Wt<float>(L.w + ".b", rows, &rc);
    if (rc) return rc;
    GemmParams p;
    p.A = L.A; p.lda = L.lda; p.B = w; p.ldb = L.K;
    p.M = L.M; p.N = L.N; p.K = L.K;
    p.bias = b; p.act = L.act; p.act_from = L.act_from;
    p.C = L.C; p.ldc = L.ldc; p.c_fp32 = L.c_fp32;
    p.C2 = L.C2; p.ldc2 = L.ldc2; p.c2_fp32 = L.c2_fp32; p.n_split = L.n_split;
    p.accumulate = L.accumulate;
    return gemm_launch(p, s);
  }

  // ---------------------------------------------------------------------------------------------
  // ResNet-50 stem + layer1..3 (FrozenBN folded) + encoder_projector.   resnet.py:178-195, aot.py:116-134
  int bottleneck(const t16* x, int Hin, int Win, int Cin, const std::string& pre, int planes, int stride, bool has_ds,
                 t16* out, cudaStream_t s, int nimg = 1) {
    int Hout = (Hin - 1) / stride + 1, Wout = (Win - 1) / stride + 1;
    RMEM_TRY(conv(x, Hin, Win, Cin, pre + ".conv1", planes, 1, 1, 0, ACT_RELU, nullptr, t1, s, nimg));
    RMEM_TRY(conv(t1, Hin, Win, planes, pre + ".conv2", planes, 3, stride, 1, ACT_RELU, nullptr, t2, s, nimg));
    const t16* idt = x;
    if (has_ds) {
      RMEM_TRY(conv(x, Hin, Win, Cin, pre + ".ds", planes * 4, 1, stride, 0, ACT_NONE, nullptr, ds, s, nimg));
      idt = ds;
    }
    RMEM_TRY(conv(t2, Hout, Wout, planes, pre + ".conv3", planes * 4, 1, 1, 0, ACT_RELU, idt, out, s, nimg));
    return RMEM_OK;
  }

  // Features of `img`: taken from a prefetch when one is outstanding for exactly this pointer, else encoded inline.
  int features(const float* img, cudaStream_t s) {
    // A prefetched frame is consumed by the very next propagate, or by the one after it when the prefetch was issued
    // before the current frame's propagate (bench / evaluator order).  Anything older was never consumed (skipped frame,
    // exception in the caller): drop it, so that a later tensor the allocator places at the same address cannot pick up
    // stale features.
    // (a group prefetch of n frames is issued n frames ahead and its last member is consumed n - 1 calls later: ttl 2n)
    for (int sl = 0; sl < kFeatSets; ++sl)
      if (pending[sl] && pf_age[sl] >= pf_ttl[sl]) pending[sl] = false;
    for (int sl = 0; sl < kFeatSets; ++sl)
      if (pending[sl]) ++pf_age[sl];
    for (int sl = 0; sl < kFeatSets; ++sl)
      if (pending[sl] && pf_img[sl] == img) {
        RMEM_CUDA_CHECK(cudaStreamWaitEvent(s, ev_done[sl], 0));
        pending[sl] = false;
        use_slot(sl);
        return RMEM_OK;
      }
    // inline: the encoder's temporaries are shared with the side stream -> order after whatever is queued there; the
    // current slot's readers were all issued on this stream before, so it can be overwritten in stream order
    if (last_enc_slot >= 0) RMEM_CUDA_CHECK(cudaStreamWaitEvent(s, ev_done[last_enc_slot], 0));
    // a pending entry may be a COMING frame (prefetch issued before this call): keep it, encode into a free slot
    const int sl = pending[fslot] ? pick_slot() : fslot;
    pending[sl] = false;
    use_slot(sl);
    const float* one[1] = {img};
    RMEM_TRY(encode_into(one, 1, s, sl));
    RMEM_CUDA_CHECK(cudaEventRecord(ev_inline, s));
    inline_valid = true;
    return RMEM_OK;
  }
  int release_features(cudaStream_t s) {     // every reader of the current feature set has been issued on s
    RMEM_CUDA_CHECK(cudaEventRecord(ev_feat_free[fslot], s));
    feat_free_valid[fslot] = true;
    return RMEM_OK;
  }

  // n = 1: one frame into feature set sl.  n = 2 | 4: n frames into the aligned run sl .. sl + n - 1 -- every GEMM / conv
  // of the encoder runs ONCE over all images (M multiplied; convolutions with a fourth tensor-map dimension so that padding
  // stays per image): the launches are latency-bound at one image, so a pair costs ~1.5x one frame, not 2x.
  int encode_into(const float* const* imgs, int n, cudaStream_t s, int sl) {
    const Geo& G = g;
    mark("begin", s);
    for (int j = 0; j < n; ++j) RMEM_TRY(pack_image_padded(imgs[j], img8 + (size_t)j * img8_elems, G.H, G.W, s));
    {
      // conv1 (7x7, stride 2, pad 3; resnet.py:178-181) on the tcgen05 kernel: one k-block per window row (gemm.cuh conv = 2)
      int rc = RMEM_OK;
      const t16* w = Wt<t16>("enc.conv1.w", (size_t)64 * 7 * 8 * 8, &rc);
      const float* b = Wt<float>("enc.conv1.b", 64, &rc);
      if (rc) return rc;
      GemmParams p;
      p.A = img8; p.B = w; p.ldb = 7 * 64;
      p.M = n * G.P1; p.N = 64; p.K = 7 * 64;
      p.conv = 2; p.Hin = G.H + 6; p.Win = G.W + 8; p.Cin = 8; p.Wout = G.W1; p.kw = 7; p.stride = 2; p.pad = 3;
      p.nimg = n;
      p.bias = b; p.act = ACT_RELU;
      p.C = c1; p.ldc = 64;
      RMEM_TRY(gemm_launch(p, s));
    }
    for (int j = 0; j < n; ++j)
      RMEM_TRY(maxpool3x3s2(c1 + (size_t)j * G.P1 * 64, x0 + (size_t)j * G.P4 * 64, G.H1, G.W1, 64, G.H4, G.W4, s));
    mark("enc.stem", s);
    t16* cur = x0;
    int Hc = G.H4, Wc = G.W4, Cc = 64;
    const int planes[3] = {64, 128, 256}, nblk[3] = {3, 4, 6}, strides[3] = {1, 2, 2};
    t16* feats[3] = {feat4s[sl], feat8s[sl], feat16s[sl]};
    for (int li = 0; li < 3; ++li) {
      for (int bi = 0; bi < nblk[li]; ++bi) {
        int st = bi == 0 ? strides[li] : 1;
        bool has_ds = bi == 0;
        t16* out = (bi == nblk[li] - 1) ? feats[li] : ((cur == x0) ? x1 : x0);
        std::string pre = "enc.layer" + std::to_string(li + 1) + "." + std::to_string(bi);
        RMEM_TRY(bottleneck(cur, Hc, Wc, Cc, pre, planes[li], st, has_ds, out, s, n));
        Hc = (Hc - 1) / st + 1; Wc = (Wc - 1) / st + 1; Cc = planes[li] * 4;
        cur = out;
      }
      mark(li == 0 ? "enc.layer1" : (li == 1 ? "enc.layer2" : "enc.layer3"), s);
    }
    Lin p;
    p.A = feat16s[sl]; p.lda = 1024; p.M = n * G.HW; p.K = 1024; p.N = kD; p.w = "proj";
    p.C = enc_tgts[sl]; p.ldc = kD; p.c_fp32 = 1;
    RMEM_TRY(linear(p, s));
    mark("enc.proj", s);
    RMEM_TRY(conv(feat16s[sl], G.h, G.w, 1024, "dec.adapter_16x", 256, 1, 1, 0, ACT_NONE, nullptr, ad16s[sl], s, n));
    RMEM_TRY(conv(feat8s[sl], G.H8, G.W8, 512, "dec.adapter_8x", 256, 1, 1, 0, ACT_NONE, nullptr, ad8s[sl], s, n));
    RMEM_TRY(conv(feat4s[sl], G.H4, G.W4, 256, "dec.adapter_4x", 128, 1, 1, 0, ACT_NONE, nullptr, ad4s[sl], s, n));
    mark("enc.adapters", s);
    return RMEM_OK;
  }

  // ---------------------------------------------------------------------------------------------
  int id_embed(Group& gr, int gi, const void* label, int label_is_f32, int use_ignore, cudaStream_t s) {
    const Geo& G = g;
    int rc = RMEM_OK;
    const float* w = Wt<float>("idbank.w", (size_t)289 * 12 * kD, &rc);
    const float* b = Wt<float>("idbank.b", kD, &rc);
    const float* pf = Wt<float>("idbank.prefix", (size_t)12 * 18 * 18 * kD, &rc);
    const float* pr = Wt<float>("idbank.prefix_rows", (size_t)17 * 12 * 18 * kD, &rc);
    const float* lg = cfg.model == 0 ? Wt<float>("id_norm.g", kD, &rc) : nullptr;   // deaot.py:65-69 only
    const float* lb = cfg.model == 0 ? Wt<float>("id_norm.b", kD, &rc) : nullptr;
    if (rc) return rc;
    RMEM_TRY(separate_label(label, label_is_f32, label8, G.H, G.W, gi, n_groups, s));
    // DeAOT: layers 1 and 2 read the embedding as the second half of their linear_ID_V input (L.cat): written by the same
    // launch instead of two copies
    t16* o2 = cfg.model == 0 ? gr.L[1].cat + kD : nullptr;
    t16* o3 = cfg.model == 0 ? gr.L[2].cat + kD : nullptr;
    RMEM_TRY(idbank_embed(label8, G.H, G.W, use_ignore, w, b, lg, lb, idemb, kD, nullptr, G.h, G.w, kD, s, pf, pr, o2, o3,
                          2 * kD));
    return RMEM_OK;
  }

  // ID_V = silu(linear_ID_V(cat[curr_ID_V, id_emb]))  -> vid[cur][:, 512:]      transformer.py:1238-1244
  int fuse_id(Group& gr, int l, cudaStream_t s) {
    const Geo& G = g;
    Lin p;
    std::string pre = "gpm." + std::to_string(l);
    if (l == 0) { p.A = idemb; p.lda = kD; p.K = kD; }
    else { p.A = gr.L[l].cat; p.lda = 2 * kD; p.K = 2 * kD; }
    p.M = G.HW; p.N = 2 * kD; p.w = pre + ".linear_ID_V"; p.act = ACT_SILU;
    p.C = gr.L[l].vid[gr.parity] + 512; p.ldc = kDv;
    return linear(p, s);
  }

  int append_long(Group& gr, cudaStream_t s) {
    const Geo& G = g;
    if (gr.free_slots.empty()) { set_error("bank overflow"); return RMEM_ERR_STATE; }
    int slot = gr.free_slots.back();
    gr.free_slots.pop_back();
    for (int l = 0; l < kLayers; ++l) {
      LayerState& L = gr.L[l];
      RMEM_TRY(copy2d_t16(L.kc[gr.parity], kDk, L.kbank + (size_t)slot * G.HWp * kDk, kDk, G.HW, kDk, s));
      RMEM_TRY(transpose_t16(L.vid[gr.parity], kDv, L.vtbank + (size_t)slot * G.HWp, (long long)nslots * G.HWp,
                              G.HW, kDv, s));
    }
    gr.slots.push_back(slot);
    return RMEM_OK;
  }

  int attention(LongAttnArgs& a, cudaStream_t s, bool seed = true) {
    if (seed) { a.seed_h = g.h; a.seed_w = g.w; }   // token grid: tc3 seeds the row maximum from the query's own neighbourhood
    if (cfg.attn_impl == RMEM_ATTN_TC4) return long_attn_tc4(a, attn_ws, attn_ws_bytes, s);
    if (cfg.attn_impl == RMEM_ATTN_TC3) return long_attn_tc3(a, attn_ws, attn_ws_bytes, s);
    if (cfg.attn_impl == RMEM_ATTN_TC2) return long_attn_tc2(a, attn_ws, attn_ws_bytes, s);
    return long_attn_dense(a, attn_ws, attn_ws_bytes, s);
  }

  // gated epilogue tail: DWConv5x5 -> Linear(1024->512) accumulated into the tgt || tgt_id residual stream
  int gated_tail(const std::string& pre, cudaStream_t s) {
    RMEM_TRY(gated_tail_dw(pre, attn_a, dwo, s));
    return gated_tail_proj(pre, dwo, s);
  }
  int gated_tail_dw(const std::string& pre, const t16* in, t16* out, cudaStream_t s, int ldy = 0) {
    const Geo& G = g;
    int rc = RMEM_OK;
    const float* dw = Wt<float>(pre + ".dw", (size_t)25 * kDv, &rc);
    if (rc) return rc;
    return dwconv5x5(in, dw, out, G.h, G.w, kDv, s, 0, ldy);
  }
  // tgt || tgt_id += long_term_attn.projection(dw_long) + short_term_attn.projection(dw_short) as ONE accumulate over the
  // concatenated K = 2048 (transformer.py:1212-1220; weights.py "tail.proj")
  int gated_tail_proj2(const std::string& pre, cudaStream_t s) {
    const Geo& G = g;
    Lin p;
    p.A = dwo2; p.lda = 2 * kDv; p.M = G.HW; p.K = 2 * kDv; p.N = 2 * kD; p.w = pre + ".tail.proj";
    p.C = res; p.ldc = 2 * kD; p.c_fp32 = 1; p.accumulate = 1;
    return linear(p, s);
  }
  int gated_tail_proj(const std::string& pre, const t16* in, cudaStream_t s) {   // res += proj(in): fp32 read-modify-write
    const Geo& G = g;
    Lin p;
    p.A = in; p.lda = kDv; p.M = G.HW; p.K = kDv; p.N = 2 * kD; p.w = pre + ".proj";
    p.C = res; p.ldc = 2 * kD; p.c_fp32 = 1; p.accumulate = 1;
    return linear(p, s);
  }

  // One GatedPropagationModule.forward (transformer.py:1091-1236).  ref_mode: curr_id_emb given.
  int gpm_layer(Group& gr, int l, bool ref_mode, cudaStream_t s) {
    const Geo& G = g;
    const std::string pre = "gpm." + std::to_string(l);
    LayerState& L = gr.L[l];
    const int cur = gr.parity, prev = gr.parity ^ 1;
    int rc = RMEM_OK;
    const float scale = 1.f / sqrtf((float)kDk);

    // t = LN1(tgt); Q|V = linear_QV(t) (silu on V); U = linear_U(t); ti = id_LN1(tgt_id); ID_U = linear_ID_U(ti).
    // Two chains of equal length (profiles/r02_timeline_*.txt: the main stream used to idle ~28 us per layer at the join):
    //   this stream:   LN1 + id_LN1 (one launch) -> linear_QV -> relative-bias GEMM of the short-term branch (needs Q only)
    //   second stream: value-major copy of the short-term V (previous frame: ready since its update_memory) ->
    //                  linear_U -> linear_ID_U
    // so that after the long-term attention kernel (which owns every SM) the short-term branch is only its windowed
    // attention + depthwise conv, the same length as the long-term tail it runs beside.
    const float* n1g = Wt<float>(pre + ".norm1.g", kD, &rc);
    const float* n1b = Wt<float>(pre + ".norm1.b", kD, &rc);
    if (rc) return rc;
    if (l > 0) {
      const float* g1 = Wt<float>(pre + ".id_norm1.g", kD, &rc);
      const float* b1 = Wt<float>(pre + ".id_norm1.b", kD, &rc);
      if (rc) return rc;
      // ti -> curr_ID_V (first half of the linear_ID_V input), also A of linear_ID_U
      RMEM_TRY(layernorm_pair(res, 2 * kD, n1g, n1b, g1, b1, t_ln, kD, G.HW, kD, s, L.cat, 2 * kD));
    } else {
      // first layer: the residual stream starts here -- tgt = projector output, tgt_id = 0 (transformer.py:786-790),
      // written by the same launch instead of a copy and a memset in front of it
      RMEM_TRY(layernorm(enc_tgt, kD, n1g, n1b, t_ln, kD, nullptr, 0, G.HW, kD, s, nullptr, res, 2 * kD));
    }
    t16* gate = (l == 0) ? cu0 : cu;
    const bool par = use_aux();
    cudaStream_t s2 = par ? aux_stream : s;          // U and ID_U do not depend on QV: second stream
    const bool tc_local = cfg.attn_impl != RMEM_ATTN_DENSE;
    const bool v_early = tc_local && !ref_mode;      // reference frame: its own V is only final after fuse_id below
    if (par) RMEM_TRY(fork(s));
    if (v_early)
      RMEM_TRY(local_attn_tc_prepare_v(L.vid[gr.parity ^ 1], kDv, G.h, G.w, kDv, local_ws, local_ws_bytes, s2));
    {
      Lin p;
      p.A = t_ln; p.lda = kD; p.M = G.HW; p.K = kD; p.N = 2 * kD; p.w = pre + ".linear_U";
      p.act = ACT_SILU; p.C = gate; p.ldc = kDv;
      RMEM_TRY(linear(p, s2));
    }
    if (l > 0) {
      Lin p;
      p.A = L.cat; p.lda = 2 * kD; p.M = G.HW; p.K = kD; p.N = 2 * kD; p.w = pre + ".linear_ID_U";
      p.act = ACT_SILU; p.C = gate + 512; p.ldc = kDv;
      RMEM_TRY(linear(p, s2));
    }
    {
      Lin p;
      p.A = t_ln; p.lda = kD; p.M = G.HW; p.K = kD; p.N = kDk + 2 * kD; p.w = pre + ".linear_QV";
      p.act = ACT_SILU; p.act_from = kDk;
      p.C = L.kc[cur]; p.ldc = kDk;
      p.C2 = L.vid[cur]; p.ldc2 = kDv; p.n_split = kDk;
      RMEM_TRY(linear(p, s));
    }
    {
      // relative-position bias of the windowed attention: rel = Q . rel_emb^T
      Lin p;
      p.A = L.kc[cur]; p.lda = kDk; p.M = G.HW; p.K = kDk; p.N = 256; p.n_weight_rows = 256;   // 225 offsets, zero-padded
      // the tensor-core kernel reads one aligned 16-float line per window row (".short.rel16": rows re-ordered at pack time)
      p.w = pre + (tc_local ? ".short.rel16" : ".short.rel"); p.C = rel; p.ldc = 256; p.c_fp32 = 1;
      RMEM_TRY(linear(p, s));
    }
    if (par) RMEM_TRY(join(s));

    mark("gpm.proj_in", s);
    // memories
    LongAttnArgs a;
    a.HW = G.HW; a.HWp = G.HWp; a.Dk = kDk; a.Dv = kDv; a.scale = scale;
    a.gate = gate; a.ldg = kDv; a.out = attn_a; a.ldo = kDv;
    const t16 *sk, *sv;
    int T;
    if (ref_mode) {
      RMEM_TRY(fuse_id(gr, l, s));                    // ID_V of the reference frame itself (:1125-1135)
      // bank := this frame -> physical slot 0
      RMEM_TRY(copy2d_t16(L.kc[cur], kDk, L.kbank, kDk, G.HW, kDk, s));
      RMEM_TRY(transpose_t16(L.vid[cur], kDv, L.vtbank, (long long)nslots * G.HWp, G.HW, kDv, s));
      T = 1;
      a.slot[0] = 0;
      sk = L.kc[cur]; sv = L.vid[cur];
    } else {
      T = (int)gr.slots.size();
      for (int t = 0; t < T; ++t) a.slot[t] = gr.slots[t];
      sk = L.kc[prev]; sv = L.vid[prev];
    }
    a.T = T; a.nslots = nslots; a.kbank = L.kbank; a.vtbank = L.vtbank;
    {
      const float* pc = Wt<float>("cur_pos_emb", kDk, &rc);
      const float* pm = Wt<float>("mem_pos_emb", 4 * kDk, &rc);
      if (rc) return rc;
      int pe_slot[kMaxBankFrames];
      temporal_pe_slots(T, 4, pe_slot);
      if ((cfg.attn_impl == RMEM_ATTN_TC3 || cfg.attn_impl == RMEM_ATTN_TC4) && fused_seed) {
        // (TC4: the launch also arms the column kernel's overflow flag, the first int of the attention workspace)
        RMEM_TRY(qprep_seed_tc3(L.kc[cur], kDk, pc, pm, pe_slot, T, scale, L.kbank, a.slot, G.HW, G.HWp, G.h, G.w, qt, qbias,
                                mseed, s, cfg.attn_impl == RMEM_ATTN_TC4 ? reinterpret_cast<int*>(attn_ws) : nullptr));
        a.mseed = mseed;
      } else {
        RMEM_TRY(qprep(L.kc[cur], kDk, pc, pm, pe_slot, T, scale, qt, qbias, G.HW, kDk, s));
      }
    }
    a.qt = qt; a.qbias = qbias;
    a.mass = (l == 0 && !ref_mode) ? gr.mass0 : nullptr;
    if (a.mass) gr.mass_T = T;
    mark("gpm.long.prep", s);
    // The short-term branch (relative-bias GEMM, windowed attention, depthwise conv) does not depend on the long-term one:
    // it runs on the second stream into its own buffers; only its final projection (an fp32 accumulate into the
    // residual stream, like the long-term one) is issued after the join, in the reference's order.
    const bool par2 = par && branch_par;
    cudaStream_t sb = par2 ? aux_stream : s;
    t16* sh_attn = par2 ? attn_b : attn_a;
    auto short_branch = [&]() -> int {
      if (!tc_local)
        RMEM_TRY(local_attn(L.kc[cur], kDk, sk, kDk, sv, kDv, rel, 256, gate, kDv, sh_attn, kDv, G.h, G.w, kDv, scale, sb));
      else
        RMEM_TRY(local_attn_tc(L.kc[cur], kDk, sk, kDk, sv, kDv, rel, 256, 16, gate, kDv, sh_attn, kDv, G.h, G.w, kDv,
                               scale, local_ws, local_ws_bytes, sb, v_early));
      mark("gpm.short.attn", sb);
      return gated_tail_dw(pre + ".short", sh_attn, dwo2 + kDv, sb, 2 * kDv);
    };
    if (par2) {
      RMEM_TRY(fork(s));
      RMEM_TRY(short_branch());
    }
    RMEM_TRY(attention(a, s));
    mark("gpm.long.attn", s);
    RMEM_TRY(gated_tail_dw(pre + ".long", attn_a, dwo2, s, 2 * kDv));
    mark("gpm.long.tail", s);
    if (par2) RMEM_TRY(join(s));
    else RMEM_TRY(short_branch());
    RMEM_TRY(gated_tail_proj2(pre, s));
    mark("gpm.short.tail", s);

    // self attention on cat(LN2(tgt), id_LN2(tgt_id))
    {
      const float* g2 = Wt<float>(pre + ".norm2.g", kD, &rc);
      const float* b2 = Wt<float>(pre + ".norm2.b", kD, &rc);
      const float* gi2 = Wt<float>(pre + ".id_norm2.g", kD, &rc);
      const float* bi2 = Wt<float>(pre + ".id_norm2.b", kD, &rc);
      if (rc) return rc;
      RMEM_TRY(layernorm_pair(res, 2 * kD, g2, b2, gi2, bi2, z, 2 * kD, G.HW, kD, s));
      // Two launches instead of five (weights.py "self.QKU" / "self.V12"): [q=k | u] = z . [W_QK ; diag(W_U1, W_U2)]^T with
      // SiLU from column 128 on the second stream, and v^T = silu(diag(W_V1, W_V2) . z^T + b) computed directly
      // value-major (bias along M) on this one.
      if (par) RMEM_TRY(fork(s));
      {
        Lin p;
        p.A = z; p.lda = 2 * kD; p.M = G.HW; p.K = 2 * kD; p.N = kDk + kDv; p.w = pre + ".self.QKU";
        p.act = ACT_SILU; p.act_from = kDk;
        p.C = qk; p.ldc = kDk;
        p.C2 = u_self; p.ldc2 = kDv; p.n_split = kDk;
        RMEM_TRY(linear(p, s2));
      }
      {
        const t16* w = Wt<t16>(pre + ".self.V12.w", (size_t)kDv * 2 * kD, &rc);
        const float* b = Wt<float>(pre + ".self.V12.b", kDv, &rc);
        if (rc) return rc;
        GemmParams q;
        q.A = w; q.lda = 2 * kD; q.B = z; q.ldb = 2 * kD;
        q.M = kDv; q.N = G.HW; q.K = 2 * kD;
        q.bias = b; q.bias_m = 1; q.act = ACT_SILU; q.pad_n_ok = 1;   // pad key columns are masked by the attention
        q.C = vt_self; q.ldc = G.HWp;
        RMEM_TRY(gemm_launch(q, s));
      }
      if (par) RMEM_TRY(join(s));
      LongAttnArgs sa;
      sa.HW = G.HW; sa.HWp = G.HWp; sa.Dk = kDk; sa.Dv = kDv; sa.scale = scale;
      sa.qt = qk; sa.kbank = qk; sa.vtbank = vt_self; sa.nslots = 1; sa.T = 1; sa.slot[0] = 0;
      sa.gate = u_self; sa.ldg = kDv; sa.out = attn_a; sa.ldo = kDv;
      mark("gpm.self.proj", s);
      RMEM_TRY(attention(sa, s, self_seed));
      mark("gpm.self.attn", s);
      RMEM_TRY(gated_tail(pre + ".self", s));
      mark("gpm.self.tail", s);
    }
    return RMEM_OK;
  }

  // DualBranchGPM.forward (transformer.py:765-824) + FPN decoder (fpn.py:36-68) for one object group.
  int lstt_decode(Group& gr, bool ref_mode, cudaStream_t s) {
    const Geo& G = g;
    int rc = RMEM_OK;
    // residual stream: tgt = projector output, tgt_id = 0 -- initialised by layer 0's LayerNorm launch (gpm_layer)
    if (!ones_ready) {
      RMEM_TRY(fill_t16(cu0 + 512, kDv, G.HW, 512, 1.0f, s));
      ones_ready = true;
    }
    for (int l = 0; l < kLayers; ++l) RMEM_TRY(gpm_layer(gr, l, ref_mode, s));
    const float* og = Wt<float>("gpm.out_norm.g", 2 * kD, &rc);
    const float* ob = Wt<float>("gpm.out_norm.b", 2 * kD, &rc);
    if (rc) return rc;
    RMEM_TRY(groupnorm_f32(res, og, ob, gpm_out, G.HW, 2 * kD, 2, 0, stats, s));
    mark("gpm.out_norm", s);
    return fpn_decode(gr, gpm_out, 2 * kD, s);
  }

  // FPNSegmentationHead.forward (fpn.py:36-68): x16 [HW, cin] -> planar fp32 logits [11, P4].
  int fpn_decode(Group& gr, const t16* x16, int cin, cudaStream_t s) {
    const Geo& G = g;
    int rc = RMEM_OK;
    // The GroupNorm statistics of every decoder convolution come out of that convolution's own epilogue (fp32 results,
    // GemmParams::gn_stats): four statistics launches less per frame and object group.  RMEM_GN_FUSED=0: separate pass.
    static const bool fused = [] { const char* e = getenv("RMEM_GN_FUSED"); return !(e && e[0] == '0'); }();
    double* st = fused ? stats : nullptr;
    auto gn = [&](const std::string& n, const t16* x, t16* y, int P, int C, const t16* add) -> int {
      int r2 = RMEM_OK;
      const float* gg = Wt<float>(n + ".gn.g", C, &r2);
      const float* gb = Wt<float>(n + ".gn.b", C, &r2);
      if (r2) return r2;
      return groupnorm_t16(x, gg, gb, y, P, C, 8, 1, stats, s, add, fused);
    };
    // x = adapter(feature) + x at every scale: the adapter maps (ad16 / ad8 / ad4) were computed with the encoder
    RMEM_TRY(conv(x16, G.h, G.w, cin, "dec.conv_in", 256, 1, 1, 0, ACT_NONE, nullptr, d0, s, 1, st, 8));
    RMEM_TRY(gn("dec.conv_in", d0, d1, G.HW, 256, ad16));                          // relu(GN(conv_in)) + adapter_16x(feat16)
    // ... and where the normalised map is only upsampled, relu(GroupNorm) is applied to the taps inside the upsample kernel
    auto up_gn = [&](const std::string& n, const t16* x, t16* tmp, t16* y, int hin, int win, int hout, int wout, int C,
                     const t16* add) -> int {
      if (!fused) {
        RMEM_TRY(gn(n, x, tmp, hin * win, C, nullptr));
        return upsample_bilinear_t16(tmp, y, hin, win, hout, wout, C, s, add);
      }
      int r2 = RMEM_OK;
      const float* gg = Wt<float>(n + ".gn.g", C, &r2);
      const float* gb = Wt<float>(n + ".gn.b", C, &r2);
      if (r2) return r2;
      return upsample_bilinear_t16(x, y, hin, win, hout, wout, C, s, add, stats, gg, gb, 8);
    };
    RMEM_TRY(conv(d1, G.h, G.w, 256, "dec.conv_16x", 256, 3, 1, 1, ACT_NONE, nullptr, d2, s, 1, st, 8));
    RMEM_TRY(up_gn("dec.conv_16x", d2, d1, d0, G.h, G.w, G.H8, G.W8, 256, ad8));      // up(relu(GN(x))) + adapter_8x(feat8)
    RMEM_TRY(conv(d0, G.H8, G.W8, 256, "dec.conv_8x", 128, 3, 1, 1, ACT_NONE, nullptr, d2, s, 1, st, 8));
    RMEM_TRY(up_gn("dec.conv_8x", d2, d1, d0, G.H8, G.W8, G.H4, G.W4, 128, ad4));     // up(relu(GN(x))) + adapter_4x(feat4)
    RMEM_TRY(conv(d0, G.H4, G.W4, 128, "dec.conv_4x", 128, 3, 1, 1, ACT_NONE, nullptr, d2, s, 1, st, 8));
    const t16* wo = Wt<t16>("dec.conv_out.w", (size_t)11 * 128, &rc);
    const float* bo = Wt<float>("dec.conv_out.b", 11, &rc);
    const float* g4 = Wt<float>("dec.conv_4x.gn.g", 128, &rc);
    const float* b4 = Wt<float>("dec.conv_4x.gn.b", 128, &rc);
    if (rc) return rc;
    // GroupNorm(8) + ReLU of the 1/4-resolution map folded into the conv_out kernel's tile load
    RMEM_TRY(conv_out_gn_logits(d2, g4, b4, 8, stats, wo, bo, gr.logits4, G.P4, 128, 11, s, fused));
    mark("decoder", s);
    return RMEM_OK;
  }

  // =============================================================================================
  // AOT (model 1): SimplifiedTransformerBlock / LongShortTermTransformer  (transformer.py:199-267, 553-692)
  // =============================================================================================
  int mha(const MhaArgs& a, cudaStream_t s) {
    return cfg.attn_impl == RMEM_ATTN_DENSE ? mha_dense(a, mha_ws, mha_ws_bytes, s) : mha_tc(a, mha_ws, mha_ws_bytes, s);
  }

  // y[HW,N] (t16 or fp32 accumulate) = x[HW,K] W^T + b
  int lin_t16(const t16* x, long long ldx, int K, const std::string& w, int N, t16* y, long long ldy, cudaStream_t s) {
    Lin p;
    p.A = x; p.lda = ldx; p.M = g.HW; p.K = K; p.N = N; p.w = w; p.C = y; p.ldc = ldy;
    return linear(p, s);
  }
  int lin_acc(const t16* x, long long ldx, int K, const std::string& w, int N, float* y, long long ldy, cudaStream_t s) {
    Lin p;
    p.A = x; p.lda = ldx; p.M = g.HW; p.K = K; p.N = N; p.w = w; p.C = y; p.ldc = ldy; p.c_fp32 = 1; p.accumulate = 1;
    return linear(p, s);
  }

  int aot_layer(Group& gr, int l, bool ref_mode, cudaStream_t s) {
    const Geo& G = g;
    const std::string pre = "lstt." + std::to_string(l);
    LayerStateA& L = gr.A[l];
    const int cur = gr.parity, prev = gr.parity ^ 1;
    const float scale = 1.f / sqrtf(32.f);
    int rc = RMEM_OK;
    auto normw = [&](const std::string& n, const float** gg, const float** bb) {
      *gg = Wt<float>(pre + "." + n + ".g", kD, &rc);
      *bb = Wt<float>(pre + "." + n + ".b", kD, &rc);
    };
    const float *g1, *b1, *g2, *b2, *g3, *b3, *g4, *b4;
    normw("norm1", &g1, &b1); normw("norm2", &g2, &b2); normw("norm3", &g3, &b3); normw("norm4", &g4, &b4);
    if (rc) return rc;

    // ---- self-attention with sine PE on q, k (:566-571) ----
    RMEM_TRY(layernorm(a_tgt, kD, g1, b1, a_tln, kD, a_qkpos, kD, G.HW, kD, s, a_pos));
    RMEM_TRY(lin_t16(a_qkpos, kD, kD, pre + ".self.linear_Q", kD, a_q, kD, s));
    RMEM_TRY(lin_t16(a_qkpos, kD, kD, pre + ".self.linear_K", kD, a_k, kD, s));
    {
      // v^T = W_V . t^T + b, value-major for the P.V contraction (bias along M)
      const t16* w = Wt<t16>(pre + ".self.linear_V.w", (size_t)kD * kD, &rc);
      const float* b = Wt<float>(pre + ".self.linear_V.b", kD, &rc);
      if (rc) return rc;
      GemmParams q;
      q.A = w; q.lda = kD; q.B = a_tln; q.ldb = kD;
      q.M = kD; q.N = G.HW; q.K = kD;
      q.bias = b; q.bias_m = 1; q.pad_n_ok = 1;
      q.C = a_vt; q.ldc = G.HWp;
      RMEM_TRY(gemm_launch(q, s));
    }
    {
      MhaArgs a;
      a.q = a_q; a.ldq = kD; a.kbank = a_k; a.vtbank = a_vt; a.nslots = 1; a.T = 1; a.slot[0] = 0;
      a.HW = G.HW; a.HWp = G.HWp; a.scale = scale; a.out = a_att; a.ldo = kD;
      RMEM_TRY(mha(a, s));
    }
    RMEM_TRY(lin_acc(a_att, kD, kD, pre + ".self.proj", kD, a_tgt, kD, s));

    // ---- long / short-term attention (:574-680) ----
    RMEM_TRY(layernorm(a_tgt, kD, g2, b2, L.v_cur, kD, nullptr, 0, G.HW, kD, s));       // V = LN2(tgt)
    RMEM_TRY(lin_t16(L.v_cur, kD, kD, pre + ".linear_Q", kD, L.k_cur, kD, s));            // Q = K
    const t16 *sk, *sv;
    int T;
    MhaArgs la;
    if (ref_mode) {
      RMEM_TRY(add_t16(L.v_cur, kD, idemb, kD, a_sum, kD, G.HW, kD, s));
      RMEM_TRY(lin_t16(a_sum, kD, kD, pre + ".linear_V", kD, L.v_lin, kD, s));            // global_V (:584)
      RMEM_TRY(copy2d_t16(L.k_cur, kD, L.kbank, kD, G.HW, kD, s));                        // bank := this frame, slot 0
      RMEM_TRY(transpose_t16(L.v_lin, kD, L.vtbank, (long long)nslots * G.HWp, G.HW, kD, s));
      T = 1;
      la.slot[0] = 0;
      sk = L.k_cur; sv = L.v_lin;
    } else {
      T = (int)gr.slots.size();
      for (int t = 0; t < T; ++t) la.slot[t] = gr.slots[t];
      sk = L.sk[prev]; sv = L.sv[prev];
    }
    {
      const float* pc = Wt<float>("cur_pos_emb", kD, &rc);
      const float* pm = Wt<float>("mem_pos_emb", 4 * kD, &rc);
      if (rc) return rc;
      int pe_slot[kMaxBankFrames];
      temporal_pe_slots(T, 4, pe_slot);
      RMEM_TRY(qprep_heads(L.k_cur, kD, pc, pm, pe_slot, T, scale, a_qt, a_qbias, G.HW, 8, s));
    }
    la.q = a_qt; la.ldq = kD; la.kbank = L.kbank; la.vtbank = L.vtbank; la.nslots = nslots; la.T = T;
    la.HW = G.HW; la.HWp = G.HWp; la.scale = scale; la.qbias = a_qbias; la.out = a_att; la.ldo = kD;
    la.mass = (l == 0 && !ref_mode) ? gr.mass0 : nullptr;
    if (la.mass) gr.mass_T = T;
    RMEM_TRY(mha(la, s));
    RMEM_TRY(lin_acc(a_att, kD, kD, pre + ".long.proj", kD, a_tgt, kD, s));               // tgt += tgt2

    // short-term: MHA(Q, LN4(local_K + K), LN4(local_V + V))  (:656-662) -- dense over the previous frame
    RMEM_TRY(add_layernorm_t16(sk, kD, L.k_cur, kD, g4, b4, a_n4k, kD, G.HW, kD, s));
    RMEM_TRY(add_layernorm_t16(sv, kD, L.v_cur, kD, g4, b4, a_n4v, kD, G.HW, kD, s));
    RMEM_TRY(transpose_t16(a_n4v, kD, a_n4vt, G.HWp, G.HW, kD, s));
    {
      MhaArgs a;
      a.q = L.k_cur; a.ldq = kD; a.kbank = a_n4k; a.vtbank = a_n4vt; a.nslots = 1; a.T = 1; a.slot[0] = 0;
      a.HW = G.HW; a.HWp = G.HWp; a.scale = scale; a.out = a_att; a.ldo = kD;
      RMEM_TRY(mha(a, s));
    }
    RMEM_TRY(lin_t16(a_att, kD, kD, pre + ".short.proj", kD, a_o3, kD, s));                // tgt3
    RMEM_TRY(accum_t16_into_f32(a_o3, kD, a_tgt, kD, G.HW, kD, s));                        // tgt += tgt3
    // next frame's short-term memory: [linear_QMem(tgt3), tgt3] (:675-678)
    RMEM_TRY(lin_t16(a_o3, kD, kD, pre + ".linear_QMem", kD, L.sk[cur], kD, s));
    if (ref_mode) {
      RMEM_TRY(add_t16(a_o3, kD, idemb, kD, a_sum, kD, G.HW, kD, s));
      RMEM_TRY(lin_t16(a_sum, kD, kD, pre + ".linear_VMem", kD, L.sv[cur], kD, s));
    } else {
      RMEM_TRY(copy2d_t16(a_o3, kD, L.sv[cur], kD, G.HW, kD, s));
    }

    // ---- feed-forward: LN3 -> linear1 -> GroupNorm(32) -> GELU -> DWConv5x5 -> linear2 (:683-687) ----
    RMEM_TRY(layernorm(a_tgt, kD, g3, b3, a_tln, kD, nullptr, 0, G.HW, kD, s));
    RMEM_TRY(lin_t16(a_tln, kD, kD, pre + ".linear1", 4 * kD, a_ff1, 4 * kD, s));
    {
      const float* gg = Wt<float>(pre + ".act.gn.g", 4 * kD, &rc);
      const float* gb = Wt<float>(pre + ".act.gn.b", 4 * kD, &rc);
      const float* dw = Wt<float>(pre + ".act.dw", (size_t)25 * 4 * kD, &rc);
      if (rc) return rc;
      RMEM_TRY(groupnorm_t16(a_ff1, gg, gb, a_ff2, G.HW, 4 * kD, 32, /*act=gelu*/ 2, stats, s));
      RMEM_TRY(dwconv5x5(a_ff2, dw, a_ff1, G.h, G.w, 4 * kD, s));
    }
    RMEM_TRY(lin_acc(a_ff1, 4 * kD, 4 * kD, pre + ".linear2", kD, a_tgt, kD, s));

    // decoder_norms[l] (:248-259) -> decoder input columns 256(l+1)..
    const float* dg = Wt<float>("lstt.dec_norm." + std::to_string(l) + ".g", kD, &rc);
    const float* db = Wt<float>("lstt.dec_norm." + std::to_string(l) + ".b", kD, &rc);
    if (rc) return rc;
    return layernorm(a_tgt, kD, dg, db, a_decin + (size_t)(l + 1) * kD, 4 * kD, nullptr, 0, G.HW, kD, s);
  }

  // LongShortTermTransformer.forward + AOT.decode_id_logits (aot.py:136-142): decoder input = cat(proj16x, l0, l1, l2)
  int lstt_decode_aot(Group& gr, bool ref_mode, cudaStream_t s) {
    const Geo& G = g;
    if (!pos_ready) {
      RMEM_TRY(sine_pos_emb(a_pos, G.h, G.w, kD, s));              // aot_engine.py:289-292, once per engine size
      pos_ready = true;
    }
    RMEM_CUDA_CHECK(cudaMemcpyAsync(a_tgt, enc_tgt, sizeof(float) * G.HW * kD, cudaMemcpyDeviceToDevice, s));
    RMEM_TRY(cvt_f32_t16(enc_tgt, kD, a_decin, 4 * kD, G.HW, kD, s));
    for (int l = 0; l < kLayers; ++l) RMEM_TRY(aot_layer(gr, l, ref_mode, s));
    return fpn_decode(gr, a_decin, 4 * kD, s);
  }

  // LongShortTermTransformer.update_short_memories (transformer.py:269-304)
  int aot_refresh(Group& gr, cudaStream_t s) {
    const Geo& G = g;
    for (int l = 0; l < kLayers; ++l) {
      const std::string pre = "lstt." + std::to_string(l);
      LayerStateA& L = gr.A[l];
      RMEM_TRY(add_t16(L.v_cur, kD, idemb, kD, a_sum, kD, G.HW, kD, s));
      RMEM_TRY(lin_t16(a_sum, kD, kD, pre + ".linear_V", kD, L.v_lin, kD, s));
      RMEM_TRY(add_t16(L.sv[gr.parity], kD, idemb, kD, a_sum, kD, G.HW, kD, s));
      RMEM_TRY(lin_t16(a_sum, kD, kD, pre + ".linear_VMem", kD, L.sv[gr.parity], kD, s));
    }
    return RMEM_OK;
  }

  int append_long_aot(Group& gr, cudaStream_t s) {
    const Geo& G = g;
    if (gr.free_slots.empty()) { set_error("bank overflow"); return RMEM_ERR_STATE; }
    int slot = gr.free_slots.back();
    gr.free_slots.pop_back();
    for (int l = 0; l < kLayers; ++l) {
      LayerStateA& L = gr.A[l];
      RMEM_TRY(copy2d_t16(L.k_cur, kD, L.kbank + (size_t)slot * G.HWp * kD, kD, G.HW, kD, s));
      RMEM_TRY(transpose_t16(L.v_lin, kD, L.vtbank + (size_t)slot * G.HWp, (long long)nslots * G.HWp, G.HW, kD, s));
    }
    gr.slots.push_back(slot);
    return RMEM_OK;
  }

  // GRU_MEMORY (transformer.py:420-430, AOT only): the frame about to be dropped goes through the layer's ConvGRU (K: 2x2,
  // V: 1x1, padding "same" = one zero row / column at the bottom / right for the 2x2 kernel, which is what the implicit
  // GEMM's out-of-bounds zero fill gives with pad = 0); the cell's output replaces bank position 1.
  int conv_same(const t16* x, int Cin, const std::string& wname, int Cout, int k, float* out, cudaStream_t s) {
    const Geo& G = g;
    int rc = RMEM_OK;
    const t16* w = Wt<t16>(wname + ".w", (size_t)Cout * k * k * Cin, &rc);
    const float* b = Wt<float>(wname + ".b", Cout, &rc);
    if (rc) return rc;
    GemmParams p;
    p.A = x; p.B = w; p.ldb = (long long)k * k * Cin;
    p.M = G.HW; p.N = Cout; p.K = k * k * Cin;
    if (k == 1) {
      p.lda = Cin;
    } else {
      p.conv = 1; p.Hin = G.h; p.Win = G.w; p.Cin = Cin; p.Wout = G.w; p.kw = k; p.stride = 1; p.pad = 0;
    }
    p.bias = b;
    p.C = out; p.ldc = Cout; p.c_fp32 = 1;
    return gemm_launch(p, s);
  }
  int gru_condense(Group& gr, int drop, cudaStream_t s) {
    const Geo& G = g;
    RMEM_REQUIRE(drop >= 2 && drop < (int)gr.slots.size(), "GRU_MEMORY: drop index %d", drop);
    const int ps_drop = gr.slots[drop], ps_one = gr.slots[1];
    const long long ldv = (long long)nslots * G.HWp;
    for (int l = 0; l < kLayers; ++l) {
      LayerStateA& L = gr.A[l];
      for (int i = 0; i < 2; ++i) {
        const std::string pre = "lstt." + std::to_string(l) + ".gru." + std::to_string(i);
        const int k = i == 0 ? 2 : 1;
        const t16* x = L.kbank + (size_t)ps_drop * G.HWp * kD;
        if (i == 1) {                                   // the V bank is value-major: back to token-major
          RMEM_TRY(transpose_t16(L.vtbank + (size_t)ps_drop * G.HWp, ldv, g_x, kD, kD, G.HW, s));
          x = g_x;
        }
        float* h = L.gru_h[i];
        RMEM_TRY(copy2d_t16(x, kD, g_comb, 2 * kD, G.HW, kD, s));
        RMEM_TRY(cvt_f32_t16(h, kD, g_comb + kD, 2 * kD, G.HW, kD, s));
        RMEM_TRY(conv_same(g_comb, 2 * kD, pre + ".gates", 2 * kD, k, g_gates, s));
        RMEM_TRY(gru_reset(g_gates, 2 * kD, h, g_comb + kD, 2 * kD, G.HW, kD, s));
        RMEM_TRY(conv_same(g_comb, 2 * kD, pre + ".can", kD, k, g_cand, s));
        RMEM_TRY(gru_blend(g_gates + kD, 2 * kD, g_cand, h, g_h16, G.HW, kD, s));
        Lin p;
        p.A = g_h16; p.lda = kD; p.M = G.HW; p.K = kD; p.N = kD; p.w = pre + ".out";
        if (i == 0) {
          p.C = L.kbank + (size_t)ps_one * G.HWp * kD; p.ldc = kD;
          RMEM_TRY(linear(p, s));
        } else {
          p.C = g_out; p.ldc = kD;
          RMEM_TRY(linear(p, s));
          RMEM_TRY(transpose_t16(g_out, kD, L.vtbank + (size_t)ps_one * G.HWp, ldv, G.HW, kD, s));
        }
      }
    }
    return RMEM_OK;
  }

  int frame_forward(Group& gr, bool ref_mode, cudaStream_t s) {
    return cfg.model == 1 ? lstt_decode_aot(gr, ref_mode, s) : lstt_decode(gr, ref_mode, s);
  }
};

// =================================================================================================
namespace rmem {

// transformer.py:907-964 -- EMA + UCB bonus + argmin, fp32 host arithmetic on T_old values.
int evict_pick_host(const float* rel_raw, int T_old, const int* idx, int former, std::map<int, float>& ema,
                    std::map<int, int>& times, int* drop, float* rel_norm_out, bool gru) {
  float sum = 0.f;
  for (int t = 0; t < T_old; ++t) sum += rel_raw[t];
  std::vector<float> rel(T_old);
  for (int t = 0; t < T_old; ++t) {
    rel[t] = rel_raw[t] / sum;
    if (rel_norm_out) rel_norm_out[t] = rel[t];
  }
  std::map<int, float> new_ema;
  for (int t = 0; t < T_old; ++t) {
    auto it = ema.find(idx[t]);
    float v = rel[t];
    if (it != ema.end()) {
      // (1 - 0.8) and 0.8 as fp32 scalars, two roundings then the add -- what torch does for `py_float * tensor`
      float x = 0.2f * it->second;
      float y = 0.8f * rel[t];
      v = x + y;
    }
    new_ema[idx[t]] = v;
  }
  ema.swap(new_ema);
  std::map<int, int> new_times;
  for (int t = 0; t <= T_old; ++t) {
    auto it = times.find(idx[t]);
    new_times[idx[t]] = 1 + (it != times.end() ? it->second : 0);
  }
  times.swap(new_times);
  std::vector<float> tt(T_old);
  float tsum = 0.f;
  for (int t = 0; t < T_old; ++t) tt[t] = (float)times[idx[t]];
  tt[0] = (float)T_old;
  if (gru && T_old > 1) tt[1] = (float)T_old;   // GRU_MEMORY: position 1 holds the condensed memory (transformer.py:395-396)
  for (int t = 0; t < T_old; ++t) tsum += tt[t];
  const float lg = logf(tsum);
  const int skip = gru ? 2 : 1;                  // positions never dropped (:406-411)
  int best = former + (gru ? 1 : 0);
  if (T_old > skip) {
    float bs = INFINITY;
    for (int t = skip; t < T_old; ++t) {
      float q = lg / (tt[t] + 8.0f);
      float bonus = 1.5f * sqrtf(q);
      float score = ema[idx[t]] + bonus;
      if (score < bs) { bs = score; best = t; }
    }
  }
  *drop = best;
  return RMEM_OK;
}

}  // namespace rmem

// =================================================================================================
extern "C" {

int rmem_engine_arena_bytes(const rmem_engine_config* cfg, size_t* bytes) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(cfg && bytes, "null argument");
  RMEM_REQUIRE(cfg->model == 0 || cfg->model == 1, "model %d: 0 = r50_deaotl, 1 = r50_aotl", cfg->model);
  RMEM_REQUIRE(cfg->gru_memory == 0 || cfg->model == 1,
               "GRU_MEMORY exists for the AOT model only (DualBranchGPM hard-codes gru_memory = False, transformer.py:728)");
  RMEM_REQUIRE(cfg->gru_memory == 0 || (cfg->former_mem_len == 1 && cfg->latter_mem_len >= 2),
               "GRU_MEMORY keeps bank positions 0 and 1: former_mem_len must be 1 and latter_mem_len >= 2");
  RMEM_REQUIRE(cfg->attn_impl == RMEM_ATTN_DENSE || cfg->attn_impl == RMEM_ATTN_TC2 || cfg->attn_impl == RMEM_ATTN_TC3 ||
                   cfg->attn_impl == RMEM_ATTN_TC4,
               "attn_impl %d: 0 = dense, 2 = tc2, 3 = tc3, 4 = tc4", cfg->attn_impl);
  RMEM_REQUIRE(cfg->H > 16 && cfg->W > 16 && (cfg->H - 1) % 16 == 0 && (cfg->W - 1) % 16 == 0,
               "input size %dx%d is not 16k+1 (snap with MultiRestrictSize first)", cfg->H, cfg->W);
  RMEM_REQUIRE(cfg->max_engines >= 1 && cfg->max_engines <= 4, "max_engines=%d out of 1..4", cfg->max_engines);
  RMEM_REQUIRE(cfg->former_mem_len >= 1 && cfg->former_mem_len + cfg->latter_mem_len + 1 <= kMaxBankFrames,
               "bank capacity %d+%d(+1) exceeds %d", cfg->former_mem_len, cfg->latter_mem_len, kMaxBankFrames);
  rmem_engine tmp;
  tmp.cfg = *cfg;
  tmp.g = make_geo(cfg->H, cfg->W);
  tmp.nslots = cfg->former_mem_len + cfg->latter_mem_len + 1;
  Arena a;
  a.dry = true;
  tmp.layout(a);
  *bytes = a.off + 256;
  return RMEM_OK;
  RMEM_API_END
}

int rmem_engine_create(const rmem_engine_config* cfg, const void* weight_blob, const rmem_weight_entry* entries,
                       int n_entries, void* arena, size_t arena_bytes, rmem_engine** out) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(cfg && weight_blob && entries && arena && out, "null argument");
  size_t need = 0;
  RMEM_TRY(rmem_engine_arena_bytes(cfg, &need));
  RMEM_REQUIRE((reinterpret_cast<uintptr_t>(arena) & 255) == 0, "arena must be 256-byte aligned");
  RMEM_REQUIRE((reinterpret_cast<uintptr_t>(weight_blob) & 255) == 0, "weight blob must be 256-byte aligned");
  std::unique_ptr<rmem_engine> holder(new rmem_engine());   // freed on every early return / exception below
  rmem_engine* e = holder.get();
  e->cfg = *cfg;
  e->g = make_geo(cfg->H, cfg->W);
  e->nslots = cfg->former_mem_len + cfg->latter_mem_len + 1;
  for (int i = 0; i < n_entries; ++i) {
    if ((entries[i].offset & 15) != 0) {
      set_error("weight '%s' is not 16-byte aligned in the blob", entries[i].name);
      return RMEM_ERR_WEIGHT;
    }
    e->weights[entries[i].name] = {reinterpret_cast<const char*>(weight_blob) + entries[i].offset, entries[i].nbytes};
  }
  Arena a;
  a.base = reinterpret_cast<char*>(arena);
  a.cap = arena_bytes;
  int rc = e->layout(a);
  if (rc) return rc;
  e->arena_base = a.base;
  // one-time clear (create is off the hot path): pad rows/columns of K / value-major buffers must be finite
  if (cudaMemset(arena, 0, a.off) != cudaSuccess) {
    set_error("arena clear failed: %s", cudaGetErrorString(cudaGetLastError()));
    return RMEM_ERR_CUDA;
  }
  rc = e->init_streams();
  if (rc) return rc;
  e->launches0 = launch_counter();
  *out = holder.release();
  return RMEM_OK;
  RMEM_API_END
}

void rmem_engine_destroy(rmem_engine* e) { delete e; }

int rmem_engine_restart(rmem_engine* e) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(e, "null engine");
  e->n_groups = 0;
  for (int sl = 0; sl < kFeatSets; ++sl) e->pending[sl] = false;
  for (auto& gr : e->groups) {
    gr.evict_pending = false;
    gr.slots.clear(); gr.free_slots.clear(); gr.long_idx.clear(); gr.ema.clear(); gr.times.clear();
    gr.frame_step = 0; gr.last_mem_step = -1; gr.has_ref = false; gr.parity = 0; gr.mass_T = 0;
    gr.last_rel.clear(); gr.last_drop = -1;
  }
  return RMEM_OK;
  RMEM_API_END
}

int rmem_engine_set_gap(rmem_engine* e, int gap) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(e && gap >= 1, "bad gap");
  e->cfg.long_term_mem_gap = gap;
  return RMEM_OK;
  RMEM_API_END
}

int rmem_engine_add_reference_frame(rmem_engine* e, const float* img, const void* label, int label_is_f32,
                                    int n_objects, int frame_step, void* stream) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(e && img && label, "null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int n = (n_objects + kMaxObj - 1) / kMaxObj;
  if (n < 1) n = 1;
  RMEM_REQUIRE(n <= e->cfg.max_engines, "%d objects need %d object groups, engine built for %d", n_objects, n,
               e->cfg.max_engines);
  if (n > e->n_groups) e->n_groups = n;
  RMEM_TRY(e->resolve_all());
  // zero the bank/state region once per clip (pad columns of the value-major bank must stay finite)
  RMEM_CUDA_CHECK(cudaMemsetAsync(e->arena_base + e->state_begin, 0, e->state_bytes, s));
  RMEM_TRY(e->features(img, s));
  for (int gi = 0; gi < e->n_groups; ++gi) {
    Group& gr = e->groups[gi];
    RMEM_TRY(e->id_embed(gr, gi, label, label_is_f32, /*use_ignore=*/0, s));
    gr.slots.clear(); gr.free_slots.clear();
    for (int sl = e->nslots - 1; sl >= 1; --sl) gr.free_slots.push_back(sl);
    gr.slots.push_back(0);
    gr.ema.clear(); gr.times.clear();
    RMEM_TRY(e->frame_forward(gr, /*ref_mode=*/true, s));
    gr.parity ^= 1;                      // this frame becomes the short-term memory
    gr.last_mem_step = frame_step < 0 ? gr.frame_step : frame_step;   // aot_engine.py:254-255: -1 = the engine's own step
    gr.long_idx.push_back(gr.frame_step);
    gr.has_ref = true;
  }
  return e->release_features(s);
  RMEM_API_END
}

int rmem_engine_propagate(rmem_engine* e, const float* img, int Ho, int Wo, float* out_logits, uint8_t* out_label,
                          void* stream) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(e && img, "null argument");
  RMEM_REQUIRE(e->n_groups >= 1 && e->groups[0].has_ref, "propagate before add_reference_frame");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  RMEM_TRY(e->features(img, s));
  RMEM_TRY(e->resolve_all());                // a deferred eviction of the previous update_memory edits the slot tables now
  const float* lg[4] = {nullptr, nullptr, nullptr, nullptr};
  for (int gi = 0; gi < e->n_groups; ++gi) {
    Group& gr = e->groups[gi];
    gr.frame_step += 1;
    RMEM_TRY(e->frame_forward(gr, /*ref_mode=*/false, s));
    lg[gi] = gr.logits4;
  }
  if (out_logits || out_label)
    RMEM_TRY(mask_head(lg, e->n_groups, e->g.H4, e->g.W4, Ho, Wo, out_logits, out_label, s));
  e->mark("mask_head", s);
  e->flush_marks(s);
  return e->release_features(s);            // the feature set of this frame may be overwritten from here on
  RMEM_API_END
}

// Encode the NEXT frame on the engine's side stream while the current frame is still being propagated (the encoder has
// no dependency on the memory bank; aot.py:116-134 is a pure function of the image).  `img` must be ready on `stream`
// and stay untouched until the rmem_engine_propagate call that consumes it has been issued; that call must pass the
// same pointer (anything else falls back to the inline encoder).  Results are bit-identical to the unprefetched path.
static int prefetch_n(rmem_engine* e, const float* const* imgs, int n, void* stream) {
  if (e->timing) return RMEM_OK;            // stage timing serialises the frame; keep it on one stream
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // a slot (pair of slots) that holds no unconsumed prefetch; the current slot's readers have all been issued
  const int sl = n > 1 ? e->pick_run(n) : e->pick_slot();
  RMEM_CUDA_CHECK(cudaEventRecord(e->ev_img_ready, s));
  RMEM_CUDA_CHECK(cudaStreamWaitEvent(e->enc_stream, e->ev_img_ready, 0));
  for (int j = 0; j < n; ++j)
    if (e->feat_free_valid[sl + j]) RMEM_CUDA_CHECK(cudaStreamWaitEvent(e->enc_stream, e->ev_feat_free[sl + j], 0));
  if (e->inline_valid) RMEM_CUDA_CHECK(cudaStreamWaitEvent(e->enc_stream, e->ev_inline, 0));
  const bool pdl_was = pdl_enabled();
  {
    static const bool side_pdl = [] { const char* e = getenv("RMEM_SIDE_PDL"); return e && e[0] == '1'; }();
    pdl_enabled() = side_pdl && pdl_was;    // see common.cuh: no early-launched grids on the side stream
  }
  const int rc = e->encode_into(imgs, n, e->enc_stream, sl);
  pdl_enabled() = pdl_was;
  if (rc) return rc;
  for (int j = 0; j < n; ++j) {
    RMEM_CUDA_CHECK(cudaEventRecord(e->ev_done[sl + j], e->enc_stream));
    e->pending[sl + j] = true;
    e->pf_img[sl + j] = imgs[j];
    e->pf_seq[sl + j] = ++e->pf_counter;
    e->pf_age[sl + j] = 0;
    e->pf_ttl[sl + j] = 2 * n;
  }
  e->last_enc_slot = sl + n - 1;
  return RMEM_OK;
}

int rmem_engine_prefetch(rmem_engine* e, const float* img, void* stream) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(e && img, "null argument");
  const float* one[1] = {img};
  return prefetch_n(e, one, 1, stream);
  RMEM_API_END
}

// Two coming frames in ONE encoder pass (see encode_into).  Meant to be issued two frames ahead -- before the propagate of
// frame i for frames i+2 and i+3, every second frame -- so that the pass has two frame periods to complete; each image
// is consumed by the rmem_engine_propagate call that passes the same pointer, exactly like a single prefetch.  The pair's
// results agree with the single-frame encoder to fp16 rounding of a differently tiled GEMM, not bit for bit.
int rmem_engine_prefetch2(rmem_engine* e, const float* img_a, const float* img_b, void* stream) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(e && img_a && img_b && img_a != img_b, "prefetch2: two distinct images");
  const float* two[2] = {img_a, img_b};
  return prefetch_n(e, two, 2, stream);
  RMEM_API_END
}

// n = 1, 2 or 4 coming frames in one encoder pass; a group of n is meant to be issued n frames ahead (before the propagate
// of frame i for frames i+n .. i+2n-1, every n-th frame).
int rmem_engine_prefetch_n(rmem_engine* e, const float* const* imgs, int n, void* stream) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(e && imgs && (n == 1 || n == 2 || n == 4), "prefetch_n: n must be 1, 2 or 4 (got %d)", n);
  for (int j = 0; j < n; ++j) {
    RMEM_REQUIRE(imgs[j], "prefetch_n: null image");
    for (int k = 0; k < j; ++k) RMEM_REQUIRE(imgs[j] != imgs[k], "prefetch_n: the images must be distinct");
  }
  return prefetch_n(e, imgs, n, stream);
  RMEM_API_END
}

int rmem_engine_update_memory(rmem_engine* e, const void* label, int label_is_f32, void* stream) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(e && label, "null argument");
  RMEM_REQUIRE(e->n_groups >= 1 && e->groups[0].has_ref, "update_memory before add_reference_frame");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const Geo& G = e->g;
  const int cap = e->cfg.former_mem_len + e->cfg.latter_mem_len;
  e->mark("begin", s);
  RMEM_TRY(e->resolve_all());
  for (int gi = 0; gi < e->n_groups; ++gi) {
    Group& gr = e->groups[gi];
    RMEM_TRY(e->id_embed(gr, gi, label, label_is_f32, /*use_ignore=*/1, s));
    e->mark("upd.id_embed", s);
    bool is_long = !e->cfg.no_long_memory && (gr.frame_step - gr.last_mem_step) >= e->cfg.long_term_mem_gap;
    if (is_long) gr.last_mem_step = gr.frame_step;
    if (e->cfg.model == 1) {
      RMEM_TRY(e->aot_refresh(gr, s));                                     // transformer.py:269-304
    } else {
      const bool par = e->use_aux();                                       // three independent GEMMs: alternate streams
      if (par) RMEM_TRY(e->fork(s));
      for (int l = 0; l < kLayers; ++l) RMEM_TRY(e->fuse_id(gr, l, (par && l == 1) ? e->aux_stream : s));
      if (par) RMEM_TRY(e->join(s));                                       // transformer.py:826-857
    }
    if (is_long) {
      RMEM_TRY(e->cfg.model == 1 ? e->append_long_aot(gr, s) : e->append_long(gr, s));
      gr.long_idx.push_back(gr.frame_step);
      if (e->cfg.model == 1 && (int)gr.slots.size() <= cap) {              // AOT only: early return before any
        gr.parity ^= 1;                                                    // statistics update (transformer.py:332-334)
        continue;
      }
      // aot_engine.py:350-369 -> transformer.py:880-991
      const int T_old = gr.mass_T;
      RMEM_REQUIRE(T_old + 1 == (int)gr.slots.size(), "attention mass is stale (T_old=%d, bank=%zu)", T_old,
                   gr.slots.size());
      RMEM_TRY(evict_relevance(gr.mass0, T_old, gr.logits4, G.H4, G.W4, G.h, G.w, e->rel_dev, s));
      // no host sync here: the T floats travel to pinned memory behind an event and the EMA / UCB pick + table edit run
      // at the next call that reads the tables (resolve_evict) -- by then the copy has long completed
      RMEM_CUDA_CHECK(cudaMemcpyAsync(e->rel_pinned + (size_t)gi * kMaxBankFrames, e->rel_dev, sizeof(float) * T_old,
                                      cudaMemcpyDeviceToHost, s));
      RMEM_CUDA_CHECK(cudaEventRecord(e->ev_rel[gi], s));
      gr.evict_pending = true;
      gr.evict_T_old = T_old;
      if (e->cfg.gru_memory) RMEM_TRY(e->resolve_evict(gi, s));   // the condensation kernels need this stream: no deferral
    }
    gr.parity ^= 1;   // current frame -> short-term memory
  }
  e->mark("upd.refresh+bank", s);
  e->flush_marks(s);
  return RMEM_OK;
  RMEM_API_END
}

int rmem_engine_num_groups(const rmem_engine* e) { return e ? e->n_groups : 0; }

int rmem_engine_long_indexes(const rmem_engine* e, int group, int* idx, int* n) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(e && idx && n && group >= 0 && group < e->n_groups, "bad argument");
  RMEM_TRY(const_cast<rmem_engine*>(e)->resolve_evict(group));
  const Group& gr = e->groups[group];
  // A mid-clip add_reference_frame re-initialises the bank but (like aot_engine.py:321-323) keeps appending to this list,
  // so it can outgrow the bank; the caller's buffer holds kMaxBankFrames + 1 entries: report the NEWEST ones.
  const int total = (int)gr.long_idx.size();
  const int cnt = total < kMaxBankFrames + 1 ? total : kMaxBankFrames + 1;
  for (int i = 0; i < cnt; ++i) idx[i] = gr.long_idx[total - cnt + i];
  *n = cnt;
  return RMEM_OK;
  RMEM_API_END
}

int rmem_engine_pred_logits(const rmem_engine* e, int group, const float** logits4, int* h4, int* w4) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(e && logits4 && group >= 0 && group < e->n_groups, "bad argument");
  *logits4 = e->groups[group].logits4;
  if (h4) *h4 = e->g.H4;
  if (w4) *w4 = e->g.W4;
  return RMEM_OK;
  RMEM_API_END
}

int rmem_engine_layer_memory(const rmem_engine* e, int group, int layer, const void** kbank, const void** vtbank,
                             const void** q_last, const void** vid_last, int* nslots, int* HWp, int* T, int* slots) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(e && group >= 0 && group < e->n_groups && layer >= 0 && layer < kLayers, "bad argument");
  RMEM_REQUIRE(e->cfg.model == 0, "layer memories are exposed for the DeAOT model only");
  RMEM_TRY(const_cast<rmem_engine*>(e)->resolve_evict(group));
  const Group& gr = e->groups[group];
  const LayerState& L = gr.L[layer];
  // update_memory flipped the parity: the last propagated frame's K / V||ID_V are the "previous frame" buffers now
  const int last = gr.parity ^ 1;
  if (kbank) *kbank = L.kbank;
  if (vtbank) *vtbank = L.vtbank;
  if (q_last) *q_last = L.kc[last];
  if (vid_last) *vid_last = L.vid[last];
  if (nslots) *nslots = e->nslots;
  if (HWp) *HWp = e->g.HWp;
  if (T) *T = (int)gr.slots.size();
  if (slots)
    for (size_t t = 0; t < gr.slots.size() && t < (size_t)kMaxBankFrames; ++t) slots[t] = gr.slots[t];
  return RMEM_OK;
  RMEM_API_END
}

int rmem_engine_last_evict(const rmem_engine* e, int group, float* rel, int* n, int* drop) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(e && rel && n && drop && group >= 0 && group < e->n_groups, "bad argument");
  RMEM_TRY(const_cast<rmem_engine*>(e)->resolve_evict(group));
  const Group& gr = e->groups[group];
  *n = (int)gr.last_rel.size();
  for (int i = 0; i < *n; ++i) rel[i] = gr.last_rel[i];
  *drop = gr.last_drop;
  return RMEM_OK;
  RMEM_API_END
}

int rmem_engine_set_timing(rmem_engine* e, int on) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(e, "null engine");
  e->timing = on != 0;
  e->stage_ms.clear();
  e->marks.clear();
  e->ev_used = 0;
  return RMEM_OK;
  RMEM_API_END
}

int rmem_engine_get_timing(rmem_engine* e, char* buf, size_t cap) {
  RMEM_API_BEGIN
  RMEM_REQUIRE(e && buf && cap > 0, "bad argument");
  std::string out;
  for (auto& kv : e->stage_ms) {
    char line[160];
    snprintf(line, sizeof(line), "%s %.6f %lld\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  snprintf(buf, cap, "%s", out.c_str());
  return RMEM_OK;
  RMEM_API_END
}

long long rmem_engine_launch_count(const rmem_engine* e) { return e ? launch_counter() - e->launches0 : 0; }

}  // extern "C"
