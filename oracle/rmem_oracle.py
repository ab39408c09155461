"""CPU oracle for the RMem VOS hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A from-scratch functional restatement (plain torch CPU, fp32) of the per-frame
propagation path of Restricted-Memory/RMem:

    AOTInferEngine.add_reference_frame / match_propogate_one_frame / update_memory
    (aot_plus/networks/engines/aot_engine.py:571-725, deaot_engine.py:20-56)
    -> AOTEngine (aot_engine.py:241-465)
    -> DeAOT / DualBranchGPM / GatedPropagationModule
       (networks/models/deaot.py:10-69, networks/layers/transformer.py:700-1249)
    -> GatedPropagation / LocalGatedPropagation (networks/layers/attention.py:93-413)
    -> ResNet-50 stem+layer1-3 (networks/encoders/resnet.py:10-195),
       FPNSegmentationHead (networks/decoders/fpn.py:7-73)
    -> restrict_long_memories (transformer.py:880-991)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product path (rmem_b200/) never does.

PARITY PINNING.  The reference ships no golden vectors, KATs or tests (SURVEY.md
section 4), so the oracle is pinned against the reference ITSELF: oracle/make_golden.py
imports the unmodified reference from /root/reference in the build container,
loads the identical state_dict, runs identical inputs through both and (a) asserts
the two agree (fp32, <=1e-4 on logits, identical labels and eviction indices) and
(b) writes the reference's outputs to tests/golden/*.pt.  tests/test_oracle_golden.py
re-checks this oracle against those committed fixtures on every run.

Weights are addressed by the reference's own state_dict key names so a reference
checkpoint drops in unchanged.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

MAX_OBJ = 10          # configs/models/default.py:17  MODEL_MAX_OBJ_NUM
LOCAL_MAX_DIS = 7     # transformer.py:1012 max_local_dis -> 15x15 window
EMA_FACTOR = 0.8      # transformer.py:919 moving_mean_factor
UCB_ADD = 8.0         # transformer.py:953 add_item
UCB_MUL = 1.5         # transformer.py:954 mul_item
PE_MAX_T = 4          # transformer.py:1144 max_T


# --------------------------------------------------------------------------------------
# configuration + synthetic weights
# --------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    """Frozen subset of configs/models/r50_deaotl.py + configs/pre_vost.py."""
    model: str = "r50_deaotl"           # "r50_deaotl" | "r50_aotl"
    former_mem_len: int = 1
    latter_mem_len: int = 8             # shipped setting (configs/models/r50_deaotl.py:8); the T=8 workloads pass 7
    no_long_memory: bool = False        # NO_LONG_MEMORY (configs/models/r50_deaotl.py:20, aot_engine.py:339)
    gru_memory: bool = False            # GRU_MEMORY (r50_aotl only): ConvGRU condensation of evicted frames into bank slot 1
    d_model: int = 256                  # MODEL_ENCODER_EMBEDDING_DIM
    n_layers: int = 3                   # MODEL_LSTT_NUM
    max_obj: int = MAX_OBJ

    @property
    def is_deaot(self) -> bool:
        return self.model == "r50_deaotl"


from rmem_b200.synth import (_resnet_blocks, make_state_dict, snap_size, synthetic_frames,  # noqa: E402,F401
                              synthetic_label)


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def silu(x: Tensor) -> Tensor:                       # attention.py:89-90
    return x * torch.sigmoid(x)


def lin(sd, name: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def ln(sd, name: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def frozen_bn(sd, name: str, x: Tensor) -> Tensor:   # normalization.py:19-43 (eps 1e-5)
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                        sd[name + ".weight"], sd[name + ".bias"], False, 0.0, 1e-5)


def tokens_to_map(x: Tensor, h: int, w: int) -> Tensor:
    """[HW, C] token-major -> [1, C, h, w]  (basic.py:73-77 seq_to_2d, bs=1)."""
    return x.view(h, w, -1).permute(2, 0, 1).unsqueeze(0).contiguous()


def map_to_tokens(x: Tensor) -> Tensor:
    """[1, C, h, w] -> [HW, C]  (utils/tensor.py:3-6 bchw_2_lbc, bs=1)."""
    return x[0].flatten(1).t().contiguous()


def encode_image(sd, img: Tensor) -> List[Tensor]:
    """resnet.py:178-195 + aot.py:116-134.  Returns [4x(256), 8x(512), 16x(1024), proj16x(256)]."""
    x = F.conv2d(img, sd["encoder.conv1.weight"], None, stride=2, padding=3)
    x = F.relu(frozen_bn(sd, "encoder.bn1", x))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    feats = []
    for lname, bi, inpl, pl, s, ds in _resnet_blocks():
        p = f"encoder.{lname}.{bi}"
        idt = x
        o = F.relu(frozen_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"])))
        o = F.relu(frozen_bn(sd, p + ".bn2", F.conv2d(o, sd[p + ".conv2.weight"], None, stride=s, padding=1)))
        o = frozen_bn(sd, p + ".bn3", F.conv2d(o, sd[p + ".conv3.weight"]))
        if ds:
            idt = frozen_bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride=s))
        x = F.relu(o + idt)
        last = {"layer1": 2, "layer2": 3, "layer3": 5}[lname]
        if bi == last:
            feats.append(x)
    proj = F.conv2d(feats[-1], sd["encoder_projector.weight"], sd["encoder_projector.bias"])
    return [feats[0], feats[1], feats[2], proj]


def one_hot_with_ignore(label: Tensor, use_ignore: bool) -> Tensor:
    """utils/image.py:69-74 one_hot_mask + aot_engine.py:208-224 assign_identity (pre-conv part).
    label [1,1,H,W] (any dtype, integer-valued). Returns [1,12,H,W] fp32."""
    idx = torch.arange(0, MAX_OBJ + 1, device=label.device).view(1, -1, 1, 1)
    oh = (label == idx).float()
    if use_ignore:
        ign = (label == 255).float()
        oh[:, 0] = oh[:, 0] * (ign[:, 0] == 0).float()
    else:                                             # add_reference_frame passes no ignore mask (:305)
        ign = torch.zeros_like(oh[:, :1])
    return torch.cat([oh, ign], dim=1)


def id_embedding(sd, cfg: OracleConfig, one_hot12: Tensor) -> Tensor:
    """aot.py:111-114 / deaot.py:65-69.  [1,12,H,W] -> [HW, 256] tokens."""
    e = F.conv2d(one_hot12, sd["patch_wise_id_bank.weight"], sd["patch_wise_id_bank.bias"],
                 stride=16, padding=8)
    t = map_to_tokens(e)
    if cfg.is_deaot:
        t = ln(sd, "id_norm", t)
    return t


def dwconv5(x_tok: Tensor, w: Tensor, h: int, wd: int) -> Tensor:
    """basic.py:38-59 DWConv2d on token-major [HW, C] (dropout = identity in eval)."""
    m = tokens_to_map(x_tok, h, wd)
    m = F.conv2d(m, w, None, padding=2, groups=m.shape[1])
    return map_to_tokens(m)


def temporal_pe_slots(T: int, n_slots: int = 4) -> List[Tuple[int, int, float]]:
    """Slot interpolation of transformer.py:1140-1170 as (lo, hi, frac) per memory frame t:
    pe_t = (1-frac)*mem_pe[lo] + frac*mem_pe[hi].  Derived by running the same F.interpolate
    sequence on a one-hot basis, so it is exact for any T."""
    eye = torch.eye(n_slots)                          # [slot, basis]
    mp = eye[:T] if T <= n_slots else eye
    if T == 1:
        wts = mp[0:1]
    else:
        x = mp.t().reshape(1, n_slots, -1)
        if T <= PE_MAX_T:
            x = F.interpolate(x, size=T, mode="linear", align_corners=True)
        else:
            x = F.interpolate(x, size=PE_MAX_T, mode="linear", align_corners=True)
            x = torch.flip(x, dims=(-1,))
            x = F.interpolate(x, size=T, mode="nearest")
            x = torch.flip(x, dims=(-1,))
        wts = x.view(n_slots, T).t()
    out = []
    for t in range(T):
        nz = torch.nonzero(wts[t]).flatten().tolist()
        if len(nz) == 1:
            out.append((nz[0], nz[0], 0.0))
        else:
            out.append((nz[0], nz[1], float(wts[t, nz[1]])))
    return out


def temporal_pe(mem_pos_emb: Tensor, T: int) -> Tensor:
    """[T, C] temporal positional embedding rows added to each memory frame's K."""
    rows = []
    for lo, hi, fr in temporal_pe_slots(T, mem_pos_emb.shape[0]):
        rows.append(mem_pos_emb[lo] if fr == 0.0 else (1 - fr) * mem_pos_emb[lo] + fr * mem_pos_emb[hi])
    return torch.stack(rows)


def long_term_attention(q: Tensor, k_bank: Tensor, v_bank: Tensor, scale_dim: int,
                        n_head: int = 1) -> Tuple[Tensor, Tensor]:
    """Dense attention over the restricted bank with per-frame mass (attention.py:174-193 /
    :45-64 and transformer.py:1185-1192 / :636-643).
    q [HW, C] (cur PE already added), k_bank [T, HW, C] (mem PE already added), v_bank [T, HW, Dv].
    Returns (out [HW, Dv], mass [HW, T])."""
    T, HW, C = k_bank.shape
    hd = C // n_head
    dv = v_bank.shape[-1] // n_head
    qh = (q / math.sqrt(scale_dim)).view(-1, n_head, hd).permute(1, 0, 2)           # [h, HWq, hd]
    kh = k_bank.reshape(T * HW, n_head, hd).permute(1, 2, 0)                        # [h, hd, THW]
    vh = v_bank.reshape(T * HW, n_head, dv).permute(1, 0, 2)                        # [h, THW, dv]
    attn = torch.softmax(qh @ kh, dim=-1)                                           # [h, HWq, THW]
    out = (attn @ vh).permute(1, 0, 2).reshape(q.shape[0], -1)
    mass = attn.view(n_head, q.shape[0], T, HW).mean(0).sum(-1)                     # [HWq, T]
    return out, mass


def local_attention(q: Tensor, k_prev: Tensor, v_prev: Tensor, rel_w: Tensor, rel_b: Tensor,
                    h: int, w: int, d_att: int = 128) -> Tensor:
    """LocalGatedPropagation core (attention.py:289-353) on token-major tensors, 1 head.
    q,k_prev [HW,128]; v_prev [HW,Dv]; rel_w [225,128]; rel_b [225].  Returns agg [HW, Dv].

    Restated as an explicit windowed gather (the reference's unfold + dense scatter is an
    implementation detail): s[i,d] = q_i/sqrt(d).k_{i+d} + rel[d,i], out-of-bounds -> -1e8."""
    HW = h * w
    md = LOCAL_MAX_DIS
    ws = 2 * md + 1
    rel = q @ rel_w.t() + rel_b                                     # [HW, 225]  (conv1x1 on UNscaled q, :314)
    qs = q / math.sqrt(d_att)
    ys = torch.arange(h).view(h, 1).expand(h, w).reshape(-1)
    xs = torch.arange(w).view(1, w).expand(h, w).reshape(-1)
    dy = torch.arange(-md, md + 1).view(ws, 1).expand(ws, ws).reshape(-1)
    dx = torch.arange(-md, md + 1).view(1, ws).expand(ws, ws).reshape(-1)
    ny = ys.view(-1, 1) + dy.view(1, -1)
    nx = xs.view(-1, 1) + dx.view(1, -1)
    valid = (ny >= 0) & (ny < h) & (nx >= 0) & (nx < w)             # [HW, 225]
    nidx = (ny.clamp(0, h - 1) * w + nx.clamp(0, w - 1))            # [HW, 225]
    kg = k_prev[nidx]                                               # [HW, 225, 128]
    s = torch.einsum("ic,idc->id", qs, kg)
    s = torch.where(valid, s, torch.zeros_like(s))                  # zero-padded K (:404-413)
    s = s + rel
    s = s - (~valid).float() * 1e8                                  # :344
    p = torch.softmax(s, dim=1)
    p = torch.where(valid, p, torch.zeros_like(p))                  # OOB cells are sliced away (:397-400)
    dense = torch.zeros(HW, HW, dtype=p.dtype)                      # local2global (:363-402)
    dense.scatter_add_(1, nidx, p)
    return dense @ v_prev


def gated_epilogue(sd, prefix: str, agg: Tensor, u: Tensor, h: int, w: int) -> Tensor:
    """attention.py:206-209 / :355-358:  Linear(DWConv5x5(agg * u))."""
    x = dwconv5(agg * u, sd[prefix + ".dw_conv.conv.weight"], h, w)
    return lin(sd, prefix + ".projection", x)


def group_norm_tokens(x: Tensor, wgt: Tensor, b: Tensor, groups: int) -> Tensor:
    """basic.py:6-12 GroupNorm1D on [HW, C] (bs = 1): stats over (C/groups x HW)."""
    return F.group_norm(x.t().unsqueeze(0), groups, wgt, b, 1e-5)[0].t().contiguous()


def sine_pos_emb(h: int, w: int, d_model: int = 256) -> Tensor:
    """position.py:35-77 PositionEmbeddingSine(num_pos_feats=d/2, normalize=True) -> [HW, d]."""
    npf = d_model // 2
    y = torch.arange(h, dtype=torch.float32).view(h, 1).expand(h, w)
    x = torch.arange(w, dtype=torch.float32).view(1, w).expand(h, w)
    y = y / (y[-1:, :] + 1e-6) * (2 * math.pi)
    x = x / (x[:, -1:] + 1e-6) * (2 * math.pi)
    dim_t = torch.arange(npf, dtype=torch.float32)
    dim_t = 10000 ** (2 * torch.div(dim_t, 2, rounding_mode="trunc") / npf)
    px = x[:, :, None] / dim_t
    py = y[:, :, None] / dim_t
    px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)
    py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
    return torch.cat((py, px), dim=2).view(h * w, d_model)


def decode_logits(sd, x16: Tensor, shortcuts: Sequence[Tensor]) -> Tensor:
    """fpn.py:36-68.  x16 [1, Cin, h, w]; shortcuts = encode_image() list."""
    def conv_gn(name, x, pad):
        x = F.conv2d(x, sd[f"decoder.{name}.conv.weight"], sd[f"decoder.{name}.conv.bias"], padding=pad)
        return F.group_norm(x, 8, sd[f"decoder.{name}.gn.weight"], sd[f"decoder.{name}.gn.bias"], 1e-5)

    def adapter(name, x):
        return F.conv2d(x, sd[f"decoder.{name}.weight"], sd[f"decoder.{name}.bias"])

    x = F.relu(conv_gn("conv_in", x16, 0))
    x = F.relu(conv_gn("conv_16x", adapter("adapter_16x", shortcuts[-2]) + x, 1))
    x = F.interpolate(x, size=shortcuts[-3].shape[-2:], mode="bilinear", align_corners=True)
    x = F.relu(conv_gn("conv_8x", adapter("adapter_8x", shortcuts[-3]) + x, 1))
    x = F.interpolate(x, size=shortcuts[-4].shape[-2:], mode="bilinear", align_corners=True)
    x = F.relu(conv_gn("conv_4x", adapter("adapter_4x", shortcuts[-4]) + x, 1))
    return F.conv2d(x, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"])


def soft_logit_aggregation(all_logits: Sequence[Tensor]) -> Tensor:
    """aot_engine.py:650-673."""
    if len(all_logits) == 1:
        return all_logits[0]
    fg, bg = [], []
    for lg in all_logits:
        p = torch.softmax(lg, dim=1)
        bg.append(p[:, 0:1]); fg.append(p[:, 1:1 + MAX_OBJ])
    bgp = torch.prod(torch.cat(bg, dim=1), dim=1, keepdim=True)
    merged = torch.cat([bgp] + fg, dim=1).clamp(1e-5, 1 - 1e-5)
    return torch.logit(merged)


def logits_to_label(logits: Tensor) -> Tensor:
    """evaluator.py:430-441 (no TTA): softmax -> argmax -> float label [1,1,H,W]."""
    return torch.argmax(torch.softmax(logits, dim=1), dim=1, keepdim=True).float()


def separate_mask(mask: Tensor, n_engines: int) -> List[Tensor]:
    """aot_engine.py:604-618 (label-map branch)."""
    if n_engines == 1:
        return [mask]
    out = []
    for i in range(n_engines):
        lo, hi = i * MAX_OBJ + 1, (i + 1) * MAX_OBJ
        fgm = ((mask >= lo) & (mask <= hi)).float()
        out.append((fgm * mask - lo + 1) * fgm)
    return out


def evict_scores(mass: Tensor, fg: Tensor) -> Tensor:
    """transformer.py:891-906:  r[t] = sum_i mass[i,t]*fg[i], normalised over t."""
    r = (mass * fg.view(-1, 1)).sum(0)
    return r / r.sum()


@dataclass
class EvictState:
    """stored_attn_weight_dict / stored_frame_times (transformer.py:993-998)."""
    ema: Dict[int, float] = field(default_factory=dict)
    times: Dict[int, int] = field(default_factory=dict)


def evict_pick(rel: Tensor, idx: List[int], st: EvictState, former: int, gru: bool = False) -> int:
    """transformer.py:907-964 (AOT twin :354-411).  rel [T_old] fp32 relevance, idx = long_memories_indexes AFTER the
    append (len T_old+1).  Updates st in place and returns the logical index to drop.
    gru (GRU_MEMORY, :337-338, 395-396, 406-411): bank position 1 holds the condensed memory and is never dropped either."""
    T_old = rel.numel()
    rel = rel.clone().float()
    new_ema = {}
    for i in range(T_old):
        f = idx[i]
        a = rel[i].clone()
        new_ema[f] = (1 - EMA_FACTOR) * st.ema[f] + EMA_FACTOR * a if f in st.ema else a
    st.ema = new_ema
    for i in range(T_old):
        rel[i] = new_ema[idx[i]]
    st.times = {f: 1 + st.times.get(f, 0) for f in idx}
    tt = torch.tensor([float(st.times[f]) for f in idx[:-1]], dtype=torch.float32)
    tt[0] = float(len(tt))
    if gru and len(tt) > 1:
        tt[1] = float(len(tt))
    bonus = UCB_MUL * torch.sqrt(torch.log(tt.sum()) / (tt + UCB_ADD))
    score = rel + bonus
    skip = 2 if gru else 1
    drop = former + (1 if gru else 0)
    if score.numel() > skip:
        drop = int(torch.argmin(score[skip:]).item()) + skip
    return drop


def conv_gru(sd, p: str, x: Tensor, h_cur: Tensor) -> Tuple[Tensor, Tensor]:
    """ConvGRUCellOutput.forward (transformer.py:84-118) on [1,C,h,w] maps: returns (h_next, output_conv(h_next)).
    padding = "same": the 2x2 kernel of the K cell pads one row / column at the bottom / right."""
    C = h_cur.shape[1]
    cc = F.conv2d(torch.cat([x, h_cur], 1), sd[p + ".conv_gru_cell.conv_gates.weight"],
                  sd[p + ".conv_gru_cell.conv_gates.bias"], padding="same")
    gamma, beta = torch.split(cc, C, dim=1)
    reset, update = torch.sigmoid(gamma), torch.sigmoid(beta)
    cnm = torch.tanh(F.conv2d(torch.cat([x, reset * h_cur], 1), sd[p + ".conv_gru_cell.conv_can.weight"],
                              sd[p + ".conv_gru_cell.conv_can.bias"], padding="same"))
    h_next = (1 - update) * h_cur + update * cnm
    return h_next, F.conv2d(h_next, sd[p + ".output_conv.weight"], sd[p + ".output_conv.bias"])


# --------------------------------------------------------------------------------------
# DeAOT GPM layer (transformer.py:1091-1244)
# --------------------------------------------------------------------------------------
@dataclass
class LayerMem:
    K: Tensor                 # [HW,128]
    V: Tensor                 # [HW,512]
    ID_V: Optional[Tensor]    # [HW,512]  (after fuse_key_value_id)
    curr_ID_V: Optional[Tensor] = None   # id_norm1(tgt_id) kept for the refresh (layer>0)


@dataclass
class Bank:
    """Per-engine memory (transformer.py:993-1007): long bank per layer as lists of frames."""
    long: List[List[LayerMem]] = field(default_factory=list)     # [layer][t]
    short: List[LayerMem] = field(default_factory=list)          # [layer]
    curr: List[LayerMem] = field(default_factory=list)           # [layer] from last forward
    short_next: list = field(default_factory=list)               # AOT: [layer] [linear_QMem(o3), o3] of the last forward
    mass0: Optional[Tensor] = None                               # layer-0 [HW,T]
    evict: EvictState = field(default_factory=EvictState)
    gru_h: list = field(default_factory=list)                    # GRU_MEMORY: [layer][K|V] hidden state [1,C,h,w]


def fuse_id(sd, l: int, curr_ID_V: Optional[Tensor], id_emb: Tensor) -> Tensor:
    """transformer.py:1238-1244."""
    p = f"LSTT.layers.{l}.linear_ID_V"
    if curr_ID_V is not None:
        return silu(lin(sd, p, torch.cat([curr_ID_V, id_emb], dim=1)))
    return silu(lin(sd, p, id_emb))


def gpm_layer(sd, l: int, tgt: Tensor, tgt_id: Optional[Tensor], bank: Bank, id_emb: Optional[Tensor],
              h: int, w: int, record_mass: bool) -> Tuple[Tensor, Tensor, LayerMem, dict]:
    p = f"LSTT.layers.{l}"
    t = ln(sd, p + ".norm1", tgt)
    qv = lin(sd, p + ".linear_QV", t)
    Q = qv[:, :128]
    V = silu(qv[:, 128:])
    U = lin(sd, p + ".linear_U", t)
    if tgt_id is None:
        cU = torch.cat([silu(U), torch.ones_like(U)], dim=1)
        curr_ID_V = None
        tgt_id = torch.zeros_like(tgt)
    else:
        ti = ln(sd, p + ".id_norm1", tgt_id)
        curr_ID_V = ti
        cU = silu(torch.cat([U, lin(sd, p + ".linear_ID_U", ti)], dim=1))

    if id_emb is not None:                                        # reference-frame mode (:1125-1135)
        gID = fuse_id(sd, l, curr_ID_V, id_emb)
        longs = [LayerMem(Q, V, gID)]
        short = LayerMem(Q, V, gID)
    else:
        longs = bank.long[l]
        short = bank.short[l]

    T = len(longs)
    pe = temporal_pe(sd["mem_pos_emb"], T)                        # [T,128]
    Kt = torch.stack([m.K for m in longs]) + pe.view(T, 1, -1)
    Vt = torch.stack([torch.cat([m.V, m.ID_V], dim=1) for m in longs])
    Qt = Q + sd["cur_pos_emb"].view(1, -1)
    A, mass = long_term_attention(Qt, Kt, Vt, 128)
    o2 = gated_epilogue(sd, p + ".long_term_attn", A, cU, h, w)

    a3 = local_attention(Q, short.K, torch.cat([short.V, short.ID_V], dim=1),
                         sd[p + ".short_term_attn.relative_emb_k.weight"].view(225, 128),
                         sd[p + ".short_term_attn.relative_emb_k.bias"], h, w)
    o3 = gated_epilogue(sd, p + ".short_term_attn", a3, cU, h, w)

    tgt = tgt + o2[:, :256] + o3[:, :256]
    tgt_id = tgt_id + o2[:, 256:] + o3[:, 256:]

    z = torch.cat([ln(sd, p + ".norm2", tgt), ln(sd, p + ".id_norm2", tgt_id)], dim=1)
    sp = p + ".self_attn"
    qk = lin(sd, sp + ".linear_QK", z)
    v = silu(torch.cat([lin(sd, sp + ".linear_V1", z[:, :256]), lin(sd, sp + ".linear_V2", z[:, 256:])], dim=1))
    u = silu(torch.cat([lin(sd, sp + ".linear_U1", z[:, :256]), lin(sd, sp + ".linear_U2", z[:, 256:])], dim=1))
    a, _ = long_term_attention(qk, qk.unsqueeze(0), v.unsqueeze(0), 128)
    o = gated_epilogue(sd, sp, a, u, h, w)
    tgt = tgt + o[:, :256]
    tgt_id = tgt_id + o[:, 256:]

    curr = LayerMem(Q, V, None, curr_ID_V)
    dbg = dict(Q=Q, V=V, U=U, cU=cU, A=A, mass=mass, o2=o2, a3=a3, o3=o3, a_self=a, o_self=o)
    if id_emb is not None:
        dbg["ref_long"] = longs[0]
    return tgt, tgt_id, curr, dbg


def gpm_forward(sd, tgt: Tensor, bank: Bank, id_emb: Optional[Tensor], h: int, w: int,
                record_mass: bool = True, collect: Optional[list] = None) -> Tensor:
    """DualBranchGPM.forward (transformer.py:765-824).  Returns final [HW,512] (GroupNorm1D(2))."""
    tgt_id = None
    bank.curr = []
    ref_long = []
    for l in range(3):
        tgt, tgt_id, curr, dbg = gpm_layer(sd, l, tgt, tgt_id, bank, id_emb, h, w, record_mass)
        bank.curr.append(curr)
        if l == 0 and id_emb is None:
            bank.mass0 = dbg["mass"]
        if id_emb is not None:
            ref_long.append(dbg["ref_long"])
        if collect is not None:
            collect.append(dbg)
    if id_emb is not None:                                        # init_memory (:993-998)
        bank.long = [[m] for m in ref_long]
        bank.short = list(ref_long)
        bank.evict = EvictState()
    cat = torch.cat([tgt, tgt_id], dim=1)
    return group_norm_tokens(cat, sd["LSTT.decoder_norms.0.gn.weight"],
                             sd["LSTT.decoder_norms.0.gn.bias"], 2)



# --------------------------------------------------------------------------------------
# AOT LSTT layer (transformer.py:553-692) -- stage pre_vost (MODEL_LINEAR_Q = False)
# --------------------------------------------------------------------------------------
N_HEAD = 8            # configs/models/default.py:19-20  MODEL_SELF_HEADS / MODEL_ATT_HEADS


@dataclass
class AotMem:
    K: Tensor                 # [HW,256]
    V: Tensor                 # [HW,256]


def mha(sd, prefix: str, Q: Tensor, K: Tensor, V: Tensor, use_linear: bool, n_head: int = N_HEAD
        ) -> Tuple[Tensor, Tensor]:
    """MultiheadAttention.forward (attention.py:28-81) on [L, C] tensors (bs = 1).
    Returns (projection(concat_h softmax(Q_h K_h^T / sqrt(d)) V_h), attn [h, Lq, Lk])."""
    if use_linear:
        Q, K, V = lin(sd, prefix + ".linear_Q", Q), lin(sd, prefix + ".linear_K", K), lin(sd, prefix + ".linear_V", V)
    C = Q.shape[1]
    d = C // n_head
    qh = (Q / math.sqrt(d)).view(-1, n_head, d).permute(1, 0, 2)
    kh = K.view(-1, n_head, d).permute(1, 2, 0)
    vh = V.view(-1, n_head, d).permute(1, 0, 2)
    attn = torch.softmax(qh @ kh, dim=-1)
    out = (attn @ vh).permute(1, 0, 2).reshape(Q.shape[0], C)
    return lin(sd, prefix + ".projection", out), attn


def gn_gelu_dwconv(sd, prefix: str, x: Tensor, h: int, w: int) -> Tensor:
    """GNActDWConv2d (basic.py:15-35): GroupNorm(32) -> GELU -> depthwise 5x5, token-major [HW, C]."""
    m = tokens_to_map(x, h, w)
    m = F.group_norm(m, 32, sd[prefix + ".gn.weight"], sd[prefix + ".gn.bias"], 1e-5)
    m = F.gelu(m)
    m = F.conv2d(m, sd[prefix + ".conv.weight"], None, padding=2, groups=m.shape[1])
    return map_to_tokens(m)


def aot_layer(sd, l: int, tgt: Tensor, bank: "Bank", id_emb: Optional[Tensor], pos: Tensor, h: int, w: int
              ) -> Tuple[Tensor, AotMem, AotMem, dict]:
    """SimplifiedTransformerBlock.forward.  Returns (tgt', curr=[K, V], short'=[local_K, local_V], dbg)."""
    p = f"LSTT.layers.{l}"
    # self-attention with sine PE on q, k (:566-571)
    t = ln(sd, p + ".norm1", tgt)
    qk = t + pos
    o1, _ = mha(sd, p + ".self_attn", qk, qk, t, use_linear=True)
    tgt = tgt + o1
    # long / short term attention (:574-680)
    t = ln(sd, p + ".norm2", tgt)
    Q = lin(sd, p + ".linear_Q", t)
    K, V = Q, t
    if id_emb is not None:                                        # reference frame (:582-588)
        gV = lin(sd, p + ".linear_V", V + id_emb)
        longs = [AotMem(K, gV)]
        short = AotMem(K, gV)
    else:
        longs, short = bank.long[l], bank.short[l]
    T = len(longs)
    pe = temporal_pe(sd["mem_pos_emb"], T)                        # [T,256]
    Kt = torch.stack([m.K for m in longs]) + pe.view(T, 1, -1)
    Vt = torch.stack([m.V for m in longs])
    Qt = Q + sd["cur_pos_emb"].view(1, -1)
    o2, attn = mha(sd, p + ".long_term_attn", Qt, Kt.flatten(0, 1), Vt.flatten(0, 1), use_linear=False)
    HW = Q.shape[0]
    mass = attn.view(N_HEAD, HW, T, HW).mean(0).sum(-1)           # [HW, T]  (:636-643)
    o3, _ = mha(sd, p + ".short_term_attn", Q, ln(sd, p + ".norm4", short.K + K),
                ln(sd, p + ".norm4", short.V + V), use_linear=False)        # (:656-662)
    local_K = lin(sd, p + ".linear_QMem", o3)
    local_V = o3
    if id_emb is not None:
        local_V = lin(sd, p + ".linear_VMem", local_V + id_emb)   # (:677-678)
    tgt = tgt + o2 + o3
    # feed-forward (:683-687)
    t = ln(sd, p + ".norm3", tgt)
    ff = lin(sd, p + ".linear2", gn_gelu_dwconv(sd, p + ".activation", lin(sd, p + ".linear1", t), h, w))
    tgt = tgt + ff
    dbg = dict(Q=Q, V=V, o1=o1, o2=o2, o3=o3, mass=mass, ff=ff)
    if id_emb is not None:
        dbg["ref_long"] = longs[0]
    return tgt, AotMem(K, V), AotMem(local_K, local_V), dbg


def lstt_forward(sd, tgt: Tensor, bank: "Bank", id_emb: Optional[Tensor], pos: Tensor, h: int, w: int,
                 collect: Optional[list] = None) -> List[Tensor]:
    """LongShortTermTransformer.forward (transformer.py:199-267): three LayerNormed layer outputs."""
    bank.curr, bank.short_next = [], []
    outs, ref_long = [], []
    for l in range(3):
        tgt, curr, short_next, dbg = aot_layer(sd, l, tgt, bank, id_emb, pos, h, w)
        bank.curr.append(curr)
        bank.short_next.append(short_next)
        if l == 0 and id_emb is None:
            bank.mass0 = dbg["mass"]
        if id_emb is not None:
            ref_long.append(dbg["ref_long"])
        if collect is not None:
            collect.append(dbg)
        outs.append(ln(sd, f"LSTT.decoder_norms.{l}", tgt))
    if id_emb is not None:                                        # init_memory (:438-443)
        bank.long = [[m] for m in ref_long]
        bank.short = list(bank.short_next)
        bank.evict = EvictState()
    return outs

# --------------------------------------------------------------------------------------
# engine state machine
# --------------------------------------------------------------------------------------
class OracleSubEngine:
    """One AOTEngine (<=10 objects) with its own bank (reference + per-engine deepcopy)."""

    def __init__(self, sd, cfg: OracleConfig, gap: int):
        self.sd, self.cfg, self.gap = sd, cfg, gap
        self.bank = Bank()
        self.frame_step = 0
        self.last_mem_step = -1
        self.long_memories_indexes: List[int] = []
        self.pred_id_logits: Optional[Tensor] = None
        self.hw: Tuple[int, int] = (0, 0)
        self.dbg: Optional[list] = None
        self.pos: Optional[Tensor] = None

    def _lstt(self, feats, id_emb):
        h, w = feats[-1].shape[-2:]
        self.hw = (h, w)
        tgt = map_to_tokens(feats[-1])
        if self.cfg.is_deaot:
            out = gpm_forward(self.sd, tgt, self.bank, id_emb, h, w, collect=self.dbg)
        else:                                                     # aot.py:136-142: cat(proj16x, l0, l1, l2)
            if self.pos is None or self.pos.shape[0] != h * w:
                self.pos = sine_pos_emb(h, w, self.cfg.d_model)   # aot_engine.py:289-292 (once per clip)
            outs = lstt_forward(self.sd, tgt, self.bank, id_emb, self.pos, h, w, collect=self.dbg)
            out = torch.cat([tgt] + outs, dim=1)
        self.pred_id_logits = decode_logits(self.sd, tokens_to_map(out, h, w), feats)
        return self.pred_id_logits

    def add_reference_frame(self, feats, mask: Tensor, frame_step: int):
        oh = one_hot_with_ignore(mask, use_ignore=False)
        id_emb = id_embedding(self.sd, self.cfg, oh)
        self._lstt(feats, id_emb)
        if self.cfg.gru_memory:                                   # init_memory (transformer.py:444-453, aot_engine.py:322)
            h, w = self.hw
            self.bank.gru_h = [[torch.zeros(1, self.cfg.d_model, h, w), torch.zeros(1, self.cfg.d_model, h, w)]
                               for _ in range(3)]
        self.last_mem_step = frame_step
        self.long_memories_indexes.append(self.frame_step)

    def propagate(self, feats, output_size) -> Tensor:
        self.frame_step += 1
        lg = self._lstt(feats, None)
        if output_size is not None:
            lg = F.interpolate(lg, size=output_size, mode="bilinear", align_corners=True)
        return lg

    def update_memory(self, label: Tensor):
        sd, cfg, bank = self.sd, self.cfg, self.bank
        h, w = self.hw
        oh = one_hot_with_ignore(label, use_ignore=True)
        id_emb = id_embedding(sd, cfg, oh)
        is_long = (not cfg.no_long_memory) and self.frame_step - self.last_mem_step >= self.gap
        if is_long:
            self.last_mem_step = self.frame_step
        if cfg.is_deaot:
            for l in range(3):                                    # transformer.py:826-857
                c = bank.curr[l]
                c.ID_V = fuse_id(sd, l, c.curr_ID_V, id_emb)
            bank.short = [LayerMem(c.K, c.V, c.ID_V) for c in bank.curr]
        else:
            for l in range(3):                                    # transformer.py:269-304
                p = f"LSTT.layers.{l}"
                bank.curr[l].V = lin(sd, p + ".linear_V", bank.curr[l].V + id_emb)
                bank.short_next[l].V = lin(sd, p + ".linear_VMem", bank.short_next[l].V + id_emb)
            bank.short = list(bank.short_next)
        if is_long:
            for l in range(3):
                c = bank.curr[l]
                bank.long[l].append(LayerMem(c.K, c.V, c.ID_V) if cfg.is_deaot else AotMem(c.K, c.V))
            self.long_memories_indexes.append(self.frame_step)
            if (not cfg.is_deaot) and len(bank.long[0]) <= cfg.former_mem_len + cfg.latter_mem_len:
                return                                            # AOT only: early return, no state change (:332-334)
            lg = F.interpolate(self.pred_id_logits, size=(h, w), mode="bilinear", align_corners=True)
            fg = 1 - torch.softmax(lg, dim=1)[0, 0].flatten()     # aot_engine.py:355-362
            rel = evict_scores(bank.mass0, fg)
            drop = evict_pick(rel, self.long_memories_indexes, bank.evict, cfg.former_mem_len, gru=cfg.gru_memory)
            self.last_rel, self.last_drop = rel, drop
            if len(bank.long[0]) > cfg.former_mem_len + cfg.latter_mem_len:
                for l in range(3):
                    if cfg.gru_memory:                            # transformer.py:420-430: the dropped frame goes through
                        for i, nm in enumerate(("K", "V")):       # the layer's ConvGRU, its output replaces bank entry 1
                            x = tokens_to_map(getattr(bank.long[l][drop], nm), h, w)
                            bank.gru_h[l][i], out = conv_gru(sd, f"LSTT.layers.{l}.memory_grus.{i}", x, bank.gru_h[l][i])
                            setattr(bank.long[l][1], nm, out.flatten(2)[0].t().contiguous())
                    del bank.long[l][drop]
                self.long_memories_indexes.pop(drop)


class OracleEngine:
    """Mirror of DeAOTInferEngine (deaot_engine.py:20-56 + aot_engine.py:571-725)."""

    def __init__(self, sd, cfg: OracleConfig, long_term_mem_gap: int = 9999):
        self.sd, self.cfg = sd, cfg
        self.long_term_mem_gap = long_term_mem_gap
        self.restart_engine()

    def restart_engine(self):
        self.aot_engines: List[OracleSubEngine] = []
        self.input_size_2d = None

    def add_reference_frame(self, img: Tensor, mask: Tensor, obj_nums, frame_step: int = -1):
        if isinstance(obj_nums, (list, tuple)):
            obj_nums = obj_nums[0]
        n = max(int(math.ceil(obj_nums / MAX_OBJ)), 1)
        while n > len(self.aot_engines):
            self.aot_engines.append(OracleSubEngine(self.sd, self.cfg, self.long_term_mem_gap))
        feats = encode_image(self.sd, img)
        fs = 0 if frame_step == -1 else frame_step
        for e, m in zip(self.aot_engines, separate_mask(mask, len(self.aot_engines))):
            e.gap = self.long_term_mem_gap
            e.add_reference_frame(feats, m, fs)
        self.input_size_2d = tuple(img.shape[-2:])
        self.enc_size_2d = self.aot_engines[0].hw
        self.enc_hw = self.enc_size_2d[0] * self.enc_size_2d[1]

    def match_propogate_one_frame(self, img: Tensor, output_size=None) -> Tensor:
        feats = encode_image(self.sd, img)
        lgs = []
        for e in self.aot_engines:
            e.gap = self.long_term_mem_gap
            lgs.append(e.propagate(feats, output_size))
        return soft_logit_aggregation(lgs)

    def update_memory(self, label: Tensor):
        for e, m in zip(self.aot_engines, separate_mask(label, len(self.aot_engines))):
            e.update_memory(m)


# --------------------------------------------------------------------------------------
# synthetic clips
# --------------------------------------------------------------------------------------
def run_clip(engine, frames: Tensor, label0: Tensor, n_obj: int, out_size=None, on_frame=None,
             forced_labels: Optional[Tensor] = None):
    """The evaluator's per-clip loop (evaluator.py:337-527) without IO.  `engine` is anything with
    the reference's AOTInferEngine surface.  Returns list of uint8 labels per propagated frame.

    forced_labels [F-1,Ho,Wo] uint8 (teacher forcing): the memory update of frame f uses
    forced_labels[f-1] instead of the engine's own argmax.  The label feedback loop is chaotic
    under random weights (one flipped near-tie pixel grows to thousands within a few frames), so
    cross-implementation parity is checked in lock-step on identical label histories."""
    H, W = frames.shape[-2:]
    out_size = out_size or (H, W)
    engine.restart_engine()
    engine.add_reference_frame(frames[0:1], label0, obj_nums=[n_obj], frame_step=0)
    labels = []
    for f in range(1, frames.shape[0]):
        logit = engine.match_propogate_one_frame(frames[f:f + 1], output_size=out_size)
        lab = logits_to_label(logit)
        labels.append(lab.to(torch.uint8))
        if forced_labels is not None:
            lab = forced_labels[f - 1].to(lab.device).float().view(1, 1, *out_size)
        lab_in = F.interpolate(lab, size=(H, W), mode="nearest")
        engine.update_memory(lab_in)
        if on_frame is not None:
            on_frame(f, logit, lab)
    return labels
