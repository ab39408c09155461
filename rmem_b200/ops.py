"""Op-level Python mirrors of the reference's module forwards on this path, each a thin
call into the C ABI.  Tensors are CUDA torch tensors; activations token-major [pixels, channels].

    GatedPropagation.forward core      networks/layers/attention.py:139-211   -> long_attention / self use
    LocalGatedPropagation.forward core networks/layers/attention.py:289-361   -> local_attention
    nn.Linear / nn.Conv2d call sites                                          -> gemm / conv2d_nhwc
    LayerNorm / GroupNorm / DWConv2d   networks/layers/basic.py               -> layernorm / groupnorm / dwconv5x5
    patch_wise_id_bank (+id_norm)      networks/models/aot.py:63-74,111-114   -> id_embedding
    logits -> mask IDs                 networks/engines/aot_engine.py:457-463,650-673 -> mask_head
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import _capi

ACT_NONE, ACT_RELU, ACT_SILU = 0, 1, 2


def _bf(t: torch.Tensor) -> torch.Tensor:
    return t.to(_capi.op_dtype()).contiguous()


def round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def gemm(A: torch.Tensor, B: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
         act_from: int = 0, residual: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None,
         out_f32: bool = False, alpha: float = 1.0, bias_along_m: bool = False,
         accumulate_into: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C = epilogue(alpha * A[M,K] @ B[N,K]^T).  A, B bf16."""
    lib = _capi.load()
    M, K = A.shape
    N = B.shape[0]
    assert A.dtype == _capi.op_dtype() and B.dtype == _capi.op_dtype() and B.shape[1] == K
    if accumulate_into is not None:
        out = accumulate_into
        assert out.dtype == torch.float32
    else:
        out = torch.empty(M, N, dtype=torch.float32 if out_f32 else _capi.op_dtype(), device=A.device)
    d = _capi.GemmDesc()
    d.A, d.lda, d.B, d.ldb = A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0)
    d.M, d.N, d.K = M, N, K
    d.alpha = alpha
    d.bias = 0 if bias is None else bias.data_ptr()
    d.bias_along_m = int(bias_along_m)
    d.act, d.act_from_col = act, act_from
    if residual is not None:
        d.residual, d.ldr = residual.data_ptr(), residual.stride(0)
    if gate is not None:
        d.gate, d.ldg = gate.data_ptr(), gate.stride(0)
    d.accumulate = int(accumulate_into is not None)
    d.C, d.ldc, d.c_is_f32 = out.data_ptr(), out.stride(0), int(out.dtype == torch.float32)
    d.n_split = 1 << 30
    _capi.check(lib.rmem_gemm_fwd(C.byref(d), _capi.stream_ptr()))
    return out


def conv2d_nhwc(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, stride: int, pad: int, act: int = ACT_NONE,
                residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x t16 [Hin,Win,Cin] (or a stack [n,Hin,Win,Cin]); w t16 [Cout,kh,kw,Cin] -> t16 [(n,)Hout,Wout,Cout]."""
    lib = _capi.load()
    nimg = x.shape[0] if x.dim() == 4 else 1
    Hin, Win, Cin = x.shape[-3:]
    Cout, kh, kw, _ = w.shape
    Hout, Wout = (Hin + 2 * pad - kh) // stride + 1, (Win + 2 * pad - kw) // stride + 1
    oshape = (nimg, Hout, Wout, Cout) if x.dim() == 4 else (Hout, Wout, Cout)
    out = torch.empty(*oshape, dtype=_capi.op_dtype(), device=x.device)
    d = _capi.GemmDesc()
    d.A, d.B, d.ldb = x.data_ptr(), w.data_ptr(), kh * kw * Cin
    d.M, d.N, d.K = nimg * Hout * Wout, Cout, kh * kw * Cin
    d.n_images = nimg
    d.conv, d.Hin, d.Win, d.Cin, d.Wout, d.kw, d.stride, d.pad = 1, Hin, Win, Cin, Wout, kw, stride, pad
    d.alpha = 1.0
    d.bias = bias.data_ptr()
    d.act = act
    if residual is not None:
        d.residual, d.ldr = residual.data_ptr(), Cout
    d.C, d.ldc, d.c_is_f32 = out.data_ptr(), Cout, 0
    d.n_split = 1 << 30
    _capi.check(lib.rmem_gemm_fwd(C.byref(d), _capi.stream_ptr()))
    return out


def stem_conv(img: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, act: int = ACT_RELU) -> torch.Tensor:
    """conv1 of the ResNet stem (7x7, stride 2, pad 3; resnet.py:178-181) on the tcgen05 GEMM: img fp32 [1,3,H,W] is
    packed into the zero-padded NHWC8 layout, w t16 [Cout,7,8,8] (weights.py "enc.conv1.w") -> t16 [H1,W1,Cout]."""
    lib = _capi.load()
    H, W = int(img.shape[-2]), int(img.shape[-1])
    H1, W1 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    Cout = w.shape[0]
    pad = torch.zeros(H + 6, W + 8, 8, dtype=_capi.op_dtype(), device=img.device)
    _capi.check(lib.rmem_pack_image_padded_fwd(_capi.ptr(img.contiguous()), _capi.ptr(pad), H, W, _capi.stream_ptr()))
    out = torch.empty(H1, W1, Cout, dtype=_capi.op_dtype(), device=img.device)
    d = _capi.GemmDesc()
    d.A, d.B, d.ldb = pad.data_ptr(), w.data_ptr(), 7 * 64
    d.M, d.N, d.K = H1 * W1, Cout, 7 * 64
    d.conv, d.Hin, d.Win, d.Cin, d.Wout, d.kw, d.stride, d.pad = 2, H + 6, W + 8, 8, W1, 7, 2, 3
    d.alpha = 1.0
    d.bias = bias.data_ptr()
    d.act = act
    d.C, d.ldc, d.c_is_f32 = out.data_ptr(), Cout, 0
    d.n_split = 1 << 30
    _capi.check(lib.rmem_gemm_fwd(C.byref(d), _capi.stream_ptr()))
    return out


def layernorm(x: torch.Tensor, g: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    lib = _capi.load()
    P, Cc = x.shape
    y = torch.empty(P, Cc, dtype=_capi.op_dtype(), device=x.device)
    _capi.check(lib.rmem_layernorm_fwd(_capi.ptr(x), C.c_longlong(x.stride(0)), _capi.ptr(g), _capi.ptr(b),
                                       _capi.ptr(y), C.c_longlong(Cc), P, Cc, _capi.stream_ptr()))
    return y


def groupnorm(x: torch.Tensor, g: torch.Tensor, b: torch.Tensor, groups: int, relu: bool) -> torch.Tensor:
    lib = _capi.load()
    P, Cc = x.shape
    y = torch.empty(P, Cc, dtype=_capi.op_dtype(), device=x.device)
    stats = torch.zeros(72 + 148 * 4 * 64, dtype=torch.float64, device=x.device)
    _capi.check(lib.rmem_groupnorm_fwd(_capi.ptr(x), int(x.dtype == torch.float32), _capi.ptr(g), _capi.ptr(b),
                                       _capi.ptr(y), P, Cc, groups, int(relu), _capi.ptr(stats), _capi.stream_ptr()))
    return y


def dwconv5x5(x: torch.Tensor, w25: torch.Tensor, h: int, w: int) -> torch.Tensor:
    lib = _capi.load()
    y = torch.empty_like(x)
    _capi.check(lib.rmem_dwconv5x5_fwd(_capi.ptr(x), _capi.ptr(w25), _capi.ptr(y), h, w, x.shape[1],
                                       _capi.stream_ptr()))
    return y


def upsample_bilinear(x: torch.Tensor, hout: int, wout: int) -> torch.Tensor:
    lib = _capi.load()
    hin, win, Cc = x.shape
    y = torch.empty(hout, wout, Cc, dtype=_capi.op_dtype(), device=x.device)
    _capi.check(lib.rmem_upsample_bilinear_fwd(_capi.ptr(x), _capi.ptr(y), hin, win, hout, wout, Cc,
                                               _capi.stream_ptr()))
    return y


def maxpool3x3s2(x: torch.Tensor) -> torch.Tensor:
    lib = _capi.load()
    Hin, Win, Cc = x.shape
    Hout, Wout = (Hin - 1) // 2 + 1, (Win - 1) // 2 + 1
    y = torch.empty(Hout, Wout, Cc, dtype=_capi.op_dtype(), device=x.device)
    _capi.check(lib.rmem_maxpool3x3s2_fwd(_capi.ptr(x), _capi.ptr(y), Hin, Win, Cc, Hout, Wout, _capi.stream_ptr()))
    return y


def transpose(x: torch.Tensor, ldy: int) -> torch.Tensor:
    lib = _capi.load()
    P, Cc = x.shape
    y = torch.zeros(Cc, ldy, dtype=_capi.op_dtype(), device=x.device)
    _capi.check(lib.rmem_transpose_fwd(_capi.ptr(x), C.c_longlong(x.stride(0)), _capi.ptr(y), C.c_longlong(ldy), P, Cc,
                                       _capi.stream_ptr()))
    return y


def id_embedding(label_u8: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, ln_g, ln_b,
                 use_ignore: bool, prefix: Optional[torch.Tensor] = None,
                 prefix_rows: Optional[torch.Tensor] = None) -> torch.Tensor:
    """label uint8 [H,W] -> fp32 [hw, C].  prefix / prefix_rows: optional [12,18,18,C] / [17,12,18,C] tables (weights.py)
    enabling the rectangle / dominant-class / row-run decompositions of the per-pixel gather."""
    lib = _capi.load()
    H, W = label_u8.shape
    h, w = (H - 1) // 16 + 1, (W - 1) // 16 + 1
    Cc = bias.numel()
    out = torch.empty(h * w, Cc, dtype=torch.float32, device=label_u8.device)
    _capi.check(lib.rmem_idbank_fwd(_capi.ptr(label_u8), H, W, int(use_ignore), _capi.ptr(w_packed),
                                    _capi.ptr(prefix), _capi.ptr(prefix_rows), _capi.ptr(bias),
                                    _capi.ptr(ln_g), _capi.ptr(ln_b), None, C.c_longlong(0), _capi.ptr(out), h, w, Cc,
                                    _capi.stream_ptr()))
    return out


def mask_head(logits4: Sequence[torch.Tensor], Ho: int, Wo: int, want_logits: bool = True):
    """k x fp32 [11,h4,w4] -> (logits fp32 [1+10k,Ho,Wo] | None, label uint8 [Ho,Wo])."""
    lib = _capi.load()
    k = len(logits4)
    _, h4, w4 = logits4[0].shape
    ptrs = (C.c_void_p * k)(*[t.data_ptr() for t in logits4])
    dev = logits4[0].device
    out = torch.empty(1 + 10 * k, Ho, Wo, dtype=torch.float32, device=dev) if want_logits else None
    lab = torch.empty(Ho, Wo, dtype=torch.uint8, device=dev)
    _capi.check(lib.rmem_mask_head_fwd(ptrs, k, h4, w4, Ho, Wo, _capi.ptr(out), _capi.ptr(lab), _capi.stream_ptr()))
    return out, lab


def tta_head(logits4: Sequence[Sequence[torch.Tensor]], flips: Sequence[bool], Ho: int, Wo: int, want_prob: bool = False):
    """Test-time-augmentation head.  logits4[a][e]: fp32 [11,h4_a,w4_a] of augmentation a, object group e.
    Returns (prob fp32 [1+10k,Ho,Wo] | None, label uint8 [Ho,Wo])."""
    lib = _capi.load()
    n_aug, k = len(logits4), len(logits4[0])
    flat = [t for per in logits4 for t in per]
    ptrs = (C.c_void_p * len(flat))(*[t.data_ptr() for t in flat])
    h4 = (C.c_int * n_aug)(*[int(per[0].shape[-2]) for per in logits4])
    w4 = (C.c_int * n_aug)(*[int(per[0].shape[-1]) for per in logits4])
    fl = (C.c_int * n_aug)(*[int(bool(f)) for f in flips])
    dev = flat[0].device
    prob = torch.empty(1 + 10 * k, Ho, Wo, dtype=torch.float32, device=dev) if want_prob else None
    lab = torch.empty(Ho, Wo, dtype=torch.uint8, device=dev)
    _capi.check(lib.rmem_tta_head_fwd(ptrs, n_aug, k, h4, w4, fl, Ho, Wo, _capi.ptr(prob), _capi.ptr(lab),
                                      _capi.stream_ptr()))
    return prob, lab


def preprocess(img_u8: torch.Tensor, nh: int, nw: int, bgr: bool = True, flip: bool = False) -> torch.Tensor:
    """uint8 [H,W,3] device frame -> normalised fp32 [1,3,nh,nw] (MultiRestrictSize + MultiToTensor on the GPU)."""
    lib = _capi.load()
    H, W, _ = img_u8.shape
    out = torch.empty(1, 3, nh, nw, dtype=torch.float32, device=img_u8.device)
    _capi.check(lib.rmem_preprocess_fwd(_capi.ptr(img_u8.contiguous()), H, W, int(bgr), nh, nw, int(flip), _capi.ptr(out),
                                        _capi.stream_ptr()))
    return out


def evict_relevance(mass: torch.Tensor, logits4: torch.Tensor, h: int, w: int) -> torch.Tensor:
    lib = _capi.load()
    T = mass.shape[1]
    _, h4, w4 = logits4.shape
    rel = torch.empty(T, dtype=torch.float32, device=mass.device)
    _capi.check(lib.rmem_evict_relevance_fwd(_capi.ptr(mass), T, _capi.ptr(logits4), h4, w4, h, w, _capi.ptr(rel),
                                             _capi.stream_ptr()))
    return rel


def evict_pick(rel: Sequence[float], idx: Sequence[int], former: int, ema: dict, times: dict) -> int:
    """Host EMA/UCB/argmin (transformer.py:907-964).  Updates `ema` / `times` in place, returns drop index."""
    lib = _capi.load()
    T_old = len(rel)
    cap = _capi.MAX_BANK_FRAMES + 1
    relc = (C.c_float * T_old)(*rel)
    idxc = (C.c_int * len(idx))(*idx)
    ek, ev = (C.c_int * cap)(*ema.keys()), (C.c_float * cap)(*ema.values())
    tk, tv = (C.c_int * cap)(*times.keys()), (C.c_int * cap)(*times.values())
    ne, nt, drop = C.c_int(len(ema)), C.c_int(len(times)), C.c_int(-1)
    _capi.check(lib.rmem_evict_pick(relc, T_old, idxc, former, ek, ev, C.byref(ne), tk, tv, C.byref(nt),
                                    C.byref(drop)))
    ema.clear(); times.clear()
    for i in range(ne.value):
        ema[ek[i]] = ev[i]
    for i in range(nt.value):
        times[tk[i]] = tv[i]
    return drop.value


def temporal_pe_slots(T: int, n_slots: int = 4) -> List[int]:
    lib = _capi.load()
    out = (C.c_int * T)()
    _capi.check(lib.rmem_temporal_pe_slots(T, n_slots, out))
    return list(out)


def build_bank(k_frames: torch.Tensor, v_frames: torch.Tensor, nslots: int, slots: Sequence[int]):
    """Lay T frames out the way the engine's ring bank does.  k [T,HW,Dk], v [T,HW,Dv] (any float dtype) ->
    (kbank t16 [nslots,HWp,Dk], vtbank t16 [Dv, nslots*HWp], HWp)."""
    T, HW, Dk = k_frames.shape
    Dv = v_frames.shape[-1]
    HWp = round_up(HW, 128)
    dev = k_frames.device
    kbank = torch.zeros(nslots, HWp, Dk, dtype=_capi.op_dtype(), device=dev)
    vtbank = torch.zeros(Dv, nslots * HWp, dtype=_capi.op_dtype(), device=dev)
    lib = _capi.load()
    for t, s in enumerate(slots):
        kbank[s, :HW] = k_frames[t].to(_capi.op_dtype())
        vt = _bf(v_frames[t])
        _capi.check(lib.rmem_transpose_fwd(_capi.ptr(vt), C.c_longlong(Dv), C.c_void_p(vtbank.data_ptr() + 2 * s * HWp),
                                           C.c_longlong(nslots * HWp), HW, Dv, _capi.stream_ptr()))
    return kbank, vtbank, HWp


def long_attention(q: torch.Tensor, kbank: torch.Tensor, vtbank: torch.Tensor, slots: Sequence[int], HW: int,
                   pe_cur: Optional[torch.Tensor] = None, mem_pos_emb: Optional[torch.Tensor] = None,
                   gate: Optional[torch.Tensor] = None, impl: int = _capi.ATTN_DENSE,
                   want_mass: bool = True, grid: Optional[Tuple[int, int]] = None
                   ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """q t16 [HW,Dk] (no PE, unscaled).  Returns (out t16 [HW,Dv], mass fp32 [HW,T]).
    grid = (h, w) token grid (h*w == HW): lets ATTN_TC3 seed the row maximum from the query's own neighbourhood."""
    lib = _capi.load()
    nslots, HWp, Dk = kbank.shape
    Dv = vtbank.shape[0]
    T = len(slots)
    dev = q.device
    scale = 1.0 / math.sqrt(Dk)
    qt = torch.empty(HW, Dk, dtype=_capi.op_dtype(), device=dev)
    qbias = torch.zeros(HW, max(T, 1), dtype=torch.float32, device=dev)
    if mem_pos_emb is not None:
        pes = (C.c_int * T)(*temporal_pe_slots(T, mem_pos_emb.shape[0]))
        _capi.check(lib.rmem_qprep_fwd(_capi.ptr(q), C.c_longlong(q.stride(0)), _capi.ptr(pe_cur),
                                       _capi.ptr(mem_pos_emb), pes, T, C.c_float(scale), _capi.ptr(qt),
                                       _capi.ptr(qbias), HW, Dk, _capi.stream_ptr()))
    else:
        qt.copy_(q)
    nbytes = C.c_size_t()
    _capi.check(lib.rmem_long_attn_workspace_bytes(impl, HW, HWp, nslots, Dv, C.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    out = torch.empty(HW, Dv, dtype=_capi.op_dtype(), device=dev)
    mass = torch.empty(HW, T, dtype=torch.float32, device=dev) if want_mass else None
    sl = (C.c_int * T)(*slots)
    gh, gw = grid if grid is not None else (0, 0)
    _capi.check(lib.rmem_long_attn_grid_fwd(impl, _capi.ptr(qt), _capi.ptr(qbias) if mem_pos_emb is not None else None,
                                            _capi.ptr(kbank), _capi.ptr(vtbank), nslots, T, sl, HW, HWp, Dk, Dv,
                                            C.c_float(scale), _capi.ptr(gate),
                                            C.c_longlong(gate.stride(0) if gate is not None else 0),
                                            _capi.ptr(out), C.c_longlong(Dv), _capi.ptr(mass), int(gh), int(gw),
                                            _capi.ptr(ws), C.c_size_t(nbytes.value), _capi.stream_ptr()))
    global last_attn_overflow
    # ATTN_TC4: first int of the workspace = the column kernel's overflow flag (1: the guarded tc3 fallback produced `out`)
    last_attn_overflow = int(ws[:4].view(torch.int32).item()) if (impl == _capi.ATTN_TC4 and T >= 3 and grid is not None) else None
    return out, mass


last_attn_overflow = None


def local_attention(q: torch.Tensor, k_prev: torch.Tensor, v_prev: torch.Tensor, rel_w: torch.Tensor,
                    rel_b: torch.Tensor, h: int, w: int, gate: Optional[torch.Tensor] = None,
                    impl: str = "tc", rel_pitch: int = 16) -> torch.Tensor:
    """q,k t16 [HW,128]; v t16 [HW,Dv]; rel_w fp32/t16 [225,128]; rel_b [225] -> t16 [HW,Dv].
    impl "tc" lays the 225 offsets out as 15 window rows of `rel_pitch` floats (16 = what the engine's packed weights
    use, 15 = the reference order); the CUDA-core kernel always reads the reference order."""
    lib = _capi.load()
    HW, Dk = q.shape
    Dv = v_prev.shape[1]
    pitch = rel_pitch if impl == "tc" else 15
    wpad = torch.zeros(256, Dk, dtype=_capi.op_dtype(), device=q.device)
    bpad = torch.zeros(256, dtype=torch.float32, device=q.device)
    for dy in range(15):
        wpad[dy * pitch: dy * pitch + 15] = rel_w[dy * 15: dy * 15 + 15].to(_capi.op_dtype())
        bpad[dy * pitch: dy * pitch + 15] = rel_b[dy * 15: dy * 15 + 15].float()
    rel = gemm(q, wpad, bpad, out_f32=True)
    out = torch.empty(HW, Dv, dtype=_capi.op_dtype(), device=q.device)
    if impl == "tc":
        nbytes = C.c_size_t()
        _capi.check(lib.rmem_local_attn_tc_workspace_bytes(h, w, Dv, C.byref(nbytes)))
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=q.device)
        _capi.check(lib.rmem_local_attn_tc_fwd(
            _capi.ptr(q), C.c_longlong(q.stride(0)), _capi.ptr(k_prev), C.c_longlong(k_prev.stride(0)),
            _capi.ptr(v_prev), C.c_longlong(v_prev.stride(0)), _capi.ptr(rel), C.c_longlong(256), int(pitch),
            _capi.ptr(gate), C.c_longlong(gate.stride(0) if gate is not None else 0), _capi.ptr(out), C.c_longlong(Dv),
            h, w, Dv, C.c_float(1.0 / math.sqrt(Dk)), _capi.ptr(ws), C.c_size_t(nbytes.value), _capi.stream_ptr()))
        return out
    _capi.check(lib.rmem_local_attn_fwd(_capi.ptr(q), C.c_longlong(q.stride(0)), _capi.ptr(k_prev),
                                        C.c_longlong(k_prev.stride(0)), _capi.ptr(v_prev),
                                        C.c_longlong(v_prev.stride(0)), _capi.ptr(rel), C.c_longlong(256),
                                        _capi.ptr(gate), C.c_longlong(gate.stride(0) if gate is not None else 0),
                                        _capi.ptr(out), C.c_longlong(Dv), h, w, Dv, C.c_float(1.0 / math.sqrt(Dk)),
                                        _capi.stream_ptr()))
    return out


def multihead_attention(q: torch.Tensor, kbank: torch.Tensor, vtbank: torch.Tensor, slots: Sequence[int], HW: int,
                        n_head: int = 8, qbias: Optional[torch.Tensor] = None, impl: int = _capi.ATTN_TC3,
                        want_mass: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """AOT MultiheadAttention over the bank (attention.py:28-81): q t16 [HW, C] (any PE already added), kbank
    [nslots, HWp, C], vtbank [C, nslots*HWp], qbias fp32 [H, HW, T] (already scaled) or None.
    Returns (out t16 [HW, C], mass fp32 [HW, T] = head mean of the per-frame probability mass)."""
    lib = _capi.load()
    nslots, HWp, Cc = kbank.shape
    T = len(slots)
    dh = Cc // n_head
    dev = q.device
    scale = 1.0 / math.sqrt(dh)
    nbytes = C.c_size_t()
    _capi.check(lib.rmem_mha_workspace_bytes(impl, HW, HWp, nslots, n_head, C.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    out = torch.empty(HW, Cc, dtype=_capi.op_dtype(), device=dev)
    mass = torch.empty(HW, T, dtype=torch.float32, device=dev) if want_mass else None
    sl = (C.c_int * T)(*slots)
    _capi.check(lib.rmem_mha_fwd(impl, _capi.ptr(q), C.c_longlong(q.stride(0)), _capi.ptr(kbank), _capi.ptr(vtbank),
                                 nslots, T, sl, HW, HWp, n_head, dh, C.c_float(scale), _capi.ptr(qbias),
                                 _capi.ptr(out), C.c_longlong(Cc), _capi.ptr(mass), _capi.ptr(ws),
                                 C.c_size_t(nbytes.value), _capi.stream_ptr()))
    return out, mass
