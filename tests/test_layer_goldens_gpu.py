"""Layer- and op-level parity against vectors recorded from the UNMODIFIED reference with forward hooks
(oracle/make_layer_golden.py -> tests/golden/layer_tiny.npz): R50_DeAOTL + RMem, 129x161, 3 objects, bank 1 + 3.

  layer level  after replaying the clip (teacher-forced with the reference's own labels) the engine's per-layer memories
               -- short-term K (= Q) and V of the last frame, the ID_V written by update_memory, the restricted bank's K / V /
               ID_V in logical order -- equal what GatedPropagationModule.forward returned / LSTT.long_term_memories hold
               (transformer.py:1091-1236, 993-1007).  These are the tensors every later layer and frame is built on.
  op level     GatedPropagation.forward (attention.py:140-213) and LocalGatedPropagation.forward (attention.py:289-361)
               of layer 1 on the recorded inputs: fused attention kernel (TC3 / TC2 / dense) or windowed attention kernel
               -> gate -> depthwise 5x5 -> projection, against the recorded module output.
Tolerances (fp16 operands, fp32 accumulate vs the fp32 reference): rel-Frobenius 4e-3 on memories, 6e-3 on module outputs."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import rmem_oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def relfro(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(HERE, "golden", "layer_tiny.npz"))
    return z, json.loads(str(z["meta"]))


def test_layer_memories_match_reference_hooks(cuda_device, gold):
    from rmem_b200.engine import DeAOTModel, RmemConfig, build_engine
    z, m = gold
    sd = O.make_state_dict(m["model"], seed=m["seed"], sharpen=m["sharpen"])
    frames = O.synthetic_frames(m["n_frames"], m["H"], m["W"], seed=m["seed"] + 1).to(cuda_device)
    label0 = O.synthetic_label(m["H"], m["W"], m["n_obj"])
    labels = torch.from_numpy(z["labels"])
    cfg = RmemConfig(former_mem_len=m["former"], latter_mem_len=m["latter"], max_engines=1)
    eng = build_engine("deaotengine", aot_model=DeAOTModel(sd, cfg, cuda_device), long_term_mem_gap=m["gap"])
    eng.restart_engine()
    eng.add_reference_frame(frames[0:1], label0.int().to(cuda_device), obj_nums=[m["n_obj"]], frame_step=0)
    for f in range(1, m["n_frames"]):
        eng.match_propogate_one_frame(frames[f:f + 1], output_size=(m["H"], m["W"]))
        eng.update_memory(labels[f - 1].view(1, 1, m["H"], m["W"]).to(cuda_device))
    sub = eng.aot_engines[0]
    assert sub.long_memories_indexes == m["idx"]
    HW = 9 * 11
    worst = {}
    for l in range(3):
        mem = sub.layer_memory(l)
        worst[f"l{l}.K"] = relfro(mem["q_last"], torch.from_numpy(z[f"l{l}.curr_K"]))
        worst[f"l{l}.V"] = relfro(mem["vid_last"][:, :512], torch.from_numpy(z[f"l{l}.curr_V"]))
        worst[f"l{l}.ID_V"] = relfro(mem["vid_last"][:, 512:], torch.from_numpy(z[f"short{l}.ID_V"].astype(np.float32)))
        kb = mem["kbank"][mem["slots"], :HW]                                    # logical frame order
        worst[f"l{l}.bankK"] = relfro(kb, torch.from_numpy(z[f"bank{l}.K"].astype(np.float32)))
        if l == 1:
            T, HWp = len(mem["slots"]), mem["HWp"]
            vt = mem["vtbank"].view(1024, mem["nslots"], HWp)[:, mem["slots"], :HW]        # [1024, T, HW]
            ref = torch.cat([torch.from_numpy(z["bank1.V"].astype(np.float32)),
                             torch.from_numpy(z["bank1.ID_V"].astype(np.float32))], dim=2)  # [T, HW, 1024]
            worst["l1.bankV"] = relfro(vt.permute(1, 2, 0), ref)
    logit_err = float((sub.pred_id_logits[0].cpu() - torch.from_numpy(z["final_logits4"])).abs().max() /
                      np.abs(z["final_logits4"]).max())
    print(json.dumps(dict(worst, logit_err=logit_err)))
    from parity_report import report
    report("layer/layer_tiny/memories", **worst, final_logit_rel_err=logit_err, tolerance=4e-3, vs="reference forward hooks")
    assert max(worst.values()) < 4e-3, worst
    assert logit_err < 1.5e-2


@pytest.mark.parametrize("impl", [4, 3, 2, 0], ids=["tc4", "tc3", "tc2", "dense"])
def test_gated_propagation_matches_reference_module(cuda_device, gold, impl):
    """long_term_attn of layer 1: out = projection(DWConv5x5((softmax(Q K^T / sqrt(128)) V) * U))."""
    from rmem_b200 import _capi, ops as K
    z, m = gold
    OP = _capi.op_dtype()
    sd = O.make_state_dict(m["model"], seed=m["seed"], sharpen=m["sharpen"])
    HW, h, w = 99, 9, 11
    Q = torch.from_numpy(z["long.Q"]).to(cuda_device)
    Kb = torch.from_numpy(z["long.K"].astype(np.float32)).view(-1, HW, 128).to(cuda_device)       # [T, HW, 128] (PE added)
    Vb = torch.from_numpy(z["long.V"].astype(np.float32)).view(-1, HW, 1024).to(cuda_device)
    U = torch.from_numpy(z["long.U"].astype(np.float32)).to(cuda_device).to(OP)
    T = Kb.shape[0]
    slots = list(range(T))
    kb, vtb, HWp = K.build_bank(Kb, Vb, T + 1, slots)
    agg, _ = K.long_attention(Q.to(OP), kb, vtb, slots, HW, gate=U, impl=impl, grid=(h, w) if impl in (3, 4) else None)
    p = "LSTT.layers.1.long_term_attn"
    dw = sd[p + ".dw_conv.conv.weight"].view(1024, 25).t().contiguous().to(cuda_device)
    x = K.dwconv5x5(agg, dw, h, w)
    out = K.gemm(x, sd[p + ".projection.weight"].to(cuda_device).to(OP), sd[p + ".projection.bias"].to(cuda_device),
                 out_f32=True)
    err = relfro(out, torch.from_numpy(z["long.out"]))
    print(f"GatedPropagation impl {impl}: rel-Frobenius {err:.3e}")
    from parity_report import report
    report(f"op/GatedPropagation/attn_impl_{impl}", rel_frobenius=err, tolerance=6e-3, vs="reference forward hooks")
    assert err < 6e-3


def test_local_gated_propagation_matches_reference_module(cuda_device, gold):
    """short_term_attn of layer 1 (15x15 window, relative_emb_k bias, zero-padded keys)."""
    from rmem_b200 import _capi, ops as K
    z, m = gold
    OP = _capi.op_dtype()
    sd = O.make_state_dict(m["model"], seed=m["seed"], sharpen=m["sharpen"])
    HW, h, w = 99, 9, 11
    q = torch.from_numpy(z["short.q"]).flatten(1).t().contiguous().to(cuda_device).to(OP)         # [HW,128]
    k = torch.from_numpy(z["short.k"]).flatten(1).t().contiguous().to(cuda_device).to(OP)
    v = torch.from_numpy(z["short.v"].astype(np.float32)).flatten(1).t().contiguous().to(cuda_device).to(OP)
    u = torch.from_numpy(z["short.u"].astype(np.float32)).to(cuda_device).to(OP)
    p = "LSTT.layers.1.short_term_attn"
    rel_w = sd[p + ".relative_emb_k.weight"].view(225, 128).to(cuda_device)
    rel_b = sd[p + ".relative_emb_k.bias"].to(cuda_device)
    dw = sd[p + ".dw_conv.conv.weight"].view(1024, 25).t().contiguous().to(cuda_device)
    for impl in ("tc", "cuda"):
        agg = K.local_attention(q, k, v, rel_w, rel_b, h, w, gate=u, impl=impl)
        x = K.dwconv5x5(agg, dw, h, w)
        out = K.gemm(x, sd[p + ".projection.weight"].to(cuda_device).to(OP), sd[p + ".projection.bias"].to(cuda_device),
                     out_f32=True)
        err = relfro(out, torch.from_numpy(z["short.out"]))
        print(f"LocalGatedPropagation impl {impl}: rel-Frobenius {err:.3e}")
        from parity_report import report
        report(f"op/LocalGatedPropagation/{impl}", rel_frobenius=err, tolerance=6e-3, vs="reference forward hooks")
        assert err < 6e-3
