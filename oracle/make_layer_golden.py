"""Layer- and op-level golden vectors recorded from the UNMODIFIED reference with forward hooks
--  TEST INFRASTRUCTURE (build container only; needs /root/reference).

    python oracle/make_layer_golden.py

oracle/make_golden.py pins frames and clips; this script records what the reference computes INSIDE one propagated frame
of a small R50_DeAOTL + RMem clip (129x161 -> 9x11 tokens, 3 objects, bank 1 + 3 full), so that the layer- and op-level
GPU tests are pinned to the reference itself and not only to the oracle:

  GatedPropagationModule.forward   (networks/layers/transformer.py:1091-1236), every layer: the returned tgt / tgt_id
                                   and the layer's curr memories (K, V) -> engine's per-layer short-term memory
  GatedPropagation.forward         (networks/layers/attention.py:140-213) = long_term_attn of layer 1: Q (+ temporal PE),
                                   flattened bank K (+ PE) / V||ID_V, gate U -> projected output
  LocalGatedPropagation.forward    (attention.py:289-361) = short_term_attn of layer 1: q, k, v, u -> projected output
  LSTT.long_term_memories          after the last update: bank K / V / ID_V per layer -> engine's ring bank

Output: tests/golden/layer_tiny.npz (fp16 for the large activations, fp32 for outputs), plus the clip's reference labels
for teacher forcing.  tests/test_layer_goldens_gpu.py consumes it.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import rmem_oracle as O  # noqa: E402
from oracle.make_golden import build_reference  # noqa: E402

CASE = dict(model="r50_deaotl", seed=11, sharpen=4.0, H=129, W=161, n_obj=3, n_frames=8, former=1, latter=3, gap=1)


def main():
    torch.set_num_threads(4)
    c = CASE
    sd = O.make_state_dict(c["model"], seed=c["seed"], sharpen=c["sharpen"])
    frames = O.synthetic_frames(c["n_frames"], c["H"], c["W"], seed=c["seed"] + 1)
    label0 = O.synthetic_label(c["H"], c["W"], c["n_obj"])
    net, eng = build_reference(c["model"], sd, c["former"], c["latter"], c["gap"])
    H, W = c["H"], c["W"]
    rec = {}
    on = {"v": False}

    def hook_layer(l):
        def fn(mod, inp, out):
            if on["v"]:
                tgt, tgt_id, mems = out
                rec[f"l{l}.tgt"] = tgt[:, 0].clone()
                rec[f"l{l}.tgt_id"] = tgt_id[:, 0].clone()
                rec[f"l{l}.curr_K"] = mems[0][0][:, 0].clone()
                rec[f"l{l}.curr_V"] = mems[0][1][:, 0].clone()
        return fn

    def hook_long(mod, inp, out):
        if on["v"]:
            Q, K, V, U = inp[0], inp[1], inp[2], inp[3]
            rec["long.Q"], rec["long.K"], rec["long.V"], rec["long.U"] = Q[:, 0].clone(), K[:, 0].clone(), V[:, 0].clone(), U[:, 0].clone()
            rec["long.out"] = out[0][:, 0].clone()

    def hook_short(mod, inp, out):
        if on["v"]:
            q, k, v, u = inp[0], inp[1], inp[2], inp[3]
            rec["short.q"], rec["short.k"], rec["short.v"] = q[0].clone(), k[0].clone(), v[0].clone()      # [C,h,w]
            rec["short.u"] = u[:, 0].clone()
            rec["short.out"] = out[0][:, 0].clone()

    layers = net.LSTT.layers
    handles = [layers[l].register_forward_hook(hook_layer(l)) for l in range(3)]
    handles.append(layers[1].long_term_attn.register_forward_hook(hook_long))
    handles.append(layers[1].short_term_attn.register_forward_hook(hook_short))
    labels = []
    sink = io.StringIO()
    with torch.no_grad(), contextlib.redirect_stdout(sink):
        eng.restart_engine()
        eng.add_reference_frame(frames[0:1], label0.int(), obj_nums=[c["n_obj"]], frame_step=0)
        for f in range(1, c["n_frames"]):
            on["v"] = f == c["n_frames"] - 1
            logit = eng.match_propogate_one_frame(frames[f:f + 1], output_size=(H, W))
            on["v"] = False
            lab = torch.argmax(torch.softmax(logit, dim=1), dim=1, keepdim=True).float()
            eng.update_memory(lab)
            labels.append(lab[0, 0].to(torch.uint8))
        rec["final_logits4"] = eng.aot_engines[0].pred_id_logits[0].clone()
        lstt = eng.aot_engines[0].AOT.LSTT
        for l in range(3):
            mem = lstt.long_term_memories[l]           # [K [T,HW,1,128], V [T,HW,1,512], None, ID_V [T,HW,1,512]]
            rec[f"bank{l}.K"] = mem[0][:, :, 0].clone()
            if l == 1:                                  # values of one layer only (fixture size)
                rec[f"bank{l}.V"] = mem[1][:, :, 0].clone()
                rec[f"bank{l}.ID_V"] = mem[3][:, :, 0].clone()
            sm = lstt.short_term_memories[l]           # after the update: [K [1,128,h,w], V [1,512,h,w], None, ID_V]
            rec[f"short{l}.ID_V"] = sm[3][0].flatten(1).t().clone()
    for h in handles:
        h.remove()
    idx = list(eng.aot_engines[0].long_memories_indexes)
    big = ("long.K", "long.V", "long.U", "short.v", "short.u", "bank", "short")
    arrays = {}
    for k, v in rec.items():
        a = v.detach().float().numpy()
        arrays[k] = a.astype(np.float16) if (k.startswith(big) and not k.endswith(".out")) else a.astype(np.float32)
    out = os.path.join(ROOT, "tests", "golden", "layer_tiny.npz")
    np.savez_compressed(out, meta=json.dumps(dict(CASE, idx=idx, reference_commit="431cde18", torch=torch.__version__)),
                        labels=torch.stack(labels).numpy(), **arrays)
    print("wrote", out, f"{os.path.getsize(out) / 1e6:.2f} MB", "bank idx", idx, {k: tuple(v.shape) for k, v in arrays.items()})


if __name__ == "__main__":
    main()
