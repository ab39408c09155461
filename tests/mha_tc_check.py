"""Stand-alone check of the fused tcgen05 multi-head attention (rmem_b200/csrc/mha_tc.cu, AOT 8 heads x 32) against a plain
torch fp32 statement of MultiheadAttention over the bank (reference: networks/layers/attention.py:28-81 and the per-frame
mass of transformer.py:636-643) and against the materialised-score CUDA path.  Prints one JSON line per case; run by
tests/test_mha_tc_gpu.py in a subprocess (a protocol bug traps the kernel instead of hanging pytest)."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rmem_b200 import _capi, ops  # noqa: E402


def reference(q, k_frames, v_frames, bias, H):
    """q [HW,C] float, k_frames / v_frames [T,HW,C] float, bias [H,HW,T] or None -> (out [HW,C], mass [HW,T])"""
    HW, C = q.shape
    T = k_frames.shape[0]
    dh = C // H
    qh = q.view(HW, H, dh).permute(1, 0, 2)                        # [H,HW,dh]
    kh = k_frames.reshape(T * HW, H, dh).permute(1, 0, 2)          # [H,T*HW,dh]
    vh = v_frames.reshape(T * HW, H, dh).permute(1, 0, 2)
    s = torch.matmul(qh, kh.transpose(1, 2)) / math.sqrt(dh)      # [H,HW,T*HW]
    if bias is not None:
        s = s + bias.repeat_interleave(HW, dim=2)
    p = torch.softmax(s, dim=-1)
    out = torch.matmul(p, vh).permute(1, 0, 2).reshape(HW, C)
    mass = p.view(H, HW, T, HW).sum(-1).mean(0)
    return out, mass


def main():
    dev = torch.device("cuda:0")
    dt = _capi.op_dtype()
    H, C = 8, 256
    cases = [  # HW, T, nslots, slots, bias, sharpen
        (99, 1, 1, [0], False, 1.0),
        (300, 3, 5, [3, 0, 4], True, 1.0),
        (1620, 4, 5, [0, 2, 3, 1], True, 4.0),
        (1674, 8, 9, [0, 5, 6, 7, 8, 1, 2, 3], True, 6.0),
        (200, 16, 16, list(range(15, -1, -1)), True, 2.0),
        (1620, 1, 1, [0], False, 3.0),
    ]
    for ci, (HW, T, nslots, slots, use_bias, sharp) in enumerate(cases):
        g = torch.Generator(device="cpu").manual_seed(100 + ci)
        q = (torch.randn(HW, C, generator=g) * sharp).to(dev).to(dt)
        kf = torch.randn(T, HW, C, generator=g).to(dev).to(dt)
        vf = torch.randn(T, HW, C, generator=g).to(dev).to(dt)
        bias = (torch.randn(H, HW, T, generator=g) * 2.0).to(dev) if use_bias else None
        kbank, vtbank, HWp = ops.build_bank(kf, vf, nslots, slots)
        ref_out, ref_mass = reference(q.float(), kf.float(), vf.float(), bias, H)
        rec = {"case": ci, "HW": HW, "T": T}
        out, mass = ops.multihead_attention(q, kbank, vtbank, slots, HW, H, bias, impl=_capi.ATTN_TC3)
        torch.cuda.synchronize()
        rec["finite"] = bool(torch.isfinite(out.float()).all() and torch.isfinite(mass).all())
        rec["tc_vs_ref"] = float((out.float() - ref_out).norm() / ref_out.norm())
        rec["mass_err"] = float((mass - ref_mass).abs().max())
        rec["mass_sum_err"] = float((mass.sum(1) - 1).abs().max())
        if HW * T <= 8000:
            d_out, d_mass = ops.multihead_attention(q, kbank, vtbank, slots, HW, H, bias, impl=_capi.ATTN_DENSE)
            rec["tc_vs_dense"] = float((out.float() - d_out.float()).norm() / d_out.float().norm())
        # second call on the same operands: bit-identical (static schedule, no atomics)
        out2, _ = ops.multihead_attention(q, kbank, vtbank, slots, HW, H, bias, impl=_capi.ATTN_TC3)
        rec["deterministic"] = bool(torch.equal(out, out2))
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
