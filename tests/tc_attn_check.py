"""Stand-alone check of the fused tcgen05 attention kernel against the dense path and the CPU oracle.
Run in its own process (tests/test_attn_tc_gpu.py spawns it): the kernel's deadlock watchdog traps, which
poisons the CUDA context, and that must not take the rest of the GPU test session with it.
Prints one JSON line per case."""
import json
import math
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import rmem_oracle as O  # noqa: E402
from rmem_b200 import _capi, ops as K  # noqa: E402

OP = _capi.op_dtype()
IMPL = int(os.environ.get("RMEM_ATTN_IMPL", str(_capi.ATTN_TC3)))
SEED = os.environ.get("RMEM_ATTN_SEED", "1") != "0"      # pass the token grid (tc3: seeded row maximum)
GRIDS = {289: (17, 17), 357: (17, 21), 1674: (31, 54), 3726: (46, 81), 70: (7, 10)}


def bfr(t):
    return t.to(OP).float()


def relfro(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def main():
    dev = torch.device("cuda:0")
    cases = [
        # name, T, HW, slots, sharp, use_pe, use_gate
        ("t1_small", 1, 289, [0], 1.0, True, True),
        ("t3_ragged", 3, 357, [2, 0, 3], 2.0, True, True),
        ("self_like", 1, 1674, [0], 1.0, False, True),
        ("c3_t8", 8, 1674, [0, 5, 1, 2, 8, 3, 4, 6], 2.0, True, True),
        ("c3_t8_sharp", 8, 1674, [0, 5, 1, 2, 8, 3, 4, 6], 6.0, True, False),
        ("t9_720p", 9, 3726, [8, 0, 5, 1, 2, 7, 3, 4, 6], 2.0, True, True),
        ("t2_tiny", 2, 70, [1, 0], 3.0, True, True),
    ]
    only = sys.argv[1:] or None
    for name, T, HW, slots, sharp, use_pe, use_gate in cases:
        if only and name not in only:
            continue
        g = torch.Generator().manual_seed(11)
        q = bfr(torch.randn(HW, 128, generator=g) * sharp)
        k = bfr(torch.randn(T, HW, 128, generator=g))
        # make later keys progressively "hotter" so the running max keeps moving (exercises the lazy rescale)
        k = bfr(k * torch.linspace(0.5, 1.5, HW).view(1, HW, 1))
        v = bfr(torch.randn(T, HW, 1024, generator=g))
        pe_cur = torch.randn(128, generator=g) * 0.1
        pe_mem = torch.randn(4, 128, generator=g) * 0.5
        gate = bfr(torch.randn(HW, 1024, generator=g)) if use_gate else None
        nslots = max(slots) + 1
        if use_pe:
            qt = bfr(q + pe_cur)
            kt = k + O.temporal_pe(pe_mem, T).view(T, 1, -1)
        else:
            qt, kt = q, k
        ref, ref_mass = O.long_term_attention(qt, kt, v, 128)
        if gate is not None:
            ref = ref * gate
        kb, vtb, HWp = K.build_bank(k.to(dev), v.to(dev), nslots, slots)
        args = dict(pe_cur=pe_cur.to(dev) if use_pe else None, mem_pos_emb=pe_mem.to(dev) if use_pe else None,
                    gate=gate.to(dev).to(OP) if gate is not None else None)
        rec = dict(case=name)
        try:
            od, md = K.long_attention(q.to(dev).to(OP), kb, vtb, slots, HW, impl=_capi.ATTN_DENSE, **args)
            cnt = torch.zeros(1, dtype=torch.int32, device=dev)
            _capi.check(_capi.load().rmem_debug_attn_rescale_counter(_capi.ptr(cnt)))
            ot, mt = K.long_attention(q.to(dev).to(OP), kb, vtb, slots, HW, impl=IMPL,
                                      grid=GRIDS[HW] if SEED else None, **args)
            torch.cuda.synchronize()
            _capi.check(_capi.load().rmem_debug_attn_rescale_counter(None))
            rec.update(ok=True, impl=IMPL, seeded=SEED, rescales=int(cnt.item()), fallback=K.last_attn_overflow, tc_vs_oracle=relfro(ot, ref), dense_vs_oracle=relfro(od, ref),
                       tc_vs_dense=relfro(ot, od), mass_err=float((mt.cpu() - ref_mass).abs().max()),
                       mass_sum_err=float((mt.sum(1).cpu() - 1).abs().max()),
                       finite=bool(torch.isfinite(ot.float()).all()))
        except Exception as e:  # noqa: BLE001
            rec.update(ok=False, error=str(e)[:300])
            print(json.dumps(rec), flush=True)
            break
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
