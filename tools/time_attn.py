"""Times the long-term attention kernels (tc2 single-CTA, tc3 CTA-pair seeded / unseeded) on one c3 layer (T = 8):
random-normal operands (the round-1 bench operands) and, when scratch/attn_cap_layer{0,1,2}.npz exist, the REAL per-layer
Q / K-bank of a steady-state c3 frame (captured from the CPU oracle by tools/capture_attn_operands.py; peaked scores,
sigma ~ 15-20 log2 units in layers 1-2).  Kernel time = CUDA events recorded by the library immediately around the main
kernel; an L2 flush (256 MB write) separates launches.  One JSON line per case."""
import ctypes as C
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rmem_b200 import _capi, ops as K  # noqa: E402

HW, T, NSLOTS, H, W = 1674, 8, 9, 31, 54
FLOPS = 2.0 * HW * (T * HW) * (128 + 1024)


def main():
    dev = torch.device("cuda:0")
    lib = _capi.load()
    OP = _capi.op_dtype()
    g = torch.Generator().manual_seed(0)
    cases = [("randn", torch.randn(HW, 128, generator=g), torch.randn(T, HW, 128, generator=g))]
    for l in range(3):
        p = os.path.join(ROOT, "scratch", f"attn_cap_layer{l}.npz")
        if os.path.exists(p):
            d = np.load(p)
            cases.append((f"real_layer{l}", torch.from_numpy(d["q"].astype(np.float32)),
                          torch.from_numpy(d["k"].astype(np.float32))))
    v = torch.randn(T, HW, 1024, generator=g).to(dev)
    gate = torch.randn(HW, 1024, generator=g).to(dev).to(OP)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    slots = list(range(T))
    sl = (C.c_int * T)(*slots)
    st = _capi.stream_ptr()
    scale = 1.0 / math.sqrt(128)
    iters = int(os.environ.get("ITERS", "10"))
    for name, q, k in cases:
        kb, vtb, HWp = K.build_bank(k.to(dev), v, NSLOTS, slots)
        qt = q.to(dev).to(OP).contiguous()
        outs = {}
        for tag, impl, grid in (("tc2", 2, (0, 0)), ("tc3_unseeded", 3, (0, 0)), ("tc3_seeded", 3, (H, W))):
            nbytes = C.c_size_t()
            _capi.check(lib.rmem_long_attn_workspace_bytes(impl, HW, HWp, NSLOTS, 1024, C.byref(nbytes)))
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
            out = torch.empty(HW, 1024, dtype=OP, device=dev)
            mass = torch.empty(HW, T, dtype=torch.float32, device=dev)
            cnt = torch.zeros(1, dtype=torch.int32, device=dev)

            def launch():
                _capi.check(lib.rmem_long_attn_grid_fwd(impl, _capi.ptr(qt), None, _capi.ptr(kb), _capi.ptr(vtb), NSLOTS,
                                                        T, sl, HW, HWp, 128, 1024, C.c_float(scale), _capi.ptr(gate),
                                                        C.c_longlong(1024), _capi.ptr(out), C.c_longlong(1024),
                                                        _capi.ptr(mass), grid[0], grid[1], _capi.ptr(ws),
                                                        C.c_size_t(nbytes.value), st))
            rec = dict(case=name, impl=tag)
            try:
                _capi.check(lib.rmem_debug_attn_rescale_counter(_capi.ptr(cnt)))
                launch()
                torch.cuda.synchronize()
                _capi.check(lib.rmem_debug_attn_rescale_counter(None))
                rec["rescale_events"] = int(cnt.item())
                for _ in range(2):
                    launch()
                pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
                for a, b in pairs:
                    a.record(); b.record()
                torch.cuda.synchronize()
                for a, b in pairs:
                    flush.zero_()
                    _capi.check(lib.rmem_debug_attn_events(C.c_void_p(a.cuda_event), C.c_void_p(b.cuda_event)))
                    launch()
                _capi.check(lib.rmem_debug_attn_events(None, None))
                torch.cuda.synchronize()
                ts = sorted(a.elapsed_time(b) * 1e3 for a, b in pairs)
                rec.update(us_min=round(ts[0], 1), us_med=round(ts[len(ts) // 2], 1), us_max=round(ts[-1], 1),
                           tflops_med=round(FLOPS / ts[len(ts) // 2] / 1e6, 1),
                           finite=bool(torch.isfinite(out.float()).all()), mass_sum_err=float((mass.sum(1) - 1).abs().max()))
                outs[tag] = (out.float().clone(), mass.clone())
                if "tc2" in outs and tag != "tc2":
                    o2, m2 = outs["tc2"]
                    rec["vs_tc2_relfro"] = float((out.float() - o2).norm() / o2.norm())
                    rec["mass_vs_tc2"] = float((mass - m2).abs().max())
            except Exception as e:  # noqa: BLE001
                rec.update(ok=False, error=str(e)[:300])
                print(json.dumps(rec), flush=True)
                return
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
