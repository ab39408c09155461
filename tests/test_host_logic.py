"""CPU: host-side logic of the engine (no GPU): the evict-pick state machine and the temporal-PE slot map in
the C++ library against the oracle's restatement of transformer.py:907-964 / :1140-1170, weight packing, and the
synthetic-shape rules."""
import random

import torch

from oracle import rmem_oracle as O
from rmem_b200 import ops as K
from rmem_b200 import synth
from rmem_b200.weights import pack_deaot


def test_temporal_pe_slots_match_oracle():
    for T in range(1, 17):
        ref = [lo if fr == 0.0 else None for lo, hi, fr in O.temporal_pe_slots(T)]
        assert K.temporal_pe_slots(T) == ref, T
    assert K.temporal_pe_slots(8) == [0, 0, 1, 1, 2, 2, 3, 3]
    assert K.temporal_pe_slots(9) == [0, 0, 1, 1, 2, 2, 3, 3, 3]


def test_evict_pick_state_machine_matches_oracle():
    for seed in range(5):
        random.seed(seed)
        g = torch.Generator().manual_seed(seed)
        st = O.EvictState()
        ema, times = {}, {}
        idx, former, cap, step = [0], 1, 3 + seed, 0
        for it in range(60):
            step += random.randint(1, 6)
            T_old = len(idx)
            idx.append(step)
            rel = torch.rand(T_old, generator=g)
            d_ref = O.evict_pick(rel / rel.sum(), idx, st, former)
            d = K.evict_pick(rel.tolist(), idx, former, ema, times)
            assert d == d_ref, (seed, it, d, d_ref)
            assert set(ema) == set(st.ema) and times == st.times
            assert all(abs(ema[k] - float(st.ema[k])) < 1e-6 for k in ema)
            if len(idx) > cap:
                idx.pop(d)
                assert 0 in idx and step in idx       # reference frame and newest frame are never dropped


def test_weight_packing_shapes_and_bn_folding():
    sd = synth.make_state_dict("r50_deaotl", seed=0)
    pk = pack_deaot({"module." + k: v for k, v in sd.items()})         # checkpoint-style prefix is stripped
    # stem layout: [Cout][7 window rows][8 pixels][8 channels]; channels 3.. and the 8th pixel carry zero weights
    assert pk["enc.conv1.w"].shape == (64, 7, 8, 8) and float(pk["enc.conv1.w"][..., 3:].abs().max()) == 0.0
    assert float(pk["enc.conv1.w"][:, :, 7].abs().max()) == 0.0
    assert pk["enc.layer3.0.ds.w"].shape == (1024, 1, 1, 512)
    assert pk["idbank.w"].shape == (289, 12, 256) and pk["gpm.1.linear_ID_V.w"].shape == (512, 512)
    assert pk["gpm.0.linear_ID_V.w"].shape == (512, 256) and pk["gpm.2.short.rel.w"].shape == (256, 128)
    assert pk["gpm.0.long.dw"].shape == (25, 1024) and pk["dec.conv_out.w"].shape == (11, 1, 1, 128)
    # folded conv+BN == conv then frozen BN
    x = torch.randn(1, 256, 9, 9)
    p = "encoder.layer1.1"
    ref = O.frozen_bn(sd, p + ".bn1", torch.nn.functional.conv2d(x, sd[p + ".conv1.weight"]))
    w = pk["enc.layer1.1.conv1.w"].float().permute(0, 3, 1, 2)
    out = torch.nn.functional.conv2d(x, w, pk["enc.layer1.1.conv1.b"])
    assert float((out - ref).abs().max() / ref.abs().max()) < 3e-3       # 16-bit weight rounding only


def test_synthetic_shapes_follow_the_reference_size_rule():
    assert (synth.snap_size(480), synth.snap_size(854)) == (481, 849)   # video_transforms.py:607-615
    assert (synth.snap_size(720), synth.snap_size(1280)) == (721, 1281)
    assert synth.snap_size(256) == 257
    lab = synth.synthetic_label(481, 849, 10)
    assert sorted(lab.unique().tolist()) == list(range(11))


def test_attention_static_schedule_invariants():
    """Host side of the fused attention kernel's stream-K schedule (attn_tc2.cu make_bounds): the cut points are
    ascending, cover every (unit, tile) step exactly once, no CTA spans more than two units, and the modelled cost
    (tiles + 6 per segment, +1 for a second one) is flat across CTAs for long launches.  No GPU involved."""
    import ctypes as C
    from rmem_b200 import _capi
    lib = _capi.load()
    for (HW, T, Dv) in [(1674, 8, 1024), (1674, 1, 1024), (1674, 5, 1024), (3726, 8, 1024), (289, 1, 1024),
                        (289, 3, 256), (540, 9, 512), (65, 2, 1024)]:
        n_units, tpu, n_cta = C.c_int(), C.c_int(), C.c_int()
        b = (C.c_int * 256)()
        _capi.check(lib.rmem_debug_attn_schedule(_capi.ATTN_TC2, HW, T, Dv, C.byref(n_units), C.byref(tpu), C.byref(n_cta), b, 256))
        n, L, TPU = n_cta.value, n_units.value * tpu.value, tpu.value
        assert n_units.value == -(-HW // 128) * (Dv // 256) and TPU == T * -(-HW // 64)
        bounds = [b[i] for i in range(n + 1)]
        assert bounds[0] == 0 and bounds[-1] == L and all(x <= y for x, y in zip(bounds, bounds[1:])), (HW, T, Dv)
        costs = []
        for lo, hi in zip(bounds, bounds[1:]):
            if hi == lo:
                continue
            segs = (hi - 1) // TPU - lo // TPU + 1
            assert segs <= 2, (HW, T, Dv, lo, hi)
            costs.append(hi - lo + 6 * segs + (1 if segs == 2 else 0))
        if L // n >= 16:
            assert max(costs) - min(costs) <= 8, (HW, T, Dv, min(costs), max(costs))
            assert len(costs) == n                                   # no idle CTA


def test_pair_attention_static_schedule_invariants():
    """Host side of the CTA-pair attention kernel's schedule (attn_tc3.cu make_bounds): clusters of two CTAs own
    contiguous ranges of (unit, 64-key sub-tile) steps; a unit is a PAIR of query tiles x one Dv chunk."""
    import ctypes as C
    from rmem_b200 import _capi
    lib = _capi.load()
    for (HW, T, Dv) in [(1674, 8, 1024), (1674, 9, 1024), (1674, 1, 1024), (1674, 5, 1024), (3726, 8, 1024),
                        (289, 1, 1024), (289, 3, 256), (540, 9, 512), (65, 2, 1024), (70, 2, 1024)]:
        n_units, tpu, n_cl = C.c_int(), C.c_int(), C.c_int()
        b = (C.c_int * 256)()
        _capi.check(lib.rmem_debug_attn_schedule(_capi.ATTN_TC3, HW, T, Dv, C.byref(n_units), C.byref(tpu),
                                                 C.byref(n_cl), b, 256))
        n, L, TPU = n_cl.value, n_units.value * tpu.value, tpu.value
        qtiles = -(-HW // 128)
        assert n_units.value == -(-qtiles // 2) * (Dv // 256) and TPU == T * -(-HW // 64)
        assert 1 <= n <= 74
        bounds = [b[i] for i in range(n + 1)]
        assert bounds[0] == 0 and bounds[-1] == L and all(x <= y for x, y in zip(bounds, bounds[1:])), (HW, T, Dv)
        costs = []
        for lo, hi in zip(bounds, bounds[1:]):
            if hi == lo:
                continue
            segs = (hi - 1) // TPU - lo // TPU + 1
            assert segs <= 2, (HW, T, Dv, lo, hi)
            costs.append(hi - lo + 2 * segs)
        # combine3 resolves at most 24 segments per unit
        for u in range(n_units.value):
            k = sum(1 for lo, hi in zip(bounds, bounds[1:]) if hi > lo and lo < (u + 1) * TPU and hi > u * TPU)
            assert k <= 24, (HW, T, Dv, u, k)
        if L // n >= 16:
            assert max(costs) - min(costs) <= 4, (HW, T, Dv, min(costs), max(costs))
            assert len(costs) == n                                   # no idle cluster
