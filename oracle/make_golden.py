"""Pin the oracle against the UNMODIFIED reference and write golden fixtures.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference, read-only):

    python oracle/make_golden.py            # validate oracle == reference, write tests/golden/*.pt

The reference (aot_plus/) is imported in place with the shims of SURVEY.md section 8(c):
stub timm / matplotlib modules, cfg built without init_dir(), the pre_vost stage deltas,
a CPU device patch for AOTEngine.assign_identity (aot_engine.py:209-213), a per-engine
deepcopy of the model for >10 objects (shared-LSTT-state bug, aot_engine.py:680-684) and
stdout silencing for restrict_long_memories' prints.  Nothing of the reference is copied.

The fixtures hold only seeds, small outputs and integer sequences; weights and inputs are
regenerated from seeds by oracle.rmem_oracle.make_state_dict / synthetic_*.
"""
from __future__ import annotations

import contextlib
import json
import copy
import io
import os
import sys
import types

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("RMEM_REFERENCE", "/root/reference")

from oracle import rmem_oracle as O  # noqa: E402


def import_reference():
    tl = types.ModuleType("timm.models.layers")
    tl.trunc_normal_ = torch.nn.init.trunc_normal_
    tl.DropPath = torch.nn.Identity
    tl.to_2tuple = lambda x: (x, x)
    sys.modules.update({"timm": types.ModuleType("timm"), "timm.models": types.ModuleType("timm.models"),
                        "timm.models.layers": tl})
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": mpl.pyplot})
    sys.path.insert(0, os.path.join(REF, "aot_plus"))
    from configs.default import DefaultEngineConfig
    from networks.models import build_vos_model
    from networks.engines import build_engine
    from networks.engines.aot_engine import AOTEngine

    orig_assign = AOTEngine.assign_identity

    def assign_identity(self, one_hot_mask, ignore_mask=None):
        if ignore_mask is None:
            ignore_mask = torch.zeros(one_hot_mask.shape[0], 1, one_hot_mask.shape[2], one_hot_mask.shape[3],
                                      device=one_hot_mask.device)
        return orig_assign(self, one_hot_mask, ignore_mask)

    AOTEngine.assign_identity = assign_identity

    orig_init = AOTEngine.__init__
    seen = {}

    def init(self, aot_model, *a, **k):                 # per-engine deepcopy for the 2nd+ engine
        key = id(aot_model)
        if seen.get(key, 0) > 0:
            aot_model = copy.deepcopy(aot_model)
        seen[key] = seen.get(key, 0) + 1
        orig_init(self, aot_model, *a, **k)

    AOTEngine.__init__ = init
    return DefaultEngineConfig, build_vos_model, build_engine, seen


def build_reference(model: str, sd, former: int, latter: int, gap: int, knobs: dict = None):
    DefaultEngineConfig, build_vos_model, build_engine, seen = import_reference()
    cfg = DefaultEngineConfig("golden", model)
    cfg.MODEL_LINEAR_Q = False
    cfg.MODEL_IGNORE_TOKEN = True
    cfg.FORMER_MEM_LEN, cfg.LATTER_MEM_LEN = former, latter
    for k, v in (knobs or {}).items():                  # ablation knobs of configs/models/r50_deaotl.py:9-28
        setattr(cfg, k, v)
    net = build_vos_model(cfg.MODEL_VOS, cfg).eval()
    missing = net.load_state_dict(sd, strict=True)
    seen.clear()
    eng = build_engine(cfg.MODEL_ENGINE, phase="eval", aot_model=net, gpu_id=0, long_term_mem_gap=gap).eval()
    return net, eng


def run_reference_clip(eng, frames, label0, n_obj, out_size):
    H, W = frames.shape[-2:]
    rec = dict(labels=[], idx=[], logits4=[], logits_out=[])
    sink = io.StringIO()
    with torch.no_grad(), contextlib.redirect_stdout(sink):
        eng.restart_engine()
        eng.add_reference_frame(frames[0:1], label0.int(), obj_nums=[n_obj], frame_step=0)
        rec["ref_logits4"] = eng.aot_engines[0].pred_id_logits.clone()
        for f in range(1, frames.shape[0]):
            logit = eng.match_propogate_one_frame(frames[f:f + 1], output_size=out_size)
            lab = torch.argmax(torch.softmax(logit, dim=1), dim=1, keepdim=True).float()
            lab_in = F.interpolate(lab, size=(H, W), mode="nearest")
            eng.update_memory(lab_in)
            rec["labels"].append(lab.to(torch.uint8))
            rec["idx"].append([list(e.long_memories_indexes) for e in eng.aot_engines])
            rec["logits4"].append(eng.aot_engines[0].pred_id_logits.clone())
            rec["logits_out"].append(logit.clone())
    return rec


def run_oracle_clip(sd, cfg, gap, frames, label0, n_obj, out_size, forced_labels=None):
    eng = O.OracleEngine(sd, cfg, long_term_mem_gap=gap)
    rec = dict(labels=[], idx=[], logits4=[], logits_out=[])

    def on_frame(f, logit, lab):
        rec["idx"].append([list(e.long_memories_indexes) for e in eng.aot_engines])
        rec["logits4"].append(eng.aot_engines[0].pred_id_logits.clone())
        rec["logits_out"].append(logit.clone())

    with torch.no_grad():
        rec["labels"] = O.run_clip(eng, frames, label0, n_obj, out_size=out_size, on_frame=on_frame,
                                   forced_labels=forced_labels)
    return rec, eng


# Ablation knobs on the deaot_small_xavier clip.  NO_LONG_MEMORY changes the result (fixture
# deaot_no_long_memory.npz); REVERSE_INFER and TIME_ENCODE are asserted to leave every inference output of the
# UNMODIFIED reference bit-identical (aot_engine.py:371-396 only feeds a training loss; the sin/cos encoding of
# aot_engine.py:293-303, 413-421 is stored and never read) -- the result is recorded in tests/golden/knobs.json.
KNOB_CASE = "deaot_small_xavier"
KNOBS_IDENTICAL = [{"REVERSE_INFER": True}, {"TIME_ENCODE": True, "TIME_ENCODE_NORM": False},
                   {"TIME_ENCODE": True, "TIME_ENCODE_NORM": True}]

CASES = {
    # name: (model, seed, sharpen, H, W, n_obj, n_frames, former, latter, gap, out_size)
    "deaot_small_10obj": ("r50_deaotl", 0, 4.0, 257, 321, 10, 14, 1, 3, 2, (256, 320)),
    "deaot_small_xavier": ("r50_deaotl", 1, 1.0, 257, 321, 3, 8, 1, 2, 2, (257, 321)),
    "deaot_13obj_2engines": ("r50_deaotl", 2, 4.0, 193, 257, 13, 7, 1, 2, 2, (193, 257)),
    # BASELINE.json configs[0] (c1): R50_AOTL, 256x256 -> 257x257, 1 object, T = 1 (reference frame only)
    "aot_c1_256_t1": ("r50_aotl", 3, 1.0, 257, 257, 1, 3, 1, 1, 9999, (256, 256)),
    # AOT + RMem: restricted bank (1 + 2), eviction active, sharpened attention
    "aot_small_rmem": ("r50_aotl", 4, 4.0, 193, 257, 3, 11, 1, 2, 2, (193, 257)),
    # GRU_MEMORY ablation (r50_aotl only): bank 1 + 3, an eviction with ConvGRU condensation on every frame
    # from the fifth on (the ConvGRU weights come from make_state_dict(..., gru_memory=True))
    "aot_gru_memory": ("r50_aotl", 5, 4.0, 193, 257, 3, 13, 1, 3, 1, (193, 257), {"GRU_MEMORY": True}),
}


def check_knobs():
    import numpy as np
    model, seed, sharpen, H, W, n_obj, nfr, former, latter, gap, out_size = CASES[KNOB_CASE][:11]
    sd = O.make_state_dict(model, seed=seed, sharpen=sharpen)
    frames = O.synthetic_frames(nfr, H, W, seed=seed + 1)
    label0 = O.synthetic_label(H, W, n_obj)
    _, eng = build_reference(model, sd, former, latter, gap)
    base = run_reference_clip(eng, frames, label0, n_obj, out_size)
    report = {"case": KNOB_CASE, "identical_to_default": []}
    for knobs in KNOBS_IDENTICAL:
        _, e2 = build_reference(model, sd, former, latter, gap, knobs)
        r = run_reference_clip(e2, frames, label0, n_obj, out_size)
        same = all(torch.equal(a, b) for a, b in zip(base["logits_out"], r["logits_out"])) and base["idx"] == r["idx"]
        print(f"[knobs] {knobs}: inference outputs identical to the default config = {same}")
        assert same, knobs
        report["identical_to_default"].append(knobs)
    # NO_LONG_MEMORY (aot_engine.py:339): the long-term bank never grows beyond the reference frame
    knobs = {"NO_LONG_MEMORY": True}
    _, e3 = build_reference(model, sd, former, latter, gap, knobs)
    ref = run_reference_clip(e3, frames, label0, n_obj, out_size)
    cfg = O.OracleConfig(model=model, former_mem_len=former, latter_mem_len=latter, no_long_memory=True)
    ref_labels = torch.stack([l[0, 0] for l in ref["labels"]])
    orc, _ = run_oracle_clip(sd, cfg, gap, frames, label0, n_obj, out_size, forced_labels=ref_labels)
    worst = max((a - b).abs().max().item() for a, b in zip(ref["logits_out"], orc["logits_out"]))
    print(f"[knobs] NO_LONG_MEMORY: max|logit_ref - logit_oracle| = {worst:.3e}, idx = {ref['idx'][-1]}")
    assert worst < 2e-4 and ref["idx"] == orc["idx"] and all(i == [[0]] for i in ref["idx"])
    differs = not all(torch.equal(a, b) for a, b in zip(base["logits_out"], ref["logits_out"]))
    assert differs, "NO_LONG_MEMORY should change the outputs of a clip that appends long-term frames"
    meta = dict(case="deaot_no_long_memory", model=model, seed=seed, sharpen=sharpen, H=H, W=W, n_obj=n_obj, n_frames=nfr,
                former=former, latter=latter, gap=gap, out_size=list(out_size), idx=ref["idx"], keep=[len(ref["logits4"]) - 1],
                knobs=knobs, reference_commit="431cde18", torch=torch.__version__)
    k = len(ref["logits4"]) - 1
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "deaot_no_long_memory.npz"), meta=json.dumps(meta),
                        labels=ref_labels.numpy(),
                        logits4_sub=torch.stack([x[0, :, ::4, ::4] for x in ref["logits4"]]).numpy(),
                        ref_logits4=ref["ref_logits4"][0].numpy().astype(np.float16),
                        **{f"logits4_{k}": ref["logits4"][k][0].numpy().astype(np.float16)})
    report["changes_outputs"] = [knobs]
    json.dump(report, open(os.path.join(ROOT, "tests", "golden", "knobs.json"), "w"), indent=1)


def main():
    torch.set_num_threads(8)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    only = sys.argv[1:]
    for name, case in CASES.items():
        model, seed, sharpen, H, W, n_obj, nfr, former, latter, gap, out_size = case[:11]
        knobs = case[11] if len(case) > 11 else {}
        gru = bool(knobs.get("GRU_MEMORY", False))
        if only and name not in only:
            continue
        sd = O.make_state_dict(model, seed=seed, sharpen=sharpen, gru_memory=gru)
        frames = O.synthetic_frames(nfr, H, W, seed=seed + 1)
        label0 = O.synthetic_label(H, W, n_obj)
        net, eng = build_reference(model, sd, former, latter, gap, knobs)
        ref = run_reference_clip(eng, frames, label0, n_obj, out_size)
        cfg = O.OracleConfig(model=model, former_mem_len=former, latter_mem_len=latter, gru_memory=gru)
        ref_labels = torch.stack([l[0, 0] for l in ref["labels"]])                # uint8 [F-1,Ho,Wo]
        orc, oeng = run_oracle_clip(sd, cfg, gap, frames, label0, n_obj, out_size, forced_labels=ref_labels)

        worst = 0.0
        for f, (a, b) in enumerate(zip(ref["logits_out"], orc["logits_out"])):
            worst = max(worst, (a - b).abs().max().item())
        lab_mismatch = sum(int((a != b).sum()) for a, b in zip(ref["labels"], orc["labels"]))
        npix = ref_labels.numel()
        idx_ok = ref["idx"] == orc["idx"]
        print(f"[{name}] lock-step max|logit_ref - logit_oracle| = {worst:.3e}  label mismatches = "
              f"{lab_mismatch}/{npix}  idx identical = {idx_ok}  final idx = {ref['idx'][-1]}  logit range = "
              f"{float(ref['logits_out'][-1].min()):.2f}..{float(ref['logits_out'][-1].max()):.2f}")
        assert worst < 2e-4, "oracle diverges from the reference"
        assert lab_mismatch <= 1e-5 * npix + 2 and idx_ok

        import numpy as np
        keep = sorted(set([0, len(ref["logits4"]) // 2, len(ref["logits4"]) - 1]))
        meta = dict(case=name, model=model, seed=seed, sharpen=sharpen, H=H, W=W, n_obj=n_obj, n_frames=nfr,
                    former=former, latter=latter, gap=gap, out_size=list(out_size), idx=ref["idx"],
                    keep=keep, reference_commit="431cde18", torch=torch.__version__)
        if knobs:
            meta["knobs"] = knobs
        arrays = dict(
            labels=ref_labels.numpy(),
            # strided sample of every frame's 1/4-res logits (engine 0), fp32
            logits4_sub=torch.stack([x[0, :, ::4, ::4] for x in ref["logits4"]]).numpy(),
            ref_logits4=ref["ref_logits4"][0].numpy().astype(np.float16),
        )
        for k in keep:
            arrays[f"logits4_{k}"] = ref["logits4"][k][0].numpy().astype(np.float16)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), meta=json.dumps(meta), **arrays)
    if not only or "knobs" in only:
        check_knobs()
    print("golden fixtures written")


if __name__ == "__main__":
    main()
