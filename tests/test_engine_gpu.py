"""End-to-end parity of the CUDA engine (through the Python mirror of AOTInferEngine and the C ABI) against
the golden fixtures recorded from the UNMODIFIED reference (oracle/make_golden.py).

Lock-step protocol: the label feedback loop is chaotic under random weights, so every memory update uses
the reference's own label history (teacher forcing) -- both implementations then see identical inputs at
every frame and are compared on (a) 1/4-res logits, (b) argmax labels, (c) long_memories_indexes after each
update (integer, exact).

Tolerance (stated): tensor-core operands are fp16 (fp32 accumulate, fp32 residual stream / norms / logits);
the reference is fp32.  1/4-res logits: max-abs error <= 1.5e-2 * max|logit| (achieved: 2-4e-3; SURVEY.md 8c asked for
5e-2); label agreement >= 99.5 %; eviction index sequence identical.  The achieved numbers of every case are appended
to gpurun_out/r02_parity.json (committed under profiles/ by the round script).
"""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import rmem_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOGIT_TOL = 1.5e-2
LABEL_AGREE = 0.995

from parity_report import report  # noqa: E402


def load_case(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return meta, z


def run_engine_lockstep(meta, z, device, attn_impl=0):
    from rmem_b200.engine import RmemModel, RmemConfig, build_engine
    knobs = meta.get("knobs", {})
    sd = O.make_state_dict(meta["model"], seed=meta["seed"], sharpen=meta["sharpen"],
                           gru_memory=bool(knobs.get("GRU_MEMORY", False)))
    H, W, n_obj = meta["H"], meta["W"], meta["n_obj"]
    frames = O.synthetic_frames(meta["n_frames"], H, W, seed=meta["seed"] + 1)
    label0 = O.synthetic_label(H, W, n_obj)
    cfg = RmemConfig(model=meta["model"], former_mem_len=meta["former"], latter_mem_len=meta["latter"],
                     attn_impl=attn_impl, no_long_memory=bool(knobs.get("NO_LONG_MEMORY", False)),
                     reverse_infer=bool(knobs.get("REVERSE_INFER", False)), time_encode=bool(knobs.get("TIME_ENCODE", False)),
                     gru_memory=bool(knobs.get("GRU_MEMORY", False)))
    model = RmemModel(sd, cfg, device)
    eng = build_engine("deaotengine" if meta["model"] == "r50_deaotl" else "aotengine", phase="eval", aot_model=model,
                       gpu_id=0, long_term_mem_gap=meta["gap"])
    out_size = tuple(meta["out_size"])
    forced = torch.from_numpy(z["labels"])                      # [F-1,Ho,Wo] uint8
    eng.restart_engine()
    eng.add_reference_frame(frames[0:1].to(device), label0.int().to(device), obj_nums=[n_obj], frame_step=0)
    rec = dict(ref_logits4=eng.aot_engines[0].pred_id_logits.cpu(), logits4=[], labels=[], idx=[])
    for f in range(1, frames.shape[0]):
        logit, lab = eng.match_propogate_one_frame(frames[f:f + 1].to(device), output_size=out_size,
                                                   return_label=True)
        rec["logits4"].append(eng.aot_engines[0].pred_id_logits.cpu())
        rec["labels"].append(lab.cpu()[0, 0])
        lab_in = F.interpolate(forced[f - 1].float().view(1, 1, *out_size), size=(H, W), mode="nearest")
        eng.update_memory(lab_in.to(device))
        rec["idx"].append([list(e.long_memories_indexes) for e in eng.aot_engines])
    return rec, eng


@pytest.mark.parametrize("impl", [4, 3, 2, 0], ids=["tcgen05_column", "tcgen05_pair", "tcgen05_v2", "dense"])
@pytest.mark.parametrize("name", ["deaot_small_10obj", "deaot_small_xavier", "deaot_13obj_2engines"])
def test_engine_matches_reference_goldens(cuda_device, name, impl):
    meta, z = load_case(name)
    rec, eng = run_engine_lockstep(meta, z, cuda_device, attn_impl=impl)
    # (c) integer state: exact
    assert rec["idx"] == meta["idx"], f"long_memories_indexes diverged:\n ours {rec['idx']}\n ref  {meta['idx']}"
    # (a) logits
    gold_ref = torch.from_numpy(z["ref_logits4"].astype(np.float32))
    scale = float(gold_ref.abs().max())
    err0 = float((rec["ref_logits4"][0] - gold_ref).abs().max())
    worst = err0 / scale
    sub = torch.from_numpy(z["logits4_sub"])                    # [F-1,11,h4/4,w4/4]
    for f, lg in enumerate(rec["logits4"]):
        s = float(sub[f].abs().max())
        e = float((lg[0, :, ::4, ::4] - sub[f]).abs().max())
        worst = max(worst, e / s)
    for kf in meta["keep"]:
        full = torch.from_numpy(z[f"logits4_{kf}"].astype(np.float32))
        e = float((rec["logits4"][kf][0] - full).abs().max())
        worst = max(worst, (e - 2e-3 * float(full.abs().max())) / float(full.abs().max()))   # fixture is fp16
    print(f"[{name}] worst relative logit error = {worst:.3e}", flush=True)
    assert worst < LOGIT_TOL
    # (b) labels
    ours = torch.stack(rec["labels"])
    agree = float((ours == torch.from_numpy(z["labels"])).float().mean())
    print(f"[{name}] label agreement = {agree:.5f}")
    report(f"golden/{name}/attn_impl_{impl}", worst_rel_logit_err=worst, label_agreement=agree, idx_identical=True,
           frames=len(rec["logits4"]), tolerances=dict(logit=LOGIT_TOL, label=LABEL_AGREE), vs="unmodified reference")
    assert agree >= LABEL_AGREE


def test_ablation_knobs_match_reference(cuda_device):
    """NO_LONG_MEMORY against its own reference golden; REVERSE_INFER / TIME_ENCODE are inference no-ops in
    the reference (tests/golden/knobs.json, asserted by oracle/make_golden.py), so the engine accepts them and must
    reproduce the default golden bit-for-bit in its integer state and within tolerance in its logits."""
    test_engine_matches_reference_goldens(cuda_device, "deaot_no_long_memory", 3)
    kn = json.load(open(os.path.join(GOLD, "knobs.json")))
    assert {"REVERSE_INFER": True} in kn["identical_to_default"] and {"NO_LONG_MEMORY": True} in kn["changes_outputs"]
    meta, z = load_case(kn["case"])
    meta = dict(meta, knobs={"REVERSE_INFER": True, "TIME_ENCODE": True})
    rec, _ = run_engine_lockstep(meta, z, cuda_device, attn_impl=3)
    assert rec["idx"] == meta["idx"]
    from rmem_b200.engine import RmemModel, RmemConfig, build_engine
    with pytest.raises(Exception):
        sd = O.make_state_dict("r50_deaotl", seed=1)
        e = build_engine("deaotengine", aot_model=RmemModel(sd, RmemConfig(gru_memory=True), cuda_device))
        e.add_reference_frame(torch.zeros(1, 3, 129, 161), torch.zeros(1, 1, 129, 161).int(), obj_nums=[1], frame_step=0)


@pytest.mark.parametrize("name", ["aot_c1_256_t1", "aot_small_rmem", "aot_gru_memory"])
def test_aot_engine_matches_reference_goldens(cuda_device, name):
    """R50_AOTL (+RMem): BASELINE.json configs[0] (c1), a restricted-bank clip with eviction, and the GRU_MEMORY ablation
    (ConvGRU condensation of every evicted frame into bank position 1, transformer.py:420-430) against its own golden."""
    test_engine_matches_reference_goldens(cuda_device, name, 0)
    test_engine_matches_reference_goldens(cuda_device, name, 3)      # attn_impl != dense: the fused tcgen05 MHA kernel


def test_engine_restart_is_deterministic(cuda_device):
    meta, z = load_case("deaot_small_xavier")
    rec1, eng = run_engine_lockstep(meta, z, cuda_device, attn_impl=4)
    rec2, _ = run_engine_lockstep(meta, z, cuda_device, attn_impl=4)
    assert rec1["idx"] == rec2["idx"]
    for a, b in zip(rec1["labels"], rec2["labels"]):
        assert torch.equal(a, b)
    assert eng.launch_count > 0


def _free_run(eng, frames_dev, label0, n_obj, gap, mode):
    """Self-fed clip (engine's own labels).  mode: None = no prefetch, 'before' = prefetch(i+1) issued before
    propagate(i), 'after' = issued after it, 'wrong' = prefetch of a frame that is not the next one."""
    eng.restart_engine()
    eng.long_term_mem_gap = gap
    eng.add_reference_frame(frames_dev[0:1], label0, obj_nums=[n_obj], frame_step=0)
    n = frames_dev.shape[0]
    labs, idx = [], []
    for f in range(1, n):
        nxt = frames_dev[f + 1:f + 2] if f + 1 < n else None
        if mode == "before" and nxt is not None:
            eng.prefetch(nxt)
        if mode == "wrong":
            eng.prefetch(frames_dev[0:1])
        lab = eng.propagate_label(frames_dev[f:f + 1])
        if mode == "after" and nxt is not None:
            eng.prefetch(nxt)
        eng.update_memory(lab)
        labs.append(lab.clone())
        idx.append(list(eng.aot_engines[0].long_memories_indexes))
    torch.cuda.synchronize()
    return torch.cat(labs), idx


@pytest.mark.parametrize("model", ["r50_deaotl", "r50_aotl"])
def test_prefetched_encoder_is_bit_identical(cuda_device, model):
    """rmem_engine_prefetch only moves the encoder of the next frame onto a side stream: labels (integer, exact) and the
    eviction index sequence must not change, whatever the call order, and a prefetch of the wrong frame is ignored."""
    from rmem_b200.engine import RmemModel, RmemConfig, build_engine
    H, W, n_obj = 257, 321, 3
    sd = O.make_state_dict(model, seed=3, sharpen=4.0)
    frames = O.synthetic_frames(12, H, W, seed=11).to(cuda_device)
    label0 = O.synthetic_label(H, W, n_obj).int().to(cuda_device)
    cfg = RmemConfig(model=model, former_mem_len=1, latter_mem_len=2)
    eng = build_engine("deaotengine" if model == "r50_deaotl" else "aotengine", phase="eval",
                       aot_model=RmemModel(sd, cfg, cuda_device), gpu_id=0, long_term_mem_gap=2)
    base, idx0 = _free_run(eng, frames, label0, n_obj, 2, None)
    for mode in ("before", "after", "wrong", "before"):
        labs, idx = _free_run(eng, frames, label0, n_obj, 2, mode)
        assert torch.equal(labs, base), mode
        assert idx == idx0, mode


@pytest.mark.parametrize("model,group", [("r50_deaotl", 2), ("r50_aotl", 2), ("r50_deaotl", 4)])
def test_pair_prefetch_matches_single_frame_encoder(cuda_device, model, group):
    """rmem_engine_prefetch2 encodes frames i+2, i+3 in one pass (every encoder GEMM / conv over both images, 4-D tensor
    maps): same math, differently tiled, so not bit-identical -- teacher-forced with the inline run's labels, every frame's
    1/4-res logits must agree within 5e-3 of the logit range (the fp16 noise floor of the path: 1.4e-3 .. 4.3e-3 against fp32; engine tolerance 1.5e-2), labels >= 99.5 % (argmax near-ties on random frames flip),
    identical eviction indices; frames the pair did not cover (odd clip end, wrong pointer) fall back to the inline encoder."""
    from rmem_b200.engine import RmemModel, RmemConfig, build_engine
    H, W, n_obj, gap = 257, 321, 3, 2
    sd = O.make_state_dict(model, seed=3, sharpen=4.0)
    frames = O.synthetic_frames(12 if group == 2 else 19, H, W, seed=11).to(cuda_device)
    label0 = O.synthetic_label(H, W, n_obj).int().to(cuda_device)
    cfg = RmemConfig(model=model, former_mem_len=1, latter_mem_len=2)
    eng = build_engine("deaotengine" if model == "r50_deaotl" else "aotengine", phase="eval",
                       aot_model=RmemModel(sd, cfg, cuda_device), gpu_id=0, long_term_mem_gap=gap)
    n = frames.shape[0]

    def run(pairs, forced=None):
        eng.restart_engine()
        eng.long_term_mem_gap = gap
        eng.add_reference_frame(frames[0:1], label0, obj_nums=[n_obj], frame_step=0)
        labs, logits, idx = [], [], []
        for f in range(1, n):
            if pairs and group == 2 and f % 2 == 1 and f + 3 < n:
                eng.prefetch2(frames[f + 2:f + 3], frames[f + 3:f + 4])
            if pairs and group == 4 and f % 4 == 1 and f + 7 < n:      # frames f+4 .. f+7, four frames ahead
                eng.prefetch_n([frames[f + 4 + j:f + 5 + j] for j in range(4)])
            lab = eng.propagate_label(frames[f:f + 1])
            logits.append(eng.logits4_views()[0].clone())
            eng.update_memory(forced[f - 1] if forced is not None else lab)
            labs.append(lab.clone())
            idx.append(list(eng.aot_engines[0].long_memories_indexes))
        torch.cuda.synchronize()
        return labs, logits, idx

    l0 = eng.launch_count
    base_l, base_g, base_i = run(False)
    l1 = eng.launch_count
    pair_l, pair_g, pair_i = run(True, forced=base_l)
    l2 = eng.launch_count
    assert pair_i == base_i
    assert l2 - l1 < l1 - l0 - 100, "the group encoder did not run"     # 4 pairs / 3 quads: >= 4 encoder passes (~50 launches each) saved
    for f in range(n - 1):
        err = float((pair_g[f] - base_g[f]).abs().max() / base_g[f].abs().max())
        agree = float((pair_l[f] == base_l[f]).float().mean())
        assert err < 5e-3 and agree >= 0.995, (f, err, agree)
    assert torch.equal(pair_g[0], base_g[0])          # frames 1, 2 were never prefetched: the inline encoder, bit for bit


def test_evaluator_shell_cuda_engine_vs_oracle_engine(cuda_device, tmp_path):
    """The same clip directory through rmem_b200.evaluator with the CUDA engine and with the CPU oracle engine: decoded
    frames, resize rule, reference frame, propagate / update loop, new-object re-reference and the PNG writer are shared;
    only the engine differs.  Free-running (each engine feeds on its own labels), so a few near-tie pixels may differ."""
    import cv2
    from PIL import Image
    from rmem_b200 import evaluator as E
    from rmem_b200.engine import RmemModel, RmemConfig, build_engine
    rng = np.random.RandomState(0)
    H, W = 97, 129
    img_dir, lab_dir = tmp_path / "JPEGImages" / "c", tmp_path / "Annotations" / "c"
    os.makedirs(img_dir), os.makedirs(lab_dir)
    base = cv2.GaussianBlur(rng.randint(0, 255, (H, W, 3)).astype(np.uint8), (0, 0), 3)
    for f in range(5):
        cv2.imwrite(str(img_dir / f"{f:05d}.png"), np.roll(base, 2 * f, axis=1))
    l0 = np.zeros((H, W), np.uint8); l0[20:60, 20:70] = 4
    Image.fromarray(l0).save(lab_dir / "00000.png")
    l2 = np.zeros((H, W), np.uint8); l2[60:90, 80:120] = 9
    Image.fromarray(l2).save(lab_dir / "00002.png")
    ds = E.ClipDataset(str(img_dir), str(lab_dir))
    sd = O.make_state_dict("r50_deaotl", seed=5, sharpen=4.0)
    eng = build_engine("deaotengine", phase="eval",
                       aot_model=RmemModel(sd, RmemConfig(model="r50_deaotl", former_mem_len=1, latter_mem_len=2),
                                           cuda_device), gpu_id=0)
    ours = E.evaluate_clip(eng, ds, out_dir=str(tmp_path / "ours"), device=cuda_device, keep_labels=True)
    with torch.no_grad():
        ref = E.evaluate_clip(O.OracleEngine(sd, O.OracleConfig(model="r50_deaotl", former_mem_len=1, latter_mem_len=2)),
                              ds, keep_labels=True)
    assert ours.frames == ref.frames == 4 and ours.seconds > 0
    agree = np.mean([float((a == b).float().mean()) for a, b in zip(ours.labels, ref.labels)])
    assert agree >= 0.97, agree
    out2 = np.array(Image.open(ours.paths[1]))                   # frame 2: dataset ids, new object pasted in
    assert (out2[60:90, 80:120] == 9).all() and set(np.unique(out2)) <= {0, 4, 9}


def test_tta_shell_cuda_engines_vs_oracle_engines(cuda_device, tmp_path):
    """Flip + two-scale test-time augmentation (evaluator.py:338-441) through rmem_b200.evaluator.evaluate_clip_tta: four
    CUDA engines on one weight blob, frames preprocessed on the GPU, logits merged by the fused TTA head -- against the
    same shell driving four CPU oracle engines with the cv2 loader.  Free-running; the averaged probabilities make the
    labels less sensitive to near-ties than the single-engine run."""
    import cv2
    from PIL import Image
    from rmem_b200 import evaluator as E
    from rmem_b200.engine import RmemModel, RmemConfig, build_engine
    rng = np.random.RandomState(1)
    H, W = 97, 129
    img_dir, lab_dir = tmp_path / "JPEGImages" / "c", tmp_path / "Annotations" / "c"
    os.makedirs(img_dir), os.makedirs(lab_dir)
    base = cv2.GaussianBlur(rng.randint(0, 255, (H, W, 3)).astype(np.uint8), (0, 0), 3)
    for f in range(4):
        cv2.imwrite(str(img_dir / f"{f:05d}.png"), np.roll(base, 2 * f, axis=1))
    l0 = np.zeros((H, W), np.uint8); l0[20:60, 20:70] = 4
    Image.fromarray(l0).save(lab_dir / "00000.png")
    ds = E.ClipDataset(str(img_dir), str(lab_dir))
    sd = O.make_state_dict("r50_deaotl", seed=5, sharpen=4.0)
    model = RmemModel(sd, RmemConfig(model="r50_deaotl", former_mem_len=1, latter_mem_len=2), cuda_device)
    engines = [build_engine("deaotengine", phase="eval", aot_model=model, gpu_id=0) for _ in range(4)]
    ours = E.evaluate_clip_tta(engines, ds, flip=True, multi_scale=(1.0, 1.3), device=cuda_device, keep_labels=True)
    ocfg = O.OracleConfig(model="r50_deaotl", former_mem_len=1, latter_mem_len=2)
    with torch.no_grad():
        ref = E.evaluate_clip_tta([O.OracleEngine(sd, ocfg) for _ in range(4)], ds, flip=True, multi_scale=(1.0, 1.3),
                                  keep_labels=True)
    assert ours.frames == ref.frames == 3 and ours.seconds > 0
    assert engines[2].input_size_2d == (129, 161)                 # int(97*1.3) = 126 -> 129, int(129*1.3) = 167 -> 161
    agree = np.mean([float((a == b).float().mean()) for a, b in zip(ours.labels, ref.labels)])
    from parity_report import report
    report("shell/tta_flip_2scales", label_agreement=float(agree), frames=3, vs="same shell with CPU oracle engines + cv2 loader")
    assert agree >= 0.97, agree
