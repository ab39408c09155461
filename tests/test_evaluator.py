"""CPU: host logic of the per-clip evaluation shell (rmem_b200/evaluator.py) against the behaviour of the reference's
evaluator (aot_plus/networks/managers/evaluator.py:300-556, dataloaders/eval_datasets.py:14-118,
dataloaders/video_transforms.py:559-666, utils/image.py:89-105): resize rule, object bookkeeping and label squeeze,
new-object merge + re-reference, palette PNG round trip, clip queue.  The engine behind the shell is a recording fake
or the CPU oracle; the CUDA engine goes through the same shell in tests/test_engine_gpu.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import rmem_oracle as O
from rmem_b200 import evaluator as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_clip(root, n_frames=5, H=65, W=81, new_at=2, seed=0):
    """frames 00000.jpg.. ; labels: frame 0 has dataset ids {3}, frame `new_at` introduces id 7."""
    import cv2
    from PIL import Image
    rng = np.random.RandomState(seed)
    img_dir, lab_dir = os.path.join(root, "JPEGImages", "clip"), os.path.join(root, "Annotations", "clip")
    os.makedirs(img_dir), os.makedirs(lab_dir)
    base = rng.randint(0, 255, (H, W, 3)).astype(np.uint8)
    for f in range(n_frames):
        cv2.imwrite(os.path.join(img_dir, f"{f:05d}.jpg"), np.roll(base, 3 * f, axis=1))
    l0 = np.zeros((H, W), np.uint8)
    l0[10:30, 10:40] = 3
    Image.fromarray(l0).save(os.path.join(lab_dir, "00000.png"))
    l2 = np.zeros((H, W), np.uint8)
    l2[40:60, 50:70] = 7
    Image.fromarray(l2).save(os.path.join(lab_dir, f"{new_at:05d}.png"))
    return img_dir, lab_dir


def test_restrict_size_rule():
    assert E.restrict_size(480, 854) == (481, 849)            # the c2/c3 geometry
    assert E.restrict_size(480, 910) == (481, 913)
    assert E.restrict_size(100, 100) == (97, 97)
    assert E.restrict_size(65, 81) == (65, 81)
    assert E.restrict_size(720, 1280) == (577, 1041)          # long edge capped at 1040, then aligned to 16k+1
    assert E.restrict_size(720, 1280, max_size=None, min_size=480) == (481, 849)
    assert E.restrict_size(480, 854, align_corners=False) == (480, 848)
    assert E.long_term_gap(100) == 5 and E.long_term_gap(2000) == 67 and E.long_term_gap(2000, True) == 17


def test_palette_and_mask_round_trip(tmp_path):
    from PIL import Image
    pal = E.davis_palette()
    assert len(pal) == 768 and pal[:12] == [0, 0, 0, 128, 0, 0, 0, 128, 0, 128, 128, 0]
    assert pal[27:30] == [191, 0, 0] and pal[66:72] == [22, 22, 22, 23, 23, 23]
    m = np.zeros((20, 30), np.uint8)
    m[2:5] = 1
    m[7:9] = 2
    E.save_mask(m, str(tmp_path / "a.png"), squeeze_idx=[0, 5, 9], background=False)
    im = Image.open(tmp_path / "a.png")
    back = np.array(im)
    assert im.mode == "P" and set(np.unique(back)) == {0, 5, 9}
    assert (back[2:5] == 5).all() and (back[7:9] == 9).all()
    assert im.getpalette()[:768] == pal


def test_clip_dataset_object_bookkeeping(tmp_path):
    img_dir, lab_dir = write_clip(str(tmp_path))
    ds = E.ClipDataset(img_dir, lab_dir)
    assert len(ds) == 5
    assert ds.obj_nums == [1, 1, 1, 2, 2]                       # eval_datasets.py:40-52 (frame 0 copies frame 1)
    assert ds.obj_indices == [[0, 3], [0, 3], [0, 3, 7], [0, 3, 7], [0, 3, 7]]
    s0, s1, s2 = ds[0], ds[1], ds[2]
    assert s0["current_img"].shape == (1, 3, 65, 81) and s0["current_img"].dtype == torch.float32
    assert "current_label" not in s1
    assert set(s0["current_label"].unique().tolist()) == {0, 1}
    assert set(s2["current_label"].unique().tolist()) == {0, 2}   # dataset id 7 squeezed to 2, id 3 absent in this file
    assert s2["meta"]["obj_idx"] == [0, 3, 7] and s2["meta"]["height"] == 65 and s2["meta"]["width"] == 81
    # MultiToTensor normalisation of the RGB-ordered frame
    import cv2
    raw = cv2.imread(os.path.join(img_dir, "00001.jpg")).astype(np.float32)[:, :, ::-1] / 255.
    ref = (raw - np.array(E.IMAGENET_MEAN)) / np.array(E.IMAGENET_STD)
    assert np.allclose(s1["current_img"][0].permute(1, 2, 0).numpy(), ref, atol=1e-5)


class FakeEngine:
    """Records the call sequence; predicts object 1 on the left half."""

    def __init__(self):
        self.calls = []
        self.long_term_mem_gap = None
        self.input_size_2d = None

    def restart_engine(self):
        self.calls.append(("restart",))

    def add_reference_frame(self, img, mask, obj_nums, frame_step=-1):
        self.input_size_2d = tuple(img.shape[-2:])
        self.calls.append(("ref", frame_step, list(obj_nums), mask.clone()))

    def match_propogate_one_frame(self, img, output_size=None):
        self.calls.append(("prop", tuple(output_size)))
        lg = torch.zeros(1, 11, *output_size)
        lg[:, 0] = 1.0
        lg[:, 1, :, : output_size[1] // 2] = 2.0
        return lg

    def update_memory(self, label):
        self.calls.append(("upd", label.clone()))


def test_evaluate_clip_sequence_merge_and_output(tmp_path):
    from PIL import Image
    img_dir, lab_dir = write_clip(str(tmp_path))
    ds = E.ClipDataset(img_dir, lab_dir)
    eng = FakeEngine()
    res = E.evaluate_clip(eng, ds, out_dir=str(tmp_path / "out"), keep_labels=True)
    kinds = [c[0] for c in eng.calls]
    assert kinds == ["restart", "ref", "prop", "upd", "prop", "ref", "prop", "upd", "prop", "upd"]
    assert eng.long_term_mem_gap == 5
    assert eng.calls[1][1] == 0 and eng.calls[1][2] == [1]
    assert eng.calls[1][3].dtype == torch.int32 and set(eng.calls[1][3].unique().tolist()) == {0, 1}
    # frame 2 introduces dataset id 7 (squeezed id 2): pasted over the prediction, engine re-referenced at step 2
    _, step, nums, merged = eng.calls[5]
    assert step == 2 and nums == [2]
    assert (merged[0, 0, 40:60, 50:70] == 2).all() and (merged[0, 0, :, :40][:, :30] == 1).all()
    assert res.frames == 4 and len(res.labels) == 4 and len(res.paths) == 4
    out2 = np.array(Image.open(res.paths[1]))                    # frame 2, dataset ids
    assert (out2[40:60, 50:70] == 7).all() and (out2[0:30, 0:40] == 3).all() and set(np.unique(out2)) <= {0, 3, 7}
    out1 = np.array(Image.open(res.paths[0]))
    assert set(np.unique(out1)) == {0, 3} and (out1[:, :40] == 3).all() and (out1[:, 41:] == 0).all()


def test_evaluate_clips_with_two_engines_in_flight(tmp_path):
    """evaluate_clips(extra_engines=...): every engine evaluates whole clips on its own thread, all drawing from one queue;
    every clip is evaluated exactly once and the per-clip call sequence is the single-engine one."""
    import shutil
    clips = []
    for k in range(5):
        d = tmp_path / f"c{k}"
        os.makedirs(d)
        img_dir, lab_dir = write_clip(str(d))
        clips.append(E.ClipDataset(img_dir, lab_dir))
    engs = [FakeEngine(), FakeEngine()]
    out = E.evaluate_clips(engs[0], clips, out_dir=None, log=None, extra_engines=engs[1:])
    assert len(out["results"]) == 5 and out["frames"] == 5 * 4
    per_clip = ["restart", "ref", "prop", "upd", "prop", "ref", "prop", "upd", "prop", "upd"]
    n0, n1 = len(engs[0].calls) // len(per_clip), len(engs[1].calls) // len(per_clip)
    assert n0 + n1 == 5
    for e, n in zip(engs, (n0, n1)):
        assert [c[0] for c in e.calls] == per_clip * n


def test_evaluate_clip_with_oracle_engine_equals_run_clip(tmp_path):
    """The shell around an engine reproduces the plain per-clip loop (oracle.run_clip) on the same tensors."""
    torch.manual_seed(0)
    img_dir, lab_dir = write_clip(str(tmp_path), n_frames=4, new_at=99)     # no new objects: plain propagation
    os.remove(os.path.join(lab_dir, "00099.png"))
    ds = E.ClipDataset(img_dir, lab_dir)
    sd = O.make_state_dict("r50_deaotl", seed=1, sharpen=1.0)
    cfg = O.OracleConfig(model="r50_deaotl", former_mem_len=1, latter_mem_len=2)
    with torch.no_grad():
        res = E.evaluate_clip(O.OracleEngine(sd, cfg), ds, keep_labels=True)
        frames = torch.cat([ds[i]["current_img"] for i in range(len(ds))])
        eng = O.OracleEngine(sd, cfg)
        eng.long_term_mem_gap = E.long_term_gap(len(ds))
        ref = O.run_clip(_GapKeeper(eng), frames, ds[0]["current_label"], 1)
    assert len(res.labels) == len(ref) == 3
    for a, b in zip(res.labels, ref):
        assert torch.equal(a, b.view_as(a))


class _GapKeeper:
    """run_clip restarts the engine; keep the evaluator's long-term gap across the restart."""

    def __init__(self, eng):
        self._e, self._gap = eng, eng.long_term_mem_gap

    def __getattr__(self, k):
        return getattr(self._e, k)

    def restart_engine(self):
        self._e.restart_engine()
        self._e.long_term_mem_gap = self._gap


def test_clip_queue_static_partition():
    for world in (1, 2, 3):
        got = sorted(i for r in range(world) for i in E.ClipQueue(10, r, world))
        assert got == list(range(10))


def _queue_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from rmem_b200 import evaluator as EV
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    store = dist.distributed_c10d._get_default_store()
    mine = list(EV.ClipQueue(9, rank, world, store=store))
    dist.barrier()
    q.put((rank, mine))
    dist.destroy_process_group()


def test_clip_queue_dynamic_over_gloo_store():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_queue_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    allc = sorted(i for _, mine in res for i in mine)
    assert allc == list(range(9))                                # every clip exactly once, drawn from one counter


def test_tta_shell_degenerate_equals_plain_and_flip_runs(tmp_path):
    """evaluate_clip_tta (evaluator.py:338-441) with a single un-flipped, un-scaled augmentation is the plain loop; with
    flip + two scales it runs one engine per augmentation (MultiRestrictSize order) and merges by probability averaging."""
    img_dir, lab_dir = write_clip(str(tmp_path), n_frames=4, new_at=2)
    ds = E.ClipDataset(img_dir, lab_dir)
    sd = O.make_state_dict("r50_deaotl", seed=1, sharpen=1.0)
    cfg = O.OracleConfig(model="r50_deaotl", former_mem_len=1, latter_mem_len=2)
    assert E.tta_augmentations(True, (1.0, 1.3)) == [(1.0, False), (1.0, True), (1.3, False), (1.3, True)]
    with torch.no_grad():
        plain = E.evaluate_clip(O.OracleEngine(sd, cfg), ds, keep_labels=True)
        one = E.evaluate_clip_tta([O.OracleEngine(sd, cfg)], ds, flip=False, multi_scale=(1.0,), keep_labels=True)
        assert one.frames == plain.frames == 3
        for a, b in zip(one.labels, plain.labels):
            assert torch.equal(a, b)
        engines = [O.OracleEngine(sd, cfg) for _ in range(4)]
        tta = E.evaluate_clip_tta(engines, ds, flip=True, multi_scale=(1.0, 1.3), keep_labels=True)
    assert tta.frames == 3 and all(l.shape == plain.labels[0].shape for l in tta.labels)
    assert engines[0].input_size_2d == (65, 81) and engines[2].input_size_2d == (81, 97)       # int(65*1.3)=84 -> 81, int(81*1.3)=105 -> 97 (np.around(6.5) = 6)
    assert (tta.labels[1][40:60, 50:70] == 2).all()             # new object pasted in at frame 2 (random weights: any id elsewhere)
    with pytest.raises(ValueError):
        E.evaluate_clip_tta(engines[:3], ds, flip=True, multi_scale=(1.0, 1.3))
