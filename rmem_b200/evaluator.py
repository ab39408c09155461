"""Per-clip evaluation shell over the engine API ("real-clip evaluator shell").

Host-side mirror of the reference's production caller:
  * clip dataset ............ aot_plus/dataloaders/eval_datasets.py:14-118  (VOSTest: object bookkeeping, label squeeze)
  * test-time resize ........ aot_plus/dataloaders/video_transforms.py:559-643 (MultiRestrictSize, single scale, no flip)
  * tensor conversion ....... aot_plus/dataloaders/video_transforms.py:646-666 (MultiToTensor: /255, ImageNet mean/std)
  * per-clip frame loop ..... aot_plus/networks/managers/evaluator.py:300-556  (reference frame, propagate, argmax,
                              new-object merge + re-reference, update_memory, long-term gap = max(round(N/30), 5))
  * mask writer ............. aot_plus/utils/image.py:89-105                   (un-squeeze object ids, palette PNG)
  * clip -> rank queue ...... aot_plus/networks/managers/evaluator.py:276-295  (dynamic queue; here an atomic counter in
                              the torch.distributed key-value store -- still no per-frame collective)

The engine is anything with the reference's method surface (`restart_engine`, `add_reference_frame`,
`match_propogate_one_frame`, `update_memory`, `long_term_mem_gap`, `input_size_2d`): the CUDA engine of
rmem_b200.engine in production, the CPU oracle in the host-logic tests.  Test-time augmentation (flip / multi-scale
engines with probability averaging, evaluator.py:338-441) is `evaluate_clip_tta`: one engine per augmentation, the
frames resized / normalised / flipped on the GPU (rmem_preprocess_fwd) and the per-augmentation logits merged by the
fused TTA head (rmem_tta_head_fwd) when the engines are CUDA engines.
"""
from __future__ import annotations

import os
import threading
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def davis_palette() -> List[int]:
    """The 256-entry palette of utils/image.py:6-74: the PASCAL-VOC bit-shuffle colours (with 191 where VOC has 192) for
    ids 0..21, a grey ramp (i, i, i) above."""
    pal: List[int] = []
    for i in range(256):
        if i >= 22:
            pal += [i, i, i]
            continue
        r = g = b = 0
        c = i
        for j in range(8):
            r |= ((c >> 0) & 1) << (7 - j)
            g |= ((c >> 1) & 1) << (7 - j)
            b |= ((c >> 2) & 1) << (7 - j)
            c >>= 3
        pal += [191 if v == 192 else v for v in (r, g, b)]
    return pal


def restrict_size(h: int, w: int, min_size: Optional[int] = None, max_size: Optional[float] = 800 * 1.3,
                  scale: float = 1.0, align_corners: bool = True, max_stride: int = 16) -> Tuple[int, int]:
    """Network input size for an h x w frame -- the rule of MultiRestrictSize (video_transforms.py:575-623; defaults =
    configs/default.py:110-112): cap the short edge at `min_size` or the long edge at `max_size` (only ever shrinking),
    apply `scale`, then snap each side to a multiple of the stride (+1 with align_corners).
    480 x 854 -> 481 x 849."""
    if min_size is not None and max_size is not None:
        raise ValueError("give min_size or max_size, not both")
    edge, limit = (min(h, w), min_size) if min_size is not None else (max(h, w), max_size)
    shrink = float(limit) / edge if (limit is not None and edge > limit) else None

    def side(v: int) -> int:
        v = int((v if shrink is None else shrink * v) * scale)
        off = 1 if align_corners else 0                 # sides are stride * k (+ 1): already-aligned values are fixed points
        return int(np.around((v - off) / max_stride) * max_stride + off)

    return side(h), side(w)


def long_term_gap(num_frames: int, no_memory_gap: bool = False) -> int:
    """evaluator.py:329-333: one long-term frame every max(round(N / 30), 5) frames."""
    gap = max(int(round(num_frames / 30)), 5)
    if no_memory_gap:
        gap = int(round(gap / 4))
    return gap


class ClipDataset:
    """One clip = a directory of frames plus label PNGs for the frames that introduce objects (VOSTest,
    eval_datasets.py:14-118).  `obj_indices[i]` lists the dataset object ids known up to frame i (0 = background);
    labels are squeezed to 0..n in order of first appearance, exactly like `read_label(..., squeeze_idx)`."""

    def __init__(self, image_dir: str, label_dir: str, images: Optional[Sequence[str]] = None,
                 labels: Optional[Sequence[str]] = None, rgb: bool = True, resolution: Optional[int] = None,
                 min_size: Optional[int] = None, max_size: Optional[float] = 800 * 1.3, seq_name: Optional[str] = None):
        self.image_dir, self.label_dir = image_dir, label_dir
        self.seq_name = seq_name or os.path.basename(os.path.normpath(image_dir))
        self.images = sorted(images if images is not None else
                             [f for f in os.listdir(image_dir) if f.lower().endswith((".jpg", ".jpeg", ".png"))])
        self.labels = set(labels if labels is not None else
                          [f for f in os.listdir(label_dir) if f.lower().endswith(".png")])
        self.rgb, self.resolution = rgb, resolution
        self.min_size, self.max_size = min_size, max_size
        assert len(self.images) >= 2, "a clip needs a reference frame and at least one frame to propagate to"
        # objects known before each frame (VOSTest, eval_datasets.py:37-52): ids enter in order of first appearance, a
        # frame's own label only counts from that frame on, and frame 0 reports the count valid at frame 1
        known: List[int] = [0]
        self.obj_nums: List[int] = []
        self.obj_indices: List[List[int]] = []
        for name in self.images:
            self.obj_nums.append(len(known) - 1)
            png = os.path.splitext(name)[0] + ".png"
            if png in self.labels:
                known += [int(v) for v in np.unique(self._read_png(png)) if int(v) not in known]
            self.obj_indices.append(list(known))
        self.obj_nums[0] = self.obj_nums[1]

    def __len__(self):
        return len(self.images)

    def _read_png(self, name: str) -> np.ndarray:
        from PIL import Image
        return np.array(Image.open(os.path.join(self.label_dir, name)), dtype=np.uint8)

    def read_label(self, name: str, squeeze_idx: Sequence[int]) -> np.ndarray:
        """Dataset ids -> 0..n by position in `squeeze_idx` (ids not listed, and id 0, map to 0)."""
        lut = np.zeros(256, np.uint8)
        for pos, obj_id in enumerate(squeeze_idx):
            if obj_id != 0:
                lut[obj_id] = pos
        return lut[self._read_png(name)]

    def read_image_u8(self, idx: int) -> np.ndarray:
        """The frame as decoded (uint8, cv2.imread channel order): input of the GPU preprocessing."""
        import cv2
        img = cv2.imread(os.path.join(self.image_dir, self.images[idx]))
        if img is None:
            raise FileNotFoundError(os.path.join(self.image_dir, self.images[idx]))
        return img

    def read_image(self, idx: int) -> np.ndarray:
        import cv2
        img = cv2.imread(os.path.join(self.image_dir, self.images[idx]))
        if img is None:
            raise FileNotFoundError(os.path.join(self.image_dir, self.images[idx]))
        img = np.array(img, dtype=np.float32)
        return img[:, :, [2, 1, 0]] if self.rgb else img

    def sample_meta(self, idx: int) -> Dict:
        """`meta` and `current_label` of __getitem__ without the frame itself (the GPU preprocessing reads it raw)."""
        from PIL import Image
        with Image.open(os.path.join(self.image_dir, self.images[idx])) as im:      # header only
            width, height = im.size
        if self.resolution is not None:
            width = int(np.ceil(float(width) * self.resolution / float(height)))
            height = int(self.resolution)
        sample: Dict = {}
        lab = os.path.splitext(self.images[idx])[0] + ".png"
        if lab in self.labels:
            sample["current_label"] = torch.from_numpy(self.read_label(lab, self.obj_indices[idx])).int()[None, None]
        sample["meta"] = {"seq_name": self.seq_name, "frame_num": len(self.images), "obj_num": self.obj_nums[idx],
                          "current_name": self.images[idx], "height": height, "width": width, "flip": False,
                          "obj_idx": self.obj_indices[idx]}
        return sample

    def __getitem__(self, idx: int) -> Dict:
        img = self.read_image(idx)
        height, width = img.shape[:2]
        if self.resolution is not None:          # output size only (eval_datasets.py:87-90); the frame is not resized here
            width = int(np.ceil(float(width) * self.resolution / float(height)))
            height = int(self.resolution)
        nh, nw = restrict_size(img.shape[0], img.shape[1], self.min_size, self.max_size)
        if (nh, nw) != img.shape[:2]:
            import cv2
            img = cv2.resize(img, dsize=(nw, nh), interpolation=cv2.INTER_CUBIC)
        img = img / 255.
        img -= IMAGENET_MEAN
        img /= IMAGENET_STD
        sample = {"current_img": torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1))).float()[None]}
        lab = os.path.splitext(self.images[idx])[0] + ".png"
        if lab in self.labels:
            sample["current_label"] = torch.from_numpy(self.read_label(lab, self.obj_indices[idx])).int()[None, None]
        sample["meta"] = {"seq_name": self.seq_name, "frame_num": len(self.images), "obj_num": self.obj_nums[idx],
                          "current_name": self.images[idx], "height": height, "width": width, "flip": False,
                          "obj_idx": self.obj_indices[idx]}
        return sample


def save_mask(mask: np.ndarray, path: str, squeeze_idx: Optional[Sequence[int]] = None, background: bool = True):
    """utils/image.py:89-105: map 0..n back to the dataset's object ids, write a palettised PNG (on a writer thread
    when `background`, like the reference's fire-and-forget thread)."""
    mask = np.asarray(mask).astype(np.uint8)

    def _write():
        from PIL import Image
        m = mask
        if squeeze_idx is not None:                   # 0..n back to the dataset's ids; anything above n becomes 0
            lut = np.zeros(256, np.uint8)
            lut[1:len(squeeze_idx)] = np.asarray(squeeze_idx[1:], np.uint8)
            m = lut[m]
        im = Image.fromarray(m).convert("P")
        im.putpalette(davis_palette())
        im.save(path)

    if background:
        t = threading.Thread(target=_write)
        t.start()
        return t
    _write()
    return None


@dataclass
class ClipResult:
    seq_name: str
    frames: int = 0                       # propagated frames (the reference frame is not timed, evaluator.py:399-404)
    seconds: float = 0.0
    labels: List[torch.Tensor] = field(default_factory=list)      # uint8 [Ho, Wo] per propagated frame (if keep_labels)
    paths: List[str] = field(default_factory=list)


def evaluate_clip(engine, dataset: ClipDataset, out_dir: Optional[str] = None, device=None, keep_labels: bool = False,
                  no_memory_gap: bool = False, timer: Optional[Callable[[], float]] = None,
                  on_frame: Optional[Callable[[int, torch.Tensor], None]] = None) -> ClipResult:
    """evaluator.py:300-556 for one clip and one (un-augmented) engine.  Frame 0 must carry a label."""
    res = ClipResult(dataset.seq_name)
    engine.restart_engine()
    engine.long_term_mem_gap = long_term_gap(len(dataset), no_memory_gap)
    if out_dir is not None:
        os.makedirs(os.path.join(out_dir, dataset.seq_name), exist_ok=True)
    use_cuda = device is not None and torch.device(device).type == "cuda"
    writers = []
    timers = []
    now = timer or time.perf_counter
    # Host -> device copies of the frames run on their own stream (a 4.9 MB copy in the caller's stream sat between two
    # frames of the propagation chain); an event per frame orders its consumers.
    h2d_stream = torch.cuda.Stream(device=device) if use_cuda else None

    def fetch(i):
        smp = dataset[i]
        im, lb = smp["current_img"], smp.get("current_label")
        ev = None
        if use_cuda:
            with torch.cuda.stream(h2d_stream):
                im = im.to(device, non_blocking=True).contiguous()
                lb = lb.to(device, non_blocking=True).float() if lb is not None else None
                ev = torch.cuda.Event()
                ev.record()
            cur_stream = torch.cuda.current_stream()
            im.record_stream(cur_stream)                   # freed only once the consuming stream is done with it
            if lb is not None:
                lb.record_stream(cur_stream)
        elif device is not None:
            im = im.to(device).contiguous()
            lb = lb.to(device).float() if lb is not None else None
        elif lb is not None:
            lb = lb.float()
        return smp["meta"], im, lb, ev

    can_prefetch = use_cuda and hasattr(engine, "prefetch")
    d2h_stream = torch.cuda.Stream(device=device) if use_cuda else None
    fast = use_cuda and hasattr(engine, "propagate_label")
    ring: List = [None] * 4                # (copy event, pinned label buffer, frame index, meta) of the frames in flight
    ring_bufs: List = [None] * 4

    def emit(frame_idx, meta, lab8):
        if on_frame is not None:
            on_frame(frame_idx, lab8)
        if keep_labels:
            res.labels.append(lab8.clone())
        if out_dir is not None:
            path = os.path.join(out_dir, dataset.seq_name, os.path.splitext(meta["current_name"])[0] + ".png")
            writers.append(save_mask(lab8.numpy().copy(), path, meta["obj_idx"]))
            res.paths.append(path)

    def drain(entry):
        done, buf, fi, mt = entry
        done.synchronize()
        emit(fi, mt, buf)

    def flush_ring():
        for e in sorted((e for e in ring if e is not None), key=lambda e: e[2]):
            drain(e)
        for k in range(len(ring)):
            ring[k] = None

    # Frames are known ahead of time: up to 2G - 1 are fetched (and copied to the device) ahead of the one being propagated,
    # and their image encoder runs on the engine's side stream meanwhile -- G frames per encoder pass (engine.prefetch_n:
    # frames i+G .. i+2G-1, every G-th frame; G = 2, RMEM_EVAL_ENC_GROUP = 1 | 2 | 4 -- 4 measured no faster than 2), a single frame (engine.prefetch) where
    # no group covers it (clip start, clip end).  Same-size frames only: a size change rebuilds the engine.
    EG = int(os.environ.get("RMEM_EVAL_ENC_GROUP", "2")) if (can_prefetch and hasattr(engine, "prefetch_n")) else 1
    EG = EG if EG in (2, 4) else 1
    look: List = []                        # fetched frames frame_idx + 1 ..
    fetched = 0
    covered = set()                        # frames whose encoding has been issued

    def ensure(n):
        nonlocal fetched
        while len(look) < n and fetched < len(dataset):
            look.append(fetch(fetched))
            fetched += 1

    for frame_idx in range(len(dataset)):
        ensure(1)
        meta, img, label, copied = look.pop(0)
        if copied is not None:
            torch.cuda.current_stream().wait_event(copied)
        ensure(2 * EG - 1 if EG > 1 else 1)
        if can_prefetch and frame_idx >= 1:
            # (stream = the copy stream: the encoder waits for the copies issued so far, not the caller's stream)
            if look and frame_idx + 1 not in covered and look[0][1].shape == img.shape:
                engine.prefetch(look[0][1], stream=h2d_stream)
                covered.add(frame_idx + 1)
            if EG > 1 and len(look) >= 2 * EG - 1 and frame_idx + EG not in covered and \
                    all(look[EG - 1 + j][1].shape == img.shape for j in range(EG)):
                engine.prefetch_n([look[EG - 1 + j][1] for j in range(EG)], stream=h2d_stream)
                covered.update(range(frame_idx + EG, frame_idx + 2 * EG))
        if frame_idx == 0:
            if label is None:
                raise ValueError(f"{dataset.seq_name}: the first frame has no label")
            ref = F.interpolate(label, size=img.shape[2:], mode="nearest").int()
            engine.add_reference_frame(img, ref, obj_nums=[int(meta["obj_num"])], frame_step=0)
            continue
        # per-frame timing like the reference (evaluator.py:399-404, 525-527): CUDA events around propagate + update,
        # read after the clip -- no device-wide sync inside the loop, so the prefetched encoder really overlaps
        if use_cuda:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        else:
            t0 = now()
        out_size = (int(meta["height"]), int(meta["width"]))
        if fast and label is None:
            # CUDA engine, ordinary frame: the mask-ID assignment (softmax -> argmax, evaluator.py:430-441) is fused into the
            # engine's mask head; only the uint8 label map leaves the GPU, through a ring of pinned buffers and without a
            # per-frame synchronisation (consumers see a frame once its copy event has completed)
            lab_dev = engine.propagate_label(img, output_size=out_size)
            cur = lab_dev if out_size == tuple(engine.input_size_2d) else \
                F.interpolate(lab_dev.float(), size=engine.input_size_2d, mode="nearest")
            engine.update_memory(cur)
            ev1.record()
            timers.append((ev0, ev1))
            res.frames += 1
            slot = res.frames % len(ring)
            if ring[slot] is not None:
                drain(ring[slot])
            buf = ring_bufs[slot] if ring_bufs[slot] is not None and ring_bufs[slot].shape == lab_dev.shape[2:] else \
                torch.empty(lab_dev.shape[2:], dtype=torch.uint8).pin_memory()
            ring_bufs[slot] = buf
            # the label map leaves on its own stream: a device->host copy in the caller's stream would sit between two
            # frames of the propagation chain
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(ready)
                buf.copy_(lab_dev[0, 0], non_blocking=True)
                lab_dev.record_stream(d2h_stream)
                done = torch.cuda.Event()
                done.record()
            ring[slot] = (done, buf, frame_idx, meta)
            continue
        logit = engine.match_propogate_one_frame(img, output_size=out_size)
        prob = torch.softmax(logit, dim=1)
        pred = torch.argmax(prob, dim=1, keepdim=True).float()
        if label is not None:              # a frame that introduces objects: paste them in and re-reference
            keep = (label == 0).float()
            pred = pred * keep + label * (1 - keep)
            new_obj_nums = [int(pred.max().item())]
            cur = F.interpolate(pred, size=engine.input_size_2d, mode="nearest")
            engine.add_reference_frame(img, cur, obj_nums=new_obj_nums, frame_step=frame_idx)
        else:
            cur = F.interpolate(pred, size=engine.input_size_2d, mode="nearest")
            engine.update_memory(cur)
        if use_cuda:
            ev1.record()
            timers.append((ev0, ev1))
        else:
            res.seconds += now() - t0
        res.frames += 1
        flush_ring()                                         # keep the output order: frames still in flight first
        emit(frame_idx, meta, pred[0, 0].to(torch.uint8).cpu())
    flush_ring()
    if timers:
        torch.cuda.synchronize()
        res.seconds += sum(a.elapsed_time(b) for a, b in timers) / 1e3
    for t in writers:
        if t is not None:
            t.join()
    return res


def tta_augmentations(flip: bool, multi_scale: Sequence[float]) -> List[Tuple[float, bool]]:
    """(scale, flipped) in the order MultiRestrictSize emits its samples (video_transforms.py:575-650)."""
    out: List[Tuple[float, bool]] = []
    for sc in multi_scale:
        out.append((float(sc), False))
        if flip:
            out.append((float(sc), True))
    return out


def evaluate_clip_tta(engines: Sequence, dataset: ClipDataset, flip: bool = True, multi_scale: Sequence[float] = (1.0,),
                      out_dir: Optional[str] = None, device=None, keep_labels: bool = False,
                      gpu_preprocess: Optional[bool] = None, no_memory_gap: bool = False) -> ClipResult:
    """evaluator.py:300-556 with test-time augmentation: `engines[a]` serves augmentation a of `tta_augmentations`
    (the reference deep-copies the model per augmentation, :342-352; engines built on one RmemModel share the weight
    blob).  Per frame every engine propagates its own resized / flipped view, the soft-maxed logits are flipped back
    and averaged (:426-441), the argmax label (flipped again where needed) refreshes every engine's memory (:484-522)."""
    augs = tta_augmentations(flip, multi_scale)
    if len(engines) != len(augs):
        raise ValueError(f"{len(augs)} augmentations need {len(augs)} engines, got {len(engines)}")
    res = ClipResult(dataset.seq_name)
    gap = long_term_gap(len(dataset), no_memory_gap)
    for e in engines:
        e.restart_engine()
        e.long_term_mem_gap = gap
    use_cuda = device is not None and torch.device(device).type == "cuda"
    cuda_engines = use_cuda and all(hasattr(e, "propagate_only") for e in engines)
    if gpu_preprocess is None:
        gpu_preprocess = cuda_engines
    if out_dir is not None:
        os.makedirs(os.path.join(out_dir, dataset.seq_name), exist_ok=True)
    writers = []
    t_sum = 0.0
    timers = []

    def views(idx):
        """[(img fp32 [1,3,nh,nw] on `device`, flipped)] for every augmentation + (meta, label)."""
        meta_smp = dataset[idx] if not gpu_preprocess else None
        out = []
        if gpu_preprocess:
            from . import ops as K
            raw = torch.from_numpy(dataset.read_image_u8(idx)).to(device, non_blocking=True)
            H0, W0 = raw.shape[:2]
            for sc, fl in augs:
                nh, nw = restrict_size(H0, W0, dataset.min_size, dataset.max_size, scale=sc)
                out.append((K.preprocess(raw, nh, nw, bgr=dataset.rgb, flip=fl), fl))
            smp = dataset.sample_meta(idx)
        else:
            import cv2
            base = dataset.read_image(idx)
            H0, W0 = base.shape[:2]
            for sc, fl in augs:
                nh, nw = restrict_size(H0, W0, dataset.min_size, dataset.max_size, scale=sc)
                im = base if (nh, nw) == (H0, W0) else cv2.resize(base, dsize=(nw, nh), interpolation=cv2.INTER_CUBIC)
                if fl:
                    im = im[:, ::-1]
                im = (im / 255. - IMAGENET_MEAN) / IMAGENET_STD
                t = torch.from_numpy(np.ascontiguousarray(im.transpose(2, 0, 1))).float()[None]
                out.append((t.to(device) if device is not None else t, fl))
            smp = {"meta": meta_smp["meta"], "current_label": meta_smp.get("current_label")}
        lab = smp.get("current_label")
        if lab is not None:
            lab = lab.float()
            if device is not None:
                lab = lab.to(device)
        return smp["meta"], out, lab

    def to_engine(label, eng, fl):
        """Label map [1,1,Ho,Wo] -> the engine's input size, mirrored for a flipped augmentation."""
        if fl:
            label = torch.flip(label, dims=(3,))
        return F.interpolate(label.float(), size=eng.input_size_2d, mode="nearest")

    for frame_idx in range(len(dataset)):
        meta, imgs, label = views(frame_idx)
        out_size = (int(meta["height"]), int(meta["width"]))
        if frame_idx == 0:
            if label is None:
                raise ValueError(f"{dataset.seq_name}: the first frame has no label")
            for (img, fl), eng in zip(imgs, engines):
                lab = torch.flip(label, dims=(3,)) if fl else label
                ref = F.interpolate(lab, size=img.shape[2:], mode="nearest").int()
                eng.add_reference_frame(img, ref, obj_nums=[int(meta["obj_num"])], frame_step=0)
            continue
        if use_cuda:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        else:
            t0 = time.perf_counter()
        if cuda_engines:
            from . import ops as K
            for (img, fl), eng in zip(imgs, engines):
                eng.propagate_only(img)
            _, lab8 = K.tta_head([eng.logits4_views() for eng in engines], [fl for _, fl in imgs], *out_size)
            pred = lab8.float()[None, None]
        else:
            probs = []
            for (img, fl), eng in zip(imgs, engines):
                lg = eng.match_propogate_one_frame(img, output_size=out_size)
                if fl:
                    lg = torch.flip(lg, dims=(3,))
                probs.append(torch.softmax(lg, dim=1))
            pred = torch.argmax(torch.mean(torch.cat(probs, dim=0), dim=0, keepdim=True), dim=1, keepdim=True).float()
        if label is not None:              # a frame that introduces objects: paste them in and re-reference every engine
            keep = (label == 0).float()
            pred = pred * keep + label * (1 - keep)
            new_obj_nums = [int(pred.max().item())]
            for (img, fl), eng in zip(imgs, engines):
                eng.add_reference_frame(img, to_engine(pred, eng, fl), obj_nums=new_obj_nums, frame_step=frame_idx)
        else:
            for (img, fl), eng in zip(imgs, engines):
                eng.update_memory(to_engine(pred, eng, fl))
        if use_cuda:
            ev1.record()
            timers.append((ev0, ev1))
        else:
            t_sum += time.perf_counter() - t0
        res.frames += 1
        lab_host = pred[0, 0].to(torch.uint8).cpu()
        if keep_labels:
            res.labels.append(lab_host)
        if out_dir is not None:
            path = os.path.join(out_dir, dataset.seq_name, os.path.splitext(meta["current_name"])[0] + ".png")
            writers.append(save_mask(lab_host.numpy(), path, meta["obj_idx"]))
            res.paths.append(path)
    if timers:
        torch.cuda.synchronize()
        t_sum += sum(a.elapsed_time(b) for a, b in timers) / 1e3
    res.seconds = t_sum
    for t in writers:
        if t is not None:
            t.join()
    return res


class ClipQueue:
    """Clip -> rank assignment.  With a torch.distributed store every rank draws the next clip index from one atomic
    counter (the reference's dynamic multiprocessing queue, evaluator.py:276-295); without one, static round-robin."""

    def __init__(self, n_clips: int, rank: int = 0, world: int = 1, store=None, key: str = "rmem_next_clip"):
        self.n, self.rank, self.world, self.store, self.key = n_clips, rank, world, store, key
        self._static = iter(range(rank, n_clips, world))

    def __iter__(self):
        return self

    def __next__(self) -> int:
        if self.store is None:
            return next(self._static)
        i = int(self.store.add(self.key, 1)) - 1
        if i >= self.n:
            raise StopIteration
        return i


def evaluate_clips(engine, clips: Sequence[ClipDataset], out_dir: Optional[str] = None, device=None, rank: int = 0,
                   world: int = 1, store=None, log: Optional[Callable[[str], None]] = print,
                   extra_engines: Sequence = ()) -> Dict:
    """evaluator.py:265-613: every rank pulls clips until the queue is empty; (frames, seconds) are gathered once at
    the end.  Returns this rank's per-clip results and the job-wide all-frame FPS.

    extra_engines: further engines of this rank (same model, own memory bank: `build_engine(..., aot_model=model)`) --
    each one evaluates its own clip on its own host thread and CUDA stream, all drawing from the same queue.  One clip's
    frame is a chain of small latency-bound launches; a second clip in flight fills the SMs it leaves idle (+16 % job
    throughput at c3 with one extra engine; per-clip latency rises accordingly)."""
    from .sharding import gather_stats
    results: List[ClipResult] = []
    engines = [engine, *extra_engines]
    if len(engines) == 1:
        for i in ClipQueue(len(clips), rank, world, store):
            r = evaluate_clip(engine, clips[i], out_dir=out_dir, device=device)
            results.append(r)
            if log:
                log(f"rank {rank} - Seq {r.seq_name} [{i + 1}/{len(clips)}] - FPS: {r.frames / max(r.seconds, 1e-9):.2f}")
    else:
        lock = threading.Lock()
        queue = ClipQueue(len(clips), rank, world, store)
        errors: List[BaseException] = []

        def next_clip():
            with lock:
                try:
                    return next(queue)
                except StopIteration:
                    return None

        def worker(eng):
            try:
                if device is not None and torch.device(device).type == "cuda":
                    torch.cuda.set_device(device)
                    ctx = torch.cuda.stream(torch.cuda.Stream(device=device))
                else:
                    import contextlib
                    ctx = contextlib.nullcontext()
                with ctx:
                    while True:
                        i = next_clip()
                        if i is None:
                            break
                        r = evaluate_clip(eng, clips[i], out_dir=out_dir, device=device)
                        with lock:
                            results.append(r)
                            if log:
                                log(f"rank {rank} - Seq {r.seq_name} [{i + 1}/{len(clips)}] - FPS: "
                                    f"{r.frames / max(r.seconds, 1e-9):.2f}")
                    if device is not None and torch.device(device).type == "cuda":
                        torch.cuda.current_stream().synchronize()
            except BaseException as ex:  # noqa: BLE001 -- re-raised on the calling thread
                errors.append(ex)

        threads = [threading.Thread(target=worker, args=(e,)) for e in engines]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
    frames = sum(r.frames for r in results)
    seconds = sum(r.seconds for r in results)
    stats = gather_stats(frames, seconds, device if device is not None else "cpu", world)
    tot_f = sum(f for f, _ in stats)
    tot_s = sum(s for _, s in stats)
    return {"results": results, "frames": frames, "seconds": seconds, "all_frames": tot_f,
            "all_frame_fps": tot_f / max(tot_s, 1e-9), "per_rank": stats}
