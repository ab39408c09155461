"""Python surface of the B200-native RMem engine -- a drop-in for the reference's

    build_engine(...)                                        networks/engines/__init__.py:5-21
    AOTInferEngine / DeAOTInferEngine                        networks/engines/aot_engine.py:571-725,
      .restart_engine / .add_reference_frame /                 deaot_engine.py:20-56
      .match_propogate_one_frame / .update_memory
    AOTEngine.{long_memories_indexes, pred_id_logits}        networks/engines/aot_engine.py:533-569

(same names, argument meaning and state attributes; paths relative to /root/reference/aot_plus/).  All the
arithmetic runs in hand-written sm_100a CUDA behind the C ABI of include/rmem_b200.h; torch only owns device
memory and the stream.  There is no CPU fallback: without the extension and a GPU these classes raise.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

import torch

from . import _capi
from .weights import WeightBlob, pack_model

MAX_OBJ = 10   # configs/models/default.py MODEL_MAX_OBJ_NUM


@dataclass
class RmemConfig:
    """The frozen subset of configs/models/r50_deaotl.py + configs/pre_vost.py this path depends on."""
    model: str = "r50_deaotl"
    former_mem_len: int = 1          # FORMER_MEM_LEN
    # LATTER_MEM_LEN: 8 = what the reference ships (configs/models/r50_deaotl.py:8, eval_vost.sh `--latter_mem_len 8`:
    # bank capacity 1 + 8 = 9 frames).  BASELINE.json's "T=8" workloads (bench.py, the c3 / c4 tests) pass 7 explicitly.
    latter_mem_len: int = 8
    max_obj_num: int = MAX_OBJ
    attn_impl: int = _capi.ATTN_TC4   # column kernel for the long-term attention, tc3 pair kernel for everything else
    max_engines: int = 4
    # ablation knobs (configs/models/r50_deaotl.py:9-28; all False in the shipped configs, see include/rmem_b200.h)
    no_long_memory: bool = False     # NO_LONG_MEMORY: the long-term bank stays at the reference frame
    reverse_infer: bool = False      # REVERSE_INFER: no effect at inference (training-loss pass only); accepted
    time_encode: bool = False        # TIME_ENCODE / TIME_ENCODE_NORM: no effect at inference (stored, never read); accepted
    gru_memory: bool = False         # GRU_MEMORY (r50_aotl only; needs the memory_grus weights in the state dict)


MODEL_IDS = {"r50_deaotl": 0, "r50_aotl": 1}


class RmemModel:
    """Weights of R50_DeAOTL / R50_AOTL resident in HBM (stands in for `build_vos_model(...)` + `load_state_dict`,
    networks/models/__init__.py:5, networks/models/{aot,deaot}.py)."""

    def __init__(self, state_dict, cfg: Optional[RmemConfig] = None, device: Union[str, torch.device] = "cuda:0"):
        self.cfg = cfg or RmemConfig()
        if self.cfg.model not in MODEL_IDS:
            raise NotImplementedError(f"model {self.cfg.model!r}: r50_deaotl and r50_aotl are built")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _capi.RmemError("rmem_b200 runs on CUDA devices only (no CPU path)")
        _capi.load()
        self.weights = WeightBlob(pack_model(state_dict, self.cfg.model), self.device)

    def eval(self):
        return self


DeAOTModel = RmemModel


def AOTModel(state_dict, cfg: Optional[RmemConfig] = None, device="cuda:0") -> RmemModel:
    cfg = cfg or RmemConfig(model="r50_aotl")
    cfg.model = "r50_aotl"
    return RmemModel(state_dict, cfg, device)


def build_vos_model(name: str, cfg: RmemConfig, state_dict=None, device="cuda:0") -> RmemModel:
    """networks/models/__init__.py:5-13: 'aot' | 'deaot'."""
    if name in ("deaot", "r50_deaotl"):
        cfg.model = "r50_deaotl"
    elif name in ("aot", "r50_aotl"):
        cfg.model = "r50_aotl"
    else:
        raise NotImplementedError(name)
    return RmemModel(state_dict, cfg, device)


class _SubEngineView:
    """Read-only view of one object group's state: mirrors the AOTEngine attributes callers/tests read."""

    def __init__(self, owner: "DeAOTInferEngine", index: int):
        self._o, self._i = owner, index

    @property
    def long_memories_indexes(self) -> List[int]:
        lib = _capi.load()
        idx = (C.c_int * (_capi.MAX_BANK_FRAMES + 1))()
        n = C.c_int(0)
        _capi.check(lib.rmem_engine_long_indexes(self._o._h, self._i, idx, C.byref(n)))
        return [idx[i] for i in range(n.value)]

    @property
    def pred_id_logits(self) -> torch.Tensor:
        """[1,11,H/4,W/4] fp32 logits of the last decode (a copy; the engine reuses its buffer)."""
        lib = _capi.load()
        p = C.c_void_p()
        h4, w4 = C.c_int(), C.c_int()
        _capi.check(lib.rmem_engine_pred_logits(self._o._h, self._i, C.byref(p), C.byref(h4), C.byref(w4)))
        n = 11 * h4.value * w4.value
        off = p.value - self._o._arena.data_ptr()
        return self._o._arena[off:off + 4 * n].view(torch.float32).view(1, 11, h4.value, w4.value).clone()

    def layer_memory(self, layer: int) -> dict:
        """Views (no copies) of one GPM layer's memories after update_memory: the restricted long-term bank in the
        engine's layout (kbank [nslots,HWp,128], vtbank [1024,nslots*HWp]), the logical->physical slot list, and the
        short-term memory (last frame's K = Q [HW,128], V||ID_V [HW,1024])."""
        lib = _capi.load()
        kb, vb, ql, vl = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        ns, hwp, T = C.c_int(), C.c_int(), C.c_int()
        slots = (C.c_int * _capi.MAX_BANK_FRAMES)()
        _capi.check(lib.rmem_engine_layer_memory(self._o._h, self._i, int(layer), C.byref(kb), C.byref(vb), C.byref(ql),
                                                 C.byref(vl), C.byref(ns), C.byref(hwp), C.byref(T), slots))
        arena, base = self._o._arena, self._o._arena.data_ptr()
        HW = self._o.enc_hw
        op = _capi.op_dtype()

        def view(p, n, shape):
            off = p.value - base
            return arena[off:off + 2 * n].view(op).view(*shape)
        return dict(kbank=view(kb, ns.value * hwp.value * 128, (ns.value, hwp.value, 128)),
                    vtbank=view(vb, 1024 * ns.value * hwp.value, (1024, ns.value * hwp.value)),
                    q_last=view(ql, HW * 128, (HW, 128)), vid_last=view(vl, HW * 1024, (HW, 1024)),
                    slots=[slots[i] for i in range(T.value)], nslots=ns.value, HWp=hwp.value)

    @property
    def last_evict(self) -> Tuple[List[float], int]:
        lib = _capi.load()
        rel = (C.c_float * _capi.MAX_BANK_FRAMES)()
        n, drop = C.c_int(), C.c_int()
        _capi.check(lib.rmem_engine_last_evict(self._o._h, self._i, rel, C.byref(n), C.byref(drop)))
        return [rel[i] for i in range(n.value)], drop.value


class DeAOTInferEngine:
    def __init__(self, aot_model: DeAOTModel, gpu_id: int = 0, long_term_mem_gap: int = 9999,
                 max_aot_obj_num: Optional[int] = None):
        self.AOT = aot_model
        self.cfg = aot_model.cfg
        self.gpu_id = gpu_id
        self.device = aot_model.device
        self.long_term_mem_gap = long_term_mem_gap
        self.max_aot_obj_num = max_aot_obj_num or self.cfg.max_obj_num
        self._h = None
        self._arena = None
        self._size = None
        self.restart_engine()

    # ---- lifetime -----------------------------------------------------------------------------
    def eval(self):
        return self

    def __del__(self):
        self._destroy()

    def _destroy(self):
        if getattr(self, "_h", None):
            _capi.load().rmem_engine_destroy(self._h)
            self._h = None

    def _ensure(self, H: int, W: int):
        if self._h is not None and self._size == (H, W):
            return
        self._destroy()
        lib = _capi.load()
        cc = _capi.EngineConfig(MODEL_IDS[self.cfg.model], H, W, self.cfg.former_mem_len, self.cfg.latter_mem_len, self.cfg.max_engines,
                                self.cfg.attn_impl, max(int(self.long_term_mem_gap), 1), int(self.cfg.no_long_memory),
                                int(self.cfg.reverse_infer), int(self.cfg.time_encode), int(self.cfg.gru_memory))
        nbytes = C.c_size_t()
        _capi.check(lib.rmem_engine_arena_bytes(C.byref(cc), C.byref(nbytes)))
        self._arena = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        h = C.c_void_p()
        w = self.AOT.weights
        _capi.check(lib.rmem_engine_create(C.byref(cc), _capi.ptr(w.blob), w.entries, w.n_entries,
                                           _capi.ptr(self._arena), C.c_size_t(nbytes.value), C.byref(h)))
        self._h = h
        self._size = (H, W)

    # ---- reference API ------------------------------------------------------------------------
    def restart_engine(self):
        """aot_engine.py:598-602."""
        if self._h is not None:
            _capi.check(_capi.load().rmem_engine_restart(self._h))
        self.aot_engines: List[_SubEngineView] = []
        self.input_size_2d = None
        self.enc_size_2d = None
        self.enc_hw = None

    def _dev(self, t: torch.Tensor, dtype=None) -> torch.Tensor:
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        if t.device != self.device:
            t = t.to(self.device, non_blocking=True)
        return t.contiguous()

    def _label_arg(self, mask: torch.Tensor):
        """Reference callers pass int (reference frame) or float-valued (update_memory) label maps [1,1,H,W]."""
        if mask.dtype == torch.uint8:
            return self._dev(mask), 0
        return self._dev(mask, torch.float32), 1

    def add_reference_frame(self, img: torch.Tensor, mask: torch.Tensor, obj_nums, frame_step: int = -1):
        """aot_engine.py:675-702 / deaot_engine.py:30-56."""
        if isinstance(obj_nums, (list, tuple)):
            obj_nums = obj_nums[0]
        assert img.shape[0] == 1, "the inference engine runs batch 1 (transformer.py:1190 asserts the same)"
        H, W = int(img.shape[-2]), int(img.shape[-1])
        self._ensure(H, W)
        lib = _capi.load()
        _capi.check(lib.rmem_engine_set_gap(self._h, max(int(self.long_term_mem_gap), 1)))
        img = self._dev(img, torch.float32)
        lab, is_f32 = self._label_arg(mask)
        assert lab.numel() == H * W, "label map must match the frame size"
        _capi.check(lib.rmem_engine_add_reference_frame(self._h, _capi.ptr(img), _capi.ptr(lab), is_f32,
                                                        int(obj_nums), int(frame_step), _capi.stream_ptr()))
        n = lib.rmem_engine_num_groups(self._h)
        self.aot_engines = [_SubEngineView(self, i) for i in range(n)]
        self.input_size_2d = (H, W)
        self.enc_size_2d = ((H - 1) // 16 + 1, (W - 1) // 16 + 1)
        self.enc_hw = self.enc_size_2d[0] * self.enc_size_2d[1]

    def match_propogate_one_frame(self, img: torch.Tensor = None, mask=None, output_size=None,
                                  return_label: bool = False):
        """aot_engine.py:704-712.  Returns logits [1, 1+10k, Ho, Wo] fp32 (and the uint8 argmax label map
        [1,1,Ho,Wo] of evaluator.py:430-441 when return_label)."""
        lib = _capi.load()
        _capi.check(lib.rmem_engine_set_gap(self._h, max(int(self.long_term_mem_gap), 1)))
        img = self._dev(img, torch.float32)
        Ho, Wo = output_size if output_size is not None else self.input_size_2d
        k = len(self.aot_engines)
        logits = torch.empty(1, 1 + MAX_OBJ * k, Ho, Wo, dtype=torch.float32, device=self.device)
        label = torch.empty(1, 1, Ho, Wo, dtype=torch.uint8, device=self.device) if return_label else None
        _capi.check(lib.rmem_engine_propagate(self._h, _capi.ptr(img), int(Ho), int(Wo), _capi.ptr(logits),
                                              _capi.ptr(label), _capi.stream_ptr()))
        return (logits, label) if return_label else logits

    def propagate_label(self, img: torch.Tensor, output_size=None) -> torch.Tensor:
        """Fast path: mask IDs only (no full-resolution logits written).  uint8 [1,1,Ho,Wo]."""
        lib = _capi.load()
        _capi.check(lib.rmem_engine_set_gap(self._h, max(int(self.long_term_mem_gap), 1)))
        img = self._dev(img, torch.float32)
        Ho, Wo = output_size if output_size is not None else self.input_size_2d
        label = torch.empty(1, 1, Ho, Wo, dtype=torch.uint8, device=self.device)
        _capi.check(lib.rmem_engine_propagate(self._h, _capi.ptr(img), int(Ho), int(Wo), None, _capi.ptr(label),
                                              _capi.stream_ptr()))
        return label

    def propagate_only(self, img: torch.Tensor):
        """Run the propagation of one frame and leave the 1/4-res logits of every object group in the engine
        (`logits4_views`): the test-time-augmentation head merges several engines' logits itself."""
        lib = _capi.load()
        _capi.check(lib.rmem_engine_set_gap(self._h, max(int(self.long_term_mem_gap), 1)))
        img = self._dev(img, torch.float32)
        _capi.check(lib.rmem_engine_propagate(self._h, _capi.ptr(img), 0, 0, None, None, _capi.stream_ptr()))

    def logits4_views(self) -> List[torch.Tensor]:
        """Views (no copy) of the fp32 [11,H/4,W/4] logits of the last decode, one per object group."""
        lib = _capi.load()
        out = []
        for i in range(len(self.aot_engines)):
            p = C.c_void_p()
            h4, w4 = C.c_int(), C.c_int()
            _capi.check(lib.rmem_engine_pred_logits(self._h, i, C.byref(p), C.byref(h4), C.byref(w4)))
            n = 11 * h4.value * w4.value
            off = p.value - self._arena.data_ptr()
            out.append(self._arena[off:off + 4 * n].view(torch.float32).view(11, h4.value, w4.value))
        return out

    def prefetch(self, img: torch.Tensor, stream: Optional[torch.cuda.Stream] = None):
        """Encode the NEXT frame on the engine's side stream while the current one is propagated (the image encoder,
        aot.py:116-134, does not depend on the memory bank).  `img` must be a contiguous fp32 CUDA tensor that is ready
        on `stream` (default: the current stream) and is passed unchanged to the next match_propogate_one_frame /
        propagate_label call.  Optional: results are bit-identical without it."""
        if self._h is None:
            return
        assert img.is_cuda and img.dtype == torch.float32 and img.is_contiguous(), "prefetch needs the engine-ready tensor"
        sp = C.c_void_p(stream.cuda_stream) if stream is not None else _capi.stream_ptr()
        _capi.check(_capi.load().rmem_engine_prefetch(self._h, _capi.ptr(img), sp))

    def prefetch2(self, img_a: torch.Tensor, img_b: torch.Tensor, stream: Optional[torch.cuda.Stream] = None):
        """Pair prefetch: the two frames AFTER the next one (i+2, i+3 when issued before frame i is propagated) go through
        the image encoder in one pass -- every encoder GEMM / conv launch covers both images.  Issue it every second
        frame; each tensor is passed unchanged to its own propagate call later.  Same math, differently tiled GEMMs:
        agrees with the single-frame encoder to fp16 rounding, not bit for bit."""
        if self._h is None:
            return
        for img in (img_a, img_b):
            assert img.is_cuda and img.dtype == torch.float32 and img.is_contiguous(), "prefetch needs the engine-ready tensor"
        sp = C.c_void_p(stream.cuda_stream) if stream is not None else _capi.stream_ptr()
        _capi.check(_capi.load().rmem_engine_prefetch2(self._h, _capi.ptr(img_a), _capi.ptr(img_b), sp))

    def prefetch_n(self, imgs, stream: Optional[torch.cuda.Stream] = None):
        """1, 2 or 4 coming frames through the image encoder in one pass (see prefetch2).  A group of n is issued n frames
        ahead: before frame i is propagated for frames i+n .. i+2n-1, every n-th frame."""
        if self._h is None:
            return
        n = len(imgs)
        for img in imgs:
            assert img.is_cuda and img.dtype == torch.float32 and img.is_contiguous(), "prefetch needs the engine-ready tensor"
        arr = (C.c_void_p * n)(*[img.data_ptr() for img in imgs])
        sp = C.c_void_p(stream.cuda_stream) if stream is not None else _capi.stream_ptr()
        _capi.check(_capi.load().rmem_engine_prefetch_n(self._h, arr, n, sp))

    def update_memory(self, label: torch.Tensor):
        """aot_engine.py:714-720 -> AOTEngine.update_short_term_memory (:327-396)."""
        lib = _capi.load()
        lab, is_f32 = self._label_arg(label)
        H, W = self.input_size_2d
        assert lab.numel() == H * W, "update_memory expects the label at the input size (evaluator.py:518-523)"
        _capi.check(lib.rmem_engine_update_memory(self._h, _capi.ptr(lab), is_f32, _capi.stream_ptr()))

    def set_timing(self, on: bool):
        """Profiling aid: per-stage CUDA-event timing inside the engine (adds a sync per call while on)."""
        _capi.check(_capi.load().rmem_engine_set_timing(self._h, int(on)))

    def get_timing(self):
        """{stage: (mean_ms, count)} accumulated since set_timing(True)."""
        buf = C.create_string_buffer(16384)
        _capi.check(_capi.load().rmem_engine_get_timing(self._h, buf, C.c_size_t(16384)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, tot, n = line.split()
            out[name] = (float(tot) / max(int(n), 1), int(n))
        return out

    @property
    def launch_count(self) -> int:
        return int(_capi.load().rmem_engine_launch_count(self._h)) if self._h else 0


class AOTInferEngine(DeAOTInferEngine):
    """aot_engine.py:571-725 over R50_AOTL (same per-clip state machine; the reference's DeAOTInferEngine derives from it)."""


def build_engine(name: str, phase: str = "eval", aot_model: DeAOTModel = None, gpu_id: int = 0,
                 long_term_mem_gap: int = 9999, **kwargs):
    """networks/engines/__init__.py:5-21."""
    if phase != "eval":
        raise NotImplementedError("rmem_b200 builds the inference path only (training is out of scope)")
    if name in ("deaotengine", "deaot_engine"):
        assert aot_model.cfg.model == "r50_deaotl", "deaotengine needs a DeAOT model"
        return DeAOTInferEngine(aot_model, gpu_id=gpu_id, long_term_mem_gap=long_term_mem_gap, **kwargs)
    if name in ("aotengine", "aot_engine"):
        assert aot_model.cfg.model == "r50_aotl", "aotengine needs an AOT model"
        return AOTInferEngine(aot_model, gpu_id=gpu_id, long_term_mem_gap=long_term_mem_gap, **kwargs)
    raise NotImplementedError(f"engine {name!r}")
