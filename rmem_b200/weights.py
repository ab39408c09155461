"""Pack a reference `state_dict` (DeAOT, key names of networks/models/deaot.py -- SURVEY.md section 8b) into
the flat device blob the C++ engine reads: FrozenBatchNorm2d folded into the conv weights
(networks/layers/normalization.py:19-43), conv weights re-ordered to [Cout, ky, kx, Cin] (the K order of the
implicit GEMM), GEMM operands in the 16-bit operand type (fp16 default), norm / bias / depthwise / positional parameters in fp32, the ID-bank
conv (networks/models/aot.py:63-74) re-ordered to [17*17, 12, C] for the label-indexed gather.
`module.` prefixes are stripped like utils/checkpoint.py:75-101 does.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import torch

from . import _capi

BN_EPS = 1e-5


def _resnet_blocks():
    out = []
    inplanes = 64
    for li, (planes, nblk, stride) in enumerate([(64, 3, 1), (128, 4, 2), (256, 6, 2)], start=1):
        for bi in range(nblk):
            s = stride if bi == 0 else 1
            out.append((li, bi, inplanes, planes, s, bi == 0))
            inplanes = planes * 4
    return out


def pack_model(sd: Dict[str, torch.Tensor], model: str = "r50_deaotl") -> Dict[str, torch.Tensor]:
    """name -> packed CPU tensor (t16 or fp32, contiguous).  model: "r50_deaotl" | "r50_aotl"."""
    deaot = model == "r50_deaotl"
    sd = {(k[7:] if k.startswith("module.") else k): v.detach().float().cpu() for k, v in sd.items()}
    out: Dict[str, torch.Tensor] = {}

    def conv_bn(dst, conv, bn, cin_pad=None):
        w = sd[conv + ".weight"]
        scale = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + BN_EPS)
        b = sd[bn + ".bias"] - sd[bn + ".running_mean"] * scale
        w = (w * scale.view(-1, 1, 1, 1)).permute(0, 2, 3, 1)             # [Cout, ky, kx, Cin]
        if cin_pad is not None and cin_pad > w.shape[-1]:
            w = torch.nn.functional.pad(w, (0, cin_pad - w.shape[-1]))
        out[dst + ".w"] = w.contiguous().to(_capi.op_dtype())
        out[dst + ".b"] = b.contiguous()

    def conv(dst, src):
        w = sd[src + ".weight"].permute(0, 2, 3, 1)
        out[dst + ".w"] = w.contiguous().to(_capi.op_dtype())
        out[dst + ".b"] = sd[src + ".bias"].contiguous()

    def linear(dst, src, pad_rows=None):
        w, b = sd[src + ".weight"], sd[src + ".bias"]
        w = w.reshape(w.shape[0], -1)
        if pad_rows is not None and pad_rows > w.shape[0]:
            w = torch.nn.functional.pad(w, (0, 0, 0, pad_rows - w.shape[0]))
            b = torch.nn.functional.pad(b, (0, pad_rows - b.shape[0]))
        out[dst + ".w"] = w.contiguous().to(_capi.op_dtype())
        out[dst + ".b"] = b.contiguous()

    def norm(dst, src):
        out[dst + ".g"] = sd[src + ".weight"].contiguous()
        out[dst + ".b"] = sd[src + ".bias"].contiguous()

    def dw(dst, src):
        w = sd[src + ".dw_conv.conv.weight"]                                # [C,1,5,5]
        out[dst] = w.view(w.shape[0], 25).t().contiguous()                  # [25, C]

    conv_bn("enc.conv1", "encoder.conv1", "encoder.bn1", cin_pad=8)
    # stem layout (gemm.cuh conv = 2): [Cout][7 window rows][8 pixels][8 channels], zero weights for the 8th pixel
    out["enc.conv1.w"] = torch.nn.functional.pad(out["enc.conv1.w"], (0, 0, 0, 1)).contiguous()
    for li, bi, inpl, pl, s, ds in _resnet_blocks():
        p, q = f"encoder.layer{li}.{bi}", f"enc.layer{li}.{bi}"
        for i in (1, 2, 3):
            conv_bn(f"{q}.conv{i}", f"{p}.conv{i}", f"{p}.bn{i}")
        if ds:
            conv_bn(f"{q}.ds", f"{p}.downsample.0", f"{p}.downsample.1")
    linear("proj", "encoder_projector")

    wb = sd["patch_wise_id_bank.weight"]                                    # [C,12,17,17]
    out["idbank.w"] = wb.permute(2, 3, 1, 0).reshape(17 * 17, wb.shape[1], wb.shape[0]).contiguous()
    out["idbank.b"] = sd["patch_wise_id_bank.bias"].contiguous()
    # per-class inclusive 2-D prefix sums over (ky, kx): P[k][ky][kx][c] = sum_{y<ky, x<kx} w[c, k, y, x]  (fp64 -> fp32)
    pre = torch.zeros(wb.shape[1], 18, 18, wb.shape[0], dtype=torch.float64)
    pre[:, 1:, 1:, :] = wb.double().permute(1, 2, 3, 0).cumsum(1).cumsum(2)
    out["idbank.prefix"] = pre.float().contiguous()
    # per-(ky, class) 1-D prefix sums over kx: R[ky][k][kx][c] = sum_{x<kx} w[c, k, ky, x]
    rows = torch.zeros(17, wb.shape[1], 18, wb.shape[0], dtype=torch.float64)
    rows[:, :, 1:, :] = wb.double().permute(2, 1, 3, 0).cumsum(2)
    out["idbank.prefix_rows"] = rows.float().contiguous()
    if deaot:
        norm("id_norm", "id_norm")
    out["cur_pos_emb"] = sd["cur_pos_emb"].reshape(-1).contiguous()
    out["mem_pos_emb"] = sd["mem_pos_emb"].contiguous()

    # ---- AOT LSTT (transformer.py:466-697; linear_KMem is never used by the reference) ----
    for l in range(0 if deaot else 3):
        p, q = f"LSTT.layers.{l}", f"lstt.{l}"
        for nm in ("norm1", "norm2", "norm3", "norm4"):
            norm(f"{q}.{nm}", f"{p}.{nm}")
        for nm in ("linear_Q", "linear_K", "linear_V"):
            linear(f"{q}.self.{nm}", f"{p}.self_attn.{nm}")
        linear(f"{q}.self.proj", f"{p}.self_attn.projection")
        for nm in ("linear_Q", "linear_V", "linear_QMem", "linear_VMem", "linear1", "linear2"):
            linear(f"{q}.{nm}", f"{p}.{nm}")
        linear(f"{q}.long.proj", f"{p}.long_term_attn.projection")
        linear(f"{q}.short.proj", f"{p}.short_term_attn.projection")
        norm(f"{q}.act.gn", f"{p}.activation.gn")
        w = sd[f"{p}.activation.conv.weight"]                               # [C,1,5,5]
        out[f"{q}.act.dw"] = w.view(w.shape[0], 25).t().contiguous()
        norm(f"lstt.dec_norm.{l}", f"LSTT.decoder_norms.{l}")
        # GRU_MEMORY ablation (transformer.py:529-545): ConvGRU of the K (2x2) and V (1x1) memories, when the checkpoint has it
        for i in (0, 1):
            g = f"{p}.memory_grus.{i}"
            if g + ".conv_gru_cell.conv_gates.weight" in sd:
                conv(f"{q}.gru.{i}.gates", g + ".conv_gru_cell.conv_gates")
                conv(f"{q}.gru.{i}.can", g + ".conv_gru_cell.conv_can")
                linear(f"{q}.gru.{i}.out", g + ".output_conv")

    for l in range(3 if deaot else 0):
        p, q = f"LSTT.layers.{l}", f"gpm.{l}"
        norm(q + ".norm1", p + ".norm1")
        linear(q + ".linear_QV", p + ".linear_QV")
        linear(q + ".linear_U", p + ".linear_U")
        linear(q + ".linear_ID_V", p + ".linear_ID_V")
        if l > 0:
            norm(q + ".id_norm1", p + ".id_norm1")
            linear(q + ".linear_ID_U", p + ".linear_ID_U")
        for dst, src in (("long", "long_term_attn"), ("short", "short_term_attn"), ("self", "self_attn")):
            dw(f"{q}.{dst}.dw", f"{p}.{src}")
            linear(f"{q}.{dst}.proj", f"{p}.{src}.projection")
        linear(q + ".short.rel", p + ".short_term_attn.relative_emb_k", pad_rows=256)
        # the same 225 offsets as 15 window rows of 16 (last one zero): one aligned 64-byte line per row for the
        # tensor-core kernel (local_attn_tc.cu, rel_pitch 16)
        for sfx in (".w", ".b"):
            src = out[q + ".short.rel" + sfx]
            dst = torch.zeros_like(src)
            for dy in range(15):
                dst[dy * 16: dy * 16 + 15] = src[dy * 15: dy * 15 + 15]
            out[q + ".short.rel16" + sfx] = dst
        norm(q + ".norm2", p + ".norm2")
        norm(q + ".id_norm2", p + ".id_norm2")
        linear(q + ".self.linear_QK", p + ".self_attn.linear_QK")
        for nm in ("linear_V1", "linear_V2", "linear_U1", "linear_U2"):
            linear(f"{q}.self.{nm}", f"{p}.self_attn.{nm}")
        # Fused launches of the engine (same arithmetic, fewer small GEMMs on the per-frame critical path):
        #  * self.QKU  [128 + 1024, 512]: rows 0..127 = linear_QK (all 512 inputs), then linear_U1 on the tgt half and
        #    linear_U2 on the tgt_id half of z = cat(LN2(tgt), id_LN2(tgt_id)) as a block-diagonal weight
        #  * self.V12  [1024, 512]: block-diagonal linear_V1 / linear_V2 (computed value-major: W . z^T)
        #  * tail.proj [512, 2048]: long_term_attn.projection | short_term_attn.projection along K, biases summed
        #    (tgt += o2 + o3 is one accumulate of the concatenated depthwise-conv outputs, transformer.py:1212-1220)
        od = _capi.op_dtype()
        z = torch.zeros(512, 256, dtype=od)
        u1, u2 = out[f"{q}.self.linear_U1.w"], out[f"{q}.self.linear_U2.w"]
        out[f"{q}.self.QKU.w"] = torch.cat([out[f"{q}.self.linear_QK.w"], torch.cat([u1, z], 1), torch.cat([z, u2], 1)], 0).contiguous()
        out[f"{q}.self.QKU.b"] = torch.cat([out[f"{q}.self.linear_QK.b"], out[f"{q}.self.linear_U1.b"], out[f"{q}.self.linear_U2.b"]]).contiguous()
        v1, v2 = out[f"{q}.self.linear_V1.w"], out[f"{q}.self.linear_V2.w"]
        out[f"{q}.self.V12.w"] = torch.cat([torch.cat([v1, z], 1), torch.cat([z, v2], 1)], 0).contiguous()
        out[f"{q}.self.V12.b"] = torch.cat([out[f"{q}.self.linear_V1.b"], out[f"{q}.self.linear_V2.b"]]).contiguous()
        out[f"{q}.tail.proj.w"] = torch.cat([out[f"{q}.long.proj.w"], out[f"{q}.short.proj.w"]], 1).contiguous()
        out[f"{q}.tail.proj.b"] = (out[f"{q}.long.proj.b"] + out[f"{q}.short.proj.b"]).contiguous()
    if deaot:
        norm("gpm.out_norm", "LSTT.decoder_norms.0.gn")

    for nm in ("conv_in", "conv_16x", "conv_8x", "conv_4x"):
        conv("dec." + nm, f"decoder.{nm}.conv")
        out[f"dec.{nm}.gn.g"] = sd[f"decoder.{nm}.gn.weight"].contiguous()
        out[f"dec.{nm}.gn.b"] = sd[f"decoder.{nm}.gn.bias"].contiguous()
    for nm in ("adapter_16x", "adapter_8x", "adapter_4x", "conv_out"):
        conv("dec." + nm, "decoder." + nm)
    return out


def load_checkpoint(ckpt, model: str = "r50_deaotl", template: Dict[str, torch.Tensor] = None,
                    widened_init: str = "template") -> Tuple[Dict[str, torch.Tensor], List[str]]:
    """`load_network` of the reference (utils/checkpoint.py:75-101) for this path: turn a released checkpoint into the
    state_dict `pack_model` / `RmemModel` take.

    ckpt      path (torch.load, CPU) or an already loaded dict.  `'state_dict'` / `'model'` wrappers are unwrapped (:78-83).
    template  the freshly built model's state_dict (names -> tensors of the model's shapes; the reference calls it
              `model_dict`, :84).  Default: rmem_b200.synth.make_state_dict(model, seed=0).
    Per checkpoint entry, in the reference's order (:87-98):
      * conv weight whose input-channel count is ONE LESS than the model's (`patch_wise_id_bank.weight` saved with 11
        channels, the model built with MODEL_IGNORE_TOKEN has 12): copied into the first channels, the extra channel keeps
        the template's value (`widened_init="template"`, what the reference does with its freshly initialised model) or is
        zeroed (`"zeros"`: the ignore channel then contributes nothing);
      * same name and shape: taken;  `module.`-prefixed name (DDP) whose stripped name and shape match: taken;
      * anything else: reported in the returned list (the reference's `pretrained_dict_remove`), not loaded.
    Returns (state_dict, dropped_keys)."""
    from .synth import make_state_dict
    if isinstance(ckpt, (str, bytes)) or hasattr(ckpt, "__fspath__"):
        ckpt = torch.load(ckpt, map_location="cpu")
    if "state_dict" in ckpt:
        pretrained = ckpt["state_dict"]
    elif "model" in ckpt:
        pretrained = ckpt["model"]
    else:
        pretrained = ckpt
    model_dict = {k: v.detach().clone().float() for k, v in (template or make_state_dict(model, seed=0)).items()}
    update, dropped = {}, []
    for k, v in pretrained.items():
        if not torch.is_tensor(v):
            dropped.append(k)
            continue
        v = v.detach().float().cpu()
        if (k in model_dict and v.dim() > 2 and v.shape[0] == model_dict[k].shape[0]
                and v.shape[1] == model_dict[k].shape[1] - 1):
            if widened_init == "zeros":
                model_dict[k][:, -1:] = 0
            model_dict[k][:, :-1] = v
            continue
        if k in model_dict and v.shape == model_dict[k].shape:
            update[k] = v
        elif k[:7] == "module.":
            if k[7:] in model_dict and v.shape == model_dict[k[7:]].shape:
                update[k[7:]] = v
        else:
            dropped.append(k)
    model_dict.update(update)
    return model_dict, dropped


def pack_deaot(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return pack_model(sd, "r50_deaotl")


def pack_aot(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return pack_model(sd, "r50_aotl")


class WeightBlob:
    """Packed weights resident in HBM + the (name, offset, nbytes) table handed to rmem_engine_create."""

    def __init__(self, packed: Dict[str, torch.Tensor], device):
        offs: List[Tuple[str, int, int]] = []
        total = 0
        for name, t in packed.items():
            nbytes = t.numel() * t.element_size()
            offs.append((name, total, nbytes))
            total += (nbytes + 255) // 256 * 256
        host = torch.zeros(total, dtype=torch.uint8)
        for (name, off, nbytes), t in zip(offs, packed.values()):
            host[off:off + nbytes] = t.contiguous().view(torch.uint8).view(-1)
        self.blob = host.to(device)
        self.offsets = {n: (o, b) for n, o, b in offs}
        self._names = [n.encode() for n, _, _ in offs]      # keep the byte strings alive
        self.entries = (_capi.WeightEntry * len(offs))()
        for i, (n, o, b) in enumerate(offs):
            self.entries[i].name = self._names[i]
            self.entries[i].offset = o
            self.entries[i].nbytes = b
        self.n_entries = len(offs)
        self.nbytes = total

    def view(self, name: str, dtype, shape):
        """Typed device view of one packed tensor (op-level tests)."""
        o, b = self.offsets[name]
        return self.blob[o:o + b].view(dtype).view(*shape)
