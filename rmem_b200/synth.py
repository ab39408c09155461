"""Synthetic weights and clips for the RMem path (no dataset or checkpoint is reachable offline).

Deterministic generators shared by bench.py, the tests and the CPU oracle so that every implementation sees
identical inputs: weights under the reference's own state_dict names/shapes and frame /
label tensors of the shapes of SURVEY.md section 8(d).  No arithmetic of the propagation path lives here.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

Tensor = torch.Tensor
MAX_OBJ = 10          # configs/models/default.py:17  MODEL_MAX_OBJ_NUM


def _resnet_blocks():
    """(layer_name, block_idx, inplanes, planes, stride, has_downsample) for ResNet-50
    truncated after layer3 (encoders/resnet.py:83-132, 134-176; layers=[3,4,6])."""
    out = []
    inplanes = 64
    for li, (planes, nblk, stride) in enumerate([(64, 3, 1), (128, 4, 2), (256, 6, 2)], start=1):
        for bi in range(nblk):
            s = stride if bi == 0 else 1
            ds = bi == 0 and (s != 1 or inplanes != planes * 4)
            out.append((f"layer{li}", bi, inplanes, planes, s, ds))
            inplanes = planes * 4
    return out


def make_state_dict(model: str = "r50_deaotl", seed: int = 0, sharpen: float = 1.0,
                    dtype=torch.float32, gru_memory: bool = False) -> Dict[str, Tensor]:
    """Deterministic synthetic weights under the reference's state_dict names/shapes.

    Magnitudes follow the reference initialisers (resnet.py:197-205 He-normal convs,
    xavier-uniform for LSTT/decoder matrices transformer.py:1246-1249 / fpn.py:70-73,
    aot.py:170-177 for the ID bank, deaot.py:45-55 temporal PE) but frozen-BN statistics,
    norm affines and biases are perturbed so that every parameter matters in parity tests.
    `sharpen` scales the Q/K projections so the attention is peaked and the eviction
    argmin is not tie-sensitive (SURVEY.md section 7 'hard parts').
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=dtype) * std

    def rand(*shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=g, dtype=dtype) * (hi - lo) + lo

    def conv(name, cout, cin, k, bias=False, he=True, groups=1):
        fan_out = k * k * cout
        fan_in = k * k * cin
        std = math.sqrt(2.0 / fan_out) if he else math.sqrt(2.0 / (fan_in + fan_out))
        sd[name + ".weight"] = randn(cout, cin // groups, k, k, std=std)
        if bias:
            b = 1.0 / math.sqrt(fan_in)
            sd[name + ".bias"] = rand(cout, lo=-b, hi=b)

    def bn(name, c, lo=0.7, hi=1.3):
        sd[name + ".weight"] = rand(c, lo=lo, hi=hi)
        sd[name + ".bias"] = randn(c, std=0.1)
        sd[name + ".running_mean"] = randn(c, std=0.1)
        sd[name + ".running_var"] = rand(c, lo=0.6, hi=1.4)

    def linear(name, cout, cin, scale=1.0):
        a = math.sqrt(6.0 / (cin + cout)) * scale
        sd[name + ".weight"] = rand(cout, cin, lo=-a, hi=a)
        b = 1.0 / math.sqrt(cin)
        sd[name + ".bias"] = rand(cout, lo=-b, hi=b)

    def norm(name, c):
        sd[name + ".weight"] = rand(c, lo=0.8, hi=1.2)
        sd[name + ".bias"] = randn(c, std=0.05)

    # ---- encoder (215 keys) ----
    conv("encoder.conv1", 64, 3, 7)
    bn("encoder.bn1", 64)
    for lname, bi, inpl, pl, s, ds in _resnet_blocks():
        p = f"encoder.{lname}.{bi}"
        conv(p + ".conv1", pl, inpl, 1); bn(p + ".bn1", pl)
        conv(p + ".conv2", pl, pl, 3); bn(p + ".bn2", pl)
        conv(p + ".conv3", pl * 4, pl, 1); bn(p + ".bn3", pl * 4, 0.15, 0.35)   # keeps the residual stream O(1)
        if ds:
            conv(p + ".downsample.0", pl * 4, inpl, 1); bn(p + ".downsample.1", pl * 4, 0.5, 0.9)
    conv("encoder_projector", 256, 1024, 1, bias=True, he=False)

    d = 256
    if model == "r50_deaotl":
        sd["cur_pos_emb"] = randn(1, d // 2, std=0.05).clamp(-0.1, 0.1)
        sd["mem_pos_emb"] = randn(4, d // 2, std=0.05).clamp(-0.1, 0.1)
        for l in range(3):
            p = f"LSTT.layers.{l}"
            norm(p + ".norm1", d)
            linear(p + ".linear_QV", d // 2 + 2 * d, d)
            sd[p + ".linear_QV.weight"][: d // 2] *= sharpen
            sd[p + ".linear_QV.bias"][: d // 2] *= sharpen
            linear(p + ".linear_U", 2 * d, d)
            if l == 0:
                linear(p + ".linear_ID_V", 2 * d, d)
            else:
                norm(p + ".id_norm1", d)
                linear(p + ".linear_ID_V", 2 * d, 2 * d)
                linear(p + ".linear_ID_U", 2 * d, d)
            for att in ("long_term_attn", "short_term_attn", "self_attn"):
                if att == "short_term_attn":
                    conv(p + ".short_term_attn.relative_emb_k", 225, d // 2, 1, bias=True, he=False)
                if att == "self_attn":
                    linear(p + ".self_attn.linear_QK", d // 2, 2 * d, scale=sharpen)
                    for nm in ("linear_V1", "linear_V2", "linear_U1", "linear_U2"):
                        linear(p + ".self_attn." + nm, 2 * d, d)
                sd[p + f".{att}.dw_conv.conv.weight"] = randn(4 * d, 1, 5, 5, std=math.sqrt(2.0 / 50.0))
                linear(p + f".{att}.projection", 2 * d, 4 * d)
            norm(p + ".norm2", d)
            norm(p + ".id_norm2", d)
        # keep state_dict order irrelevant; reference loads by name
        norm("LSTT.decoder_norms.0.gn", 2 * d)
        dec_in = 2 * d
    elif model == "r50_aotl":
        sd["cur_pos_emb"] = randn(1, d, std=0.05).clamp(-0.1, 0.1)
        sd["mem_pos_emb"] = randn(4, d, std=0.05).clamp(-0.1, 0.1)
        for l in range(3):
            p = f"LSTT.layers.{l}"
            norm(p + ".norm1", d)
            for nm in ("linear_Q", "linear_K", "linear_V"):
                linear(p + ".self_attn." + nm, d, d)
            linear(p + ".self_attn.projection", d, d)
            norm(p + ".norm2", d)
            linear(p + ".linear_Q", d, d, scale=sharpen)
            for nm in ("linear_V", "linear_QMem", "linear_VMem", "linear_KMem"):
                linear(p + "." + nm, d, d)
            norm(p + ".norm4", d)
            linear(p + ".long_term_attn.projection", d, d)
            linear(p + ".short_term_attn.projection", d, d)
            norm(p + ".norm3", d)
            linear(p + ".linear1", 4 * d, d)
            norm(p + ".activation.gn", 4 * d)
            sd[p + ".activation.conv.weight"] = randn(4 * d, 1, 5, 5, std=math.sqrt(2.0 / 50.0))
            linear(p + ".linear2", d, 4 * d)
        for i in range(3):
            norm(f"LSTT.decoder_norms.{i}", d)
        dec_in = 4 * d
    else:
        raise ValueError(model)

    # ---- decoder (fpn.py:24-34) ----
    for nm, co, ci, k in (("conv_in", d, dec_in, 1), ("conv_16x", d, d, 3),
                          ("conv_8x", d // 2, d, 3), ("conv_4x", d // 2, d // 2, 3)):
        conv(f"decoder.{nm}.conv", co, ci, k, bias=True, he=False)
        norm(f"decoder.{nm}.gn", co)
    conv("decoder.adapter_16x", d, 1024, 1, bias=True, he=False)
    conv("decoder.adapter_8x", d, 512, 1, bias=True, he=False)
    conv("decoder.adapter_4x", d // 2, 256, 1, bias=True, he=False)
    conv("decoder.conv_out", MAX_OBJ + 1, d // 2, 1, bias=True, he=False)

    # ---- ID bank (aot.py:63-74, 170-177): rows of norm 17^-2 ----
    fan = 12 * 17 * 17
    sd["patch_wise_id_bank.weight"] = randn(d, 12, 17, 17, std=10.0 * (17.0 ** -2) / math.sqrt(fan))
    sd["patch_wise_id_bank.bias"] = rand(d, lo=-1.0 / math.sqrt(fan), hi=1.0 / math.sqrt(fan)) * 0.01
    if model == "r50_deaotl":
        norm("id_norm", d)
    if gru_memory:
        # GRU_MEMORY ablation (transformer.py:35-119, 529-545; AOT only): per layer a ConvGRU for K (2x2) and V (1x1).
        # Generated last, so the rest of the dict is identical to the default one of the same seed.
        if model != "r50_aotl":
            raise ValueError("GRU_MEMORY exists for r50_aotl only (DualBranchGPM hard-codes gru_memory = False)")
        for l in range(3):
            for i, k in ((0, 2), (1, 1)):
                q = f"LSTT.layers.{l}.memory_grus.{i}"
                conv(q + ".conv_gru_cell.conv_gates", 2 * d, 2 * d, k, bias=True, he=False)
                conv(q + ".conv_gru_cell.conv_can", d, 2 * d, k, bias=True, he=False)
                conv(q + ".output_conv", d, d, 1, bias=True, he=False)
    return sd


def snap_size(s: int) -> int:
    """video_transforms.py:607-615: each side -> round((s-1)/16)*16+1."""
    return int(round((s - 1) / 16.0)) * 16 + 1 if (s - 1) % 16 else s


def synthetic_label(H: int, W: int, n_obj: int) -> Tensor:
    """n_obj disjoint rectangles with ids 1..n on a zero background, [1,1,H,W] float."""
    lab = torch.zeros(1, 1, H, W)
    cols = int(math.ceil(math.sqrt(n_obj * W / H)))
    rows = int(math.ceil(n_obj / cols))
    ch, cw = H // rows, W // cols
    for i in range(n_obj):
        r, c = divmod(i, cols)
        y0, x0 = r * ch + ch // 6, c * cw + cw // 6
        lab[0, 0, y0:y0 + max(2 * ch // 3, 1), x0:x0 + max(2 * cw // 3, 1)] = i + 1
    return lab


def synthetic_frames(n: int, H: int, W: int, seed: int = 1) -> Tensor:
    """n frames [n,3,H,W] of N(0,1) noise (ImageNet-normalised range) with temporal correlation so
    consecutive frames resemble each other like a video."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(1, 3, H, W, generator=g)
    out = [base]
    for _ in range(n - 1):
        out.append(0.9 * out[-1] + math.sqrt(1 - 0.81) * torch.randn(1, 3, H, W, generator=g))
    return torch.cat(out, 0)
