"""Checkpoint ingest: rmem_b200.weights.load_checkpoint against the reference's load_network
(aot_plus/utils/checkpoint.py:75-101) -- 'state_dict' / 'model' unwrapping, `module.` stripping, ID-bank widening from
11 to 12 input channels, dropped-key report.  The comparison with the UNMODIFIED reference function runs wherever
/root/reference exists (the build container); the semantic checks run everywhere.  CPU only."""
import os
import sys
import types

import pytest
import torch

from rmem_b200 import synth
from rmem_b200.weights import load_checkpoint

REF = os.environ.get("RMEM_REFERENCE", "/root/reference")


def synthetic_checkpoint(seed=5):
    """What a released DDP checkpoint looks like: 'state_dict' wrapper, 'module.' prefixes, an 11-channel ID bank, one
    stale key the model no longer has and one key of the wrong shape."""
    sd = synth.make_state_dict("r50_deaotl", seed=seed)
    ck = {}
    for i, (k, v) in enumerate(sd.items()):
        ck[("module." + k) if i % 2 else k] = v.clone()
    name = "module.patch_wise_id_bank.weight" if "module.patch_wise_id_bank.weight" in ck else "patch_wise_id_bank.weight"
    ck[name] = ck[name][:, :11].clone()
    ck["LSTT.layers.0.some_removed_buffer"] = torch.zeros(3)
    ck.pop("module.decoder.conv_out.bias", None)
    ck["decoder.conv_out.bias"] = torch.zeros(7)            # wrong shape: neither branch takes it
    return {"state_dict": ck, "epoch": 3}, sd


def test_load_checkpoint_semantics(tmp_path):
    ck, sd = synthetic_checkpoint()
    template = synth.make_state_dict("r50_deaotl", seed=0)
    for wrapper in ("state_dict", "model", None):
        c = ck["state_dict"] if wrapper is None else {wrapper: ck["state_dict"]}
        out, dropped = load_checkpoint(c, "r50_deaotl", template=template)
        assert set(out) == set(template)
        for k in template:
            if k == "patch_wise_id_bank.weight":
                assert torch.equal(out[k][:, :11], sd[k][:, :11]) and torch.equal(out[k][:, 11:], template[k][:, 11:])
            elif k == "decoder.conv_out.bias":
                assert torch.equal(out[k], template[k])     # wrong-shaped entry ignored, model value kept
            else:
                assert torch.equal(out[k], sd[k]), k
        assert dropped == ["LSTT.layers.0.some_removed_buffer", "decoder.conv_out.bias"]
    out0, _ = load_checkpoint(ck, "r50_deaotl", template=template, widened_init="zeros")
    assert float(out0["patch_wise_id_bank.weight"][:, 11:].abs().max()) == 0.0
    path = tmp_path / "ckpt.pth"
    torch.save(ck, path)
    out1, dropped1 = load_checkpoint(str(path), "r50_deaotl", template=template)
    assert all(torch.equal(out1[k], out[k]) for k in out) and dropped1 == dropped
    from rmem_b200.weights import pack_model
    packed = pack_model(out1, "r50_deaotl")                  # the widened table packs into the engine layout
    assert packed["idbank.w"].shape == (289, 12, 256)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "aot_plus")), reason="reference tree not present")
def test_load_checkpoint_matches_reference_load_network(tmp_path, monkeypatch):
    tl = types.ModuleType("timm.models.layers")
    tl.trunc_normal_ = torch.nn.init.trunc_normal_
    tl.DropPath = torch.nn.Identity
    tl.to_2tuple = lambda x: (x, x)
    for name, mod in (("timm", types.ModuleType("timm")), ("timm.models", types.ModuleType("timm.models")),
                      ("timm.models.layers", tl)):
        monkeypatch.setitem(sys.modules, name, mod)
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    monkeypatch.setitem(sys.modules, "matplotlib", mpl)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", mpl.pyplot)
    monkeypatch.syspath_prepend(os.path.join(REF, "aot_plus"))
    from configs.default import DefaultEngineConfig
    from networks.models import build_vos_model
    import utils.checkpoint as ref_ckpt

    cfg = DefaultEngineConfig("probe", "r50_deaotl")
    cfg.MODEL_LINEAR_Q = False
    cfg.MODEL_IGNORE_TOKEN = True
    torch.manual_seed(0)
    net = build_vos_model(cfg.MODEL_VOS, cfg).eval()
    template = {k: v.detach().clone() for k, v in net.state_dict().items()}
    ck, _ = synthetic_checkpoint()
    ck["state_dict"] = {k: v for k, v in ck["state_dict"].items()
                        if (k[7:] if k.startswith("module.") else k) in template or "removed" in k or "conv_out.bias" in k}
    path = tmp_path / "ckpt.pth"
    torch.save(ck, path)
    # CPU patches of the reference call: torch.load onto the CPU, Module.cuda a no-op
    real_load = torch.load
    monkeypatch.setattr(ref_ckpt.torch, "load", lambda f, map_location=None: real_load(f, map_location="cpu"))
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    net2, removed = ref_ckpt.load_network(net, str(path), 0)
    ours, dropped = load_checkpoint(str(path), "r50_deaotl", template=template)
    ref_sd = net2.state_dict()
    assert set(ours) == set(ref_sd)
    for k in ref_sd:
        assert torch.equal(ours[k], ref_sd[k].float()), k
    assert dropped == removed
