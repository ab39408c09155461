"""CPU: the oracle (oracle/rmem_oracle.py) against the golden fixtures recorded from the UNMODIFIED reference by
oracle/make_golden.py (lock-step on the reference's label history).  This is what pins the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import rmem_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["deaot_small_xavier", "deaot_13obj_2engines", "deaot_small_10obj",
                                  "aot_c1_256_t1", "aot_small_rmem", "deaot_no_long_memory", "aot_gru_memory"])
def test_oracle_reproduces_reference_goldens(name):
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 1))
    z = np.load(os.path.join(GOLD, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    gru = bool(meta.get("knobs", {}).get("GRU_MEMORY", False))
    sd = O.make_state_dict(meta["model"], seed=meta["seed"], sharpen=meta["sharpen"], gru_memory=gru)
    frames = O.synthetic_frames(meta["n_frames"], meta["H"], meta["W"], seed=meta["seed"] + 1)
    label0 = O.synthetic_label(meta["H"], meta["W"], meta["n_obj"])
    cfg = O.OracleConfig(model=meta["model"], former_mem_len=meta["former"], latter_mem_len=meta["latter"],
                         no_long_memory=bool(meta.get("knobs", {}).get("NO_LONG_MEMORY", False)), gru_memory=gru)
    eng = O.OracleEngine(sd, cfg, long_term_mem_gap=meta["gap"])
    forced = torch.from_numpy(z["labels"])
    rec = dict(idx=[], logits4=[])

    def on_frame(f, logit, lab):
        rec["idx"].append([list(e.long_memories_indexes) for e in eng.aot_engines])
        rec["logits4"].append(eng.aot_engines[0].pred_id_logits.clone())

    with torch.no_grad():
        labels = O.run_clip(eng, frames, label0, meta["n_obj"], out_size=tuple(meta["out_size"]), on_frame=on_frame,
                            forced_labels=forced)
    assert rec["idx"] == meta["idx"]
    sub = torch.from_numpy(z["logits4_sub"])
    for f, lg in enumerate(rec["logits4"]):
        assert float((lg[0, :, ::4, ::4] - sub[f]).abs().max()) < 2e-4, f
    ours = torch.stack([l[0, 0] for l in labels])
    mism = int((ours != forced).sum())
    assert mism <= 1e-5 * forced.numel() + 2, mism
