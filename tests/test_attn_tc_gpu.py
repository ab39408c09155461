"""Parity of the fused tcgen05/TMA long-term attention kernels (rmem_b200/csrc/attn_tc3.cu: CTA pairs, seeded and
unseeded row maximum; attn_tc2.cu: single CTAs) against the CPU oracle
and the dense CUDA path, through the C ABI.  Tolerance: rel-Frobenius <= 8e-3 on the attention output (16-bit
operands and P, fp32 accumulate), per-frame mass max-abs <= 2e-3."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("impl,seed", [(4, 1), (3, 1), (3, 0), (2, 0)], ids=["tc4_column", "tc3_seeded", "tc3_unseeded", "tc2"])
def test_tc_attention_matches_oracle(cuda_device, impl, seed):
    r = subprocess.run([sys.executable, os.path.join(HERE, "tc_attn_check.py")], capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, RMEM_ATTN_IMPL=str(impl), RMEM_ATTN_SEED=str(seed)))
    print(r.stdout)
    print(r.stderr[-2000:])
    recs = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and len(recs) == 7, "tcgen05 attention check crashed"
    for rec in recs:
        assert rec["ok"], rec
        assert rec["finite"], rec
        assert rec["tc_vs_oracle"] < 8e-3, rec
        assert rec["mass_err"] < 2e-3 and rec["mass_sum_err"] < 2e-3, rec
