"""Parity of the fused tcgen05 multi-head attention of the AOT model (rmem_b200/csrc/mha_tc.cu) through the C ABI
(rmem_mha_fwd) against a torch fp32 statement of attention.py:28-81 + transformer.py:636-643.  Tolerance: rel-Frobenius
<= 8e-3 on the output (16-bit operands and P, fp32 accumulate), per-frame mass max-abs <= 2e-3."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_mha_tc_matches_reference(cuda_device):
    r = subprocess.run([sys.executable, os.path.join(HERE, "mha_tc_check.py")], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    print(r.stderr[-2000:])
    recs = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and len(recs) == 6, "tcgen05 multi-head attention check crashed"
    for rec in recs:
        assert rec["finite"], rec
        assert rec["tc_vs_ref"] < 8e-3, rec
        assert rec["mass_err"] < 2e-3 and rec["mass_sum_err"] < 2e-3, rec
        assert rec["deterministic"], rec
        if "tc_vs_dense" in rec:
            assert rec["tc_vs_dense"] < 8e-3, rec
