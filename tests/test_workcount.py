"""CPU: profiles/workcount.json (the algorithmic work per unit used by bench.py's roofline) matches a fresh count of
field multiplications executed by the kernels' own per-proof routines on the host build."""
import json
import os
import sys

from conftest import ROOT


def test_committed_workcount_is_current():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import count_work
    fresh = count_work.main(write=False)
    committed = json.load(open(os.path.join(ROOT, "profiles", "workcount.json")))
    assert fresh == committed
    assert committed["plonk_early_reject_mul"] < committed["plonk_full_path_mul"] // 50
