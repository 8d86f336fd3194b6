"""One PlonK batch through the C ABI (profiling target): python tools/probe/plonk_only.py [log2 batch] [repeats]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
import workloads
pkg = ge.load_package(); pkg.init([0])
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 14)
p, i, r, e = workloads.plonk_workload(n, seed=3)
b = pkg.PlonkDeviceBatch(workloads.plonk_vk_bytes(), p, i, r)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    st, ms = b.verify()
    assert (st == e).all()
print("ok", ms, pkg.last_stage_ms())
