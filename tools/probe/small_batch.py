"""Device time of small batches (three-lanes-per-proof kernels): python tools/probe/small_batch.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
import workloads
pkg = ge.load_package(); pkg.init([0])
for n in (1024, 4096, 16384):
    vk, p, i, e = pkg.groth16_synth(11, n)
    b = pkg.Groth16DeviceBatch(vk, p, i)
    best = 1e9
    for _ in range(4):
        st, ms = b.verify(); best = min(best, ms)
    assert (st == e).all()
    print("groth16", n, "ms", round(best, 3), pkg.last_kernel_split())
    b.free()
n = 1 << 14
p, i, r, e = workloads.plonk_workload(n, seed=3)
b = pkg.PlonkDeviceBatch(workloads.plonk_vk_bytes(), p, i, r)
for _ in range(3):
    st, ms = b.verify()
assert (st == e).all()
print("plonk", n, "ms", ms, pkg.last_stage_ms())
