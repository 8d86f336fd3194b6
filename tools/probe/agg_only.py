"""Aggregate Groth16 check against the per-proof path: python tools/probe/agg_only.py [log2 valid proofs] [repeats]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); pkg.init([0])
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 16)
vk, proofs, inputs, expected = pkg.groth16_synth(41, 2 * n)
keep = expected == pkg.OK_TRUE
vp, vi = np.ascontiguousarray(proofs[keep]), np.ascontiguousarray(inputs[keep])
assert vp.shape[0] == n
ver = pkg.Groth16Verifier
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    t0 = time.perf_counter(); ok = ver.batch_all_valid(vp, vk, vi); t1 = time.perf_counter()
    ms = pkg.last_stage_ms()
    st = ver.verify_batch(vp, vk, vi); t2 = time.perf_counter()
    assert ok and (st == 0).all()
    print("n", n, "all_valid call %.2f ms (per-proof half %.2f ms, batch half %.2f ms)" % ((t1 - t0) * 1e3, ms[0], ms[1]), "verify_batch call %.2f ms" % ((t2 - t1) * 1e3))
