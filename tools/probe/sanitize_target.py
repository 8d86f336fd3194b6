"""compute-sanitizer target: every kernel family once, small batches, malformed proofs mixed in.
python tools/probe/sanitize_target.py [big]   (big: also one 50 000-proof Groth16 batch = the two-launch 448-thread kernels)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import __graft_entry__ as ge
import workloads
pkg = ge.load_package(); pkg.init([0])
big = len(sys.argv) > 1 and sys.argv[1] == "big"
n = 4096
vk, proofs, inputs, expected = pkg.groth16_synth(5, n)
proofs = proofs.copy(); expected = expected.copy()
proofs[7, 0] = 0xff; expected[7] = pkg.PANIC_FIELD_NOT_MEMBER      # A.x >= p
proofs[40, 63] ^= 1; expected[40] = pkg.PANIC_NOT_ON_CURVE         # A off the curve, inside a warp of good proofs
st, dbg = pkg.Groth16Verifier.verify_batch(proofs, vk, inputs, debug=True)
assert (st == expected).all()
if big:
    vk, proofs, inputs, expected = pkg.groth16_synth(6, 50000)
    assert (pkg.Groth16Verifier.verify_batch(proofs, vk, inputs) == expected).all()
p, i, r, e = workloads.plonk_workload(1024, seed=3)
st, dbg = pkg.PlonkVerifier.verify_batch(p, workloads.plonk_vk_bytes(), i, rnd=r, debug=True)
assert (st == e).all()
for k in (1, 4):
    g1, g2, exp1 = pkg.pairing_synth(11, 512, k=k)
    g1 = g1.copy(); g1[3, 0] = 0
    one, ml, gt = pkg.pairing_product_batch(g1, g2, k, want_values=True)
    assert (np.delete(one, 3) == np.delete(exp1, 3)).all()
print("SANITIZE-TARGET-OK trio_max=%s big=%s launches=%d" % (os.environ.get("BN254V_TRIO_MAX", "default"), big, pkg.launch_count()))
