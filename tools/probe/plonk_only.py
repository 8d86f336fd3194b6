import sys, time
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge, workloads
pkg = ge.load_package(); pkg.init([0])
n = 1 << 14
proofs, inputs, rnd, expected = workloads.plonk_workload(n, seed=3)
vk = workloads.plonk_vk_bytes()
for it in range(3):
    t0 = time.perf_counter(); st = pkg.PlonkVerifier.verify_batch(proofs, vk, inputs, rnd=rnd); dt = time.perf_counter() - t0
    assert (st == expected).all()
print("plonk 2^14 ms", dt * 1e3)
