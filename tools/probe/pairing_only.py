"""Times the raw k=4 pairing-product path (2^17 sets) three times; BN254V_LIB selects an experiment build."""
import sys, time
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge
pkg = ge.load_package(); pkg.init([0])
m = 1 << 17
g1, g2, exp1 = pkg.pairing_synth(11, m, k=4)
for it in range(3):
    t0 = time.perf_counter(); one = pkg.pairing_product_batch(g1, g2, 4); dt = time.perf_counter() - t0
    assert (one == exp1).all()
    print("pairing product k=4, 2^17 sets: %.2f ms" % (dt * 1e3), flush=True)
