"""One raw 4-pair product batch through the C ABI (profiling target): python tools/probe/pairing_only.py [log2 batch] [repeats]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); pkg.init([0])
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 17)
g1, g2, e = pkg.pairing_synth(11, n, k=4)
b = pkg.PairingDeviceBatch(g1, g2, 4)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    one, ms = b.verify()
    assert (one == e).all()
print("ok", ms)
