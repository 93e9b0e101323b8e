"""Check the tensor-core PSM kernel against numpy and the byte-compare kernel; time both (RCB200_VERBOSE lines)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
os.environ["RCB200_VERBOSE"] = "1"
rng = np.random.default_rng(5)
ok = True
for (n, S, K) in [(300, 37, 20), (129, 5, 3), (1000, 130, 64), (777, 64, 128), (513, 257, 33), (2000, 300, 50)]:
    L = rng.integers(1, K + 1, size=(S, n)).astype(np.int64)
    L[:, 0] = 1; L[0, :K] = np.arange(1, K + 1)           # every label 1..K appears; first-appearance order is not identity in general
    want = np.zeros((n, n), dtype=np.int64)
    for s in range(S):
        want += (L[s][:, None] == L[s][None, :])
    res = {}
    for mode in ("tc", "compare"):
        os.environ["RCB200_PSM"] = mode
        res[mode] = pkg.psm(L)
    for mode in res:
        good = np.array_equal(res[mode], want / S)
        ok &= good
        print(n, S, K, mode, "OK" if good else "MISMATCH max|d|=%g" % np.abs(res[mode] - want / S).max(), flush=True)
if ok and len(sys.argv) > 1:
    n, S, K = 10000, 2000, 50
    lab = rng.integers(1, K + 1, size=n)
    L = np.tile(lab, (S, 1)); flip = rng.random(L.shape) < 0.03; L[flip] = rng.integers(1, K + 1, size=int(flip.sum()))
    out = {}
    for mode in ("tc", "compare"):
        os.environ["RCB200_PSM"] = mode
        t = time.perf_counter(); out[mode] = pkg.psm(L); print(mode, "wall", time.perf_counter() - t, flush=True)
    print("big equal:", np.array_equal(out["tc"], out["compare"]))
print("ALL OK" if ok else "FAILED")
