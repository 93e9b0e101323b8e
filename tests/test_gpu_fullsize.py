"""Parity at BASELINE.json's full size (configs[2]: n = 10 000, K = 50, dim = 100 -- the bench workload): the chain
kernel against the oracle bit for bit on a start that forces hundreds of moves per sweep (permutation patches,
rebuilds, corrections, several row tiles, two chains per CTA plus a half-filled CTA), and size-independent
properties of the PSM (symmetry, unit diagonal, row-sum checksum) on both count kernels."""
import ctypes as C
import os
import sys
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _workload():
    import bench
    return bench.synth(10000, 50, 100, 0.1, 50, 44)


def test_bench_workload_n10000_matches_oracle(pkg, orc):
    X, lab = _workload()
    data = pkg.MCMCData.from_points(X)
    D = data.D
    params = pkg.params_from_labels(D, lab)
    P = orc.make_params(**{k: getattr(params, k) for k in params._fields})
    g = np.random.default_rng(9)
    init = lab.copy()
    idx = g.choice(lab.size, size=300, replace=False)              # 3 % of the points start in a wrong cluster
    init[idx] = g.integers(1, 51, size=idx.size)
    nch, iters = 3, 3
    rp = [pkg.init_rp(params, 5, c) for c in range(nch)]
    opts = pkg.MCMCOptionsList(numiters=iters, burnin=0, thin=1, numGibbs=5, numMH=1)
    smp = pkg.Sampler(data, opts, params, np.tile(init, (nch, 1)), [a for a, _ in rp], [b for _, b in rp], seed=5)
    smp.run(1); smp.run(-1)                                         # two launches == one run
    from redclust_jl_b200._lib import lib, check, ptr
    q = params._c()
    moved = 0
    for c in range(nch):
        got = smp.samples(c)
        ref = orc.run_chain(D, orc.Options(iters, 0, 1, 5, 1), P, init, rp[c][0], rp[c][1], seed=5, chain=c)
        for k in ("labels", "K", "r", "p", "loglik", "logposterior", "r_acc", "sm_acc", "sm_split"):
            assert np.array_equal(got[k], ref[k]), (c, k)
        st = smp.state(c)
        assert np.array_equal(st.clusts, ref["final_labels"]) and (st.r, st.p) == tuple(ref["final_rp"])
        moved += int((ref["final_labels"] != init).sum())
        # the recorded log-likelihood (incrementally maintained block sums) equals a from-scratch evaluation
        out = C.c_double()
        last = np.ascontiguousarray(got["labels"][-1], dtype=np.int64)
        check(lib().rc_loglik(data._h, C.byref(q), ptr(last), C.byref(out)))
        assert out.value == got["loglik"][-1]
    assert moved >= 300                                             # the start really was off-equilibrium


def test_psm_fullsize_properties(pkg, monkeypatch):
    _, lab = _workload()
    n, S, K = lab.size, 300, 50
    g = np.random.default_rng(3)
    L = np.tile(lab, (S, 1))
    flip = g.random(L.shape) < 0.05
    L[flip] = g.integers(1, K + 1, size=int(flip.sum()))
    want_rowsum = np.zeros(n, np.int64)                             # sum_j counts[i, j] = sum_s |cluster of i in sample s|
    for s in range(S):
        want_rowsum += np.bincount(L[s], minlength=K + 1)[L[s]]
    res = {}
    for mode in ("tc", "compare"):
        monkeypatch.setenv("RCB200_PSM", mode)
        P = pkg.psm(L)
        cnt = np.rint(P * S).astype(np.int64)
        assert np.array_equal(cnt / S, P)                            # entries are exact multiples of 1 / S
        assert np.array_equal(P, P.T) and np.all(np.diag(P) == 1.0)
        assert np.array_equal(cnt.sum(axis=1), want_rowsum)
        res[mode] = P
    assert np.array_equal(res["tc"], res["compare"])
