"""CPU tests of the oracle (test infrastructure) against the reference's only numeric fixture
(data/example_datasets.h5 -> tests/golden/example{1,2,3}.npz) and of its internal consistency."""
import math
import numpy as np
import pytest


def test_distm_matches_reference_fixture(orc, golden):
    """pairwise(Euclidean(), X, dims=2) as stored by RedClust.jl itself: <= 1e-10 relative (north-star tolerance)."""
    for k in (1, 2, 3):
        D = orc.distm(golden[k]["points"])
        ref = golden[k]["distance_matrix"]
        assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0)              # test/test_datagen.jl:13
        off = ~np.eye(100, dtype=bool)
        rel = np.abs(D[off] - ref[off]) / ref[off]
        assert rel.max() < 1e-10, rel.max()
        assert rel.max() < 5e-15                                                # observed: <= 10 ulp


def test_logdist_definition(orc, golden):
    D = golden[1]["distance_matrix"]
    L = orc.logdist(D)
    off = ~np.eye(100, dtype=bool)
    assert np.all(np.diag(L) == 0)                                              # log.(D - Diagonal(D) + I), types.jl:155
    assert np.max(np.abs(L[off] - np.log(D[off]))) < 1e-15 * 8


def test_math_accuracy(orc):
    L = orc.lib()
    rng = np.random.default_rng(1)
    from scipy.special import gammaln, erfc, ndtri
    for x in np.concatenate([rng.uniform(1e-3, 50, 2000), 10 ** rng.uniform(-8, 9, 2000)]):
        x = float(x)
        assert abs(L.rco_log(x) - math.log(x)) <= 4e-16 * max(1.0, abs(math.log(x)))
        assert abs(L.rco_lgamma(x) - gammaln(x)) <= 2e-14 * max(1.0, abs(gammaln(x)))
    for x in rng.uniform(-700, 700, 2000):
        x = float(x)
        assert abs(L.rco_exp(x) - math.exp(x)) <= 4e-16 * math.exp(x)
    for x in rng.uniform(-6, 6, 500):
        x = float(x)
        assert abs(L.rco_erfc(x) - erfc(x)) <= 2e-15        # absolute (rc_math.h): only used through normcdf
    for p in rng.uniform(1e-9, 1 - 1e-9, 500):
        assert abs(L.rco_norminv(float(p)) - ndtri(p)) <= 1e-9
    assert L.rco_log(0.0) == -math.inf and math.isnan(L.rco_log(-1.0)) and L.rco_exp(800.0) == math.inf


def test_philox_known_answer(orc):
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors)."""
    assert list(orc.philox([0, 0, 0, 0], [0, 0])) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert list(orc.philox([0xffffffff] * 4, [0xffffffff] * 2)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert list(orc.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def _params(orc, pkg, D, lab, **kw):
    p = pkg.params_from_labels(D, lab, **kw)
    return orc.make_params(**{k: getattr(p, k) for k in p._fields})


def test_exact_and_plain_summation_agree(orc, pkg, golden):
    """sum_mode 0 (exact fixed-point sums, what the GPU reproduces bit for bit) vs sum_mode 1 (ascending fp64
    summation, closest to the reference's matsum): log-likelihoods agree to 1e-10 relative."""
    for k in (1, 2, 3):
        D, lab = golden[k]["distance_matrix"], golden[k]["cluster_labels"]
        P = _params(orc, pkg, D, lab)
        rng = np.random.default_rng(k)
        for labels in (lab, rng.integers(1, 8, size=100), np.ones(100, np.int64), np.arange(1, 101)):
            a, b = orc.loglik(D, P, labels, 0), orc.loglik(D, P, labels, 1)
            assert abs(a - b) <= 1e-10 * abs(b)


def test_loglik_matches_direct_formula(orc, pkg, golden):
    """loglik restated independently in numpy from the model definition (docs/src/index.md:37-68)."""
    from scipy.special import gammaln
    D, lab = golden[2]["distance_matrix"], golden[2]["cluster_labels"]
    p = pkg.params_from_labels(D, lab)
    P = orc.make_params(**{k: getattr(p, k) for k in p._fields})
    logD = np.log(D + np.eye(100))
    L1 = L2 = 0.0
    ks = np.unique(lab)
    for a, k in enumerate(ks):
        ik = np.where(lab == k)[0]
        nk = ik.size; pairs = nk * (nk - 1) // 2
        A = p.alpha + p.delta1 * pairs; B = p.beta + D[np.ix_(ik, ik)].sum() / 2
        L1 += (p.delta1 - 1) * logD[np.ix_(ik, ik)].sum() / 2 - pairs * gammaln(p.delta1) + p.alpha * np.log(p.beta) - gammaln(p.alpha) + gammaln(A) - A * np.log(B)
        for t in ks[a + 1:]:
            it = np.where(lab == t)[0]
            pr = nk * it.size
            Z = p.zeta + p.delta2 * pr; G = p.gamma + D[np.ix_(ik, it)].sum()
            L2 += (p.delta2 - 1) * logD[np.ix_(ik, it)].sum() - pr * gammaln(p.delta2) + p.zeta * np.log(p.gamma) - gammaln(p.zeta) + gammaln(Z) - Z * np.log(G)
    got = orc.loglik(D, P, lab, 1)
    assert abs(got - (L1 + L2)) <= 1e-9 * abs(L1 + L2)


def test_quirk_q1_accepted_proposals_do_not_move_the_chain(orc, pkg, golden):
    """`state = finalstate` rebinds a local; in an iteration with an accepted proposal the caller's
    labels do not change (single-cluster start: splits are accepted often, K stays 1)."""
    D = golden[3]["distance_matrix"]
    P = _params(orc, pkg, D, golden[3]["cluster_labels"], repulsion=False)
    out = orc.run_chain(D, orc.Options(60, 0, 1, 3, 1), P, np.ones(100, np.int64), 1.0, 0.5, seed=5)
    acc = out["sm_acc"].astype(bool)
    assert acc.sum() > 10
    lab = np.vstack([np.ones((1, 100), np.int64), out["labels"]])
    for i in np.where(acc)[0]:
        assert np.array_equal(lab[i + 1], lab[i])


def test_numMH0_is_pure_gibbs_and_deterministic(orc, pkg, golden):
    D, lab = golden[1]["distance_matrix"], golden[1]["cluster_labels"]
    P = _params(orc, pkg, D, lab)
    a = orc.run_chain(D, orc.Options(30, 5, 2, 5, 0), P, lab, 1.2, 0.4, seed=9)
    b = orc.run_chain(D, orc.Options(30, 5, 2, 5, 0), P, lab, 1.2, 0.4, seed=9)
    assert a["labels"].shape == (12, 100) and np.array_equal(a["labels"], b["labels"])
    assert a["sm_acc"].size == 0
    for s in range(12):   # recorded labels are sortlabels'd: first appearances ascend
        first = [np.where(a["labels"][s] == k)[0][0] for k in range(1, a["K"][s] + 1)]
        assert first == sorted(first) and a["labels"][s].max() == a["K"][s]
    c = orc.run_chain(D, orc.Options(30, 5, 2, 5, 0), P, lab, 1.2, 0.4, seed=10)
    assert not np.array_equal(a["labels"], c["labels"])


def test_prior_restatements(orc, golden):
    """k-medoids and the within / between split of fitprior (prior.jl:55-75) on the reference's fixture."""
    import numpy as np
    D, lab = golden[1]["distance_matrix"], golden[1]["cluster_labels"]
    n = D.shape[0]
    a, med, conv, cost = orc.kmedoids_fixed_point(D, 40, [3, 50, 77, 90])
    assert conv and set(np.unique(a)) == {1, 2, 3, 4}
    assert np.array_equal(a[med], np.arange(1, 5))                       # every medoid sits in its own cluster
    for c in range(4):                                                   # no member beats the medoid (fixed point)
        mem = np.where(a == c + 1)[0]
        assert D[np.ix_(mem, mem)].sum(0).min() >= D[np.ix_(mem, [med[c]])].sum() - 1e-9
    assert np.array_equal(a - 1, np.argmin(D[med], axis=0))
    st = orc.pair_stats(D, lab)
    assert st["nA"] + st["nB"] == n * (n - 1) // 2
    assert st["nA"] == sum(int(s) * (int(s) - 1) // 2 for s in np.bincount(lab))
    assert abs(st["sA"] + st["sB"] - np.triu(D, 1).sum()) < 1e-8
