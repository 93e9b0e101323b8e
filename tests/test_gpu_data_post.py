"""GPU tests of the data kernels (distance matrix, log D, validation), the PSM and the MPEL search, through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_distance_matrix_golden_and_oracle(pkg, orc, golden, monkeypatch):
    monkeypatch.setenv("RCB200_DISTM", "exact")                                # the ascending-coordinate kernel: the oracle's operation order
    for k in (1, 2, 3):
        pts, ref = golden[k]["points"], golden[k]["distance_matrix"]
        data = pkg.MCMCData.from_points(pts)
        D = data.D
        assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0)               # exact symmetry, zero diagonal
        off = ~np.eye(100, dtype=bool)
        assert (np.abs(D[off] - ref[off]) / ref[off]).max() < 1e-10            # reference fixture (Julia / Distances.jl)
        assert np.array_equal(D, orc.distm(pts))                               # same operation order as the oracle
        assert np.array_equal(data.logD, orc.logdist(D))
        import ctypes as C
        qd, ql = C.c_int(), C.c_int()
        orc.lib().rco_fixedpoint_scales(D.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(100), C.byref(qd), C.byref(ql))
        assert data.scales() == (qd.value, ql.value)
    # vector-of-vectors constructor (MCMCData(points), types.jl:159-162) and a ragged dimension / size
    data = pkg.MCMCData([list(p) for p in golden[1]["points"]])
    assert np.array_equal(data.D, orc.distm(golden[1]["points"]))
    X = np.random.default_rng(0).normal(size=(333, 37))
    assert np.array_equal(pkg.MCMCData.from_points(X).D, orc.distm(X))


def test_data_validation_errors(pkg, golden):
    D = golden[1]["distance_matrix"].copy()
    D[3, 7] += 1e-9
    with pytest.raises(pkg.RCError, match="D must be symmetric."):              # types.jl:149-151
        pkg.MCMCData(D)
    with pytest.raises(pkg.RCError, match="D must be a square matrix."):        # types.jl:152-154
        pkg.MCMCData(np.zeros((3, 4)))
    D = golden[1]["distance_matrix"].copy()
    D[2, 5] = D[5, 2] = 0.0
    with pytest.raises(pkg.RCError) as e:
        pkg.MCMCData(D)
    assert e.value.status == -4
    assert pkg.MCMCData(np.zeros((1, 1))).n == 1


def test_psm_exact_counts(pkg, orc):
    rng = np.random.default_rng(2)
    for S, n, K in ((1, 5, 2), (37, 130, 6), (200, 257, 40)):
        L = rng.integers(1, K + 1, size=(S, n))
        got = pkg.psm(L)
        assert np.array_equal(got, orc.psm_counts(L) / S)          # exact integer counts, one fp64 divide per entry
        assert np.array_equal(got, got.T) and np.all(np.diag(got) == 1.0)
    # labels need not be compact (any positive ids)
    L = rng.integers(1, 5, size=(9, 40)) * 1000 + 7
    assert np.array_equal(pkg.psm(L), orc.psm_counts(L) / 9)


def test_mpel_losses_match_oracle(pkg, orc):
    rng = np.random.default_rng(5)
    L = np.stack([orc.sortlabels(rng.integers(1, rng.integers(2, 9), size=80)) for _ in range(24)])
    L[7] = L[3]                                                                # a duplicate sample
    for loss in ("binder", "omARI", "VI", "ID"):
        sums, best = pkg.mpel_loss_sums(L, loss)
        ref = orc.mpel_loss_sums(L, loss)
        assert np.allclose(sums, ref, rtol=1e-10, atol=1e-12), loss
        # ties between duplicate samples are broken by fp noise in the reference: compare the attained loss
        assert abs(ref[best] - ref.min()) <= 1e-10 * max(1.0, abs(ref.min()))
    a, b = L[0], L[1]
    assert abs(pkg.binderloss(a, b) - orc.binderloss(a, b)) < 1e-12
    assert abs(pkg.binderloss(a, b, normalised=False) - orc.binderloss(a, b, normalised=False)) < 1e-9
    assert abs(pkg.infodist(a, b) - orc.infodist(a, b)) < 1e-12
    assert abs(pkg.infodist(a, b, normalised=False) - orc.infodist(a, b, normalised=False)) < 1e-12
    assert abs(pkg.binderloss(a, a)) < 1e-9 and abs(pkg.infodist(a, a)) < 1e-9   # test/test_pointestimates.jl:3-8


def test_runsampler_and_pointestimate_api(pkg, golden):
    """The README flow: MCMCData(points) -> runsampler(data, options, params) -> getpointestimate, all 7 method / loss
    combinations of test/test_pointestimates.jl:10-16 (here with explicit params / init: fitprior is host code)."""
    pts, lab = golden[1]["points"], golden[1]["cluster_labels"]
    data = pkg.MCMCData.from_points(pts)
    params = pkg.params_from_labels(data.D, lab)
    opts = pkg.MCMCOptionsList(numiters=120, burnin=20, thin=2)
    init = pkg.MCMCState(lab, 1.5, 0.5)
    res = pkg.runsampler(data, opts, params, init, verbose=False)
    assert len(res.clusts) == 50 and res.posterior_coclustering.shape == (100, 100)
    assert np.all(np.diag(res.posterior_coclustering) == 1.0)
    assert res.K.shape == (50,) and res.r_acceptances.shape == (120,) and res.splitmerge_splits.shape == (120,)
    assert 0 <= res.splitmerge_acceptance_rate <= 1 and res.runtime > 0 and abs(res.mean_iter_time - res.runtime / 120) < 1e-12
    assert len(res.K_acf) == min(49, round(10 * np.log10(50))) + 1 and res.K_acf[0] == pytest.approx(1.0)
    for kw in (dict(method="MAP"), dict(method="MLE"), dict(method="MPEL", loss="VI"), dict(method="MPEL", loss="binder"),
               dict(method="MPEL", loss="omARI"), dict(method="MPEL", loss="ID"),
               dict(method="MPEL", loss=lambda x, y: float(np.mean(np.asarray(x) != np.asarray(y))))):
        clust, i = pkg.getpointestimate(res, **kw)
        assert 0 <= i < 50 and np.array_equal(clust, res.clusts[i])
    # numMH = 0 (pure Gibbs, test/test_sampler.jl:7-10) and multi-chain
    rs = pkg.runsampler(data, pkg.MCMCOptionsList(numiters=30, numMH=0), params, init, verbose=False, nchains=3, seed=4)
    assert len(rs) == 3 and not np.array_equal(rs[0].K, rs[1].K) or not np.array_equal(rs[0].r, rs[1].r)
    assert rs[0].splitmerge_acceptance_rate == 0


def test_statistical_agreement_with_independent_streams(pkg, orc, golden):
    """Second check of the north star: with INDEPENDENT random streams the posterior K distribution and the PSM of the
    GPU chains agree with the oracle's within Monte Carlo error."""
    D, lab = golden[1]["distance_matrix"], golden[1]["cluster_labels"]
    params = pkg.params_from_labels(D, lab)
    P = orc.make_params(**{k: getattr(params, k) for k in params._fields})
    nch, iters, burn = 16, 400, 100
    opts = pkg.MCMCOptionsList(numiters=iters, burnin=burn, thin=1)
    smp = pkg.Sampler(pkg.MCMCData(D), opts, params, np.tile(lab, (nch, 1)), [1.5] * nch, [0.5] * nch, seed=1234)
    smp.run(-1)
    gK = np.concatenate([smp.samples(c)["K"] for c in range(nch)])
    gpsm = smp.psm(0, nch)
    oK, ocnt = [], np.zeros((100, 100))
    for c in range(nch):
        o = orc.run_chain(D, orc.Options(iters, burn, 1, 5, 1), P, lab, 1.5, 0.5, seed=99, chain=c)
        oK.append(o["K"]); ocnt += orc.psm_counts(o["labels"])
    oK = np.concatenate(oK); opsm = ocnt / (nch * (iters - burn))
    assert abs(gK.mean() - oK.mean()) < 0.6
    assert np.abs(gpsm - opsm).mean() < 0.03


@pytest.mark.parametrize("mode", [None, "dmma32"])
def test_distance_matrix_fp64_tensor_core_path(pkg, orc, golden, monkeypatch, mode):
    """The default distance build: Gram blocks on the FP64 tensor cores (mma.sync m8n8k4 f64 -> DMMA, 128 x 128 tiles;
    "dmma32": round 1's small-tile kernel).  Different summation order than the oracle, so 1e-10 relative (north-star
    tolerance) against the reference's fixture and the oracle; symmetry and zero diagonal stay exact; sizes that are
    not multiples of the tile or of the staged chunk."""
    if mode:
        monkeypatch.setenv("RCB200_DISTM", mode)
    else:
        monkeypatch.delenv("RCB200_DISTM", raising=False)
    for n, dim in ((1, 3), (7, 1), (129, 16), (300, 5), (257, 130)):
        X = np.random.default_rng(n).normal(size=(n, dim))
        D = pkg.MCMCData.from_points(X).D if n > 1 else None
        if n > 1:
            ref = orc.distm(X)
            off = ~np.eye(n, dtype=bool)
            assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0)
            assert (np.abs(D[off] - ref[off]) / ref[off]).max() < 1e-10
    for k in (1, 2, 3):
        pts, ref = golden[k]["points"], golden[k]["distance_matrix"]
        D = pkg.MCMCData.from_points(pts).D
        assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0)
        off = ~np.eye(100, dtype=bool)
        assert (np.abs(D[off] - ref[off]) / ref[off]).max() < 1e-10
    X = np.random.default_rng(1).normal(size=(1000, 77))
    D = pkg.MCMCData.from_points(X).D
    ref = orc.distm(X)
    off = ~np.eye(1000, dtype=bool)
    assert np.array_equal(D, D.T) and (np.abs(D[off] - ref[off]) / ref[off]).max() < 1e-10


def test_psm_tensor_core_and_compare_kernels_agree(pkg, orc, monkeypatch):
    """The tcgen05 int8 one-hot kernel (slot widths 32 / 64 / 128, ragged tiles, sample counts that are not a
    multiple of a stage) and the byte-compare kernel both give the exact co-clustering counts (src/mcmc.jl:560)."""
    rng = np.random.default_rng(5)
    for n, S, K in ((300, 37, 20), (129, 5, 3), (1000, 130, 64), (777, 64, 128), (513, 257, 33), (2000, 300, 50), (260, 3, 200)):
        L = rng.integers(1, K + 1, size=(S, n)).astype(np.int64)
        L[0, :K] = np.arange(1, K + 1)
        want = orc.psm_counts(L) / S
        for mode in ("tc", "compare"):
            monkeypatch.setenv("RCB200_PSM", mode)
            assert np.array_equal(pkg.psm(L), want), (n, S, K, mode)


def test_mpel_larger_sample_sets(pkg, orc):
    """Column blocks (more samples than one CTA's block), ragged n (unaligned label rows) and sorted labels."""
    rng = np.random.default_rng(11)
    for n, S, K in ((203, 21, 5), (640, 19, 30)):
        base = np.sort(rng.integers(1, K + 1, size=n))
        L = np.tile(base, (S, 1)); flip = rng.random(L.shape) < 0.1; L[flip] = rng.integers(1, K + 1, size=int(flip.sum()))
        for loss in ("binder", "omARI", "VI", "ID"):
            sums, best = pkg.mpel_loss_sums(L, loss)
            ref = orc.mpel_loss_sums(np.stack([orc.sortlabels(l) for l in L]), loss)
            assert np.allclose(sums, ref, rtol=1e-10, atol=1e-12), (n, S, loss)
            assert abs(ref[best] - ref.min()) <= 1e-10 * max(1.0, abs(ref.min()))


def test_kmedoids_device_matches_fixed_point_restatement(pkg, orc, golden):
    """rc_kmedoids (fitprior's elbow scan and runsampler's default init, prior.jl:55-71, mcmc.jl:519-527)."""
    from redclust_jl_b200.prior import kmedoids_device
    rng = np.random.default_rng(4)
    cases = [(golden[1]["distance_matrix"], 10), (golden[2]["distance_matrix"], 3)]
    X = rng.normal(size=(700, 6)); X[:350] += 2.5
    cases.append((None, 7))
    for D, k in cases:
        data = pkg.MCMCData(D) if D is not None else pkg.MCMCData.from_points(X)
        Dh = data.D
        qD, _ = data.scales()
        init = rng.choice(data.n, size=k, replace=False)
        got = kmedoids_device(data, k, init_medoids=init)
        a, med, conv, cost_q = orc.kmedoids_fixed_point(Dh, qD, init)
        assert np.array_equal(got["assignments"], a) and np.array_equal(got["medoids"], med) and got["converged"] == conv
        assert got["totalcost"] == cost_q / 2.0 ** qD
        assert abs(got["totalcost"] - Dh[med[a - 1], np.arange(data.n)].sum()) <= 1e-10 * got["totalcost"]
    # seeded path (k-medoids++ from medoid rows) and the degenerate k = 1 / k = n
    data = pkg.MCMCData(golden[1]["distance_matrix"])
    r = pkg.kmedoids(data, 10, rng=3)
    assert r["converged"] and sorted(np.unique(r["assignments"])) == list(range(1, 11))
    assert np.all(pkg.kmedoids(data, 1)["assignments"] == 1)
    assert sorted(pkg.kmedoids(data, data.n)["assignments"]) == list(range(1, data.n + 1))


def test_pair_stats_and_fitprior_on_device(pkg, orc, golden):
    """Within / between sufficient statistics (prior.jl:73-110) against numpy; fitprior end to end on the device."""
    D, lab = golden[1]["distance_matrix"], golden[1]["cluster_labels"]
    data = pkg.MCMCData(D)
    st = pkg.pair_stats(data, lab)
    ref = orc.pair_stats(D, lab)
    assert st["nA"] == ref["nA"] and st["nB"] == ref["nB"]
    for f in ("sA", "lA", "sB", "lB"):
        assert abs(st[f] - ref[f]) <= 1e-10 * abs(ref[f]), f
    p_dev = pkg.params_from_labels(data, lab)
    p_host = pkg.params_from_labels(D, lab)
    for f in ("delta1", "alpha", "beta", "delta2", "zeta", "gamma"):
        assert abs(getattr(p_dev, f) - getattr(p_host, f)) <= 1e-8 * abs(getattr(p_host, f)), f
    params = pkg.fitprior(data, "k-medoids", True, Kmin=1, Kmax=20, verbose=False, rng=1)
    assert 2 <= params.K_initial <= 20 and params.delta1 > 0 and params.delta2 > 0 and params.beta > 0 and params.gamma > 0
    pts = golden[1]["points"] if "points" in golden[1] else None
    if pts is not None:
        p2 = pkg.fitprior(list(np.asarray(pts)), "k-means", False, Kmin=1, Kmax=15, verbose=False, rng=1)   # vector of observations
        assert p2.K_initial >= 2 and p2.alpha > 0


def test_sharded_entry_points_single_rank(pkg, golden):
    """Row-block distance build and row-block MPEL (the multi-GPU paths of SURVEY 8e) with one rank: bit-equal to the
    plain calls.  tools/multigpu_check.py runs the same comparison across ranks over NCCL."""
    rng = np.random.default_rng(12)
    X = rng.normal(size=(333, 17))
    a, b = pkg.MCMCData.from_points(X), pkg.MCMCData.from_points_sharded(X)
    assert np.array_equal(a.D, b.D) and np.array_equal(a.logD, b.logD) and a.scales() == b.scales()
    L = np.stack([rng.integers(1, 7, size=90) for _ in range(37)])
    assert np.array_equal(pkg.psm_sharded(L), pkg.psm(L))
    for loss in ("binder", "omARI", "VI", "ID"):
        s1, b1 = pkg.mpel_loss_sums(L, loss)
        s2, b2 = pkg.mpel_loss_sums_sharded(L, loss)
        assert np.array_equal(s1, s2) and b1 == b2, loss


def test_fitprior2_and_summaries(pkg, orc, golden):
    """fitprior2 (prior.jl:152-277) on the device pieces; evaluateclustering / summarise (summaries.jl:13-44)."""
    import io
    D, lab = golden[1]["distance_matrix"], golden[1]["cluster_labels"]
    data = pkg.MCMCData(D)
    p = pkg.fitprior2(data, "k-medoids", True, Kmin=2, Kmax=12, verbose=False, rng=2)
    assert 2 <= p.K_initial <= 12
    for f in ("delta1", "delta2", "alpha", "beta", "zeta", "gamma", "eta", "sigma", "u", "v", "proposalsd_r"):
        assert getattr(p, f) > 0 and np.isfinite(getattr(p, f)), f
    # a single K: the shapes are those of that clustering's within / between fits
    p1 = pkg.fitprior2(data, "k-medoids", True, Kmin=6, Kmax=6, verbose=False, rng=2)
    notional = pkg.kmedoids(data, 6, rng=2)["assignments"]
    st = pkg.pair_stats(data, notional)
    from redclust_jl_b200.prior import gamma_shape_from_stats
    assert abs(p1.delta1 - gamma_shape_from_stats(st["sA"] / st["nA"], st["lA"] / st["nA"])) < 1e-9 * p1.delta1
    assert abs(p1.delta2 - gamma_shape_from_stats(st["sB"] / st["nB"], st["lB"] / st["nB"])) < 1e-9 * p1.delta2
    rng = np.random.default_rng(0)
    other = lab.copy(); idx = rng.choice(lab.size, 15, replace=False); other[idx] = rng.integers(1, lab.max() + 1, size=15)
    ev = pkg.evaluateclustering(other, lab)
    assert abs(ev["nbloss"] - orc.binderloss(other, lab)) < 1e-12 and abs(ev["id"] - orc.infodist(other, lab, normalised=False)) < 1e-12
    assert 0 < ev["ari"] < 1 and 0 < ev["nmi"] < 1 and abs(ev["nvi"] - ev["vi"] / np.log(lab.size)) < 1e-15
    same = pkg.evaluateclustering(lab, lab)
    assert abs(same["nbloss"]) < 1e-12 and abs(same["ari"] - 1) < 1e-12 and abs(same["nmi"] - 1) < 1e-12 and abs(same["vi"]) < 1e-9
    buf = io.StringIO(); pkg.summarise(other, lab, io=buf)
    assert "Adjusted Rand Index" in buf.getvalue() and "Number of clusters" in buf.getvalue()
    with pytest.raises(ValueError):
        pkg.evaluateclustering(lab[:-1], lab)


def test_reference_fitprior_and_sampler_cases(pkg, golden):
    """The reference's own test/test_fitprior.jl:1-22 and test/test_sampler.jl:1-11, case by case."""
    import warnings
    pts = np.asarray(golden[1]["points"]); distM = golden[1]["distance_matrix"]
    pnts = list(pts)                                                     # Vector{Vector{Float64}}
    N = len(pnts)
    for fit in (pkg.fitprior, pkg.fitprior2):
        kw = dict(verbose=False, rng=3)
        if fit is pkg.fitprior2:
            kw.update(Kmin=1, Kmax=12)                                   # the reference scans 1:N/2; keep the test short
        fit(pnts, "k-means", False, **kw)                                # kmeans with points
        fit(pnts, "k-medoids", False, **kw)                              # kmedoids with points
        fit(distM, "k-medoids", True, **kw)                              # kmedoids with distances
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            p = fit(pnts, "k-means", False, verbose=False, Kmin=1, Kmax=1, rng=3)
            assert p.K_initial == 1 and len(w) >= 1                      # single cluster: repulsion defaults + warning
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            p = fit(pnts, "k-means", False, verbose=False, Kmin=N, Kmax=N, rng=3)
            assert p.K_initial == N and len(w) >= 1                      # all singletons: cohesion defaults + warning
        for bad in (lambda: fit(distM, "hierarchical", True, verbose=False),
                    lambda: fit(pnts, "k-means", True, verbose=False),
                    lambda: fit(distM, "k-means", True, verbose=False),
                    lambda: fit(pnts, "k-means", False, Kmin=0, verbose=False),
                    lambda: fit(pnts, "k-means", False, Kmin=N, Kmax=1, verbose=False),
                    lambda: fit(pnts, "k-means", False, Kmax=N + 1, verbose=False)):
            with pytest.raises(ValueError):                               # ArgumentError
                bad()
    repr(p)
    # test_sampler.jl: defaults (fitprior inside, k-medoids init, 5000 iterations), then pure Gibbs
    for data in (pkg.MCMCData.from_points(pts), pkg.MCMCData(distM)):
        repr(data)
        result = pkg.runsampler(data, verbose=False)
        assert len(result.clusts) == 4000 and result.posterior_coclustering.shape == (N, N)
    options = pkg.MCMCOptionsList(numMH=0)
    repr(options)
    result = pkg.runsampler(data, options, verbose=False)
    repr(result)
    assert np.all(np.diag(result.posterior_coclustering) == 1.0)


def test_reference_datagen_cases(pkg):
    """test/test_datagen.jl:1-15, plus the oracle co-clustering against a direct numpy evaluation of utils.jl:130-143."""
    K, N, dim, sig = 10, 100, 10, 0.25
    data = pkg.generatemixture(N, K, alpha=10, sigma=sig, dim=dim, rng=5)
    pnts, distM, clusts, probs, occ = (data[k] for k in ("points", "distancematrix", "clusts", "probs", "oracle_coclustering"))
    assert len(pnts) == N and len(pnts[0]) == dim and len(np.unique(clusts)) == K
    assert distM.shape == (N, N) and np.array_equal(distM, distM.T) and abs(probs.sum() - 1) < 1e-12
    assert occ.shape == (N, N) and np.allclose(occ, occ.T) and np.all(occ >= 0) and np.all(occ <= 1 + 1e-12)
    same = clusts[:, None] == clusts[None, :]
    assert occ[same].mean() > 0.5 > occ[~same].mean()                     # co-clustered pairs look co-clustered
    # one Dirichlet draw evaluated as the reference does: P[j, i] = w_j pdf_j(x_i) / sum_j, P' P
    from redclust_jl_b200.host import _oracle_coclustering
    X = np.asarray(pnts)
    g = np.random.default_rng(9)
    got = _oracle_coclustering(X, K, 10.0, 1.0, sig, np.random.default_rng(9), 0, numiters=3)
    W = g.dirichlet(np.full(K, 10.0), size=3)
    C = np.eye(K, dim)
    ref = np.zeros((N, N))
    for w in W:
        pdf = np.exp(-((X[None, :, :] - C[:, None, :]) ** 2).sum(2) / (2 * sig * sig)) * w[:, None]
        P = pdf / pdf.sum(0, keepdims=True)
        ref += P.T @ P
    assert np.allclose(got, ref / 3, rtol=1e-10, atol=1e-14)


def test_sample_rp_matches_oracle(pkg, orc):
    """sample_rp (mcmc.jl:592-636, the (r, p)-only chain of fitprior) on the device: bit-identical to the oracle's
    restatement on the same structured stream, with the default and with fitted hyperparameters."""
    from redclust_jl_b200.prior import sample_rp
    sizes = [10, 0, 12, 9, 11, 0, 58]
    for params, kw in ((None, dict(eta=1.0, sigma=1.0, u=1.0, v=1.0, proposalsd_r=1.0)),
                       (pkg.PriorHyperparamsList(eta=4.0, sigma=2.0, u=2.0, v=20.0), dict(eta=4.0, sigma=2.0, u=2.0, v=20.0))):
        got = sample_rp(sizes, numiters=600, burnin=100, thin=3, params=params, rng=17)
        ref = orc.sample_rp(sizes, orc.Options(600, 100, 3, 5, 1), orc.make_params(**kw), seed=17)
        assert len(got["r"]) == 166
        for k in ("r", "p", "r_acc"):
            assert np.array_equal(got[k], ref[k]), k
        assert np.all(got["r"] > 0) and np.all((got["p"] > 0) & (got["p"] < 1)) and 0 < got["r_acc"].mean() < 1


def test_generatemixture_oracle_coclustering(pkg, golden):
    """generatemixture's oracle co-clustering matrix (utils.jl:130-143) on the FP64 tensor cores: (i) against a direct
    numpy evaluation with the same Dirichlet draws, (ii) against the matrix the reference ships for the same points in
    data/example_datasets.h5 (5000 different draws there: Monte Carlo error only)."""
    import ctypes as C
    from redclust_jl_b200._lib import lib, check, ptr
    g = np.random.default_rng(3)
    for k, sigma in ((1, 0.25), (3, 0.18)):
        pts, cc = golden[k]["points"], golden[k]["oracle_coclustering_probabilities"]
        n, dim = pts.shape
        K, T = 10, 700
        W = np.ascontiguousarray(g.dirichlet(np.full(K, 10.0), size=T))
        out = np.zeros((n, n))
        check(lib().rc_oracle_coclustering(ptr(np.ascontiguousarray(pts)), dim, n, K, 1.0, sigma, ptr(W), T, 0, ptr(out)))
        centres = np.eye(K, dim)
        d2 = ((pts[:, None, :] - centres[None, :, :]) ** 2).sum(-1)                   # n x K
        ref = np.zeros((n, n))
        for t in range(T):
            P = W[t][None, :] * np.exp(-d2 / (2 * sigma * sigma))
            P /= P.sum(1, keepdims=True)
            ref += P @ P.T
        ref /= T
        assert np.allclose(out, ref, rtol=1e-10, atol=1e-13)
        assert np.array_equal(out, out.T)
        assert np.abs(out - cc).max() < 0.05 and np.abs(out - cc).mean() < 0.005
    mix = pkg.generatemixture(300, 6, alpha=6, sigma=0.2, dim=8, rng=5)
    occ = mix["oracle_coclustering"]
    same = mix["clusts"][:, None] == mix["clusts"][None, :]
    assert occ.shape == (300, 300) and np.array_equal(occ, occ.T) and occ[same].mean() > occ[~same].mean() + 0.3


@pytest.mark.gpu
def test_kmeans_device_matches_lloyd_restatement(pkg, orc, golden):
    """rc_kmeans (fitprior's elbow scan with algo = "k-means", prior.jl:63-69) against the numpy Lloyd iterations from the
    same seeds: same labels, centres and objective to 1e-10; a fixed point of both steps; k-means++ seeding reproducible."""
    rng = np.random.default_rng(11)
    for n_ex, k in ((1, 4), (2, 7), (3, 12)):
        P = np.ascontiguousarray(golden[n_ex]["points"])                  # n x dim
        init = rng.choice(P.shape[0], size=k, replace=False)
        got = pkg.kmeans(P.T, k, init=init)
        a, cent, cost, conv, its = orc.kmeans_lloyd(P, init)
        assert np.array_equal(got["assignments"], a) and got["converged"] == conv and got["iterations"] == its
        assert np.allclose(got["centers"].T, cent, rtol=1e-10, atol=1e-12) and abs(got["totalcost"] - cost) <= 1e-10 * cost
    # larger, well separated mixture: converged result is a fixed point and recovers the truth
    g = np.random.default_rng(5)
    K, n, dim = 20, 5000, 30
    mu = g.normal(size=(K, dim)) * 10
    z = g.integers(0, K, n)
    X = mu[z] + 0.3 * g.normal(size=(n, dim))
    r = pkg.kmeans(X.T, K, rng=7)
    r2 = pkg.kmeans(X.T, K, rng=7)
    assert r["converged"] and np.array_equal(r["assignments"], r2["assignments"]) and r["totalcost"] == r2["totalcost"]
    C = r["centers"].T
    d = ((X[:, None, :] - C[None]) ** 2).sum(2)
    assert np.array_equal(np.argmin(d, 1) + 1, r["assignments"])
    for c in range(K):
        m = r["assignments"] == c + 1
        if m.any():
            assert np.allclose(C[c], X[m].mean(0), rtol=1e-10, atol=1e-12)
    assert abs(r["totalcost"] - d.min(1).sum()) <= 1e-10 * r["totalcost"]
    # k-means++ from the same uniforms never does worse than 3 x the truth's own cost here, and k = 1 / k = n are exact
    truth_cost = ((X - np.stack([X[z == c].mean(0) for c in range(K)])[z]) ** 2).sum()
    assert r["totalcost"] < 3 * truth_cost
    one = pkg.kmeans(X.T, 1, rng=0)
    assert np.all(one["assignments"] == 1) and np.allclose(one["centers"][:, 0], X.mean(0), rtol=1e-10)
    small = X[:50]
    alln = pkg.kmeans(small.T, 50, rng=1)
    assert alln["totalcost"] < 1e-18 * 50 or len(set(alln["assignments"])) >= 45
    with pytest.raises(pkg.ArgumentError):
        pkg.kmeans(small.T, 51)


@pytest.mark.gpu
def test_kmedoids_host_matrix_runs_on_device(pkg, golden):
    D = golden[1]["distance_matrix"]
    r = pkg.kmedoids(D, 10, rng=0)
    assert r["assignments"].min() == 1 and r["assignments"].max() == 10 and r["converged"]
    data = pkg.MCMCData(D)
    r2 = pkg.kmedoids(data, 10, rng=0)
    assert np.array_equal(r["assignments"], r2["assignments"]) and r["totalcost"] == r2["totalcost"]


@pytest.mark.gpu
def test_kmedoids_seeding_on_device(pkg, golden):
    """rc_kmedoids_seed: the first medoid is point floor(u0 n); medoid t is the first point whose cumulative cost (distance to
    the nearest medoid so far) exceeds u_t x the total -- checked against numpy cumulative sums of the same rows."""
    import ctypes as C
    from redclust_jl_b200._lib import lib, check, ptr
    D = golden[2]["distance_matrix"]
    data = pkg.MCMCData(D)
    n, k = D.shape[0], 9
    u = np.random.default_rng(4).random(k)
    med = np.zeros(k, np.int64)
    check(lib().rc_kmedoids_seed(data._h, k, ptr(u), ptr(med)))
    assert med[0] == int(u[0] * n) and len(set(med.tolist())) == k
    mind = D[med[0]].copy()
    for t in range(1, k):
        cum = np.cumsum(mind); target = u[t] * cum[-1]; j = med[t]
        assert cum[j] >= target * (1 - 1e-12) and (j == 0 or cum[j - 1] <= target * (1 + 1e-12)) and mind[j] > 0
        mind = np.minimum(mind, D[j])
    # the draw follows the weights: far points are preferred
    hits = np.zeros(n)
    g = np.random.default_rng(0)
    for _ in range(300):
        uu = np.array([0.0, g.random()]); m2 = np.zeros(2, np.int64)
        check(lib().rc_kmedoids_seed(data._h, 2, ptr(uu), ptr(m2)))
        hits[m2[1]] += 1
    w = D[0] / D[0].sum()
    assert abs((hits / 300 * D[0]).sum() - (w * D[0]).sum()) < 0.25 * (w * D[0]).sum()
