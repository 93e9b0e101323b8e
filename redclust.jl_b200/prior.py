"""Prior fitting: fitprior / sampledist / sampleK and the notional clustering helpers
(/root/reference/src/prior.jl:22-367, src/mcmc.jl:592-636 sample_rp).  SURVEY.md section 8(f) rank 1, the caller
of the hot path.  The O(n^2) parts run on the device-resident matrix behind MCMCData (prior.jl:51,180): the
distance build, k-medoids (rc_kmedoids) and the within / between sufficient statistics of the Gamma fits
(rc_pair_stats), k-means on the points (rc_kmeans) and the (r, p) chain (rc_sample_rp).  The O(K) fits stay on the host;
k-means / k-medoids are third-party Clustering.jl code in the reference and are restated here."""
import math
import warnings
from functools import partial
import numpy as np
from scipy.special import digamma, polygamma, gammaln, betaln


from .host import ArgumentError          # Julia's ArgumentError, one class for the package


# ---- third-party pieces restated (Clustering.jl kmeans / kmedoids, Distributions.fit_mle) -----------
def _gen(rng):
    """One random stream per call tree: a Generator is used as is (fitprior threads a single one through the elbow scan,
    the notional clustering and sample_rp), an int seeds a new one, None draws fresh entropy."""
    return rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)


def kmedoids_device(data, k, maxiter=1000, rng=None, init_medoids=None):
    """Clustering.kmedoids on a device-resident MCMCData (librcb200 rc_kmedoids): same seeding and iteration as
    kmedoids(), the O(n^2) medoid updates run on the GPU over the exact fixed-point image of D."""
    import ctypes as C
    from ._lib import lib, check, ptr
    n = data.n
    g = _gen(rng)
    if init_medoids is None:                          # k-medoids++ on the device, driven by k uniforms of the host stream
        u = np.ascontiguousarray(g.random(k))
        init_medoids = np.zeros(k, np.int64)
        check(lib().rc_kmedoids_seed(data._h, k, ptr(u), ptr(init_medoids)))
    init = np.ascontiguousarray(np.asarray(init_medoids, dtype=np.int64))
    assign = np.zeros(n, np.int64); med = np.zeros(k, np.int64)
    cost, conv, its = C.c_double(), C.c_int32(), C.c_int64()
    check(lib().rc_kmedoids(data._h, k, ptr(init), maxiter, ptr(assign), ptr(med), C.byref(cost), C.byref(conv), C.byref(its)))
    return dict(assignments=assign, medoids=med, totalcost=cost.value, converged=bool(conv.value), iterations=its.value)


def kmedoids(D, k, maxiter=1000, rng=None, device=0):
    """Clustering.kmedoids(D, k; maxiter): alternate assignment / medoid update from a k-medoids++ seeding, on the device.
    D: a device-resident MCMCData, or a host matrix (uploaded into one first)."""
    from .host import MCMCData
    if not isinstance(D, MCMCData):
        D = MCMCData(np.ascontiguousarray(np.asarray(D, dtype=np.float64)), device=device)
    return kmedoids_device(D, k, maxiter=maxiter, rng=rng)


def kmeans(X, k, maxiter=1000, rng=None, init=None, tol=1e-6, device=0):
    """Clustering.kmeans(X, k; maxiter) with X dim x n (columns are observations), on the device (librcb200 rc_kmeans):
    k-means++ seeding driven by k uniforms of `rng` (or the 0-based points `init`), Lloyd iterations until no label
    changes or the objective moves by less than tol (Clustering.jl's default 1e-6).  centers: dim x k."""
    import ctypes as C
    from ._lib import lib, check, ptr
    P = np.ascontiguousarray(np.asarray(X, dtype=np.float64).T)           # n x dim, a point per row
    n, dim = P.shape
    if not 1 <= k <= n:
        raise ArgumentError("Number of clusters must satisfy 1 ≤ k ≤ n")
    idx = None if init is None else np.ascontiguousarray(np.asarray(init, dtype=np.int64))
    u = None if init is not None else np.ascontiguousarray(_gen(rng).random(k))
    assign = np.zeros(n, np.int64); cent = np.zeros((k, dim))
    cost, conv, its = C.c_double(), C.c_int32(), C.c_int64()
    check(lib().rc_kmeans(ptr(P), n, dim, k, ptr(idx), ptr(u), maxiter, float(tol), device, ptr(assign), ptr(cent),
                          C.byref(cost), C.byref(conv), C.byref(its)))
    return dict(assignments=assign, centers=np.ascontiguousarray(cent.T), totalcost=cost.value, converged=bool(conv.value),
                iterations=its.value)


def gamma_shape_from_stats(mx, mlx):
    """Gamma MLE shape from the sample mean and the mean of logs (Newton on log(a) - digamma(a) = log(mx) - mlx)."""
    s = math.log(mx) - mlx
    if not s > 0:
        return 1e8
    a = (3 - s + math.sqrt((s - 3) ** 2 + 24 * s)) / (12 * s)
    for _ in range(1000):
        f = math.log(a) - float(digamma(a)) - s
        fp = 1 / a - float(polygamma(1, a))
        an = a - f / fp
        if an <= 0:
            an = a / 2
        if abs(an - a) <= 1e-16 * max(1.0, a) * 16:
            a = an
            break
        a = an
    return float(a)


def gamma_mle_shape(x, w=None):
    """shape(fit_mle(Gamma, x[, w]))."""
    x = np.asarray(x, dtype=np.float64)
    if w is None:
        mx, mlx = x.mean(), np.log(x).mean()
    else:
        w = np.asarray(w, dtype=np.float64)
        mx, mlx = (w * x).sum() / w.sum(), (w * np.log(x)).sum() / w.sum()
    return gamma_shape_from_stats(mx, mlx)


def gamma_mle(x, w=None):
    """(shape, rate) of fit_mle(Gamma, x)."""
    a = gamma_mle_shape(x, w)
    x = np.asarray(x, dtype=np.float64)
    m = x.mean() if w is None else float((np.asarray(w) * x).sum() / np.sum(w))
    return a, a / m


def beta_mle(x, maxiter=1000, tol=1e-14):
    """Distributions.fit_mle(Beta, x): Newton on the digamma equations from moment-matched start."""
    x = np.asarray(x, dtype=np.float64)
    m, v = x.mean(), x.var()
    t = m * (1 - m) / v - 1 if v > 0 else 1.0
    a, b = max(m * t, 1e-3), max((1 - m) * t, 1e-3)
    g1, g2 = np.log(x).mean(), np.log1p(-x).mean()
    for _ in range(maxiter):
        f = np.array([digamma(a) - digamma(a + b) - g1, digamma(b) - digamma(a + b) - g2])
        t3 = polygamma(1, a + b)
        J = np.array([[polygamma(1, a) - t3, -t3], [-t3, polygamma(1, b) - t3]])
        step = np.linalg.solve(J, f)
        an, bn = a - step[0], b - step[1]
        if an <= 0: an = a / 2
        if bn <= 0: bn = b / 2
        done = abs(an - a) + abs(bn - b) < tol * (a + b)
        a, b = float(an), float(bn)
        if done:
            break
    return a, b


def detectknee(x, y):
    """prior.jl:340-360."""
    x = np.asarray(x, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
    ind = np.argsort(x, kind="stable")
    x, y = x[ind], y[ind]
    a = (y[-1] - y[0]) / (x[-1] - x[0]) if x[-1] != x[0] else float("nan")
    b = y[0] - a * x[0]
    dist = np.abs(a * x + b - y) / math.sqrt(a * a + 1)
    if np.isnan(dist).all():
        return x[0], y[0]
    i = int(np.nanargmax(dist)) if not np.isnan(dist[0]) else 0
    return x[i], y[i]


def _trunc_logcdf_ratio(r, cand, sd):
    from scipy.stats import norm
    return norm.logcdf(r / sd) - norm.logcdf(cand / sd)


def sample_rp(clustsizes, numiters=5000, burnin=None, thin=1, params=None, rng=None, device=0):
    """sample_rp(clustsizes, options, params) (mcmc.jl:592-636): the (r, p)-only chain for fixed cluster sizes, run on
    the device by rc_sample_rp (one warp on the sampler's own sample_r / sample_p; default hyperparameters when called
    from fitprior: eta = sigma = u = v = 1, proposalsd_r = 1).  `rng`: an int seed, a numpy Generator (one draw seeds the
    device stream) or None (fresh entropy)."""
    from .host import MCMCOptionsList, PriorHyperparamsList
    from ._lib import lib, check, ptr
    import ctypes as C
    burnin = int(math.floor(0.2 * numiters)) if burnin is None else burnin
    opts = MCMCOptionsList(numiters=numiters, burnin=burnin, thin=thin)
    params = PriorHyperparamsList() if params is None else params
    if isinstance(rng, np.random.Generator):
        seed = int(rng.integers(0, 2 ** 63 - 1))
    elif rng is None:
        seed = int(np.random.default_rng().integers(0, 2 ** 63 - 1))
    else:
        seed = int(rng)
    cs = np.ascontiguousarray(np.asarray(clustsizes, dtype=np.int64))
    S = opts.numsamples
    out = dict(r=np.zeros(S), p=np.zeros(S), r_acc=np.zeros(numiters, np.uint8))
    o, q = opts._c(), params._c()
    check(lib().rc_sample_rp(ptr(cs), cs.size, C.byref(o), C.byref(q), C.c_uint64(seed), device, ptr(out["r"]), ptr(out["p"]), ptr(out["r_acc"])))
    return out


def _prepare(data, algo, diss, Kmin, Kmax, device):
    from .host import MCMCData, makematrix
    if isinstance(data, MCMCData):                   # an already resident dissimilarity matrix
        if algo == "k-means":
            raise ArgumentError("Cannot use algorithm `k-means` with a dissimilarity matrix.")
        if algo != "k-medoids":
            raise ArgumentError("Algo must be 'k-means' or 'k-medoids'.")
        N = data.n
        Kmax = N // 2 if Kmax is None else Kmax
        if not (1 <= Kmin <= Kmax <= N):
            raise ArgumentError("Kmin and Kmax must satisfy 1 ≤ Kmin ≤ Kmax ≤ N")
        return None, N, data, Kmax
    if isinstance(data, (list, tuple)):
        if diss:
            raise ArgumentError("diss = true but data is not a dissimilarity matrix. Assuming that the data is a vector of observations.")
        x = makematrix(data)
        diss = False
    else:
        x = np.asarray(data, dtype=np.float64)
    N = x.shape[1]
    if diss and x.shape[0] != x.shape[1]:
        raise ArgumentError("Supplied dissimilarity matrix is not square.")
    if algo == "k-means" and diss:
        raise ArgumentError("Cannot use algorithm `k-means` with a dissimilarity matrix.")
    if algo not in ("k-means", "k-medoids"):
        raise ArgumentError("Algo must be 'k-means' or 'k-medoids'.")
    Kmax = N // 2 if Kmax is None else Kmax
    if not (1 <= Kmin <= Kmax <= N):
        raise ArgumentError("Kmin and Kmax must satisfy 1 ≤ Kmin ≤ Kmax ≤ N")
    dev = MCMCData(np.ascontiguousarray(x), device=device) if diss else MCMCData.from_points(x.T, device=device)   # pairwise(Euclidean(), x, dims=2)
    return x, N, dev, Kmax


def fitprior(data, algo, diss=False, Kmin=1, Kmax=None, verbose=True, device=0, rng=None):
    """fitprior(data, algo, diss = false; Kmin, Kmax, verbose) -> PriorHyperparamsList   (prior.jl:22-128)."""
    from .host import PriorHyperparamsList, pair_stats
    x, N, dev, Kmax = _prepare(data, algo, diss, Kmin, Kmax, device)
    rng = _gen(rng)                                   # one stream for the elbow scan, the notional clustering and sample_rp
    if verbose:
        print("Fitting prior hyperparameters")
    clustfn, inp = (partial(kmeans, device=device), x) if algo == "k-means" else (kmedoids, dev)   # both on the device
    objective = np.zeros(Kmax - Kmin + 1)
    for k in range(1, Kmax - Kmin + 2):            # quirk Q11: clusters with k = loop index (prior.jl:63-64)
        t = clustfn(inp, k, maxiter=1000, rng=rng)
        objective[k - 1] = t["totalcost"]
        if not t["converged"]:
            warnings.warn(f"Clustering did not converge at K = {k}")
    K = int(detectknee(np.arange(Kmin, Kmax + 1), objective)[0])
    notional = clustfn(inp, K, maxiter=1000, rng=rng)["assignments"]
    st = pair_stats(dev, notional)          # A = within-cluster, B = between-cluster upper-triangle dissimilarities (:73-75)
    sizes = np.bincount(notional)[1:]
    t = sample_rp(sizes, rng=rng)
    proposalsd_r = float(np.std(t["r"], ddof=1))
    eta, sigma = gamma_mle(t["r"])
    u, v = beta_mle(t["p"])
    if K == N:
        warnings.warn("Got a notional clustering of entirely singletons. Falling back to defaults for cohesion parameters.")
        d1, al, be = 1.0, 1.0, 1.0
    else:
        d1 = gamma_shape_from_stats(st["sA"] / st["nA"], st["lA"] / st["nA"]); al = st["nA"] * d1; be = float(st["sA"])
    if K == 1:
        warnings.warn("Got a notional clustering with a single cluster. Falling back to defaults for repulsion parameters.")
        d2, ze, ga = 1.0, 1.0, 1.0
    else:
        d2 = gamma_shape_from_stats(st["sB"] / st["nB"], st["lB"] / st["nB"]); ze = st["nB"] * d2; ga = float(st["sB"])
    return PriorHyperparamsList(delta1=d1, delta2=d2, alpha=al, beta=be, zeta=ze, gamma=ga, eta=eta, sigma=sigma,
                                proposalsd_r=proposalsd_r, u=u, v=v, K_initial=K)


def fitprior2(data, algo, diss=False, Kmin=1, Kmax=None, verbose=True, device=0, rng=None):
    """fitprior2(data, algo, diss = false; Kmin, Kmax, verbose) -> PriorHyperparamsList   (prior.jl:152-277): partition
    prior as in fitprior; cohesion / repulsion parameters from the within / between dissimilarities of the clusterings
    with K = Kmin..Kmax clusters, weighted by the prior predictive distribution of K.  The weighted Gamma fits only
    need the weighted counts, sums and log-sums, which come from rc_pair_stats per clustering; k-medoids runs on
    the device.  (The reference clusters for every k in 1:N and then uses Kmin:Kmax; only that range is computed.)"""
    from .host import PriorHyperparamsList, pair_stats
    x, N, dev, Kmax = _prepare(data, algo, diss, Kmin, Kmax, device)
    rng = _gen(rng)                                   # one stream for the elbow scan, the notional clustering and sample_rp
    if verbose:
        print("Fitting prior hyperparameters")
    clustfn, inp = (partial(kmeans, device=device), x) if algo == "k-means" else (kmedoids, dev)
    ks = list(range(Kmin, Kmax + 1))
    objective = np.zeros(len(ks))
    stats = []
    for q, k in enumerate(ks):
        t = clustfn(inp, k, maxiter=1000, rng=rng)
        objective[q] = t["totalcost"]
        if not t["converged"]:
            warnings.warn(f"Clustering did not converge at K = {k}")
        stats.append(pair_stats(dev, t["assignments"]))
    K = int(detectknee(np.arange(Kmin, Kmax + 1), objective)[0])
    notional = clustfn(inp, K, maxiter=1000, rng=rng)["assignments"]
    sizes = np.bincount(notional)[1:]
    t = sample_rp(sizes, rng=rng)
    proposalsd_r = float(np.std(t["r"], ddof=1))
    eta, sigma = gamma_mle(t["r"])
    u, v = beta_mle(t["p"])
    Ks = sampleK(eta, sigma, u, v, max(10000, 100 * N), N, rng=rng)           # prior on K (:232-233)
    Kprior = np.bincount(Ks, minlength=N + 1)[1:N + 1] / Ks.size
    w = np.array([Kprior[k - 1] for k in ks])
    nA = np.array([s["nA"] for s in stats], dtype=np.float64); nB = np.array([s["nB"] for s in stats], dtype=np.float64)
    sA = np.array([s["sA"] for s in stats]); lA = np.array([s["lA"] for s in stats])
    sB = np.array([s["sB"] for s in stats]); lB = np.array([s["lB"] for s in stats])
    if nA.sum() == 0:
        warnings.warn("The ensemble of clusterings has only one clustering, consisting of all singletons. Falling back to defaults for cohesion parameters.")
        d1, al, be = 1.0, 1.0, 1.0
    else:
        W = float((w * nA).sum())                                            # fit_mle(Gamma, A, wtsA) from weighted statistics
        d1 = gamma_shape_from_stats(float((w * sA).sum()) / W, float((w * lA).sum()) / W) if W > 0 else 1e8
        al = W * d1; be = float((w * sA).sum())
    if nB.sum() == 0:
        warnings.warn("The ensemble of clusterings has only one clustering, consisting of a single cluster. Falling back to defaults for repulsion parameters.")
        d2, ze, ga = 1.0, 1.0, 1.0
    else:
        W = float((w * nB).sum())
        d2 = gamma_shape_from_stats(float((w * sB).sum()) / W, float((w * lB).sum()) / W) if W > 0 else 1e8
        ze = W * d2; ga = float((w * sB).sum())
    return PriorHyperparamsList(delta1=d1, delta2=d2, alpha=al, beta=be, zeta=ze, gamma=ga, eta=eta, sigma=sigma,
                                proposalsd_r=proposalsd_r, u=u, v=v, K_initial=K)


def sampledist(params, type, numsamples=1, rng=None):
    """prior.jl:284-308."""
    if type not in ("intercluster", "intracluster"):
        raise ArgumentError('type must be either "intercluster" or "intracluster".')
    if numsamples < 1:
        raise ArgumentError("numsamples must be a positive integer.")
    g = np.random.default_rng(rng)
    a, b, d = (params.alpha, params.beta, params.delta1) if type == "intracluster" else (params.zeta, params.gamma, params.delta2)
    lam = g.gamma(a, 1 / b, size=numsamples)
    return g.gamma(d, 1 / lam)


def sampleK(params_or_eta, *args, rng=None):
    """sampleK(params, numsamples, n) / sampleK(eta, sigma, u, v, numsamples, n)   (prior.jl:316-338)."""
    if hasattr(params_or_eta, "eta"):
        q = params_or_eta
        eta, sigma, u, v = q.eta, q.sigma, q.u, q.v
        numsamples, n = args
    else:
        eta = params_or_eta
        sigma, u, v, numsamples, n = args
    if n < 1:
        raise ArgumentError("n must be a positive integer.")
    if numsamples < 1:
        raise ArgumentError("numsamples must be a positive integer.")
    g = np.random.default_rng(rng)
    K = np.arange(1, n)
    out = np.zeros(numsamples, np.int64)
    for i in range(numsamples):
        r, p = g.gamma(eta, 1 / sigma), g.beta(u, v)
        lp = np.zeros(n)
        with np.errstate(divide="ignore", invalid="ignore"):
            lp[:n - 1] = (r * K) * np.log1p(-p) + (n - K) * np.log(p) - np.log(n - K) - betaln(r * K, n - K)
            lp[n - 1] = r * n * np.log1p(-p)
        lp = lp - lp.min()
        out[i] = int(np.argmax(-np.log(-np.log(g.random(n))) + lp)) + 1
    return out
