"""Throughput of the other BASELINE configs (C2: n=1000 K=20 dim=50, 64 chains; C4: n=2000, 128 chains per GPU;
C5: n=50000, S=10000 samples, PSM + MPEL) -- evidence for profiles/, not the bench line.
usage: configs_probe.py C2|C4|C5 [iters]"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g, bench
pkg = g.load_package()
which = sys.argv[1]
if which in ("C2", "C4"):
    n, K, dim, chains, iters = (1000, 20, 50, 64, 2000) if which == "C2" else (2000, 20, 50, 128, 500)
    if len(sys.argv) > 2: iters = int(sys.argv[2])
    X, lab = bench.synth(n, K, dim, 0.1, K, 44)
    data = pkg.MCMCData.from_points(X)
    params = pkg.params_from_labels(data.D, lab)
    opts = pkg.MCMCOptionsList(numiters=iters + 50, burnin=0, thin=1)
    rp = [pkg.init_rp(params, 44, c) for c in range(chains)]
    smp = pkg.Sampler(data, opts, params, np.tile(lab, (chains, 1)), [a for a, _ in rp], [b for _, b in rp], seed=44)
    smp.run(50)
    _, t0 = smp.progress()
    w0 = time.perf_counter(); smp.run(iters); w1 = time.perf_counter()
    _, t1 = smp.progress()
    s0 = smp.samples(0)
    print(json.dumps({"config": which, "n": n, "K": K, "chains": chains, "iters_in_one_launch": iters, "device_s": t1 - t0, "wall_s": w1 - w0,
                      "chain_sweeps_per_s": chains * iters / (t1 - t0), "us_per_sweep_per_chain": (t1 - t0) / iters * 1e6,
                      "algorithmic_GBps": chains * iters * 16.0 * n * (n - 1) / (t1 - t0) / 1e9, "K_last": int(s0["K"][-1]),
                      "sm_acceptance": float(s0["sm_acc"].mean())}), flush=True)
    orc = g.load_oracle()
    P = orc.make_params(**{k: getattr(params, k) for k in params._fields})
    it_cpu = 20
    ncpu = min(16, os.cpu_count() or 1)
    secs, _ = orc.time_chains(data.D, orc.Options(it_cpu, 0, 1, 5, 1), P, lab, [1.5] * ncpu, [0.9] * ncpu, seed=44, nthreads=ncpu)
    print(json.dumps({"config": which, "cpu_oracle_chain_sweeps_per_s": ncpu * it_cpu / secs, "threads": ncpu}), flush=True)
else:
    n, K, dim, S = 50000, 100, 200, 10000
    if len(sys.argv) > 2: S = int(sys.argv[2])
    os.environ["RCB200_VERBOSE"] = "1"
    rng = np.random.default_rng(0)
    lab = np.sort(rng.integers(1, K + 1, size=n)).astype(np.int64)
    t = time.perf_counter()
    L = np.tile(lab.astype(np.int64), (S, 1))
    for s in range(S):                                 # 3 % of the points move in every sample (in place, bounded memory)
        idx = rng.integers(0, n, size=n * 3 // 100)
        L[s, idx] = rng.integers(1, K + 1, size=idx.size)
    print("labels built", time.perf_counter() - t, flush=True)
    t = time.perf_counter(); P = pkg.psm(L); dt = time.perf_counter() - t
    print(json.dumps({"config": "C5", "piece": "PSM", "n": n, "S": S, "seconds": dt, "diag_ok": bool(np.all(np.diag(P) == 1.0)), "sym": bool(P[123, 4567] == P[4567, 123])}), flush=True)
    del P
    Sm = min(S, int(sys.argv[3]) if len(sys.argv) > 3 else 2000)
    for loss in ("binder", "VI"):
        t = time.perf_counter(); sums, best = pkg.mpel_loss_sums(L[:Sm], loss); dt = time.perf_counter() - t
        print(json.dumps({"config": "C5", "piece": "MPEL", "loss": loss, "n": n, "S": Sm, "seconds": dt, "best": int(best)}), flush=True)
