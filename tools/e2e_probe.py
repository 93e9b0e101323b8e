"""Phase timings of the end-to-end path of bench.py (host D -> MCMCData -> Sampler -> run(-1) -> samples)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g, bench
pkg = g.load_package()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
chains = 256
X, lab = bench.synth(10000, 50, 100, 0.1, 50, 44)
data = pkg.MCMCData.from_points(X); Dh = data.D
params = pkg.params_from_labels(Dh, lab)
rp = [pkg.init_rp(params, 44, c) for c in range(chains)]
r0 = np.array([x[0] for x in rp]); p0 = np.array([x[1] for x in rp]); labs = np.tile(lab, (chains, 1))
del data
Dp = torch.from_numpy(Dh).pin_memory()
for rep in range(6):
    T = [time.perf_counter()]
    d2 = pkg.MCMCData(Dp.numpy()); T.append(time.perf_counter())
    o2 = pkg.MCMCOptionsList(numiters=steps, burnin=0, thin=1)
    s2 = pkg.Sampler(d2, o2, params, labs, r0, p0, seed=44); T.append(time.perf_counter())
    if rep != 1:
        s2.run(-1); T.append(T[-1])
    else:
        s2.run(0); T.append(time.perf_counter())
        for _ in range(steps): s2.run(1)
    torch.cuda.synchronize(); T.append(time.perf_counter())
    outs = [s2.samples(c) for c in range(chains)]; T.append(time.perf_counter())
    print("rep", rep, "phases:", [round(b - a, 4) for a, b in zip(T, T[1:])], "device s", s2.progress(), flush=True)
    s2.close(); del d2
