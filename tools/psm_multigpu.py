"""N-rank check of the sharded PSM (BASELINE configs[4]): chains shard across GPUs, per-rank int32 counts,
one NCCL all_reduce of the n x n matrix.  torchrun --nproc-per-node N tools/psm_multigpu.py [n] [chains_per_rank] [iters]
Rank 0 compares with the single-process PSM of the same global chains and prints timings."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
cpr = int(sys.argv[2]) if len(sys.argv) > 2 else 4
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = g.load_package()
X, lab = bench.synth(n, 20, 50, 0.1, 20, 3)
data = pkg.MCMCData.from_points(X, device=local)
params = pkg.params_from_labels(data.D, lab)
opts = pkg.MCMCOptionsList(numiters=iters, burnin=0, thin=1, numGibbs=5, numMH=1)

def make(chain0, nch, dev_data):
    rp = [pkg.init_rp(params, 7, chain0 + c) for c in range(nch)]
    s = pkg.Sampler(dev_data, opts, params, np.tile(lab, (nch, 1)), [a for a, _ in rp], [b for _, b in rp], seed=7, chain_offset=chain0)
    s.run()
    return s

smp = make(rank * cpr, cpr, data)
torch.cuda.synchronize(); dist.barrier()
t = time.perf_counter()
psm = smp.psm_allreduce()
torch.cuda.synchronize(); dist.barrier()
dt = time.perf_counter() - t
if rank == 0:
    ref = make(0, cpr * world, data).psm()
    print(f"world={world} n={n} chains={cpr * world} samples/chain={iters}: psm_allreduce {dt * 1e3:.1f} ms; "
          f"equal to the single-process PSM of the same chains: {np.array_equal(psm, ref)}", flush=True)
    assert np.array_equal(psm, ref)
dist.barrier()
dist.destroy_process_group()
