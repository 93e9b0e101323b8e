"""world_size-2 `gloo` test of the multi-GPU host logic: chains shard across ranks with no data-path
collective; PSM counts of the per-rank sample shards are combined by an all-reduce(sum) of the n x n integer matrix.
Runs on CPU: the per-rank chains come from the oracle (the checker), the collective and the sharding rule are the
product's (bench.py / DESIGN.md section 6 use the same rule: global chain id = rank * chains_per_rank + local id)."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    orc, pkg = g.load_oracle(), g.load_package()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gd = np.load(os.path.join(ROOT, "tests", "golden", "example1.npz"))
    D, lab = gd["distance_matrix"], gd["cluster_labels"]
    params = pkg.params_from_labels(D, lab)
    P = orc.make_params(**{k: getattr(params, k) for k in params._fields})
    per_rank, S = 2, 20
    cnt = np.zeros((100, 100), np.int64)
    Ks = []
    for c in range(per_rank):
        chain = rank * per_rank + c                               # global chain id -> RNG stream
        r0, p0 = orc.init_rp(P, 7, chain)
        o = orc.run_chain(D, orc.Options(S, 0, 1, 5, 1), P, lab, r0, p0, seed=7, chain=chain)
        cnt += orc.psm_counts(o["labels"]); Ks.append(o["K"])
    t = torch.from_numpy(cnt.astype(np.int32))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)                      # the one collective of the PSM path
    if rank == 0:
        np.save(out, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_chain_sharding_and_psm_allreduce(tmp_path, orc, pkg, golden):
    out = str(tmp_path / "psm.npy")
    mp.spawn(_worker, args=(2, 29531, out), nprocs=2, join=True)
    got = np.load(out)
    D, lab = golden[1]["distance_matrix"], golden[1]["cluster_labels"]
    params = pkg.params_from_labels(D, lab)
    P = orc.make_params(**{k: getattr(params, k) for k in params._fields})
    ref = np.zeros((100, 100), np.int64)
    for chain in range(4):                                         # the same 4 chains run in one process
        r0, p0 = orc.init_rp(P, 7, chain)
        ref += orc.psm_counts(orc.run_chain(D, orc.Options(20, 0, 1, 5, 1), P, lab, r0, p0, seed=7, chain=chain)["labels"])
    assert np.array_equal(got, ref)
    assert np.all(np.diag(got) == 4 * 20)


def _mpel_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    orc, pkg = g.load_oracle(), g.load_package()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)                                  # the same label vectors on every rank
    L = np.stack([orc.sortlabels(rng.integers(1, 6, size=40)) for _ in range(11)])
    S = L.shape[0]
    rows, per = pkg.cyclic_rows(rank, world, S)                      # the product's sharding rule
    mine = torch.zeros((per, S), dtype=torch.float64)
    for k, i in enumerate(rows):
        for j in range(i + 1, S):
            mine[k, j] = orc.binderloss(L[i], L[j])                  # strict upper triangle of the loss matrix
    flat = torch.empty((world * per, S), dtype=torch.float64)
    dist.all_gather_into_tensor(flat, mine)
    upper = pkg.assemble_cyclic_rows(flat.view(world, per, S), S).numpy()
    if rank == 0:
        np.save(out, upper)
    dist.barrier()
    dist.destroy_process_group()


def test_mpel_candidate_sharding_and_reassembly(tmp_path, orc):
    """SURVEY 8e, MPEL row: candidates dealt out cyclically, one all_gather, column sums in ascending row order."""
    out = str(tmp_path / "upper.npy")
    mp.spawn(_mpel_worker, args=(2, 29533, out), nprocs=2, join=True)
    upper = np.load(out)
    rng = np.random.default_rng(3)
    L = np.stack([orc.sortlabels(rng.integers(1, 6, size=40)) for _ in range(11)])
    full = upper + upper.T
    sums = np.zeros(11)
    for i in range(11):                                              # ascending rows, as k_colsum_upper
        sums += full[i]
    ref = orc.mpel_loss_sums(L, "binder")
    assert np.allclose(sums, ref, rtol=1e-12, atol=1e-14)
    assert np.all(np.tril(upper) == 0) and int(np.argmin(sums)) == int(np.argmin(ref))
