"""CPU suite: the N > 1 host logic over the gloo backend, world_size 2 -- shard bounds, the [3][C] partial block, the
SUM/MIN all-reduce and the combination rule. The GPU evaluator is replaced by the CPU checker for each rank's shard,
so what is exercised is exactly the part of bcm3_b200/parallel.py that runs between the kernels and the sampler."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port_no, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from bcm3_b200 import synthetic as syn
    from bcm3_b200.parallel import allreduce_partial, combine_partials, partial_from_patient_ll, shard_bounds

    from tests.util import make_nan_inf_case

    prob, vals = make_nan_inf_case()
    chk = oracle.load("port")
    full = chk.poppk_evaluate(prob, vals, want_patient_ll=True)
    lo, hi = shard_bounds(61, rank, world)
    partial = torch.from_numpy(partial_from_patient_ll(full["patient_ll"][:, lo:hi], lo))
    allreduce_partial(partial)
    logp, status = combine_partials(partial.numpy())
    np.save(os.path.join(out_dir, f"logp_{rank}.npy"), logp)
    if rank == 0:
        np.save(os.path.join(out_dir, "want.npy"), full["logp"])
    dist.destroy_process_group()


def test_two_ranks_reproduce_the_serial_result(built, tmp_path):
    port_no = 29500 + (os.getpid() % 2000)
    mp.start_processes(_worker, args=(2, port_no, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    a = np.load(tmp_path / "logp_0.npy")
    b = np.load(tmp_path / "logp_1.npy")
    want = np.load(tmp_path / "want.npy")
    assert np.array_equal(a, b, equal_nan=True)  # every rank holds the same result
    assert want[0] == -np.inf and a[0] == -np.inf
    assert want[1] == -np.inf and a[1] == -np.inf
    assert np.isnan(want[2]) and np.isnan(a[2])


def test_shard_bounds_cover_everything():
    from bcm3_b200.parallel import shard_bounds

    for P in (0, 1, 7, 1000, 100000):
        for W in (1, 2, 3, 8):
            edges = [shard_bounds(P, r, W) for r in range(W)]
            assert edges[0][0] == 0 and edges[-1][1] == P
            assert all(edges[i][1] == edges[i + 1][0] for i in range(W - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_partial_block_matches_serial_loop():
    from bcm3_b200.parallel import combine_partials, partial_from_patient_ll

    rng = np.random.default_rng(4)
    for _ in range(200):
        P = int(rng.integers(1, 12))
        ll = -rng.uniform(1, 5, size=(1, P))
        for j in range(P):
            u = rng.uniform()
            if u < 0.15:
                ll[0, j] = -np.inf
            elif u < 0.3:
                ll[0, j] = np.nan
        # serial semantics of LikelihoodPopPKTrajectory.cpp:427-440
        s = 0.0
        for j in range(P):
            s += ll[0, j]
            if s == -np.inf:
                break
        cut = int(rng.integers(0, P + 1))
        a = partial_from_patient_ll(ll[:, :cut], 0)
        b = partial_from_patient_ll(ll[:, cut:], cut)
        tot = np.stack([a[0] + b[0], np.minimum(a[1], b[1]), np.minimum(a[2], b[2])])
        got, _ = combine_partials(tot)
        assert (np.isnan(s) and np.isnan(got[0])) or s == got[0] or abs(s - got[0]) < 1e-12
