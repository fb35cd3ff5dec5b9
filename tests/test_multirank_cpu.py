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


def _cellpop_worker(rank, world, port_no, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from bcm3_b200.parallel import cellpop_average_from_partial, cellpop_partial_from_cell_values, shard_bounds

    from tests.util import load_cellpop_golden

    prob, gold = load_cellpop_golden("cellpop_n5_late_entry")
    full = oracle.load("port").cellpop_evaluate(prob, gold["values"], want_cell_values=True, want_average=True, want_steps=True)
    lo, hi = shard_bounds(prob.num_cells, rank, world)
    status = np.ones((gold["values"].shape[0], hi - lo), dtype=np.int32)
    if rank == 1:
        status[0, 0] = 0  # one failed cell on one rank must reach every rank
    partial = torch.from_numpy(cellpop_partial_from_cell_values(full["cell_values"][:, :, lo:hi], status))
    dist.all_reduce(partial, op=dist.ReduceOp.SUM)
    avg, nfail = cellpop_average_from_partial(partial.numpy())
    np.save(os.path.join(out_dir, f"avg_{rank}.npy"), avg)
    np.save(os.path.join(out_dir, f"nfail_{rank}.npy"), nfail)
    if rank == 0:
        np.save(os.path.join(out_dir, "want_avg.npy"), full["population_average"])
    dist.destroy_process_group()


def test_cell_population_shards_combine_to_the_population_average(built, tmp_path):
    """Cells split over two ranks: per-timepoint sums and counts, one SUM all-reduce, average = sum / count on every rank
    (cells that enter late are NaN before their entry time and must not be counted)."""
    port_no = 31500 + (os.getpid() % 2000)
    mp.start_processes(_cellpop_worker, args=(2, port_no, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    a, b = np.load(tmp_path / "avg_0.npy"), np.load(tmp_path / "avg_1.npy")
    want = np.load(tmp_path / "want_avg.npy")
    assert np.array_equal(a, b)
    assert np.abs(a - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
    assert np.load(tmp_path / "nfail_0.npy").tolist() == np.load(tmp_path / "nfail_1.npy").tolist()
    assert np.load(tmp_path / "nfail_0.npy")[0] == 1
