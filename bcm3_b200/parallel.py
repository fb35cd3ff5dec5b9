"""One process per GPU: patients / cells sharded across ranks, per-chain partials combined inside the library.

The path shards naturally: given a chain's parameter vector every patient is independent, so
rank g owns a contiguous slice of the patients for ALL chains and the only exchange is the combination of the
[3][C] partial block (bcm3b200.h): SUM of the finite per-patient terms, MIN of the first -inf / first NaN patient
index.  The exchange is the LIBRARY's (csrc/comm_host.cuh: one NCCL all-gather stream-ordered right behind the
reduction kernel + a rank-order combination kernel, bit-identical on every rank); the combination rule then reproduces
the reference's serial loop (LikelihoodPopPKTrajectory.cpp:427-440) independently of the number of ranks.
torch.distributed is used for ONE thing here: handing rank 0's communicator id to the other ranks (a C++ host would use
MPI_Bcast or a file for that).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(num_patients: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous balanced slice [lo, hi) of rank `rank` -- the same formula as the C ABI (bcm3b200.cu finalize)."""
    lo = num_patients * rank // world_size
    hi = num_patients * (rank + 1) // world_size
    return lo, hi


def combine_partials(partial: np.ndarray):
    """partial[3][C] -> (logp[C], status[C]); numpy twin of bcm3b200_combine_partials for host-only use."""
    s, first_inf, first_nan = partial[0], partial[1], partial[2]
    logp = np.where(first_nan < first_inf, np.nan, np.where(np.isfinite(first_inf), -np.inf, s))
    return logp, np.isnan(logp).astype(np.int32)


def partial_from_patient_ll(patient_ll: np.ndarray, offset: int) -> np.ndarray:
    """[C][P_shard] per-patient log-likelihoods -> the [3][C] partial block a shard contributes."""
    ll = np.asarray(patient_ll, dtype=np.float64)
    C, P = ll.shape
    idx = (offset + np.arange(P, dtype=np.float64))[None, :]
    nan = np.isnan(ll) | (ll == np.inf)
    ninf = ll == -np.inf
    fin = ~(nan | ninf)
    out = np.empty((3, C))
    out[0] = np.where(fin, ll, 0.0).sum(axis=1)
    out[1] = np.where(ninf, idx, np.inf).min(axis=1, initial=np.inf)
    out[2] = np.where(nan, idx, np.inf).min(axis=1, initial=np.inf)
    return out


def share_comm_id(group=None) -> bytes | None:
    """Rank 0 asks the library for a communicator id and the process group hands it to everyone (any backend: the id is
    128 plain bytes). None when there is one rank only."""
    import torch.distributed as dist

    from . import _lib

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    box = [_lib.comm_unique_id() if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return box[0]


def allreduce_partial(partial, group=None):
    """In-place all-reduce of a torch tensor [3][C]: SUM on row 0, MIN on rows 1-2 (NCCL on device, gloo on host)."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return partial
    dist.all_reduce(partial[0], op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(partial[1:], op=dist.ReduceOp.MIN, group=group)
    return partial


class ShardedPopPKLikelihood:
    """Rank-local GPU evaluator + the cross-rank reduction. Every rank ends up with the same logp[C]
    (so the host-side, deterministic proposal / swap logic can run replicated)."""

    def __init__(self, problem, rank: int, world_size: int, device: int, group=None, block_size: int = 0, library_comm: bool = True):
        import torch

        from .poppk import PopPKEvaluator

        self.torch = torch
        self.rank, self.world_size, self.group = rank, world_size, group
        self.device = torch.device("cuda", device)
        self.evaluator = PopPKEvaluator(problem, device=device, shard_rank=rank, shard_count=world_size, block_size=block_size)
        self.nvar = problem.num_variables
        self._partial = None
        self._h_partial = None
        # the library's own communicator (default); library_comm=False keeps the exchange in torch.distributed (two all-reduces)
        self.library_comm = library_comm and world_size > 1
        if self.library_comm:
            self.evaluator.comm_init(share_comm_id(group))

    def _buffers(self, C: int):
        torch = self.torch
        if self._partial is None or self._partial.shape[1] != C:
            self._partial = torch.empty((3, C), dtype=torch.float64, device=self.device)
            self._h_partial = torch.empty((3, C), dtype=torch.float64, pin_memory=True)
        return self._partial, self._h_partial

    def enqueue(self, values_host):
        """values_host: torch CPU tensor [C][nvar] float64 (pinned for an asynchronous copy). Returns the device
        partial after the all-reduce; nothing is synchronised."""
        torch = self.torch
        C = values_host.shape[0]
        partial, _ = self._buffers(C)
        stream = torch.cuda.current_stream(self.device)
        self.evaluator.enqueue(values_host.data_ptr(), C, self.nvar, partial.data_ptr(), stream.cuda_stream)
        if self.library_comm:
            self.evaluator.exchange(partial.data_ptr(), C, stream.cuda_stream)
        else:
            allreduce_partial(partial, self.group)
        return partial

    def evaluate(self, values_host):
        """Full end-to-end call: H2D of this rank's slice, kernels, exchange, D2H, combination. With the library's
        communicator this is ONE C-ABI call (bcm3b200_evaluate_batch), exactly what a C++ host makes."""
        if self.library_comm:
            return self.evaluator.evaluate(values_host.numpy())
        partial = self.enqueue(values_host)
        _, h = self._buffers(values_host.shape[0])
        h.copy_(partial, non_blocking=True)
        self.torch.cuda.current_stream(self.device).synchronize()
        return combine_partials(h.numpy())

    def close(self):
        self.evaluator.close()


# ---- cell_population: the simulated cells are split over the ranks ----

def cellpop_partial_from_cell_values(cell_values: np.ndarray, cell_status: np.ndarray) -> np.ndarray:
    """CPU statement of cellpop_partial_kernel: cell_values [C][T][cells_local] (NaN = the cell does not exist at that
    time), cell_status [C][cells_local] (1 ok) -> partial [C][2 T + 1] = sums, counts, failed cells."""
    C, T, _ = cell_values.shape
    partial = np.zeros((C, 2 * T + 1))
    exists = ~np.isnan(cell_values)
    partial[:, :T] = np.where(exists, cell_values, 0.0).sum(axis=2)
    partial[:, T:2 * T] = exists.sum(axis=2)
    partial[:, 2 * T] = (cell_status == 0).sum(axis=1)
    return partial


def cellpop_average_from_partial(partial: np.ndarray):
    """Combined partial -> (population average [C][T], failed cells [C]); cellpop_unpack_partial_kernel."""
    T = (partial.shape[1] - 1) // 2
    n = partial[:, T:2 * T]
    with np.errstate(invalid="ignore", divide="ignore"):
        avg = np.where(n > 0, partial[:, :T] / n, 0.0)
    return avg, partial[:, 2 * T].astype(np.int64)


class ShardedCellPopLikelihood:
    """Rank-local GPU evaluator over this rank's slice of the simulated cells + one SUM all-reduce of the per-chain
    partials; every rank ends up with the same logp[C]."""

    def __init__(self, problem, rank: int, world_size: int, device: int, group=None, kernel: str = "auto", library_comm: bool = True):
        import torch

        from .cellpop import CellPopEvaluator

        self.torch = torch
        self.rank, self.world_size, self.group = rank, world_size, group
        self.device = torch.device("cuda", device)
        self.evaluator = CellPopEvaluator(problem, device=device, kernel=kernel, shard_rank=rank, shard_count=world_size)
        self.nvar = problem.num_variables
        self.width = 2 * problem.num_timepoints + 1
        self._partial = None
        self.library_comm = library_comm and world_size > 1
        if self.library_comm:
            self.evaluator.comm_init(share_comm_id(group))

    def evaluate(self, values_host):
        """values_host: torch CPU tensor [C][nvar] float64. H2D, kernels, all-reduce, data likelihood, D2H of logp."""
        torch = self.torch
        C = values_host.shape[0]
        if self.library_comm:  # one C-ABI call: partial, all-gather + rank-order sum, data likelihood, logp back
            return self.evaluator.evaluate(values_host.numpy())
        if self._partial is None or self._partial.shape[0] != C:
            self._partial = torch.empty((C, self.width), dtype=torch.float64, device=self.device)
        stream = torch.cuda.current_stream(self.device)
        self.evaluator.enqueue(values_host.data_ptr(), C, self.nvar, self._partial.data_ptr(), stream.cuda_stream)
        if self.world_size > 1:
            import torch.distributed as dist

            dist.all_reduce(self._partial, op=dist.ReduceOp.SUM, group=self.group)
        return self.evaluator.finish(self._partial.data_ptr(), C, stream.cuda_stream)

    def close(self):
        self.evaluator.close()
