"""GPU-backed PopPK likelihood: host-side wrapper of the C ABI (include/bcm3b200.h).

Mirrors the reference's plugin for this path -- ``LikelihoodPopPKTrajectory``
(src/likelihoods/LikelihoodPopPKTrajectory.{h,cpp}) behind ``bcm3::Likelihood``
(src/sampler/Likelihood.h:9-35) -- with the one added entry the batched design needs:
``EvaluateLogProbabilityBatch(values[C, nvar]) -> logp[C]``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .poppk_data import PK_ONE, PK_TYPE_NAMES, PopPKProblem

NUM_COUNTERS = 8


class PopPKEvaluator:
    """Owns one ``bcm3b200`` handle for a PopPKProblem (optionally a contiguous shard of its patients)."""

    def __init__(self, problem: PopPKProblem, device: int = 0, device_count: int = 1, shard_rank: int = 0,
                 shard_count: int = 1, diagnostics: bool = False, block_size: int = 0, sort_patients: bool | None = None,
                 type_string: str | None = None):
        """sort_patients: None = the library's policy (rank each chain's patients by absorption rate for large batches),
        True = always, False = never. type_string: the <pk_model type=> string to hand to the library instead of the one
        derived from problem.pk_type (tests of the reference's string -> model mapping)."""
        self.lib = _lib.load()
        self.problem = problem
        tr = problem.trial
        P, T = tr.num_patients, tr.num_timepoints
        desc = (f"type={type_string or PK_TYPE_NAMES[problem.pk_type]};drug={tr.drug};num_patients={P};num_timepoints={T};"
                f"num_variables={problem.num_variables};sd_ix={problem.sd_ix};max_steps={problem.max_steps};"
                f"shard_rank={shard_rank};shard_count={shard_count};device={device}")
        for name in ("n_transit_ix", "mean_transit_time_ix", "biphasic_uptake_time_ix", "mean_absorption2_ix"):
            if getattr(problem, name) >= 0:
                desc += f";{name}={getattr(problem, name)}"
        for key, v in (("volume_of_distribution", problem.fixed_vod), ("k_periphery_fwd", problem.fixed_periphery_fwd),
                       ("k_periphery_bwd", problem.fixed_periphery_bwd)):
            if v == v:  # not NaN: fixed in likelihood.xml (cpp:64-67)
                desc += f";{key}={v!r}"
        desc = desc.encode()
        h = C.c_void_p()
        kind = b"pharmacokinetic_trajectory" if getattr(problem, "single", False) else b"pop_pk_trajectory"
        _lib.check(self.lib.bcm3b200_create(kind, desc, len(desc), device_count, C.byref(h)))
        self.handle = h
        try:
            self._set("time", tr.time)
            self._set("observed_concentration", tr.observed_concentration)
            self._set("dose", tr.dose)
            self._set("dosing_interval", tr.dosing_interval)
            self._set("dose_after_dose_change", tr.dose_after_dose_change)
            self._set("dose_change_time", tr.dose_change_time)
            self._set("intermittent", tr.intermittent)
            self._set("treatment_interruptions", tr.treatment_interruptions)
            self._set("transforms", problem.transforms)
            if diagnostics:
                self.set_option("diagnostics", 1)
            if block_size:
                self.set_option("block_size", block_size)
            if sort_patients is not None:
                self.set_option("sort_patients", int(sort_patients))
                if sort_patients:
                    self.set_option("sort_min_systems", 0)
            _lib.check(self.lib.bcm3b200_finalize(self.handle))
        except Exception:
            self.close()
            raise
        self.num_patients_local = self.get_stat("num_patients_local")
        self.patient_offset = self.get_stat("patient_offset")
        self._last_C = 0

    def _set(self, name: str, arr) -> None:
        a = np.ascontiguousarray(arr, dtype=np.float64)
        shape = (C.c_size_t * a.ndim)(*a.shape)
        _lib.check(self.lib.bcm3b200_set_data(self.handle, name.encode(), a.ctypes.data, shape, a.ndim))

    def set_option(self, name: str, value: int) -> None:
        _lib.check(self.lib.bcm3b200_set_option(self.handle, name.encode(), int(value)))

    def get_stat(self, name: str) -> int:
        v = C.c_int64()
        _lib.check(self.lib.bcm3b200_get_stat(self.handle, name.encode(), C.byref(v)))
        return int(v.value)

    def evaluate(self, values: np.ndarray, logp: np.ndarray | None = None, status: np.ndarray | None = None):
        """HOST buffers in, host buffers out: the reference-facing call (H2D + kernels + D2H)."""
        values = np.ascontiguousarray(values, dtype=np.float64)
        if values.ndim == 1:
            values = values[None, :]
        nC, nvar = values.shape
        if logp is None:
            logp = np.empty(nC, dtype=np.float64)
        if status is None:
            status = np.empty(nC, dtype=np.int32)
        _lib.check(self.lib.bcm3b200_evaluate_batch(self.handle, nC, nvar, values.ctypes.data, logp.ctypes.data, status.ctypes.data))
        self._last_C = nC
        return logp, status

    def evaluate_raw(self, values_ptr: int, nC: int, nvar: int, logp_ptr: int, status_ptr: int = 0) -> None:
        _lib.check(self.lib.bcm3b200_evaluate_batch(self.handle, nC, nvar, values_ptr, logp_ptr, status_ptr or None))
        self._last_C = nC

    def evaluate_device(self, d_values_ptr: int, nC: int, nvar: int, d_partial_ptr: int, stream: int = 0) -> None:
        """DEVICE buffers, asynchronous on `stream`; d_partial is [3][C] (see include/bcm3b200.h)."""
        _lib.check(self.lib.bcm3b200_evaluate_batch_device(self.handle, nC, nvar, d_values_ptr, d_partial_ptr, stream or None))
        self._last_C = nC

    def enqueue(self, values_ptr: int, nC: int, nvar: int, d_partial_ptr: int, stream: int = 0) -> None:
        """HOST values (ideally pinned) in, DEVICE partial [3][C] out, asynchronous on `stream`."""
        _lib.check(self.lib.bcm3b200_enqueue_batch(self.handle, nC, nvar, values_ptr, d_partial_ptr, stream or None))
        self._last_C = nC

    def comm_init(self, comm_id: bytes) -> None:
        """Collective over the ranks (one process per GPU): attach the library's NCCL communicator; afterwards evaluate()
        returns the complete result on every rank. comm_id: _lib.comm_unique_id() of rank 0, handed round by the caller."""
        _lib.check(self.lib.bcm3b200_comm_init(self.handle, comm_id, len(comm_id)))

    def exchange(self, d_partial_ptr: int, nC: int, stream: int = 0) -> None:
        """In-place all-gather + rank-order combination of a device partial block, asynchronous on `stream`."""
        _lib.check(self.lib.bcm3b200_exchange_partials(self.handle, nC, d_partial_ptr, stream or None))

    def combine_partials(self, partial: np.ndarray):
        partial = np.ascontiguousarray(partial, dtype=np.float64)
        nC = partial.shape[1]
        logp = np.empty(nC, dtype=np.float64)
        status = np.empty(nC, dtype=np.int32)
        _lib.check(self.lib.bcm3b200_combine_partials(nC, partial.ctypes.data, logp.ctypes.data, status.ctypes.data))
        return logp, status

    def diagnostics(self):
        nC, Pl, T = self._last_C, self.num_patients_local, self.problem.trial.num_timepoints
        conc = np.empty((nC, Pl, T), dtype=np.float64)
        pll = np.empty((nC, Pl), dtype=np.float64)
        cnt = np.empty((nC, Pl, NUM_COUNTERS), dtype=np.int32)
        _lib.check(self.lib.bcm3b200_get_diagnostics(self.handle, conc.ctypes.data, pll.ctypes.data, cnt.ctypes.data))
        return dict(conc=conc, patient_ll=pll, counters=cnt)

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.bcm3b200_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
