"""GPU-backed cell_population likelihood: host-side wrapper of the C ABI (include/bcm3b200.h, model kind
"cell_population"). Mirrors CellPopulationLikelihood::EvaluateLogProbability (src/cellpop/CellPopulationLikelihood.cpp:82-101)
for all chains at once."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .cellpop_data import CellPopProblem


class CellPopEvaluator:
    def __init__(self, problem: CellPopProblem, device: int = 0, compile_only: bool = False, kernel: str = "auto",
                 shard_rank: int = 0, shard_count: int = 1, rhs_lanes: bool | None = None, device_count: int = 1):
        """rhs_lanes: None = the library's default (the lane-parallel right-hand side for models with 16 or 32 lanes per cell),
        False = evaluate the generated text as it stands (every lane of a cell's group runs every rate law), "always" = regroup
        the text whenever it can be parsed, whatever the number of lanes per cell."""
        self.lib = _lib.load()
        self.problem = p = problem
        kv = dict(
            num_species=p.num_species, num_constant_species=len(p.constant_species), num_variables=p.num_variables,
            num_non_sampled=len(p.non_sampled_parameters), num_cells=p.num_cells, num_timepoints=p.num_timepoints,
            num_replicates=p.num_replicates, variability_dim=p.variability_dim,
            solver_relative_tolerance=repr(p.solver_relative_tolerance), solver_absolute_tolerance=repr(p.solver_absolute_tolerance),
            solver_min_timestep=repr(p.solver_min_timestep), solver_max_steps=p.solver_max_steps, error_model=p.error_model,
            weight=repr(p.weight), missing_simulation_time_stdev=repr(p.missing_simulation_time_stdev),
            obs_species="+".join(str(s) for s in p.obs_species), device=device, compile_only=int(compile_only),
            shard_rank=shard_rank, shard_count=shard_count)
        kv["variability_distribution"] = p.variability_distribution
        if p.data_kind != "time_course_population_average":
            kv["data_kind"] = p.data_kind
        if p.optimize_offset_scale:
            kv.update(optimize_offset_scale=1, optimize_offset_min=repr(float(p.optimize_offset_range[0])), optimize_offset_max=repr(float(p.optimize_offset_range[1])),
                      optimize_scale_min=repr(float(p.optimize_scale_range[0])), optimize_scale_max=repr(float(p.optimize_scale_range[1])))
        if p.nuclear_envelope_species is not None:
            kv["nuclear_envelope_species"] = int(p.nuclear_envelope_species)
        if p.include_only_cells_that_went_through_mitosis:
            kv["include_only_cells_that_went_through_mitosis"] = 1
        if p.use_only_nondivided:
            kv["use_only_nondivided"] = 1
        if p.saturation_scale_ix is not None:
            kv["saturation_scale_ix"] = int(p.saturation_scale_ix)
        if p.value_relative_to_timepoint_ix is not None:
            kv["value_relative_to_timepoint_ix"] = int(p.value_relative_to_timepoint_ix)
        kv["relative_to_time_average"] = int(p.relative_to_time_average)
        kv["stdev_relative_to_scale"] = int(p.stdev_relative_to_scale)
        if p.treatment_species is not None:
            kv["treatment_species"] = p.treatment_species
        if p.simulation_end_time is not None:
            kv["simulation_end_time"] = repr(float(p.simulation_end_time))
        if np.isfinite(p.solver_max_timestep):
            kv["solver_max_timestep"] = repr(float(p.solver_max_timestep))
        if p.divide_cells:
            kv["divide_cells"] = 1
            kv["max_cells"] = int(p.max_cells)
            if p.cytokinesis_species is not None:
                kv["cytokinesis_species"] = int(p.cytokinesis_species)
            kv["division_reset_species"] = "+".join(str(int(i)) for i in p.division_reset_species)
        if p.apoptosis_species is not None:
            kv["apoptosis_species"] = int(p.apoptosis_species)
        for name in ("entry_time", "stdev", "offset", "scale", "proportional_stdev"):
            ix = getattr(p, name + "_ix")
            if ix is not None:
                kv[name + "_ix"] = ix
            else:
                kv[name] = repr(float(getattr(p, name)))
        # further markers of a per-cell data set: data sets @1, @2, ... of the handle that name data set 0 as theirs (marker_of);
        # after them the denominators of log ratios (use_log_ratio): value rows only, flagged denominator_of = the entry they divide
        self._denominators = []
        if p.log_ratio_denominator is not None:
            self._denominators.append((0, int(p.log_ratio_denominator)))
        for k, mk in enumerate(p.extra_markers, start=1):
            if mk.log_ratio_denominator is not None:
                self._denominators.append((k, int(mk.log_ratio_denominator)))
        if p.extra_markers or self._denominators:
            kv["num_data_sets"] = 1 + len(p.extra_markers) + len(self._denominators)
            for q, (owner, species) in enumerate(self._denominators, start=1 + len(p.extra_markers)):
                kv[f"denominator_of@{q}"] = owner
                kv[f"num_timepoints@{q}"] = p.num_timepoints
                kv[f"num_replicates@{q}"] = 1
                kv[f"obs_species@{q}"] = species
            for k, mk in enumerate(p.extra_markers, start=1):
                kv[f"marker_of@{k}"] = 0
                kv[f"num_timepoints@{k}"] = p.num_timepoints
                kv[f"num_replicates@{k}"] = int(np.asarray(mk.observed).shape[0])
                kv[f"obs_species@{k}"] = "+".join(str(s) for s in mk.obs_species)
                for name in ("stdev", "offset", "scale", "proportional_stdev"):
                    ix = getattr(mk, name + "_ix")
                    if ix is not None:
                        kv[f"{name}_ix@{k}"] = ix
                    else:
                        kv[f"{name}@{k}"] = repr(float(getattr(mk, name)))
        desc = ";".join(f"{k}={v}" for k, v in kv.items()).encode()
        h = C.c_void_p()
        _lib.check(self.lib.bcm3b200_create(b"cell_population", desc, len(desc), device_count, C.byref(h)))
        self.handle = h
        try:
            self._set("initial_conditions", p.initial_conditions)
            self._set("constant_species", p.constant_species)
            self._set("non_sampled_parameters", p.non_sampled_parameters)
            self._set("timepoints", p.timepoints)
            self._set("observed", p.observed)
            for k, mk in enumerate(p.extra_markers, start=1):
                self._set(f"timepoints@{k}", p.timepoints)
                self._set(f"observed@{k}", np.asarray(mk.observed, dtype=np.float64))
            for q in range(1 + len(p.extra_markers), 1 + len(p.extra_markers) + len(self._denominators)):
                self._set(f"timepoints@{q}", p.timepoints)
                self._set(f"observed@{q}", np.zeros((1, p.num_timepoints)))  # a denominator has no observations of its own
            self._set("transforms", p.transforms)
            if p.treatment_species is not None and len(p.treatment_times):
                self._set("treatment_times", np.asarray(p.treatment_times, dtype=np.float64))
            if p.variability_dim:
                self._set("sobol", np.asarray(p.sobol, dtype=np.float64).reshape(-1, p.variability_dim))
                self._set("variability", p.variability_rows())
                if p.variability_distribution == "full_gaussian" and p.variability_dim > 1:
                    self._set("variability_covariance", p.covariance_rows())
            code = p.derivative_code.encode()
            _lib.check(self.lib.bcm3b200_set_text(self.handle, b"derivative_code", code, len(code)))
            _lib.check(self.lib.bcm3b200_set_option(self.handle, b"cellpop_kernel", {"auto": 0, "warp": 1, "thread": 2, "group": 3}[kernel]))
            if rhs_lanes is not None:
                _lib.check(self.lib.bcm3b200_set_option(self.handle, b"cellpop_rhs_lanes", 2 if rhs_lanes == "always" else int(bool(rhs_lanes))))
            _lib.check(self.lib.bcm3b200_finalize(self.handle))
        except Exception:
            self.close()
            raise
        self._last_C = 0

    def _set(self, name: str, arr) -> None:
        a = np.ascontiguousarray(arr, dtype=np.float64)
        if a.size == 0 and a.ndim == 1:
            shape = (C.c_size_t * 1)(0)
            buf = np.zeros(1)
            _lib.check(self.lib.bcm3b200_set_data(self.handle, name.encode(), buf.ctypes.data, shape, 1))
            return
        shape = (C.c_size_t * a.ndim)(*a.shape)
        _lib.check(self.lib.bcm3b200_set_data(self.handle, name.encode(), a.ctypes.data, shape, a.ndim))

    def evaluate(self, values: np.ndarray):
        values = np.ascontiguousarray(values, dtype=np.float64)
        if values.ndim == 1:
            values = values[None, :]
        nC, nvar = values.shape
        logp = np.empty(nC)
        status = np.empty(nC, dtype=np.int32)
        _lib.check(self.lib.bcm3b200_evaluate_batch(self.handle, nC, nvar, values.ctypes.data, logp.ctypes.data, status.ctypes.data))
        self._last_C = nC
        return logp, status

    def enqueue(self, values_ptr: int, nC: int, nvar: int, d_partial_ptr: int, stream: int) -> None:
        """Sharded handles: HOST values in, this shard's DEVICE partial [nC][2 T + 1] out, enqueued on `stream`."""
        _lib.check(self.lib.bcm3b200_enqueue_batch(self.handle, nC, nvar, values_ptr, d_partial_ptr, stream))
        self._last_C = nC

    def finish(self, d_partial_ptr: int, nC: int, stream: int):
        """Combined (summed over shards) partial -> logp, status. Synchronises the stream."""
        logp = np.empty(nC)
        status = np.empty(nC, dtype=np.int32)
        _lib.check(self.lib.bcm3b200_cellpop_finish(self.handle, nC, d_partial_ptr, logp.ctypes.data, status.ctypes.data, stream))
        return logp, status

    def comm_init(self, comm_id: bytes) -> None:
        """Collective over the ranks (one process per GPU): attach the library's NCCL communicator; afterwards evaluate()
        returns the complete result on every rank."""
        _lib.check(self.lib.bcm3b200_comm_init(self.handle, comm_id, len(comm_id)))

    def exchange(self, d_partial_ptr: int, nC: int, stream: int = 0) -> None:
        _lib.check(self.lib.bcm3b200_exchange_partials(self.handle, nC, d_partial_ptr, stream or None))

    def get_stat(self, name: str) -> int:
        v = C.c_int64()
        _lib.check(self.lib.bcm3b200_get_stat(self.handle, name.encode(), C.byref(v)))
        return int(v.value)

    def diagnostics(self):
        # the library writes "value_rows" rows per chain: every data set's (and every further marker's) timepoints one after the other;
        # the first num_timepoints rows are this problem's first marker
        nC, nc, T, rows = self._last_C, self.get_stat("cell_columns"), self.problem.num_timepoints, self.get_stat("value_rows")
        vals = np.empty((nC, rows, nc))
        status = np.empty((nC, nc), dtype=np.int32)
        steps = np.empty((nC, nc), dtype=np.int32)
        avg = np.empty((nC, rows))
        _lib.check(self.lib.bcm3b200_get_cell_diagnostics(self.handle, vals.ctypes.data, status.ctypes.data, steps.ctypes.data, avg.ctypes.data))
        out = dict(cell_values=np.ascontiguousarray(vals[:, :T]), cell_status=status, cell_steps=steps, population_average=np.ascontiguousarray(avg[:, :T]))
        if rows > T:  # further markers, then the denominators of log ratios
            out["marker_values"] = [np.ascontiguousarray(vals[:, T * l:T * (l + 1)]) for l in range(1, rows // T)]
        return out

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.bcm3b200_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
