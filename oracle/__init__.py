"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the CPU checkers (oracle/oracle_api.h).

``load("ref")``  -> oracle/_ref/libbcm3ref.so   the reference's own compiled CVODE/odecommon stack
``load("port")`` -> oracle/_build/liboracle.so  the plain-C restatement (oracle/cvode_bdf.c, oracle/poppk_oracle.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; nothing under bcm3_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libbcm3ref.so")
PORT_LIB = os.path.join(HERE, "_build", "liboracle.so")
REFERENCE_ROOT = "/root/reference"

NUM_COUNTERS = 8
CNT_STEPS, CNT_NFE, CNT_NSETUPS, CNT_NJE, CNT_NETF, CNT_NCFN, CNT_NNI, CNT_OK = range(8)


class _PopPKProblem(C.Structure):
    _fields_ = [
        ("pk_type", C.c_int32),
        ("num_patients", C.c_int32),
        ("num_timepoints", C.c_int32),
        ("num_variables", C.c_int32),
        ("sd_ix", C.c_int32),
        ("max_steps", C.c_int32),
        ("fixed_vod", C.c_double),
        ("fixed_periphery_fwd", C.c_double),
        ("fixed_periphery_bwd", C.c_double),
        ("mol_weight", C.c_double),
        ("rtol", C.c_double),
        ("atol", C.c_double),
        ("time", C.c_void_p),
        ("observed_concentration", C.c_void_p),
        ("dose", C.c_void_p),
        ("dosing_interval", C.c_void_p),
        ("dose_after_dose_change", C.c_void_p),
        ("dose_change_time", C.c_void_p),
        ("intermittent", C.c_void_p),
        ("skipped_days", C.c_void_p),
        ("simulate_until", C.c_void_p),
        ("transforms", C.c_void_p),
        ("n_transit_ix", C.c_int32),
        ("mean_transit_time_ix", C.c_int32),
        ("biphasic_uptake_time_ix", C.c_int32),
        ("mean_absorption2_ix", C.c_int32),
        ("single", C.c_int32),
    ]


def build_port() -> str:
    subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)
    return PORT_LIB


def build_ref() -> str | None:
    """Compile the reference's own sources in place; only possible where /root/reference is mounted."""
    if not os.path.isdir(REFERENCE_ROOT):
        return REF_LIB if os.path.exists(REF_LIB) else None
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "ref")], check=True)
    return REF_LIB


def available(kind: str) -> bool:
    return os.path.exists(REF_LIB if kind == "ref" else PORT_LIB)


class _CellPopMarker(C.Structure):
    _fields_ = [("num_obs_species", C.c_int32), ("obs_species", C.c_int32 * 8), ("stdev_ix", C.c_int32), ("offset_ix", C.c_int32),
                ("scale_ix", C.c_int32), ("proportional_stdev_ix", C.c_int32), ("stdev", C.c_double), ("offset", C.c_double),
                ("scale", C.c_double), ("proportional_stdev", C.c_double), ("observed", C.c_void_p), ("log_ratio_denominator", C.c_int32)]


class _CellPopProblem(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("num_species", "num_constant_species", "num_variables", "num_non_sampled", "num_cells",
                                         "num_timepoints", "num_replicates", "variability_dim", "entry_time_ix", "max_steps",
                                         "error_model", "stdev_ix", "offset_ix", "scale_ix", "proportional_stdev_ix", "full_gaussian",
                                         "num_obs_species")] + [
        ("obs_species", C.c_int32 * 8)] + [(n, C.c_double) for n in ("entry_time", "rel_tol", "abs_tol", "min_dt", "weight", "stdev",
                                                                     "offset", "scale", "missing_stdev", "proportional_stdev")] + [
        (n, C.c_void_p) for n in ("covariance", "initial_conditions", "constant_species", "non_sampled", "sobol", "timepoints", "observed",
                                  "variability", "transforms", "derivative")] + [
        ("treatment_species", C.c_int32), ("treatment_num_pulses", C.c_int32), ("treatment_times", C.c_void_p),
        ("relative_to_time_average", C.c_int32), ("have_sim_end_time", C.c_int32), ("sim_end_time", C.c_double),
        ("stdev_relative_to_scale", C.c_int32), ("divide_cells", C.c_int32), ("max_cells", C.c_int32), ("sobol_rows", C.c_int32),
        ("cytokinesis_ix", C.c_int32), ("apoptosis_ix", C.c_int32), ("reset_ix", C.c_int32 * 7), ("max_dt", C.c_double), ("data_kind", C.c_int32),
        ("value_relative_to_timepoint_ix", C.c_int32), ("optimize_offset_scale", C.c_int32), ("optimize_offset_min", C.c_double),
        ("optimize_offset_max", C.c_double), ("optimize_scale_min", C.c_double), ("optimize_scale_max", C.c_double),
        ("saturation_scale_ix", C.c_int32), ("num_extra_markers", C.c_int32), ("extra_markers", C.c_void_p), ("log_ratio_denominator", C.c_int32), ("use_only_nondivided", C.c_int32), ("include_only_mitotic", C.c_int32),
        ("nuclear_envelope_ix", C.c_int32)]


_derivative_libs: dict[str, C.CDLL] = {}


RHS_FLAGS = {
    # strict IEEE (no FMA contraction): the device build of the same text uses -fmad=false, which makes the rate laws
    # bit-identical on both sides
    "strict": ["-O2", "-ffp-contract=off"],
    # what the reference's own CMake build of the generated code does (-O3 -march=native, contraction on): only used to
    # measure how far the reference moves under its own flags (tests/golden/measure_noise_floor.py)
    "contracted": ["-O3", "-march=x86-64-v3"],
}
rhs_build = "strict"


def compile_cellpop_derivative(code: str) -> C.CDLL:
    """Compile the generated RHS text for the HOST the way the reference does (SolverCodeGenerator.cpp:100-300,390,407-414):
    helper prelude + generated text -> shared library -> dlopen."""
    import hashlib

    key = hashlib.sha1((code + "|" + rhs_build).encode()).hexdigest()[:16]
    if key in _derivative_libs:
        return _derivative_libs[key]
    d = os.path.join(HERE, "_build", "cellpop_" + key)
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, "libgenerated_derivatives.so")
    if not os.path.exists(so):
        with open(os.path.join(d, "code.cpp"), "w") as f:
            f.write('#include <limits>\n#include "cellpop_prelude.h"\n#define EXPORT_PREFIX extern "C"\n'
                    "struct OdeMatrixReal { double dummy; double& operator()(int, int) { return dummy; } };\n\n")
            f.write(code)
        # default "strict": the reference's own flags (-O3 -march=native) would contract -- see DESIGN.md section 9
        subprocess.run(["g++"] + RHS_FLAGS[rhs_build] + ["-std=c++14", "-fPIC", "-shared", "-w", "-I", HERE, "-o", so + ".tmp",
                        os.path.join(d, "code.cpp")], check=True)
        os.replace(so + ".tmp", so)
    lib = C.CDLL(so)
    _derivative_libs[key] = lib
    return lib


class Oracle:
    def __init__(self, kind: str, path: str | None = None):
        path = path or (REF_LIB if kind == "ref" else PORT_LIB)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built (run `make -C oracle` / `make -C oracle/ref`)")
        self.kind = kind
        self.lib = C.CDLL(path)
        self.lib.oracle_poppk_evaluate.restype = C.c_int
        self.lib.oracle_poppk_evaluate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_void_p, C.c_int]
        self.lib.oracle_cellpop_evaluate.restype = C.c_int
        self.lib.oracle_cellpop_evaluate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        self.lib.oracle_kind.restype = C.c_char_p
        assert self.lib.oracle_kind().decode() == kind

    def poppk_evaluate(self, problem, values, threads: int = 1, want_conc=False, want_patient_ll=False, want_counters=False):
        """problem: bcm3_b200.poppk_data.PopPKProblem; values [C, nvar] -> dict(logp[C], conc, patient_ll, counters)."""
        tr = problem.trial
        P, T = tr.num_patients, tr.num_timepoints
        values = np.ascontiguousarray(values, dtype=np.float64)
        if values.ndim == 1:
            values = values[None, :]
        nC, nvar = values.shape
        assert nvar == problem.num_variables
        keep = dict(
            time=np.ascontiguousarray(tr.time, dtype=np.float64),
            obs=np.ascontiguousarray(tr.observed_concentration, dtype=np.float64),
            dose=np.ascontiguousarray(tr.dose, dtype=np.float64),
            di=np.ascontiguousarray(tr.dosing_interval, dtype=np.float64),
            dac=np.ascontiguousarray(tr.dose_after_dose_change, dtype=np.float64),
            dct=np.ascontiguousarray(tr.dose_change_time, dtype=np.float64),
            inter=np.ascontiguousarray(tr.intermittent, dtype=np.int32),
            skipped=np.ascontiguousarray(problem.skipped_days, dtype=np.uint32),
            su=np.ascontiguousarray(problem.simulate_until, dtype=np.int32),
            transforms=np.ascontiguousarray(problem.transforms, dtype=np.int32),
        )
        s = _PopPKProblem(
            pk_type=problem.pk_type, num_patients=P, num_timepoints=T, num_variables=nvar, sd_ix=problem.sd_ix,
            max_steps=problem.max_steps, fixed_vod=problem.fixed_vod, fixed_periphery_fwd=problem.fixed_periphery_fwd,
            fixed_periphery_bwd=problem.fixed_periphery_bwd, mol_weight=problem.mol_weight, rtol=problem.rtol,
            atol=problem.atol,
            time=keep["time"].ctypes.data, observed_concentration=keep["obs"].ctypes.data, dose=keep["dose"].ctypes.data,
            dosing_interval=keep["di"].ctypes.data, dose_after_dose_change=keep["dac"].ctypes.data,
            dose_change_time=keep["dct"].ctypes.data, intermittent=keep["inter"].ctypes.data,
            skipped_days=keep["skipped"].ctypes.data, simulate_until=keep["su"].ctypes.data,
            transforms=keep["transforms"].ctypes.data,
            n_transit_ix=problem.n_transit_ix, mean_transit_time_ix=problem.mean_transit_time_ix,
            biphasic_uptake_time_ix=problem.biphasic_uptake_time_ix, mean_absorption2_ix=problem.mean_absorption2_ix,
            single=1 if getattr(problem, "single", False) else 0,
        )
        logp = np.empty(nC, dtype=np.float64)
        conc = np.empty((nC, P, T), dtype=np.float64) if want_conc else None
        pll = np.empty((nC, P), dtype=np.float64) if want_patient_ll else None
        cnt = np.zeros((nC, P, NUM_COUNTERS), dtype=np.int64) if want_counters else None
        rc = self.lib.oracle_poppk_evaluate(
            C.byref(s), nC, values.ctypes.data, logp.ctypes.data,
            conc.ctypes.data if conc is not None else None,
            pll.ctypes.data if pll is not None else None,
            cnt.ctypes.data if cnt is not None else None, int(threads))
        if rc != 0:
            raise RuntimeError(f"oracle_poppk_evaluate failed: {rc}")
        return dict(logp=logp, conc=conc, patient_ll=pll, counters=cnt)


ERROR_MODELS = {"normal": 0, "additive_normal": 0, "student_t4": 1, "t4": 1, "proportional_normal": 2, "additive_proportional_normal": 3}


def _cellpop_struct(problem, values):
    """(struct, arrays it points into, values [C, nvar]) for a bcm3_b200.cellpop_data.CellPopProblem."""
    p = problem
    values = np.ascontiguousarray(values, dtype=np.float64)
    if values.ndim == 1:
        values = values[None, :]
    nC = values.shape[0]
    T, nc, D = p.num_timepoints, p.num_cells, p.variability_dim
    dlib = compile_cellpop_derivative(p.derivative_code)
    fn = C.cast(dlib.generated_derivative, C.c_void_p).value
    keep = dict(
        ic=np.ascontiguousarray(p.initial_conditions, dtype=np.float64), const=np.ascontiguousarray(p.constant_species, dtype=np.float64),
        ns=np.ascontiguousarray(p.non_sampled_parameters, dtype=np.float64), sobol=np.ascontiguousarray(p.sobol, dtype=np.float64),
        tp=np.ascontiguousarray(p.timepoints, dtype=np.float64), obs=np.ascontiguousarray(p.observed, dtype=np.float64),
        var=np.ascontiguousarray(p.variability_rows(), dtype=np.float64), tr=np.ascontiguousarray(p.transforms, dtype=np.int32),
        cov=np.ascontiguousarray(p.covariance_rows(), dtype=np.float64),
        treat=np.ascontiguousarray(np.sort(np.asarray(p.treatment_times, dtype=np.float64))))
    ptr = lambda a: a.ctypes.data if a.size else None
    extra = list(getattr(p, "extra_markers", []) or [])
    keep["marker_obs"] = [np.ascontiguousarray(m.observed, dtype=np.float64) for m in extra]
    opt = lambda v: -1 if v is None else int(v)
    keep["markers"] = (_CellPopMarker * max(1, len(extra)))(*[
        _CellPopMarker(num_obs_species=len(m.obs_species), obs_species=(C.c_int32 * 8)(*(list(m.obs_species) + [0] * (8 - len(m.obs_species)))),
                       stdev_ix=opt(m.stdev_ix), offset_ix=opt(m.offset_ix), scale_ix=opt(m.scale_ix), proportional_stdev_ix=opt(m.proportional_stdev_ix),
                       stdev=m.stdev, offset=m.offset, scale=m.scale, proportional_stdev=m.proportional_stdev, observed=o.ctypes.data,
                       log_ratio_denominator=opt(getattr(m, "log_ratio_denominator", None)))
        for m, o in zip(extra, keep["marker_obs"])])
    s = _CellPopProblem(
        num_species=p.num_species, num_constant_species=len(keep["const"]), num_variables=p.num_variables, num_non_sampled=len(keep["ns"]),
        num_cells=nc, num_timepoints=T, num_replicates=p.num_replicates, variability_dim=D,
        entry_time_ix=-1 if p.entry_time_ix is None else p.entry_time_ix, max_steps=p.solver_max_steps,
        error_model=ERROR_MODELS[p.error_model],
        proportional_stdev_ix=-1 if p.proportional_stdev_ix is None else p.proportional_stdev_ix,
        full_gaussian=int(p.variability_distribution == "full_gaussian"), proportional_stdev=p.proportional_stdev,
        covariance=ptr(keep["cov"]),
        treatment_species=-1 if p.treatment_species is None else p.treatment_species, treatment_num_pulses=len(keep["treat"]),
        treatment_times=ptr(keep["treat"]), relative_to_time_average=int(p.relative_to_time_average),
        have_sim_end_time=int(p.simulation_end_time is not None), sim_end_time=float(p.simulation_end_time or 0.0),
        stdev_relative_to_scale=int(p.stdev_relative_to_scale),
        divide_cells=int(p.divide_cells), max_cells=p.capacity, sobol_rows=int(keep["sobol"].shape[0]) if keep["sobol"].ndim == 2 else 0,
        cytokinesis_ix=-1 if p.cytokinesis_species is None else p.cytokinesis_species,
        apoptosis_ix=-1 if p.apoptosis_species is None else p.apoptosis_species,
        reset_ix=(C.c_int32 * 7)(*(list(p.division_reset_species) if p.divide_cells else [0] * 7)), max_dt=float(p.solver_max_timestep),
        data_kind={"time_course_population_average": 0, "time_course": 1, "time_points": 2}[getattr(p, "data_kind", "time_course_population_average")],
        value_relative_to_timepoint_ix=-1 if getattr(p, "value_relative_to_timepoint_ix", None) is None else int(p.value_relative_to_timepoint_ix),
        saturation_scale_ix=-1 if getattr(p, "saturation_scale_ix", None) is None else int(p.saturation_scale_ix),
        log_ratio_denominator=opt(getattr(p, "log_ratio_denominator", None)), use_only_nondivided=int(bool(getattr(p, "use_only_nondivided", False))),
        include_only_mitotic=int(bool(getattr(p, "include_only_cells_that_went_through_mitosis", False))),
        nuclear_envelope_ix=opt(getattr(p, "nuclear_envelope_species", None)),
        num_extra_markers=len(extra), extra_markers=C.cast(keep["markers"], C.c_void_p).value if extra else None,
        optimize_offset_scale=int(bool(getattr(p, "optimize_offset_scale", False))),
        optimize_offset_min=float(getattr(p, "optimize_offset_range", (-1.0, 1.0))[0]), optimize_offset_max=float(getattr(p, "optimize_offset_range", (-1.0, 1.0))[1]),
        optimize_scale_min=float(getattr(p, "optimize_scale_range", (0.1, 10.0))[0]), optimize_scale_max=float(getattr(p, "optimize_scale_range", (0.1, 10.0))[1]),
        stdev_ix=-1 if p.stdev_ix is None else p.stdev_ix, offset_ix=-1 if p.offset_ix is None else p.offset_ix,
        scale_ix=-1 if p.scale_ix is None else p.scale_ix, num_obs_species=len(p.obs_species),
        obs_species=(C.c_int32 * 8)(*(list(p.obs_species) + [0] * (8 - len(p.obs_species)))),
        entry_time=p.entry_time, rel_tol=p.solver_relative_tolerance, abs_tol=p.solver_absolute_tolerance, min_dt=p.solver_min_timestep,
        weight=p.weight, stdev=p.stdev, offset=p.offset, scale=p.scale, missing_stdev=p.missing_simulation_time_stdev,
        initial_conditions=ptr(keep["ic"]), constant_species=ptr(keep["const"]), non_sampled=ptr(keep["ns"]), sobol=ptr(keep["sobol"]),
        timepoints=ptr(keep["tp"]), observed=ptr(keep["obs"]), variability=ptr(keep["var"]), transforms=ptr(keep["tr"]), derivative=fn)
    return s, keep, values


def _cellpop_evaluate(self, problem, values, threads: int = 1, want_cell_values=False, want_steps=False, want_average=False):
    """problem: bcm3_b200.cellpop_data.CellPopProblem; values [C, nvar]."""
    s, keep, values = _cellpop_struct(problem, values)
    nC = values.shape[0]
    T, nc = problem.num_timepoints, problem.capacity
    logp = np.empty(nC)
    cv = np.full((nC, T, nc), np.nan) if want_cell_values else None
    st = np.zeros((nC, nc), dtype=np.int32) if want_steps else None
    avg = np.empty((nC, T)) if want_average else None
    rc = self.lib.oracle_cellpop_evaluate(C.byref(s), nC, values.ctypes.data, logp.ctypes.data,
                                          cv.ctypes.data if cv is not None else None, st.ctypes.data if st is not None else None,
                                          avg.ctypes.data if avg is not None else None, int(threads))
    if rc != 0:
        raise RuntimeError(f"oracle_cellpop_evaluate failed: {rc}")
    return dict(logp=logp, cell_values=cv, cell_steps=st, population_average=avg)




def _cellpop_counters(self, problem, values, threads: int = 1):
    """The solver's counters per cell, [C, cells, 8] = (steps, nfe, nsetups, nje, netf, ncfn, nni, ok) summed over the restarts of
    each solve: what bench.py computes the algorithmic FLOPs of the cell_population workloads from."""
    s, keep, values = _cellpop_struct(problem, values)
    nC = values.shape[0]
    fn = self.lib.oracle_cellpop_evaluate_counters
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    logp = np.empty(nC)
    cnt = np.zeros((nC, problem.capacity, NUM_COUNTERS), dtype=np.int64)
    rc = fn(C.byref(s), nC, values.ctypes.data, logp.ctypes.data, cnt.ctypes.data, int(threads))
    if rc != 0:
        raise RuntimeError(f"oracle_cellpop_evaluate_counters failed: {rc}")
    return dict(logp=logp, counters=cnt)


class _PharmacoProblem(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("num_patients", "num_timepoints", "num_variables", "use_peripheral", "num_transit", "use_bioavailability",
                                         "additive_sd_ix", "proportional_sd_ix", "mean_absorption_ix", "mean_excretion_ix", "mean_clearance_ix",
                                         "mean_vod_ix", "sigma_absorption_ix", "sigma_excretion_ix", "sigma_clearance_ix", "sigma_vod_ix",
                                         "sigma_transit_ix", "periph_fwd_ix", "periph_bwd_ix", "mean_transit_time_ix")] + [
        ("mol_weight", C.c_double)] + [(n, C.c_void_p) for n in (
            "transforms", "p_absorption_ix", "p_excretion_ix", "p_clearance_ix", "p_vod_ix", "p_transit_ix", "p_bioavailability_ix", "time",
            "observed_concentration", "dose", "dosing_interval", "dose_after_dose_change", "dose_change_time", "intermittent", "skipped_days")] + [
        (n, C.c_int32) for n in ("single", "use_biphasic", "use_metabolite", "direct_absorption_ix", "metabolite_conversion_ix")]


def _pharmaco_evaluate(self, problem, values, threads: int = 1, want_conc=False, want_patient_ll=False):
    """problem: bcm3_b200.pharmaco.PharmacoProblem; values [C, nvar]."""
    p, tr = problem, problem.trial
    values = np.ascontiguousarray(values, dtype=np.float64)
    if values.ndim == 1:
        values = values[None, :]
    nC = values.shape[0]
    P, T = tr.num_patients, tr.num_timepoints
    skipped = np.zeros(P, dtype=np.uint32)
    for d in range(29):
        skipped |= (np.asarray(tr.treatment_interruptions)[:, d] != 0).astype(np.uint32) << np.uint32(d)
    keep = dict(tr=np.ascontiguousarray(p.transforms, dtype=np.int32), time=np.ascontiguousarray(tr.time, dtype=np.float64),
                obs=np.ascontiguousarray(tr.observed_concentration, dtype=np.float64), dose=np.ascontiguousarray(tr.dose, dtype=np.float64),
                interval=np.ascontiguousarray(tr.dosing_interval, dtype=np.float64), dac=np.ascontiguousarray(tr.dose_after_dose_change, dtype=np.float64),
                dct=np.ascontiguousarray(tr.dose_change_time, dtype=np.float64), inter=np.ascontiguousarray(tr.intermittent, dtype=np.int32), skipped=skipped)
    pix = {k: np.ascontiguousarray(v, dtype=np.int32) for k, v in p.patient_indices().items()}
    ptr = lambda a: a.ctypes.data
    r = p.role_indices()
    s = _PharmacoProblem(
        num_patients=P, num_timepoints=T, num_variables=p.num_variables, use_peripheral=int(p.peripheral_compartment),
        num_transit=int(p.num_transit_compartments), use_bioavailability=int(p.bioavailability), additive_sd_ix=r["additive_sd"],
        proportional_sd_ix=r["proportional_sd"], mean_absorption_ix=r["mean_absorption"], mean_excretion_ix=r["mean_excretion"],
        mean_clearance_ix=r["mean_clearance"], mean_vod_ix=r["mean_volume_of_distribution"],
        sigma_absorption_ix=r["sigma_absorption"] if "patient_absorption_ix" in pix else -1,
        sigma_excretion_ix=r["sigma_excretion"] if "patient_excretion_ix" in pix else -1,
        sigma_clearance_ix=r["sigma_clearance"] if "patient_clearance_ix" in pix else -1,
        sigma_vod_ix=r["sigma_volume_of_distribution"] if "patient_volume_of_distribution_ix" in pix else -1,
        sigma_transit_ix=r["sigma_transit_time"] if "patient_transit_time_ix" in pix else -1,
        periph_fwd_ix=r["peripheral_forward_rate"], periph_bwd_ix=r["peripheral_backward_rate"], mean_transit_time_ix=r["mean_transit_time"],
        mol_weight=p.mol_weight, transforms=ptr(keep["tr"]),
        p_absorption_ix=ptr(pix["patient_absorption_ix"]) if "patient_absorption_ix" in pix else None,
        p_excretion_ix=ptr(pix["patient_excretion_ix"]) if "patient_excretion_ix" in pix else None,
        p_clearance_ix=ptr(pix["patient_clearance_ix"]) if "patient_clearance_ix" in pix else None,
        p_vod_ix=ptr(pix["patient_volume_of_distribution_ix"]) if "patient_volume_of_distribution_ix" in pix else None,
        p_transit_ix=ptr(pix["patient_transit_time_ix"]) if "patient_transit_time_ix" in pix else None,
        p_bioavailability_ix=ptr(pix["patient_bioavailability_ix"]) if "patient_bioavailability_ix" in pix else None,
        time=ptr(keep["time"]), observed_concentration=ptr(keep["obs"]), dose=ptr(keep["dose"]), dosing_interval=ptr(keep["interval"]),
        dose_after_dose_change=ptr(keep["dac"]), dose_change_time=ptr(keep["dct"]), intermittent=ptr(keep["inter"]), skipped_days=ptr(keep["skipped"]),
        single=int(getattr(p, "single", False)), use_biphasic=int(getattr(p, "biphasic_absorption", False)), use_metabolite=int(getattr(p, "metabolite", False)),
        direct_absorption_ix=r.get("direct_absorption", -1), metabolite_conversion_ix=r.get("metabolite_conversion_rate", -1))
    logp = np.empty(nC)
    conc = np.empty((nC, P, T)) if want_conc else None
    pll = np.empty((nC, P)) if want_patient_ll else None
    fn = self.lib.oracle_pharmaco_evaluate
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    rc = fn(C.byref(s), nC, values.ctypes.data, logp.ctypes.data, conc.ctypes.data if conc is not None else None,
            pll.ctypes.data if pll is not None else None, int(threads))
    if rc != 0:
        raise RuntimeError(f"oracle_pharmaco_evaluate failed: {rc}")
    return dict(logp=logp, conc=conc, patient_ll=pll)


Oracle.pharmaco_evaluate = _pharmaco_evaluate
def _hungarian_match(self, cost):
    """hungarianMinimumWeightPerfectMatching on a complete [n][n] cost matrix (oracle_api.h: oracle_hungarian_match); None without a matching."""
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    n = cost.shape[0]
    out = np.full(n, -1, dtype=np.int32)
    self.lib.oracle_hungarian_match(C.c_int(n), C.c_void_p(cost.ctypes.data), C.c_void_p(out.ctypes.data))
    return None if (n and out.min() < 0) else out


Oracle.cellpop_evaluate = _cellpop_evaluate
Oracle.hungarian_match = _hungarian_match
Oracle.cellpop_counters = _cellpop_counters

_cache: dict[str, Oracle] = {}


def load(kind: str = "ref") -> Oracle:
    if kind not in _cache:
        _cache[kind] = Oracle(kind)
    return _cache[kind]
