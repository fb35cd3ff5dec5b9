"""GPU-backed pharmaco_population likelihood: host-side data model and wrapper of the C ABI (include/bcm3b200.h).

Mirrors ``PharmacoLikelihoodPopulation`` (src/pharmaco/PharmacoLikelihoodPopulation.{h,cpp}) behind ``bcm3::Likelihood``: a
population PK model whose compartments are advanced with the matrix exponential (PharmacokineticModel.cpp:111-247) instead of
an ODE solver. The trial is the same NetCDF group as for pop_pk_trajectory (Patient::Load, PharmacoPatient.cpp:8-116) and is
handed in as a ``PopPKTrial``; variables are found by NAME in the prior (PostInitialize, cpp:102-188), so a problem carries the
list of variable names.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .poppk_data import MOLECULAR_WEIGHT, TRANSFORM_LOG10, TRANSFORM_NONE, PopPKTrial

# role in the C ABI's descriptor -> variable name in prior.xml (cpp:104-186)
ROLES = {
    "additive_sd": "additive_error_standard_deviation", "proportional_sd": "proportional_error_standard_deviation",
    "mean_absorption": "mean_absorption", "mean_excretion": "mean_excretion", "mean_clearance": "mean_clearance",
    "mean_volume_of_distribution": "mean_volume_of_distribution", "sigma_absorption": "sigma_absorption", "sigma_excretion": "sigma_excretion",
    "sigma_clearance": "sigma_clearance", "sigma_volume_of_distribution": "sigma_volume_of_distribution", "sigma_transit_time": "sigma_transit_time",
    "peripheral_forward_rate": "peripheral_forward_rate", "peripheral_backward_rate": "peripheral_backward_rate", "mean_transit_time": "mean_transit_time",
}
# model kind pharmaco_single (PharmacoLikelihoodSingle.cpp:75-146): key of the C ABI's descriptor -> (variable name in prior.xml, internal role)
SINGLE_ROLES = {
    "additive_sd": ("additive_error_standard_deviation", "additive_sd"), "proportional_sd": ("proportional_error_standard_deviation", "proportional_sd"),
    "absorption": ("absorption", "mean_absorption"), "excretion": ("excretion", "mean_excretion"), "clearance": ("clearance", "mean_clearance"),
    "volume_of_distribution": ("volume_of_distribution", "mean_volume_of_distribution"),
    "peripheral_forward_rate": ("peripheral_forward_rate", "peripheral_forward_rate"), "peripheral_backward_rate": ("peripheral_backward_rate", "peripheral_backward_rate"),
    "mean_transit_time": ("mean_transit_time", "mean_transit_time"), "direct_absorption": ("direct_absorption", "direct_absorption"),
    "metabolite_conversion_rate": ("metabolite_conversion_rate", "metabolite_conversion_rate"),
}
# per-patient marginals p<i>_<name> (InitializePatientMarginals, cpp:342-354): ABI array name -> (variable suffix, the sigma that switches it on)
PATIENT_ARRAYS = {
    "patient_absorption_ix": ("absorption", "sigma_absorption"), "patient_excretion_ix": ("excretion", "sigma_excretion"),
    "patient_clearance_ix": ("clearance", "sigma_clearance"), "patient_volume_of_distribution_ix": ("volume_of_distribution", "sigma_volume_of_distribution"),
    "patient_transit_time_ix": ("transit_time", "sigma_transit_time"), "patient_bioavailability_ix": ("bioavailability", None),
}


@dataclass
class PharmacoProblem:
    trial: PopPKTrial
    variable_names: list            # prior.xml order
    transforms: np.ndarray          # [nvar] VariableSet transform codes
    peripheral_compartment: bool = False
    num_transit_compartments: int = 0
    bioavailability: bool = False
    # likelihood.xml type="pharmaco_single" (src/pharmaco/PharmacoLikelihoodSingle.cpp): ONE patient, the variables are its rates;
    # <pk_model biphasic_absorption= metabolite=> exist for this likelihood only
    single: bool = False
    biphasic_absorption: bool = False
    metabolite: bool = False

    @property
    def num_variables(self) -> int:
        return len(self.variable_names)

    @property
    def mol_weight(self) -> float:
        return MOLECULAR_WEIGHT[self.trial.drug]

    def index(self, name: str) -> int:
        return self.variable_names.index(name) if name in self.variable_names else -1

    def role_indices(self) -> dict:
        if self.single:  # internal roles (the checker's struct); sigma roles do not exist
            out = {role: -1 for role in ROLES}
            out.update(direct_absorption=-1, metabolite_conversion_rate=-1)
            for name, role in SINGLE_ROLES.values():
                out[role] = self.index(name)
            return out
        return {role: self.index(name) for role, name in ROLES.items()}

    def descriptor_indices(self) -> dict:
        """key -> variable index as the C ABI's descriptor names them."""
        if self.single:
            return {key: self.index(name) for key, (name, _) in SINGLE_ROLES.items()}
        return self.role_indices()

    def patient_indices(self) -> dict:
        """ABI array name -> [P] variable indices, for the marginals the prior switches on."""
        P = self.trial.num_patients
        out = {}
        if self.single:
            return out
        for array, (suffix, sigma) in PATIENT_ARRAYS.items():
            on = self.bioavailability if sigma is None else (self.index(sigma) >= 0)
            if suffix == "excretion":
                on = on and self.index("mean_excretion") >= 0
            if suffix == "transit_time":
                on = on and self.num_transit_compartments > 0
            if on:
                out[array] = np.array([self.variable_names.index(f"p{i + 1}_{suffix}") for i in range(P)], dtype=np.int32)
        return out


class PharmacoEvaluator:
    """Owns one ``bcm3b200`` handle of kind pharmaco_population (optionally a contiguous shard of the patients)."""

    def __init__(self, problem: PharmacoProblem, device: int = 0, shard_rank: int = 0, shard_count: int = 1, diagnostics: bool = False):
        self.lib = _lib.load()
        self.problem = p = problem
        tr = p.trial
        P, T = tr.num_patients, tr.num_timepoints
        desc = (f"drug={tr.drug};num_patients={P};num_timepoints={T};num_variables={p.num_variables};peripheral_compartment={int(p.peripheral_compartment)};"
                f"num_transit_compartments={int(p.num_transit_compartments)};bioavailability={int(p.bioavailability)};"
                f"shard_rank={shard_rank};shard_count={shard_count};device={device}")
        for role, ix in p.descriptor_indices().items():
            if ix >= 0:
                desc += f";{role}_ix={ix}"
        if p.single:
            desc += f";biphasic_absorption={int(p.biphasic_absorption)};metabolite={int(p.metabolite)}"
        desc = desc.encode()
        h = C.c_void_p()
        _lib.check(self.lib.bcm3b200_create(b"pharmaco_single" if p.single else b"pharmaco_population", desc, len(desc), 1, C.byref(h)))
        self.handle = h
        try:
            for name in ("time", "observed_concentration", "dose", "dosing_interval", "dose_after_dose_change", "dose_change_time", "intermittent",
                         "treatment_interruptions"):
                self._set(name, getattr(tr, name))
            self._set("transforms", p.transforms)
            for array, ixs in p.patient_indices().items():
                self._set(array, ixs)
            if diagnostics:
                _lib.check(self.lib.bcm3b200_set_option(self.handle, b"diagnostics", 1))
            _lib.check(self.lib.bcm3b200_finalize(self.handle))
        except Exception:
            self.close()
            raise
        self.num_patients_local = self.get_stat("num_patients_local")
        self._last_C = 0

    def _set(self, name: str, arr) -> None:
        a = np.ascontiguousarray(arr, dtype=np.float64)
        shape = (C.c_size_t * a.ndim)(*a.shape)
        _lib.check(self.lib.bcm3b200_set_data(self.handle, name.encode(), a.ctypes.data, shape, a.ndim))

    def get_stat(self, name: str) -> int:
        v = C.c_int64()
        _lib.check(self.lib.bcm3b200_get_stat(self.handle, name.encode(), C.byref(v)))
        return int(v.value)

    def comm_init(self, comm_id: bytes) -> None:
        _lib.check(self.lib.bcm3b200_comm_init(self.handle, comm_id, len(comm_id)))

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

    def diagnostics(self):
        nC, Pl, T = self._last_C, self.num_patients_local, self.problem.trial.num_timepoints
        conc = np.empty((nC, Pl, T))
        pll = np.empty((nC, Pl))
        _lib.check(self.lib.bcm3b200_get_diagnostics(self.handle, conc.ctypes.data, pll.ctypes.data, None))
        return dict(conc=conc, patient_ll=pll)

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.bcm3b200_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_pharmaco_problem(P: int = 200, T: int = 10, peripheral: bool = False, num_transit: int = 0, bioavailability: bool = False,
                          heterogeneous: bool = True, missing_fraction: float = 0.1, seed: int = 1, excretion: bool = True) -> PharmacoProblem:
    """Synthetic trial (the generator of the PopPK workloads, bcm3_b200.synthetic) + a prior with population means, between-patient
    standard deviations for absorption and clearance and the per-patient marginals they need."""
    from . import synthetic as syn
    from .poppk_data import PK_ONE, PK_TWO

    base = syn.make_poppk_problem(PK_TWO if peripheral else PK_ONE, P=P, T=T, t_end=96.0, seed=seed, heterogeneous=heterogeneous,
                                  missing_fraction=missing_fraction)
    names = ["mean_absorption", "mean_clearance", "mean_volume_of_distribution", "sigma_absorption", "sigma_clearance",
             "additive_error_standard_deviation", "proportional_error_standard_deviation"]
    if excretion:
        names.append("mean_excretion")
    if peripheral:
        names += ["peripheral_forward_rate", "peripheral_backward_rate"]
    if num_transit > 0:
        names += ["mean_transit_time", "sigma_transit_time"]
    names += [f"p{i + 1}_absorption" for i in range(P)] + [f"p{i + 1}_clearance" for i in range(P)]
    if num_transit > 0:
        names += [f"p{i + 1}_transit_time" for i in range(P)]
    if bioavailability:
        names += [f"p{i + 1}_bioavailability" for i in range(P)]
    transforms = np.full(len(names), TRANSFORM_NONE, dtype=np.int32)
    for n in ("additive_error_standard_deviation", "proportional_error_standard_deviation", "peripheral_forward_rate", "peripheral_backward_rate"):
        if n in names:
            transforms[names.index(n)] = TRANSFORM_LOG10
    if "mean_transit_time" in names:
        transforms[names.index("mean_transit_time")] = TRANSFORM_LOG10
    return PharmacoProblem(trial=base.trial, variable_names=names, transforms=transforms, peripheral_compartment=peripheral,
                           num_transit_compartments=num_transit, bioavailability=bioavailability)


def make_pharmaco_values(problem: PharmacoProblem, C: int, seed: int = 20261018) -> np.ndarray:
    p = problem
    P = p.trial.num_patients
    out = np.empty((C, p.num_variables))
    for c in range(C):
        rng = np.random.default_rng(seed + c)
        v = dict(mean_absorption=rng.normal(-0.3, 0.1), mean_clearance=rng.normal(0.7, 0.1), mean_volume_of_distribution=rng.normal(1.8, 0.05),
                 sigma_absorption=0.2, sigma_clearance=0.15, additive_error_standard_deviation=rng.normal(0.3, 0.05),
                 proportional_error_standard_deviation=rng.normal(-0.7, 0.05), mean_excretion=rng.normal(-1.5, 0.1),
                 peripheral_forward_rate=rng.normal(-0.8, 0.1), peripheral_backward_rate=rng.normal(-1.0, 0.1),
                 mean_transit_time=rng.normal(0.3, 0.05), sigma_transit_time=0.1)
        for i, n in enumerate(p.variable_names):
            if n in v:
                out[c, i] = v[n]
            elif n.endswith("_bioavailability"):
                out[c, i] = rng.uniform(0.5, 1.0)
            else:
                out[c, i] = rng.uniform(0.02, 0.98)  # a patient's quantile
    return out


def make_pharmaco_single_problem(T: int = 12, peripheral: bool = False, num_transit: int = 0, biphasic_absorption: bool = False, metabolite: bool = False,
                                 excretion: bool = True, missing_fraction: float = 0.1, seed: int = 1) -> PharmacoProblem:
    """One patient of a synthetic trial with the prior of a pharmaco_single model directory: every variable a rate in log10 space."""
    from . import synthetic as syn
    from .poppk_data import PK_ONE, PK_TWO

    base = syn.make_poppk_problem(PK_TWO if peripheral else PK_ONE, P=1, T=T, t_end=120.0, seed=seed, heterogeneous=True, missing_fraction=missing_fraction)
    names = ["absorption", "clearance", "volume_of_distribution", "additive_error_standard_deviation", "proportional_error_standard_deviation"]
    if excretion:
        names.append("excretion")
    if peripheral:
        names += ["peripheral_forward_rate", "peripheral_backward_rate"]
    if num_transit > 0:
        names.append("mean_transit_time")
    if biphasic_absorption:
        names.append("direct_absorption")
    if metabolite:
        names.append("metabolite_conversion_rate")
    return PharmacoProblem(trial=base.trial, variable_names=names, transforms=np.full(len(names), TRANSFORM_LOG10, dtype=np.int32), peripheral_compartment=peripheral,
                           num_transit_compartments=num_transit, single=True, biphasic_absorption=biphasic_absorption, metabolite=metabolite)


def make_pharmaco_single_values(problem: PharmacoProblem, C: int, seed: int = 20261018) -> np.ndarray:
    out = np.empty((C, problem.num_variables))
    for c in range(C):
        rng = np.random.default_rng(seed + c)
        v = dict(absorption=rng.normal(-0.3, 0.3), clearance=rng.normal(0.7, 0.3), volume_of_distribution=rng.normal(1.8, 0.1),
                 additive_error_standard_deviation=rng.normal(0.3, 0.05), proportional_error_standard_deviation=rng.normal(-0.7, 0.05),
                 excretion=rng.normal(-1.5, 0.2), peripheral_forward_rate=rng.normal(-0.8, 0.2), peripheral_backward_rate=rng.normal(-1.0, 0.2),
                 mean_transit_time=rng.normal(0.3, 0.1), direct_absorption=rng.normal(-1.0, 0.2), metabolite_conversion_rate=rng.normal(-1.2, 0.2))
        for i, n in enumerate(problem.variable_names):
            out[c, i] = v[n]
    return out
