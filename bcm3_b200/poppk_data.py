"""Host-side data model of the PopPK likelihood.

Mirrors what ``LikelihoodPopPKTrajectory::Initialize`` leaves in the object
after parsing ``likelihood.xml`` and the NetCDF trial group
(reference: src/likelihoods/LikelihoodPopPKTrajectory.cpp:50-252).  The NetCDF
reader itself is out of scope; a trial is handed in
as numpy arrays with the same names and shapes as the NetCDF variables.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# pk_model type= strings, LikelihoodPopPKTrajectory.cpp:69-83
PK_ONE = 0
PK_TWO = 1
PK_ONE_BIPHASIC = 2
PK_TWO_BIPHASIC = 3
PK_ONE_TRANSIT = 4
PK_TWO_TRANSIT = 5
# likelihood.xml strings -> model. As in the reference BOTH biphasic strings select the two-compartment biphasic model
# (cpp:73-76); its one-compartment biphasic right-hand side (cpp:496-530) cannot be reached from XML.
PK_TYPES = {"one": PK_ONE, "two": PK_TWO, "one_biphasic_uptake": PK_TWO_BIPHASIC, "two_biphasic_uptake": PK_TWO_BIPHASIC,
            "one_transit": PK_ONE_TRANSIT, "two_transit": PK_TWO_TRANSIT}
# model -> the C ABI's type= key (the one-compartment biphasic model has a name of its own there)
PK_TYPE_NAMES = {PK_ONE: "one", PK_TWO: "two", PK_ONE_BIPHASIC: "one_compartment_biphasic_uptake", PK_TWO_BIPHASIC: "two_biphasic_uptake",
                 PK_ONE_TRANSIT: "one_transit", PK_TWO_TRANSIT: "two_transit"}


def is_two_compartment(pk_type: int) -> bool:
    return pk_type in (PK_TWO, PK_TWO_BIPHASIC, PK_TWO_TRANSIT)


def is_biphasic(pk_type: int) -> bool:
    return pk_type in (PK_ONE_BIPHASIC, PK_TWO_BIPHASIC)


def is_transit(pk_type: int) -> bool:
    return pk_type in (PK_ONE_TRANSIT, PK_TWO_TRANSIT)

# VariableSet transforms, src/sampler/VariableSet.cpp:97-124
TRANSFORM_NONE = 0
TRANSFORM_LOG = 1
TRANSFORM_LOG10 = 2
TRANSFORM_LOGIT = 3

# LikelihoodPopPKTrajectory.cpp:377-393
MOLECULAR_WEIGHT = {
    "lapatinib": 581.06,
    "dacomitinib": 469.95,
    "afatinib": 485.94,
    "trametinib": 615.404,
    "mirdametinib": 482.19,
    "selumetinib": 457.68,
}

F32_1E_6 = float(np.float32(1e-6))  # the reference passes the float literal 1e-6f, cpp:238


def num_pk_params(pk_type: int) -> int:
    """cpp:99-120, as the reference has them (7 for BOTH biphasic types: the two-compartment one has no free slot for its
    second named variable, which therefore has to alias another variable)."""
    return {PK_ONE: 4, PK_TWO: 6, PK_ONE_BIPHASIC: 7, PK_TWO_BIPHASIC: 7, PK_ONE_TRANSIT: 6, PK_TWO_TRANSIT: 8}[pk_type]


@dataclass
class PopPKTrial:
    """One NetCDF trial group (cpp:94-161)."""

    drug: str
    time: np.ndarray  # [T] hours
    observed_concentration: np.ndarray  # [P, T] nM, NaN = missing
    dose: np.ndarray  # [P]
    dosing_interval: np.ndarray  # [P]
    dose_after_dose_change: np.ndarray  # [P] NaN = none
    dose_change_time: np.ndarray  # [P]
    intermittent: np.ndarray  # [P] 0..3
    treatment_interruptions: np.ndarray  # [P, 29] flags per day

    @property
    def num_patients(self) -> int:
        return int(self.observed_concentration.shape[0])

    @property
    def num_timepoints(self) -> int:
        return int(self.time.shape[0])


@dataclass
class PopPKProblem:
    """Derived state of an initialised likelihood (cpp:122-252)."""

    pk_type: int
    trial: PopPKTrial
    transforms: np.ndarray  # [nvar] int32
    sd_ix: int
    fixed_vod: float = math.nan
    fixed_periphery_fwd: float = math.nan
    fixed_periphery_bwd: float = math.nan
    max_steps: int = 2000  # ODESolverCVODE.cpp:45
    # variables the variants look up by NAME (cpp:296-310): their indices in the variable vector
    n_transit_ix: int = -1
    mean_transit_time_ix: int = -1
    biphasic_uptake_time_ix: int = -1
    mean_absorption2_ix: int = -1
    # True: likelihood.xml type="pharmacokinetic_trajectory" (src/likelihoods/LikelihoodPharmacokineticTrajectory.cpp) -- ONE
    # patient of the trial, no population level: variables 0..5 are the rates themselves (its cpp:264-290), the biphasic pair
    # sits at positions 6 and 7 (its cpp:302-303), the whole time vector is simulated, no variable-count check (compiled out)
    single: bool = False
    # derived
    simulate_until: np.ndarray = field(init=False)
    skipped_days: np.ndarray = field(init=False)
    rtol: float = field(init=False)
    atol: float = field(init=False)
    mol_weight: float = field(init=False)

    def __post_init__(self):
        tr = self.trial
        P, T = tr.num_patients, tr.num_timepoints
        if tr.drug not in MOLECULAR_WEIGHT:
            raise ValueError(f'Unknown drug "{tr.drug}"')  # cpp:391
        self.mol_weight = MOLECULAR_WEIGHT[tr.drug]
        fixed = sum(0 if math.isnan(v) else 1 for v in (self.fixed_vod, self.fixed_periphery_fwd, self.fixed_periphery_bwd))
        expected = num_pk_params(self.pk_type) - fixed + 2 * (P + 1) + 2
        self.transforms = np.ascontiguousarray(self.transforms, dtype=np.int32)
        if self.single:
            if P != 1:
                raise ValueError("pharmacokinetic_trajectory is the likelihood of one patient")
            if is_biphasic(self.pk_type):
                self.biphasic_uptake_time_ix, self.mean_absorption2_ix = 6, 7
            last = 7 if is_biphasic(self.pk_type) else (5 if is_two_compartment(self.pk_type) and math.isnan(self.fixed_periphery_fwd) else 3)
            if self.transforms.shape[0] <= max(last, self.sd_ix + 1):
                raise ValueError("the prior is shorter than the positions the likelihood reads")
            expected = self.transforms.shape[0]
        if self.transforms.shape[0] != expected:
            raise ValueError("Incorrect number of variables in prior")  # cpp:127-130
        if is_transit(self.pk_type) and (self.n_transit_ix < 0 or self.mean_transit_time_ix < 0):
            raise ValueError('transit models need the variables "n_transit" and "mean_transit_time"')
        if is_biphasic(self.pk_type) and (self.biphasic_uptake_time_ix < 0 or self.mean_absorption2_ix < 0):
            raise ValueError('biphasic models need the variables "biphasic_uptake_time" and "mean_absorption2"')
        # fixed attributes: the reference keeps indexing the variable vector at the all-sampled positions (cpp:267-272,
        # 283-286) although the prior is `fixed` variables shorter -- reproduced as it is; with all three fixed it would
        # read past the end of the vector
        if P > 0 and not self.single and num_pk_params(self.pk_type) + 2 * P + 1 >= expected:
            raise ValueError("with these fixed pk_model attributes the reference reads past the end of the variable vector")

        inter = np.asarray(tr.treatment_interruptions).reshape(P, 29) != 0
        self.skipped_days = (inter.astype(np.uint64) << np.arange(29, dtype=np.uint64)).sum(axis=1).astype(np.uint32)

        # simulate_until, cpp:163-184
        time = np.asarray(tr.time, dtype=np.float64)
        su = np.full(P, T, dtype=np.int32)
        ge24 = np.nonzero(time >= 24.0)[0]
        first_ge24 = int(ge24[0]) if ge24.size else None
        day1 = inter[:, 1]
        if first_ge24 is not None:
            su[day1] = first_ge24
        else:
            # reference leaves simulate_until[j] value-initialised (0) when no timepoint is >= 24 h
            su[day1] = 0
        obs = np.asarray(tr.observed_concentration, dtype=np.float64)
        notnan = ~np.isnan(obs)
        has = notnan.any(axis=1)
        first = np.argmax(notnan, axis=1)
        late = has & (time[first] > 15 * 24)
        su[late] = 0
        if self.single:
            su[:] = T
        self.simulate_until = su

        dac = np.asarray(tr.dose_after_dose_change, dtype=np.float64)
        dct = np.asarray(tr.dose_change_time, dtype=np.float64)
        bad = ~np.isnan(dac) & np.isnan(dct)
        if bad.any() and not self.single:
            raise ValueError(f"Patient {int(np.nonzero(bad)[0][0])} has dose change, but time of dose change is not specified.")  # cpp:187-190
        min_dose = float(np.min(tr.dose)) if P else float(np.finfo(np.float64).max)
        if (~np.isnan(dac)).any():
            min_dose = min(min_dose, float(np.nanmin(dac)))
        # SetTolerance(1e-6f, minimum_dose * 1e-6f), cpp:238
        self.rtol = F32_1E_6
        self.atol = min_dose * F32_1E_6
        if self.single:
            self.atol = float(tr.dose[0]) * F32_1E_6  # SetTolerance(1e-6f, dose * 1e-6f), LikelihoodPharmacokineticTrajectory.cpp:226

    @property
    def num_variables(self) -> int:
        return int(self.transforms.shape[0])

    @property
    def num_states(self) -> int:
        return 2 if self.pk_type == PK_ONE else 3
