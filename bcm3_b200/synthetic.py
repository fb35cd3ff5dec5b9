"""Synthetic PopPK workloads of the shapes BASELINE.json names.

The reference ships no data files, so trials and chain parameter vectors are
generated here with fixed seeds.  Observations come from the EXACT solution of
the linear compartment model (eigen-decomposition propagator between events),
multiplied by log-normal noise -- nothing here touches the oracle or the GPU.
"""
from __future__ import annotations

import math

import numpy as np

from .poppk_data import (
    PK_ONE,
    PK_ONE_BIPHASIC,
    PK_ONE_TRANSIT,
    PK_TWO,
    PK_TWO_BIPHASIC,
    PK_TWO_TRANSIT,
    is_biphasic,
    is_transit,
    is_two_compartment,
    TRANSFORM_LOG10,
    TRANSFORM_NONE,
    PopPKProblem,
    PopPKTrial,
    MOLECULAR_WEIGHT,
    num_pk_params,
)

LN10 = 2.3025850929940459


def population_defaults(pk_type: int) -> dict:
    """Population-level parameters of the synthetic configs (SURVEY.md 8d cfg 2/5)."""
    d = dict(mu_logka=0.1, log_kex=-1.3, mu_logcl=0.9, log_vd=1.7, sigma_ka=0.3, sigma_cl=0.3, log_sd=0.0, log_sd2=math.log10(0.2))
    if is_two_compartment(pk_type):
        d.update(log_kf=-0.5, log_kb=-1.0)
    if is_biphasic(pk_type):
        d.update(log_uptake_time=0.3, log_ka2=-0.5)
    if is_transit(pk_type):
        d.update(log_n_transit=0.5, log_transit_time=0.3)
    return d


def named_variable_indices(pk_type: int) -> dict:
    """Where the synthetic prior puts the variables the variants look up by name (cpp:296-310). The reference counts 7
    "pk params" for BOTH biphasic types (cpp:105-110), so the two-compartment one has a single free slot: its
    mean_absorption2 aliases variable 1 (the excretion rate) here."""
    if pk_type == PK_ONE_BIPHASIC:
        return dict(biphasic_uptake_time_ix=4, mean_absorption2_ix=5)
    if pk_type == PK_TWO_BIPHASIC:
        return dict(biphasic_uptake_time_ix=6, mean_absorption2_ix=1)
    if pk_type == PK_ONE_TRANSIT:
        return dict(n_transit_ix=4, mean_transit_time_ix=5)
    if pk_type == PK_TWO_TRANSIT:
        return dict(n_transit_ix=6, mean_transit_time_ix=7)
    return {}


def poppk_transforms(pk_type: int, P: int) -> np.ndarray:
    """prior.xml layout: [0] mean log10 ka (none), [1] kex (logspace), [2] mean log10 CL (none), [3] Vd (logspace),
    ([4],[5] kf, kb logspace), sigma_ka, sigma_cl (none), 2 probabilities per patient (none),
    standard_deviation, proportional sd (logspace)."""
    npk = num_pk_params(pk_type)
    nvar = npk + 2 * (P + 1) + 2
    tr = np.zeros(nvar, dtype=np.int32)
    tr[1] = TRANSFORM_LOG10
    tr[3] = TRANSFORM_LOG10
    if is_two_compartment(pk_type):
        tr[4] = TRANSFORM_LOG10
        tr[5] = TRANSFORM_LOG10
    for ix in named_variable_indices(pk_type).values():
        tr[ix] = TRANSFORM_LOG10
    tr[nvar - 2] = TRANSFORM_LOG10
    tr[nvar - 1] = TRANSFORM_LOG10
    return tr


def _ndtri(p: np.ndarray) -> np.ndarray:
    from scipy.special import ndtri

    return ndtri(p)


def exact_linear_pk(pk_type, ka, kex, kel, kf, kb, dose, dosing_interval, times, give_dose=None):
    """Exact amounts in the central compartment at `times` [T] for P patients (arrays [P]).

    Bolus `dose` into the depot at t=0 and at every multiple of dosing_interval
    (optionally masked by give_dose[P, ndoses]).  Returns [P, T].
    """
    P = ka.shape[0]
    if P == 0:
        return np.zeros((0, np.asarray(times).shape[0]))
    N = 3 if is_two_compartment(pk_type) else 2
    A = np.zeros((P, N, N))
    A[:, 0, 0] = -(ka + kex)
    A[:, 1, 0] = ka
    if N == 2:
        A[:, 1, 1] = -kel
    else:
        A[:, 1, 1] = -(kel + kf)
        A[:, 1, 2] = kb
        A[:, 2, 1] = kf
        A[:, 2, 2] = -kb
    lam, V = np.linalg.eig(A)
    Vinv = np.linalg.inv(V)
    times = np.asarray(times, dtype=np.float64)
    t_end = float(times[-1])
    out = np.zeros((P, times.shape[0]))
    # superposition of single-dose responses: x(t) = sum_d expm(A (t - t_d)) e0 * dose_d
    e0 = Vinv[:, :, 0]  # V^-1 e_0
    nd_max = int(np.ceil(t_end / float(np.min(dosing_interval)))) + 1
    for d in range(nd_max):
        td = d * dosing_interval  # [P]
        s = times[None, :] - td[:, None]  # [P, T]
        # doses strictly before t contribute; a dose given exactly at an output time is applied after the output
        active = (s > 0) | ((d == 0) & (s >= 0))
        if not active.any():
            break
        amt = dose.copy()
        if give_dose is not None and d > 0:
            amt = amt * give_dose[:, d]
        sc = np.where(active, s, 0.0)
        ex = np.exp(lam[:, None, :] * sc[:, :, None])  # [P, T, N]
        x1 = np.einsum("pk,ptk,pk->pt", V[:, 1, :], ex, e0)
        out += np.where(active, np.real(x1) * amt[:, None], 0.0)
    return out


def make_poppk_problem(pk_type: int = PK_ONE, P: int = 1000, T: int = 10, t_end: float = 72.0, dosing_interval: float = 24.0,
                       dose: float = 100.0, drug: str = "lapatinib", seed: int = 1, missing_fraction: float = 0.0,
                       heterogeneous: bool = False) -> PopPKProblem:
    """cfg 2 (PK_ONE, P=1000) / cfg 5 (PK_TWO, P=100000) style trial.

    heterogeneous=True additionally varies dose / dosing interval / intermittent schedules /
    skipped days / dose changes across patients to exercise every branch of the dosing logic
    (LikelihoodPopPKTrajectory.cpp:644-690).
    """
    rng = np.random.default_rng(seed)
    time = t_end * (np.arange(T) + 1.0) / T
    pop = population_defaults(pk_type)
    p_ka = rng.uniform(0.02, 0.98, P)
    p_cl = rng.uniform(0.02, 0.98, P)
    ka = 10.0 ** (pop["mu_logka"] + pop["sigma_ka"] * _ndtri(p_ka))
    vd = 10.0 ** pop["log_vd"]
    kel = 10.0 ** (pop["mu_logcl"] + pop["sigma_cl"] * _ndtri(p_cl)) / vd
    kex = np.full(P, 10.0 ** pop["log_kex"])
    kf = np.full(P, 10.0 ** pop.get("log_kf", 0.0))
    kb = np.full(P, 10.0 ** pop.get("log_kb", 0.0))

    doses = np.full(P, dose)
    intervals = np.full(P, dosing_interval)
    dac = np.full(P, np.nan)
    dct = np.full(P, np.nan)
    intermittent = np.zeros(P, dtype=np.uint32)
    interruptions = np.zeros((P, 29), dtype=np.uint32)
    if heterogeneous:
        doses = rng.choice([50.0, 100.0, 250.0], P)
        intervals = rng.choice([12.0, 24.0, 24.0, 48.0], P)
        intermittent = rng.choice([0, 0, 1, 2, 3], P).astype(np.uint32)
        skip = rng.uniform(size=(P, 29)) < 0.05
        skip[:, 0] = False
        interruptions = skip.astype(np.uint32)
        change = rng.uniform(size=P) < 0.3
        dac[change] = rng.choice([25.0, 50.0, 150.0], int(change.sum()))
        dct[change] = intervals[change] * rng.integers(1, 4, int(change.sum()))

    conc = exact_linear_pk(pk_type, ka, kex, kel, kf, kb, doses, intervals, time) * (1e6 / MOLECULAR_WEIGHT[drug]) / vd
    obs = conc * np.exp(0.2 * rng.standard_normal((P, T)))
    if missing_fraction > 0:
        obs[rng.uniform(size=(P, T)) < missing_fraction] = np.nan

    trial = PopPKTrial(drug=drug, time=time, observed_concentration=obs, dose=doses, dosing_interval=intervals,
                       dose_after_dose_change=dac, dose_change_time=dct, intermittent=intermittent,
                       treatment_interruptions=interruptions)
    nvar = num_pk_params(pk_type) + 2 * (P + 1) + 2
    return PopPKProblem(pk_type=pk_type, trial=trial, transforms=poppk_transforms(pk_type, P), sd_ix=nvar - 2, **named_variable_indices(pk_type))


def make_chain_values(problem: PopPKProblem, C: int, seed: int = 20261018) -> np.ndarray:
    """One proposal per tempered chain: values[C, nvar] (SURVEY.md 8d cfg 2: seed 20261018+c)."""
    P = problem.trial.num_patients
    npk = num_pk_params(problem.pk_type)
    nvar = problem.num_variables
    pop = population_defaults(problem.pk_type)
    out = np.empty((C, nvar), dtype=np.float64)
    for c in range(C):
        rng = np.random.default_rng(seed + c)
        v = out[c]
        v[0] = rng.normal(pop["mu_logka"], 0.1)
        v[1] = pop["log_kex"]
        v[2] = rng.normal(pop["mu_logcl"], 0.1)
        v[3] = pop["log_vd"]
        v[4:npk] = 0.0
        if is_two_compartment(problem.pk_type):
            v[4] = pop["log_kf"]
            v[5] = pop["log_kb"]
        named = named_variable_indices(problem.pk_type)
        if is_biphasic(problem.pk_type):
            v[named["biphasic_uptake_time_ix"]] = rng.normal(pop["log_uptake_time"], 0.05)
            if named["mean_absorption2_ix"] >= 4:
                v[named["mean_absorption2_ix"]] = rng.normal(pop["log_ka2"], 0.05)
        if is_transit(problem.pk_type):
            v[named["n_transit_ix"]] = rng.normal(pop["log_n_transit"], 0.05)
            v[named["mean_transit_time_ix"]] = rng.normal(pop["log_transit_time"], 0.05)
        v[npk + 0] = pop["sigma_ka"]
        v[npk + 1] = pop["sigma_cl"]
        v[npk + 2: npk + 2 + 2 * P] = rng.uniform(0.02, 0.98, 2 * P)
        v[nvar - 2] = pop["log_sd"]
        v[nvar - 1] = pop["log_sd2"]
    return out


# ---- pharmacokinetic_trajectory: the likelihood of ONE patient (LikelihoodPharmacokineticTrajectory.cpp) ----
SINGLE_NVAR = 10  # prior layout: ka, kex, clearance, Vd, kf, kb, (switch time | n_transit), (ka2 | mean_transit_time), standard_deviation, sd2


def make_single_patient_problem(pk_type: int = PK_TWO, T: int = 12, t_end: float = 96.0, seed: int = 1, heterogeneous: bool = True,
                                missing_fraction: float = 0.1, **fixed) -> PopPKProblem:
    """One patient of a synthetic trial with the prior a pharmacokinetic_trajectory model directory has: every variable is a
    rate in log10 space (positions as the likelihood reads them, its cpp:264-305), the named transit pair at 6 and 7."""
    pop = make_poppk_problem(pk_type, P=1, T=T, t_end=t_end, seed=seed, heterogeneous=heterogeneous, missing_fraction=missing_fraction)
    named = {}
    if is_transit(pk_type):
        named = dict(n_transit_ix=6, mean_transit_time_ix=7)
    if is_biphasic(pk_type):
        named = dict(biphasic_uptake_time_ix=6, mean_absorption2_ix=7)
    transforms = np.full(SINGLE_NVAR, TRANSFORM_LOG10, dtype=np.int32)
    return PopPKProblem(pk_type=pk_type, trial=pop.trial, transforms=transforms, sd_ix=SINGLE_NVAR - 2, single=True, **named, **fixed)


def make_single_patient_values(problem: PopPKProblem, C: int, seed: int = 20261018) -> np.ndarray:
    pop = population_defaults(problem.pk_type)
    out = np.zeros((C, SINGLE_NVAR))
    for c in range(C):
        rng = np.random.default_rng(seed + c)
        v = out[c]
        v[0] = rng.normal(pop["mu_logka"], 0.3)
        v[1] = rng.normal(pop["log_kex"], 0.1)
        v[2] = rng.normal(pop["mu_logcl"], 0.3)
        v[3] = rng.normal(pop["log_vd"], 0.1)
        v[4] = rng.normal(pop.get("log_kf", -0.5), 0.1)
        v[5] = rng.normal(pop.get("log_kb", -1.0), 0.1)
        if is_biphasic(problem.pk_type):
            v[6] = rng.normal(pop["log_uptake_time"], 0.05)
            v[7] = rng.normal(pop["log_ka2"], 0.05)
        if is_transit(problem.pk_type):
            v[6] = rng.normal(pop["log_n_transit"], 0.05)
            v[7] = rng.normal(pop["log_transit_time"], 0.05)
        v[8] = pop["log_sd"]
        v[9] = pop["log_sd2"]
    return out
