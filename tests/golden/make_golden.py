"""Generates tests/golden/poppk_*.npz with the reference's OWN compiled solver stack (oracle/_ref).

Run in the build container, where /root/reference is mounted:
    python tests/golden/make_golden.py
Each fixture holds a complete small trial, C parameter vectors and what the reference computes for them:
per-chain log-likelihoods, per-patient log-likelihoods, simulated concentrations and CVODE counters.
The reference ships no fixtures for this path, so these ARE the golden vectors.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from bcm3_b200 import synthetic as syn  # noqa: E402
from bcm3_b200.poppk_data import PK_ONE, PK_ONE_BIPHASIC, PK_ONE_TRANSIT, PK_TWO, PK_TWO_BIPHASIC, PK_TWO_TRANSIT  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (pk_type, P, T, t_end, heterogeneous, missing_fraction, C, seed)
    "poppk_one_plain": (PK_ONE, 64, 10, 72.0, False, 0.0, 3, 11),
    "poppk_one_hetero": (PK_ONE, 96, 12, 120.0, True, 0.15, 3, 12),
    "poppk_two_plain": (PK_TWO, 64, 10, 72.0, False, 0.0, 3, 13),
    "poppk_two_hetero": (PK_TWO, 96, 12, 120.0, True, 0.15, 3, 14),
    # long horizon: some patients exceed max_steps = 2000 => -inf (ODESolverCVODE.cpp:440-446)
    "poppk_one_maxsteps": (PK_ONE, 32, 8, 600.0, True, 0.0, 2, 15),
    # the variants of LikelihoodPopPKTrajectory.cpp:496-642, 692-718
    "poppk_one_biphasic": (PK_ONE_BIPHASIC, 64, 10, 96.0, True, 0.1, 3, 16),
    "poppk_two_biphasic": (PK_TWO_BIPHASIC, 64, 10, 96.0, True, 0.1, 3, 17),
    "poppk_one_transit": (PK_ONE_TRANSIT, 64, 10, 96.0, True, 0.1, 3, 18),
    "poppk_two_transit": (PK_TWO_TRANSIT, 64, 10, 96.0, True, 0.1, 3, 19),
}


def problem_arrays(prob):
    tr = prob.trial
    return dict(
        pk_type=np.int32(prob.pk_type), drug=np.array(tr.drug), time=tr.time, observed_concentration=tr.observed_concentration,
        dose=tr.dose, dosing_interval=tr.dosing_interval, dose_after_dose_change=tr.dose_after_dose_change,
        dose_change_time=tr.dose_change_time, intermittent=tr.intermittent, treatment_interruptions=tr.treatment_interruptions,
        transforms=prob.transforms, sd_ix=np.int32(prob.sd_ix), n_transit_ix=np.int32(prob.n_transit_ix),
        mean_transit_time_ix=np.int32(prob.mean_transit_time_ix), biphasic_uptake_time_ix=np.int32(prob.biphasic_uptake_time_ix),
        mean_absorption2_ix=np.int32(prob.mean_absorption2_ix))


def main():
    ref = oracle.load("ref")
    only = sys.argv[1:]
    for name, (pk, P, T, t_end, het, miss, C, seed) in CASES.items():
        if only and name not in only:
            continue
        prob = syn.make_poppk_problem(pk, P=P, T=T, t_end=t_end, heterogeneous=het, missing_fraction=miss, seed=seed)
        if name.endswith("hetero"):
            # exercise simulate_until: day-1 interruption => only the first day; late first observation => nothing
            prob.trial.treatment_interruptions[3, 1] = 1
            prob.trial.observed_concentration[5, :] = np.nan
            prob = type(prob)(pk_type=prob.pk_type, trial=prob.trial, transforms=prob.transforms, sd_ix=prob.sd_ix)
        if name.endswith("maxsteps"):
            # first non-missing observation later than 15 days => simulate_until = 0 (LikelihoodPopPKTrajectory.cpp:176-184)
            prob.trial.observed_concentration[2, :5] = np.nan
            prob = type(prob)(pk_type=prob.pk_type, trial=prob.trial, transforms=prob.transforms, sd_ix=prob.sd_ix)
            assert prob.simulate_until[2] == 0
        vals = syn.make_chain_values(prob, C, seed=seed * 1000)
        r = ref.poppk_evaluate(prob, vals, threads=1, want_conc=True, want_patient_ll=True, want_counters=True)
        out = problem_arrays(prob)
        out.update(values=vals, logp=r["logp"], conc=r["conc"], patient_ll=r["patient_ll"], counters=r["counters"].astype(np.int32))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "logp", r["logp"], "ok frac", r["counters"][..., 7].mean(), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
