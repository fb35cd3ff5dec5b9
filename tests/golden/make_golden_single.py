"""Generates tests/golden/pksingle_*.npz -- the likelihood of ONE patient, likelihood.xml type="pharmacokinetic_trajectory"
(src/likelihoods/LikelihoodPharmacokineticTrajectory.cpp) -- with the reference's own compiled solver stack (oracle/_ref) behind
the restatement of that file's glue (oracle/ref/poppk_ref.cpp, `single`). Run where /root/reference is mounted:
    python tests/golden/make_golden_single.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from bcm3_b200 import synthetic as syn  # noqa: E402
from bcm3_b200.poppk_data import PK_ONE, PK_ONE_TRANSIT, PK_TWO, PK_TWO_BIPHASIC, PK_TWO_TRANSIT  # noqa: E402
from make_golden import problem_arrays  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {
    # name: (pk_type, T, t_end, seed, fixed attributes)
    "pksingle_one": (PK_ONE, 12, 96.0, 3, {}),
    "pksingle_two": (PK_TWO, 12, 96.0, 4, {}),
    "pksingle_two_biphasic": (PK_TWO_BIPHASIC, 12, 120.0, 6, {}),
    "pksingle_one_transit": (PK_ONE_TRANSIT, 12, 96.0, 7, {}),
    "pksingle_two_transit": (PK_TWO_TRANSIT, 12, 120.0, 8, {}),
    "pksingle_two_fixed": (PK_TWO, 10, 96.0, 9, dict(fixed_vod=45.0, fixed_periphery_fwd=0.3, fixed_periphery_bwd=0.1)),
    # long horizon: chains that exceed max_steps = 2000 evaluate to -inf
    "pksingle_two_maxsteps": (PK_TWO, 8, 1000.0, 10, {}),
}


def main():
    ref = oracle.load("ref")
    for name, (pk, T, t_end, seed, fixed) in CASES.items():
        prob = syn.make_single_patient_problem(pk, T=T, t_end=t_end, seed=seed, **fixed)
        vals = syn.make_single_patient_values(prob, 5, seed=seed * 1000)
        r = ref.poppk_evaluate(prob, vals, threads=1, want_conc=True, want_patient_ll=True, want_counters=True)
        out = problem_arrays(prob)
        out.update(single=np.int32(1), fixed_vod=prob.fixed_vod, fixed_periphery_fwd=prob.fixed_periphery_fwd, fixed_periphery_bwd=prob.fixed_periphery_bwd,
                   values=vals, logp=r["logp"], conc=r["conc"], patient_ll=r["patient_ll"], counters=r["counters"].astype(np.int32))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "logp", r["logp"], "steps", r["counters"][:, 0, 0], "intermittent", prob.trial.intermittent, "interval", prob.trial.dosing_interval)


if __name__ == "__main__":
    main()
