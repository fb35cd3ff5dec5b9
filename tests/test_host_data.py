"""CPU suite: host-side data model (mirror of LikelihoodPopPKTrajectory::Initialize, cpp:122-252)."""
import numpy as np
import pytest

from bcm3_b200 import synthetic as syn
from bcm3_b200.poppk_data import F32_1E_6, PK_ONE, PK_TWO, PopPKProblem


def test_tolerances_use_the_float_literal():
    prob = syn.make_poppk_problem(PK_ONE, P=4, dose=100.0)
    assert prob.rtol == float(np.float32(1e-6)) == F32_1E_6 != 1e-6
    assert prob.atol == 100.0 * F32_1E_6


def test_variable_count_is_checked():
    prob = syn.make_poppk_problem(PK_TWO, P=5)
    assert prob.num_variables == 6 + 2 * 6 + 2
    with pytest.raises(ValueError):
        PopPKProblem(pk_type=PK_ONE, trial=prob.trial, transforms=prob.transforms, sd_ix=prob.sd_ix)


def test_simulate_until_rules():
    prob = syn.make_poppk_problem(PK_ONE, P=4, T=10, t_end=400.0)  # times 40, 80, ..., 400
    tr = prob.trial
    tr.treatment_interruptions[1, 1] = 1           # day-2 interruption: simulate the first day only
    tr.observed_concentration[2, :9] = np.nan       # first observation at 400 h > 15 days: nothing simulated
    tr.observed_concentration[3, :] = np.nan        # no observation at all: everything simulated
    p2 = PopPKProblem(pk_type=PK_ONE, trial=tr, transforms=prob.transforms, sd_ix=prob.sd_ix)
    assert p2.simulate_until.tolist() == [10, 0, 0, 10]
    assert p2.skipped_days[1] == 2


def test_minimum_dose_includes_dose_changes():
    prob = syn.make_poppk_problem(PK_ONE, P=3, dose=100.0)
    prob.trial.dose_after_dose_change[1] = 25.0
    prob.trial.dose_change_time[1] = 24.0
    p2 = PopPKProblem(pk_type=PK_ONE, trial=prob.trial, transforms=prob.transforms, sd_ix=prob.sd_ix)
    assert p2.atol == 25.0 * F32_1E_6
    prob.trial.dose_change_time[1] = np.nan
    with pytest.raises(ValueError):
        PopPKProblem(pk_type=PK_ONE, trial=prob.trial, transforms=prob.transforms, sd_ix=prob.sd_ix)
