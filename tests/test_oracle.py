"""CPU suite: the checkers themselves. The plain-C restatement (oracle/cvode_bdf.c + poppk_oracle.c) is pinned
against golden vectors produced by the reference's own compiled CVODE/odecommon stack (tests/golden/make_golden.py);
where that compiled reference is present it must also reproduce its own fixtures bit for bit."""
import ctypes as C

import numpy as np
import pytest

from bcm3_b200 import synthetic as syn
from bcm3_b200.poppk_data import PK_ONE, PK_TWO
from tests.util import GOLDEN_NAMES, assert_matches_golden, counter_match_floor, load_golden, rel_err


def test_golden_fixtures_exist():
    assert len(GOLDEN_NAMES) >= 5


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_port_matches_reference_golden(port, name):
    prob, gold = load_golden(name)
    r = port.poppk_evaluate(prob, gold["values"], threads=2, want_conc=True, want_patient_ll=True, want_counters=True)
    assert_matches_golden(r["logp"], r["conc"], r["counters"], gold, logp_tol=1e-7, min_counter_match=counter_match_floor(name))
    # per-patient log-likelihoods: same -inf pattern, finite ones close
    assert (np.isneginf(r["patient_ll"]) == np.isneginf(gold["patient_ll"])).all()
    fin = np.isfinite(gold["patient_ll"])
    assert np.abs(r["patient_ll"][fin] - gold["patient_ll"][fin]).max() < 1e-2
    # trajectories: round-off flips a few step decisions (the reference itself moves by 3e-5 between an FMA and a
    # non-FMA build); the bulk must agree tightly
    m = ~np.isnan(gold["conc"])
    rel = rel_err(r["conc"][m], gold["conc"][m])
    assert np.median(rel) < 1e-9
    assert (rel > 1e-6).mean() < 1.5 * (1.0 - counter_match_floor(name))


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_compiled_reference_reproduces_golden(ref, name):
    prob, gold = load_golden(name)
    r = ref.poppk_evaluate(prob, gold["values"], threads=1, want_conc=True, want_patient_ll=True, want_counters=True)
    assert np.array_equal(r["logp"], gold["logp"])
    assert np.array_equal(r["counters"], gold["counters"])
    assert np.array_equal(r["conc"], gold["conc"], equal_nan=True)


def test_threads_do_not_change_results(port):
    prob, gold = load_golden("poppk_two_hetero")
    a = port.poppk_evaluate(prob, gold["values"], threads=1)["logp"]
    b = port.poppk_evaluate(prob, gold["values"], threads=3)["logp"]
    assert np.array_equal(a, b)


def test_ndtri_matches_scipy(port):
    from scipy.special import ndtri

    f = port.lib.oracle_ndtri
    f.restype = C.c_double
    f.argtypes = [C.c_double]
    p = np.concatenate([np.linspace(1e-12, 1 - 1e-12, 2001), 10.0 ** -np.arange(1, 300, 7.0), [0.5, 0.02, 0.98]])
    got = np.array([f(float(x)) for x in p])
    want = ndtri(p)
    assert np.abs(got - want).max() <= 4e-15 * np.maximum(1.0, np.abs(want)).max()
    assert f(0.0) == -np.inf and f(1.0) == np.inf and np.isnan(f(-0.1))


@pytest.mark.parametrize("pk", [PK_ONE, PK_TWO])
def test_cvode_tracks_exact_linear_solution(port, pk):
    """Known-answer sanity bound (not a parity target): at rtol 1e-6 CVODE stays within ~1e-3 of the exact
    matrix-exponential solution of the linear compartment model (the survey measured 2.4e-4)."""
    prob = syn.make_poppk_problem(pk, P=200, T=10, t_end=72.0, seed=5)
    vals = syn.make_chain_values(prob, 1, seed=99)
    r = port.poppk_evaluate(prob, vals, want_conc=True)
    pop = syn.population_defaults(pk)
    from scipy.special import ndtri

    npk = 4 if pk == PK_ONE else 6
    v = vals[0]
    pka = v[npk + 2::2][:200]
    pcl = v[npk + 3::2][:200]
    ka = 10.0 ** (v[0] + v[npk] * ndtri(pka))
    vd = 10.0 ** v[3]
    kel = 10.0 ** (v[2] + v[npk + 1] * ndtri(pcl)) / vd
    kex = np.full(200, 10.0 ** v[1])
    kf = np.full(200, 10.0 ** pop.get("log_kf", 0.0))
    kb = np.full(200, 10.0 ** pop.get("log_kb", 0.0))
    exact = syn.exact_linear_pk(pk, ka, kex, kel, kf, kb, prob.trial.dose, prob.trial.dosing_interval, prob.trial.time)
    exact *= (1e6 / prob.mol_weight) / vd
    rel = np.abs(r["conc"][0] - exact) / np.abs(exact).max(axis=1, keepdims=True)
    assert rel.max() < 2e-3


def test_edge_cases(port):
    # empty trial
    prob = syn.make_poppk_problem(PK_ONE, P=0, T=4)
    r = port.poppk_evaluate(prob, syn.make_chain_values(prob, 2))
    assert np.array_equal(r["logp"], [0.0, 0.0])
    # a timepoint at t = 0 returns the initial condition (ODESolver.cpp:109-118); all-missing observations give 0
    prob = syn.make_poppk_problem(PK_ONE, P=3, T=5, t_end=48.0)
    prob.trial.time[0] = 0.0
    prob.trial.observed_concentration[1, :] = np.nan
    prob = type(prob)(pk_type=prob.pk_type, trial=prob.trial, transforms=prob.transforms, sd_ix=prob.sd_ix)
    r = port.poppk_evaluate(prob, syn.make_chain_values(prob, 1), want_conc=True, want_patient_ll=True)
    assert np.all(r["conc"][0, :, 0] == 0.0)
    assert r["patient_ll"][0, 1] == 0.0
    assert np.isfinite(r["logp"]).all()


# ---- pharmacokinetic_trajectory: the likelihood of one patient (LikelihoodPharmacokineticTrajectory.cpp) ----
from tests.util import SINGLE_GOLDEN_NAMES  # noqa: E402


def test_single_patient_fixtures_exist():
    assert len(SINGLE_GOLDEN_NAMES) >= 6


@pytest.mark.parametrize("name", SINGLE_GOLDEN_NAMES)
def test_single_patient_port_matches_reference_golden(port, name):
    """The plain-C restatement against the compiled reference: the chain's variables are the patient's rates, the whole time
    vector is simulated, any intermittent schedule acts as schedule 1, the biphasic switching time is not clipped."""
    prob, gold = load_golden(name)
    assert prob.single and prob.trial.num_patients == 1
    r = port.poppk_evaluate(prob, gold["values"], threads=2, want_conc=True, want_counters=True)
    assert (np.isneginf(r["logp"]) == np.isneginf(gold["logp"])).all()
    assert rel_err(r["logp"], gold["logp"]).max() < 1e-7
    assert (r["counters"] == gold["counters"]).all(axis=-1).mean() >= 0.8  # 5 systems per fixture: at most one may differ in a counter
    m = ~np.isnan(gold["conc"]) & np.isfinite(gold["logp"])[:, None, None]
    assert np.median(rel_err(r["conc"][m], gold["conc"][m])) < 1e-9


@pytest.mark.parametrize("name", SINGLE_GOLDEN_NAMES)
def test_single_patient_reference_reproduces_golden(ref, name):
    prob, gold = load_golden(name)
    r = ref.poppk_evaluate(prob, gold["values"], threads=1, want_conc=True, want_counters=True)
    assert np.array_equal(r["logp"], gold["logp"]) and np.array_equal(r["counters"], gold["counters"])


def test_single_patient_differs_from_the_population_likelihood_where_the_reference_does(port):
    """Same trial arrays, same numbers in the variable vector: the two likelihood types read them differently."""
    from bcm3_b200.poppk_data import PK_TWO_BIPHASIC

    prob = syn.make_single_patient_problem(PK_TWO_BIPHASIC, seed=6)
    vals = syn.make_single_patient_values(prob, 3)
    a = port.poppk_evaluate(prob, vals)["logp"]
    # the schedule is a bool in this likelihood (its cpp:184-186): 2 and 3 act as 1
    for schedule in (2, 3):
        prob.trial.intermittent[:] = schedule
        b = port.poppk_evaluate(prob, vals)["logp"]
        prob.trial.intermittent[:] = 1
        assert np.array_equal(b, port.poppk_evaluate(prob, vals)["logp"])
    assert np.isfinite(a).all()
