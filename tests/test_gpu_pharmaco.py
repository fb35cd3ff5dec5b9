"""GPU suite for the pharmaco_population path (csrc/pharmaco_kernel.cuh through the C ABI): golden vectors of the reference's
compiled compartment model, fresh inputs against the CPU checker, shards, batch independence."""
import numpy as np
import pytest

from bcm3_b200 import pharmaco as ph
from tests.util import PHARMACO_GOLDEN_NAMES, load_pharmaco_golden, rel_err

pytestmark = pytest.mark.gpu
LOGP_RTOL = 1e-6  # BASELINE.json north star; measured: 1e-13


@pytest.fixture(scope="module")
def Evaluator(built):
    from bcm3_b200 import _lib

    assert _lib.device_count() > 0
    return ph.PharmacoEvaluator


@pytest.mark.parametrize("name", PHARMACO_GOLDEN_NAMES)
def test_matches_reference_golden(Evaluator, name):
    prob, gold = load_pharmaco_golden(name)
    ev = Evaluator(prob, diagnostics=True)
    logp, status = ev.evaluate(gold["values"])
    d = ev.diagnostics()
    ev.close()
    assert (status == 0).all()
    assert rel_err(logp, gold["logp"]).max() <= 1e-10  # far inside the 1e-6 bar: there is no step-size control to disagree on
    assert (np.isnan(d["conc"]) == np.isnan(gold["conc"])).all()
    m = ~np.isnan(gold["conc"])
    assert np.abs(d["conc"][m] - gold["conc"][m]).max() <= 1e-10 * max(1.0, np.abs(gold["conc"][m]).max())
    assert np.abs(d["patient_ll"] - gold["patient_ll"]).max() < 1e-8


@pytest.mark.parametrize("kw", [dict(), dict(peripheral=True), dict(peripheral=True, num_transit=5, bioavailability=True), dict(excretion=False, heterogeneous=False)])
def test_fresh_inputs_against_cpu_checker(Evaluator, checker, kw):
    prob = ph.make_pharmaco_problem(P=500, T=12, seed=17, **kw)
    vals = ph.make_pharmaco_values(prob, 8, seed=99)
    want = checker.pharmaco_evaluate(prob, vals, threads=8)
    ev = Evaluator(prob)
    got, status = ev.evaluate(vals)
    single, _ = ev.evaluate(vals[3:4])
    ev.close()
    assert (status == 0).all() and np.isfinite(got).all()
    assert rel_err(got, want["logp"]).max() <= LOGP_RTOL
    assert single[0] == got[3]  # a chain's result does not depend on its batch


def test_shards_add_up_and_failures_propagate(Evaluator, checker):
    prob = ph.make_pharmaco_problem(P=301, T=10, seed=5)
    vals = ph.make_pharmaco_values(prob, 4, seed=7)
    ev = Evaluator(prob)
    want, _ = ev.evaluate(vals)
    ev.close()
    parts = []
    for r in range(3):  # without a communicator a sharded handle returns its shard's sum
        e = Evaluator(prob, shard_rank=r, shard_count=3)
        parts.append(e.evaluate(vals)[0])
        e.close()
    assert rel_err(np.sum(parts, axis=0), want).max() < 1e-13
    # a NaN volume of distribution makes every simulated concentration of that chain NaN: the reference turns a NaN or infinite
    # concentration into a log-likelihood of -inf (cpp:224-227), not NaN, and the other chains are untouched
    bad = vals.copy()
    bad[1, prob.index("mean_volume_of_distribution")] = np.nan
    ev = Evaluator(prob)
    got, status = ev.evaluate(bad)
    ev.close()
    ref = checker.pharmaco_evaluate(prob, bad, threads=2)["logp"]
    assert np.isneginf(ref[1]) and np.isneginf(got[1]) and status[1] == 0 and np.array_equal(got[[0, 2, 3]], want[[0, 2, 3]])


def test_plugin_through_the_factory(built):
    """likelihood.xml type="pharmaco_population" -> LikelihoodFactory -> PharmacoLikelihoodPopulationB200: variables found by name,
    per-patient marginals p<i>_<name> resolved as InitializePatientMarginals does; equal to the direct ABI call."""
    from bcm3_b200 import host_api
    from bcm3_b200.poppk_data import TRANSFORM_LOG10

    prob = ph.make_pharmaco_problem(P=40, T=10, peripheral=True, num_transit=3, seed=8)
    vals = ph.make_pharmaco_values(prob, 3, seed=9)
    ev = ph.PharmacoEvaluator(prob)
    want, _ = ev.evaluate(vals)
    ev.close()
    prior = "<variableset>" + "".join(
        f'<variable name="{n}" {"logspace=" + chr(34) + "true" + chr(34) + " " if prob.transforms[i] == TRANSFORM_LOG10 else ""}distribution="uniform" lower="-5" upper="5"/>'
        for i, n in enumerate(prob.variable_names)) + "</variableset>"
    lik = ('<bcm_likelihood type="pharmaco_population"><pk_model drug="lapatinib" trial="synthetic" peripheral_compartment="true" '
           'num_transit_compartments="3"/></bcm_likelihood>')
    for batched in (True, False):
        got = host_api.pharmaco_evaluate(prior, lik, prob.trial, vals, batched=batched)
        assert np.array_equal(got, want)


# ---- pharmaco_single: PharmacoLikelihoodSingle, one patient (the goldens pharmaco_single_* run with the parametrised tests above) ----
@pytest.mark.parametrize("kw", [dict(), dict(peripheral=True, num_transit=4, metabolite=True), dict(biphasic_absorption=True, excretion=False)])
def test_single_patient_fresh_inputs(Evaluator, checker, kw):
    prob = ph.make_pharmaco_single_problem(T=16, seed=31, **kw)
    vals = ph.make_pharmaco_single_values(prob, 64, seed=7)
    vals[5, 0] = np.nan  # a NaN absorption rate: NaN state, Solve returns false, -inf (PharmacoLikelihoodSingle.cpp:215-217)
    want = checker.pharmaco_evaluate(prob, vals, threads=4)["logp"]
    ev = Evaluator(prob)
    got, status = ev.evaluate(vals)
    ev.close()
    assert np.isneginf(got[5]) and np.isneginf(want[5]) and (status == 0).all()
    assert rel_err(got, want).max() <= 1e-10


def test_single_patient_plugin_through_the_factory(built):
    """likelihood.xml type="pharmaco_single" -> LikelihoodFactory -> the plugin picks <pk_model patient=> out of the trial."""
    from bcm3_b200 import host_api
    from bcm3_b200 import synthetic as syn
    from bcm3_b200.poppk_data import PK_TWO

    trial = syn.make_poppk_problem(PK_TWO, P=4, T=12, t_end=120.0, seed=12, heterogeneous=True, missing_fraction=0.1).trial
    one = ph.make_pharmaco_single_problem(peripheral=True, biphasic_absorption=True, metabolite=True)
    pick = lambda a: np.asarray(a)[2:3]
    one.trial = type(trial)(drug=trial.drug, time=trial.time, observed_concentration=pick(trial.observed_concentration), dose=pick(trial.dose),
                            dosing_interval=pick(trial.dosing_interval), dose_after_dose_change=pick(trial.dose_after_dose_change),
                            dose_change_time=pick(trial.dose_change_time), intermittent=pick(trial.intermittent),
                            treatment_interruptions=pick(trial.treatment_interruptions))
    vals = ph.make_pharmaco_single_values(one, 5, seed=3)
    ev = ph.PharmacoEvaluator(one)
    want, _ = ev.evaluate(vals)
    ev.close()
    prior = "<variableset>" + "".join(f'<variable name="{n}" logspace="true" distribution="uniform" lower="-5" upper="5"/>' for n in one.variable_names) + "</variableset>"
    lik = ('<bcm_likelihood type="pharmaco_single"><pk_model drug="lapatinib" trial="synthetic" patient="2" peripheral_compartment="true" '
           'biphasic_absorption="true" metabolite="true"/></bcm_likelihood>')
    for batched in (True, False):
        assert np.array_equal(host_api.pharmaco_evaluate(prior, lik, trial, vals, batched=batched), want)
    with pytest.raises(RuntimeError, match="Cannot find patient"):
        host_api.pharmaco_evaluate(prior, lik.replace('patient="2"', 'patient="9"'), trial, vals)
