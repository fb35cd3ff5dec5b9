"""CPU suite: the C++ host mirror (bcm3_b200/host) on BASELINE config 1 -- examples/banana, parallel-tempered MCMC with an
analytic likelihood: the batched-evaluation plumbing must reproduce the serial sampler exactly for a fixed seed, and the
posterior must match the analytic banana (TestLikelihoodBanana.cpp:42-55)."""
import dataclasses

import numpy as np
import pytest
from scipy import stats

# the reference's examples/banana/{prior.xml,likelihood.xml,config.txt}; CONFIG swaps the proposal for global_covariance,
# GMM_CONFIG is the example's own (proposal_type=gaussian_mixture)
PRIOR = """<?xml version="1.0" encoding="utf-8"?>
<variableset>
  <variable name="x1"   distribution="uniform" lower="-6.0" upper="4.0"/>
  <variable name="x2"   distribution="uniform" lower="-6.0" upper="20.0"/>
</variableset>"""
LIKELIHOOD = """<bcm_likelihood type="banana" dimension="2" sd1="2.0" sd2="1.0">
</bcm_likelihood>"""
CONFIG = """[sampler]
num_samples=8000
use_every_nth=5

[ptmhsampler]
num_chains=6
swapping_scheme=deterministic_even_odd
num_exploration_steps=1
max_history_size=5000
proposal_type=global_covariance
adapt_proposal_times=1
adapt_proposal_samples=2000
adapt_proposal_max_history_samples=5000
stop_proposal_scaling=4000
temperature_schedule_power=3.0
"""
GMM_CONFIG = CONFIG.replace("proposal_type=global_covariance", "proposal_type=gaussian_mixture")


@pytest.fixture(scope="module")
def host(built):
    from bcm3_b200 import host_api

    host_api.load()
    return host_api


def test_variable_set_from_prior_xml(host):
    xml = """<prior>
      <variable name="mean_absorption" distribution="normal" mu="0" sigma="1"/>
      <variable name="excretion" logspace="true" distribution="uniform" lower="-3" upper="1"/>
      <variable name="patient" repeat="3" distribution="uniform" lower="0" upper="1"/>
      <variable name="frac" logistic="true" distribution="normal" mu="0" sigma="2"/>
    </prior>"""
    n, transforms, idx = host.varset_info(xml, "patient_2")
    assert n == 6 and transforms == [0, 2, 0, 0, 0, 3] and idx == 4
    assert host.varset_info(xml, "nope")[2] == np.iinfo(np.uint64).max  # VariableSet.cpp:84-95


def test_factory_and_default_batched_entry(host):
    rng = np.random.default_rng(0)
    vals = np.column_stack([rng.uniform(-6, 4, 50), rng.uniform(-6, 20, 50)])
    serial = host.evaluate(PRIOR, LIKELIHOOD, vals, batched=False)
    batch = host.evaluate(PRIOR, LIKELIHOOD, vals, batched=True)
    assert np.array_equal(serial, batch)
    want = stats.norm.logpdf(vals[:, 0], 0, 2.0) + stats.norm.logpdf(vals[:, 1], (1 + vals[:, 0]) ** 2, 1.0)
    assert np.allclose(serial, want, rtol=1e-12, atol=1e-12)  # App. D #15: reference PdfNormal is itself only ~1e-14 accurate
    with pytest.raises(RuntimeError, match="Unknown likelihood type"):
        host.evaluate(PRIOR, '<bcm_likelihood type="nope"/>', vals, batched=True)


def test_batched_run_equals_serial_run(host):
    cfg = CONFIG.replace("num_samples=8000", "num_samples=600").replace("adapt_proposal_samples=2000", "adapt_proposal_samples=200")
    a, sa = host.run_pt(PRIOR, LIKELIHOOD, cfg, batched=True, seed=7)
    b, sb = host.run_pt(PRIOR, LIKELIHOOD, cfg, batched=False, seed=7)
    assert a.shape == b.shape == (600 * 6, 5)
    assert np.array_equal(a, b)  # same proposals, same decisions, same samples
    assert sa["evaluations"] == sb["evaluations"]
    assert sb["batched_calls"] == 0 and sa["batched_calls"] >= 600 * 5  # one batched call per mutate round
    c, _ = host.run_pt(PRIOR, LIKELIHOOD, cfg, batched=True, seed=8)
    assert not np.array_equal(a, c)


def test_banana_posterior(host):
    rows, st = host.run_pt(PRIOR, LIKELIHOOD, CONFIG, batched=True, seed=20261018)
    assert st["chains"] == 6
    temps = np.unique(rows[:, 0])
    assert len(temps) == 6 and temps[0] == 0.0 and temps[-1] == 1.0
    assert np.allclose(temps[1:-1], [(i / 5) ** 3 for i in range(1, 5)])  # SamplerPT.cpp:83-93
    post = rows[rows[:, 0] == 1.0][1000:, 3:]  # discard burn-in
    x1, x2 = post[:, 0], post[:, 1]
    # analytic marginal of x1 under the uniform prior box: N(0, 2) truncated to [-6, 4], further weighted by the mass of
    # x2 | x1 ~ N((1 + x1)^2, 1) inside [-6, 20]
    grid = np.linspace(-6, 4, 4001)
    w = stats.norm.pdf(grid, 0, 2) * (stats.norm.cdf(20, (1 + grid) ** 2, 1) - stats.norm.cdf(-6, (1 + grid) ** 2, 1))
    w /= np.trapezoid(w, grid)
    m1 = np.trapezoid(w * grid, grid)
    s1 = np.sqrt(np.trapezoid(w * (grid - m1) ** 2, grid))
    assert abs(x1.mean() - m1) < 0.15 and abs(x1.std() - s1) < 0.15
    # x2 - (1 + x1)^2 is (almost) standard normal
    r = x2 - (1 + x1) ** 2
    assert abs(r.mean()) < 0.1 and abs(r.std() - 1.0) < 0.1
    # the T = 0 chain samples the prior
    pri = rows[rows[:, 0] == 0.0][:, 3:]
    assert abs(pri[:, 0].mean() - (-1.0)) < 0.15 and abs(pri[:, 1].mean() - 7.0) < 0.4


def test_cell_population_plugin_parses_the_reference_xml_surface(built, tmp_path, monkeypatch):
    """CellPopulationLikelihoodB200 behind LikelihoodFactory type="cell_population": Initialize reads the reference's
    likelihood.xml attributes, PostInitialize hands descriptor + arrays + generated text to the C ABI (compile-only here)."""
    from bcm3_b200 import host_api
    from bcm3_b200 import synthetic_cellpop as sc
    from tests.util import cellpop_xml

    monkeypatch.setenv("BCM3B200_CACHE", str(tmp_path))
    prob = sc.make_cellpop_problem(N=5, num_cells=16, T=6, data_cells=2, seed=3)
    prior, lik, species = cellpop_xml(prob)
    _, desc = host_api.cellpop_evaluate(prior, lik, prob, species, compile_only=True)
    kv = dict(item.split("=", 1) for item in desc.split(";"))
    assert kv["num_species"] == "5" and kv["num_cells"] == "16" and kv["num_timepoints"] == "6" and kv["variability_dim"] == "3"
    assert kv["stdev_ix"] == "5" and kv["obs_species"] == "4" and kv["error_model"] == "normal" and float(kv["entry_time"]) == 0.0
    assert kv["variability_distribution"] == "diagonal_gaussian" and "proportional_stdev" not in kv
    # two diagonal_gaussian blocks = one block with all their variables (successive quasi-random dimensions, applied in order)
    split = lik.replace('negate="true"/>', 'negate="true" only_initial_cells="true"/></cell_variability><cell_variability distribution="diagonal_gaussian">')
    assert split.count("<cell_variability") == 2
    assert host_api.cellpop_evaluate(prior, split, prob, species, compile_only=True)[1] == desc
    with pytest.raises(RuntimeError, match="all of them are diagonal_gaussian"):
        host_api.cellpop_evaluate(prior, split.replace('<cell_variability distribution="diagonal_gaussian">', '<cell_variability distribution="full_gaussian" covar_base_name="c">', 1),
                                  prob, species, compile_only=True)
    rel = host_api.cellpop_evaluate(prior, lik.replace('stdev="stdev"', 'stdev="stdev" stdev_relative_to_scale="true" optimize_offset_scale="false"'),
                                    prob, species, compile_only=True)[1]
    assert "stdev_relative_to_scale=1" in rel and "stdev_relative_to_scale=0" in desc
    with pytest.raises(RuntimeError, match="use_log_ratio"):  # a population average knows no log ratio
        host_api.cellpop_evaluate(prior, lik.replace('stdev="stdev"', 'stdev="stdev" use_log_ratio="true"'), prob, species, compile_only=True)
    # include_only_cells_that_went_through_mitosis needs the species Cell::Cell looks up by name (Cell.cpp:40-55)
    with pytest.raises(RuntimeError, match="nuclear_envelope"):
        host_api.cellpop_evaluate(prior, lik.replace('stdev="stdev"', 'stdev="stdev" include_only_cells_that_went_through_mitosis="true"'), prob, species, compile_only=True)
    mitotic = host_api.cellpop_evaluate(prior, lik.replace('stdev="stdev"', 'stdev="stdev" include_only_cells_that_went_through_mitosis="true"'), prob,
                                        [("nuclear_envelope" if s == "x3" else s) for s in species], compile_only=True)[1]
    assert "include_only_cells_that_went_through_mitosis=1" in mitotic and "nuclear_envelope_species=3" in mitotic
    # trailing_simulation_time (Experiment.cpp:489, 655-656) moves the end of the integration past the last timepoint
    trail = host_api.cellpop_evaluate(*cellpop_xml(prob, trailing_simulation_time="2.5")[:2], prob, species, compile_only=True)[1]
    assert float(dict(i.split("=", 1) for i in trail.split(";"))["simulation_end_time"]) == float(prob.timepoints[-1]) + 2.5
    assert "solver_max_timestep=0.5" in host_api.cellpop_evaluate(*cellpop_xml(prob, solver_max_timestep="0.5")[:2], prob, species, compile_only=True)[1]
    for attrs, message in ((dict(synchronization_time_offset="3"), "synchronization_time_offset"),):
        with pytest.raises(RuntimeError, match=message):
            host_api.cellpop_evaluate(*cellpop_xml(prob, **attrs)[:2], prob, species, compile_only=True)
    # a species named "apoptosis" ends a cell's life when it passes 1 (Cell.cpp:516-534): handed on by index
    renamed = [("apoptosis" if s == "x2" else s) for s in species]
    assert "apoptosis_species=2" in host_api.cellpop_evaluate(prior, lik, prob, renamed, compile_only=True)[1]
    with pytest.raises(RuntimeError, match="simulate_past_chromatid_separation_time"):
        host_api.cellpop_evaluate(*cellpop_xml(prob, simulate_past_chromatid_separation_time="3")[:2], prob, species, compile_only=True)
    with pytest.raises(RuntimeError, match="proportional stdev has not been specified"):
        host_api.cellpop_evaluate(prior, lik.replace('stdev="stdev"', 'stdev="stdev" error_model="proportional_normal"'), prob, species, compile_only=True)
    # an entry_time variable occupies a quasi-random dimension (kind 2 of the variability rows) and needs a wider table
    with_entry = lik.replace("</cell_variability>", '<variable entry_time="true" apply="additive" scale="0.3"/></cell_variability>')
    with pytest.raises(RuntimeError, match="SetSobolTable"):
        host_api.cellpop_evaluate(prior, with_entry, prob, species, compile_only=True)
    wider = dataclasses.replace(prob, sobol=np.concatenate([prob.sobol, prob.sobol[:, :1]], axis=1))
    assert "variability_dim=4" in host_api.cellpop_evaluate(prior, with_entry, wider, species, compile_only=True)[1]
    # the reference's default is dividing cells (Experiment.cpp:488); a model without a "cytokinesis" species never divides
    # (Cell.cpp:499), so the default and divide_cells="false" describe the same evaluator ...
    assert host_api.cellpop_evaluate(prior, lik.replace(' divide_cells="false"', ""), prob, species, compile_only=True)[1] == desc
    # ... and one with "cytokinesis" needs all seven species a daughter resets (Cell.cpp:127-133 looks them up unchecked)
    with pytest.raises(RuntimeError, match="nuclear_envelope"):
        host_api.cellpop_evaluate(prior, lik.replace(' divide_cells="false"', ""), prob, [("cytokinesis" if s == "x2" else s) for s in species], compile_only=True)
    with pytest.raises(RuntimeError, match="not supported"):
        host_api.cellpop_evaluate(prior, lik.replace("time_course_population_average", "duration"), prob, species, compile_only=True)
    # the per-cell types are taken when they do not need stored integration points ...
    with pytest.raises(RuntimeError, match="stored integration points"):
        host_api.cellpop_evaluate(prior, lik.replace('type="time_course_population_average"', 'type="time_course" synchronize="mitosis"'), prob, species, compile_only=True)
    with pytest.raises(RuntimeError, match="stored integration points"):
        host_api.cellpop_evaluate(prior, lik.replace('type="time_course_population_average"', 'type="time_points" synchronize="anaphase"'), prob, species, compile_only=True)
    with pytest.raises(RuntimeError, match="optimize_offset_scale"):
        host_api.cellpop_evaluate(prior, lik.replace('type="time_course_population_average"', 'type="time_course" optimize_offset_scale="true" error_model="proportional_normal" proportional_stdev="0.1"'),
                                  prob, species, compile_only=True)
    with pytest.raises(RuntimeError, match="saturation_scale must name a variable"):  # a number: unusable in the reference (overwritten with DBL_MAX)
        host_api.cellpop_evaluate(prior, lik.replace('type="time_course_population_average"', 'type="time_course" saturation_scale="2.0"'), prob, species, compile_only=True)
    # ... and a time_course data set needs one observed trajectory per simulated cell (here: 1 row of data for 16 cells)
    with pytest.raises(RuntimeError, match="num_cells"):
        host_api.cellpop_evaluate(prior, lik.replace("time_course_population_average", "time_course"), prob, species, compile_only=True)
    tp = host_api.cellpop_evaluate(prior, lik.replace('type="time_course_population_average"', 'type="time_points" value_relative_to_timepoint_ix="2"'), prob, species, compile_only=True)[1]
    assert "data_kind=time_points" in tp and "value_relative_to_timepoint_ix=2" in tp
    with pytest.raises(RuntimeError, match="Could not find variable"):
        host_api.cellpop_evaluate(prior, lik.replace('stdev="stdev"', 'stdev="no_such_variable"'), prob, species, compile_only=True)


def test_cell_population_plugin_takes_several_experiments_and_data_sets(built, tmp_path, monkeypatch):
    """One handle of the C ABI per experiment, shared by its data sets (the default), or one per <data> element; either way the
    data sets of an experiment carry the experiment's common simulation end (the last time any of them requests,
    Experiment.cpp:190-214, 655-656) so that each reproduces the reference's one simulation of that experiment."""
    from tests.util import cellpop_two_experiment_setup, open_cellpop_session

    monkeypatch.setenv("BCM3B200_CACHE", str(tmp_path))
    prior, lik, species, problems = cellpop_two_experiment_setup()
    # shared integration: the first data set's handle carries the second one as data set @1
    s = open_cellpop_session(prior, lik, species, problems)
    s.post_initialize(compile_only=True)
    shared = dict(item.split("=", 1) for item in s.descriptor(0, 0).split(";"))
    assert shared["num_data_sets"] == "2" and shared["num_timepoints@1"] == "6" and shared["obs_species@1"] == "2+3" and shared["error_model@1"] == "student_t4"
    assert float(shared["stdev@1"]) == 0.3 and float(shared["weight@1"]) == 0.5 and shared["stdev_ix"] == "5" and "simulation_end_time" not in shared
    assert s.descriptor(0, 1) == "" and "num_data_sets" not in s.descriptor(1, 0)
    s.close()
    s = open_cellpop_session(prior, lik, species, problems)
    s.share_integration(False)
    s.post_initialize(compile_only=True)
    kv = lambda e, d: dict(item.split("=", 1) for item in s.descriptor(e, d).split(";"))
    first, early, second = kv(0, 0), kv(0, 1), kv(1, 0)
    assert "simulation_end_time" not in first and "simulation_end_time" not in second
    assert float(early["simulation_end_time"]) == float(problems[0][0].timepoints[-1]) > float(problems[0][1].timepoints[-1])
    assert early["num_timepoints"] == "6" and early["obs_species"] == "2+3" and early["error_model"] == "student_t4"
    assert float(early["stdev"]) == 0.3 and float(early["weight"]) == 0.5 and first["stdev_ix"] == "5"
    assert first["num_cells"] == "96" and second["num_cells"] == "64" and float(second["entry_time"]) == problems[1][0].entry_time
    s.close()
    # a data set that was never supplied is an error of PostInitialize, not a crash
    s = open_cellpop_session(prior, lik, species, [[problems[0][0], problems[0][1]], problems[1]])
    s.set_data(0, 1, [], np.zeros((1, 0)))
    with pytest.raises(RuntimeError, match="consistent data set"):
        s.post_initialize(compile_only=True)
    s.close()


def test_cell_population_plugin_experiment_specific_elements(built, tmp_path, monkeypatch):
    """<experiment_specific_parameter>, <set_parameter>, <set_species> of Experiment::Load (Experiment.cpp:495-527)."""
    from bcm3_b200 import host_api
    from tests.util import cellpop_two_experiment_setup, open_cellpop_session

    monkeypatch.setenv("BCM3B200_CACHE", str(tmp_path))
    prior, lik, species, problems = cellpop_two_experiment_setup()
    prior = prior.replace("</variableset>", '<variable name="k_in_second" logspace="true" distribution="uniform" lower="-5" upper="5"/>'
                          '<variable name="shift" distribution="uniform" lower="-5" upper="5"/></variableset>')
    head = '<experiment name="second"'
    at = lik.index(">", lik.index(head)) + 1
    extra = ('<experiment_specific_parameter parameter_name="k_in" replacement_parameter_name="k_in_second"/>'
             '<set_parameter parameter_name="k_leak" value="0.125"/><set_species species_name="x3" value="2.0"/>')
    s = open_cellpop_session(prior, lik[:at] + extra + lik[at:], species, problems)
    assert s.fixed_parameters(0) == [] and s.fixed_parameters(1) == [("k_leak", 0.125)]
    s.post_initialize(compile_only=True)
    s.close()
    bad = lambda text: host_api.CellPopSession(prior, lik[:at] + text + lik[at:])
    with pytest.raises(RuntimeError, match="share a variable transform"):
        bad('<experiment_specific_parameter parameter_name="k_in" replacement_parameter_name="shift"/>')
    with pytest.raises(RuntimeError, match="is not a variable"):
        bad('<experiment_specific_parameter parameter_name="k_in" replacement_parameter_name="nope"/>')
    with pytest.raises(RuntimeError, match="a data set of the experiment reads"):
        bad('<experiment_specific_parameter parameter_name="stdev" replacement_parameter_name="k_in_second"/>')
    s = open_cellpop_session(prior, lik[:at] + '<set_species species_name="not_a_species" value="1"/>' + lik[at:], species, problems)
    with pytest.raises(RuntimeError, match="not a simulated species"):
        s.post_initialize(compile_only=True)
    s.close()


def test_symmetric_eigen_and_mixture_density(host):
    rng = np.random.default_rng(3)
    a = rng.standard_normal((7, 7))
    a = a @ a.T + 0.1 * np.eye(7)
    vals, vecs = host.symmetric_eigen(a)
    assert np.all(np.diff(vals) >= 0) and np.allclose(vals, np.linalg.eigvalsh(a), rtol=1e-12)
    assert np.allclose(vecs @ np.diag(vals) @ vecs.T, a, atol=1e-12) and np.allclose(vecs.T @ vecs, np.eye(7), atol=1e-12)
    w = np.array([0.3, 0.7])
    mu = np.array([[0.0, 1.0, -1.0], [2.0, -1.0, 0.5]])
    cov = np.stack([np.diag([1.0, 0.5, 2.0]), a[:3, :3]])
    x = rng.standard_normal((20, 3)) * 2
    lp, resp = host.gmm_evaluate(w, mu, cov, x)
    comp = np.stack([np.log(w[k]) + stats.multivariate_normal(mu[k], cov[k]).logpdf(x) for k in range(2)], axis=1)
    assert np.allclose(lp, np.logaddexp(comp[:, 0], comp[:, 1]), rtol=1e-12, atol=1e-12)
    assert np.allclose(resp, np.exp(comp - lp[:, None]), atol=1e-12) and np.allclose(resp.sum(axis=1), 1.0)


def test_mixture_fit_recovers_two_clusters(host):
    """GMM::Fit (GMM.cpp:48-158): k-means++ start + EM with the shrunk covariance estimate; AIC prefers the true component count."""
    rng = np.random.default_rng(11)
    c0 = rng.multivariate_normal([-3.0, 0.0, 1.0], np.diag([0.5, 1.0, 0.2]), 1200)
    c1 = rng.multivariate_normal([2.0, 2.0, -1.0], [[1.0, 0.6, 0.0], [0.6, 1.0, 0.0], [0.0, 0.0, 0.3]], 800)
    x = rng.permutation(np.concatenate([c0, c1]))
    one, two, three = (host.gmm_fit(x, k, seed=5) for k in (1, 2, 3))
    assert np.allclose(one["means"][0], x.mean(axis=0), atol=1e-9)
    # one component with ess_factor 1: the sample covariance up to the O(D / n) eigenvalue shrinkage
    assert np.allclose(one["covariances"][0], np.cov(x.T), rtol=0.01, atol=0.01)
    order = np.argsort(two["means"][:, 0])
    assert np.allclose(two["weights"][order], [0.6, 0.4], atol=0.02)
    assert np.allclose(two["means"][order], [[-3.0, 0.0, 1.0], [2.0, 2.0, -1.0]], atol=0.12)
    assert abs(two["covariances"][order[1]][0, 1] - 0.6) < 0.1
    assert two["aic"] < one["aic"] - 500 and two["aic"] < three["aic"] + 30 and two["logl"] > one["logl"]
    lp, _ = host.gmm_evaluate(two["weights"], two["means"], two["covariances"], x)
    assert np.isclose(lp.sum(), two["logl"], rtol=1e-3)  # the reported log-likelihood is the last expectation step's
    assert host.gmm_fit(x[:10], 2) is None  # fewer than 2 D K samples (GMM.cpp:86-90)
    assert host.gmm_fit(x, 2, seed=5)["aic"] == two["aic"]  # deterministic given the seed


def test_gaussian_mixture_proposal_on_the_banana(host):
    """BASELINE config 1 as the reference ships it: examples/banana/config.txt with proposal_type=gaussian_mixture."""
    cfg = GMM_CONFIG.replace("num_samples=8000", "num_samples=600").replace("adapt_proposal_samples=2000", "adapt_proposal_samples=200")
    a, sa = host.run_pt(PRIOR, LIKELIHOOD, cfg, batched=True, seed=7)
    b, sb = host.run_pt(PRIOR, LIKELIHOOD, cfg, batched=False, seed=7)
    assert np.array_equal(a, b) and sa["evaluations"] == sb["evaluations"]
    g, _ = host.run_pt(PRIOR, LIKELIHOOD, cfg.replace("gaussian_mixture", "global_covariance"), batched=True, seed=7)
    assert not np.array_equal(a, g)
    with pytest.raises(RuntimeError, match="proposal_type"):
        host.run_pt(PRIOR, LIKELIHOOD, cfg.replace("gaussian_mixture", "clustered_covariance"), seed=7)
    rows, st = host.run_pt(PRIOR, LIKELIHOOD, GMM_CONFIG, batched=True, seed=20261018)
    post = rows[rows[:, 0] == 1.0][1000:, 3:]
    x1, x2 = post[:, 0], post[:, 1]
    grid = np.linspace(-6, 4, 4001)
    w = stats.norm.pdf(grid, 0, 2) * (stats.norm.cdf(20, (1 + grid) ** 2, 1) - stats.norm.cdf(-6, (1 + grid) ** 2, 1))
    w /= np.trapezoid(w, grid)
    m1 = np.trapezoid(w * grid, grid)
    s1 = np.sqrt(np.trapezoid(w * (grid - m1) ** 2, grid))
    assert abs(x1.mean() - m1) < 0.15 and abs(x1.std() - s1) < 0.15
    r = x2 - (1 + x1) ** 2
    assert abs(r.mean()) < 0.1 and abs(r.std() - 1.0) < 0.1
    adj, _ = host.run_pt(PRIOR, LIKELIHOOD, cfg.replace("gaussian_mixture", "gaussian_mixture_adjustedAIC"), batched=True, seed=7)
    assert adj.shape == a.shape and np.isfinite(adj[:, 2]).all()


def test_t_distributed_proposals_on_the_banana(host):
    """ptmhsampler.proposal_t_dof > 0 (SamplerPT.cpp:63,169; ProposalGlobalCovariance.cpp:23-29): every step is scaled by
    1 / sqrt(w), w a gamma draw -- heavier-tailed proposals, the same posterior. Batched and serial runs stay identical."""
    cfg = CONFIG.replace("num_samples=8000", "num_samples=4000").replace("[ptmhsampler]", "[ptmhsampler]\nproposal_t_dof=4")
    rows, _ = host.run_pt(PRIOR, LIKELIHOOD, cfg, batched=True, seed=11)
    plain, _ = host.run_pt(PRIOR, LIKELIHOOD, cfg.replace("proposal_t_dof=4", "proposal_t_dof=0"), batched=True, seed=11)
    assert not np.array_equal(rows, plain)  # the option is read and changes the proposals
    serial, _ = host.run_pt(PRIOR, LIKELIHOOD, cfg, batched=False, seed=11)
    assert np.array_equal(rows, serial)
    post = rows[rows[:, 0] == 1.0][800:, 3:]
    r = post[:, 1] - (1 + post[:, 0]) ** 2
    assert abs(r.mean()) < 0.15 and abs(r.std() - 1.0) < 0.15


def test_complete_linkage_clustering_against_scipy(host):
    """TreeClusterCompleteLinkage (bcm3::TreeCluster around the C Clustering Library's pairwise-maximum-linkage tree) against
    scipy's complete linkage cut at the same height: the same partition."""
    from scipy.cluster.hierarchy import fcluster, linkage
    from scipy.spatial.distance import squareform

    rng = np.random.default_rng(5)
    for n in (2, 3, 7, 12, 25):
        for trial in range(4):
            x = rng.normal(size=(n, 3))
            d = np.abs(x[:, None, :] - x[None, :, :]).sum(-1)
            d = d / d.max()
            got = host.tree_cluster(d, 0.5)
            want = fcluster(linkage(squareform(d, checks=False), method="complete"), t=0.5 - 1e-12, criterion="distance")
            assert got.min() == 0 and len(np.unique(got)) == got.max() + 1
            # same partition: items share a block here iff they share one there
            assert np.array_equal(got[:, None] == got[None, :], want[:, None] == want[None, :])
    # everything further apart than the cut: singletons; everything closer: one block
    assert len(np.unique(host.tree_cluster(np.full((5, 5), 0.9) - 0.9 * np.eye(5), 0.5))) == 5
    assert len(np.unique(host.tree_cluster(np.full((5, 5), 0.1) - 0.1 * np.eye(5), 0.5))) == 1


def test_blocking_strategies_on_the_banana(host):
    """ptmhsampler.blocking_strategy (SamplerPT.cpp:48,151; SamplerPTChain.cpp:66-77,249-307): no_blocking updates one variable
    at a time (two likelihood evaluations per mutate move of a two-variable chain, block b of all chains in one batched call),
    Turek joins variables whose history correlates above 0.5. Batched and serial runs stay identical, the posterior is the
    banana's either way."""
    base = CONFIG.replace("num_samples=8000", "num_samples=3000")
    one, s_one = host.run_pt(PRIOR, LIKELIHOOD, base, batched=True, seed=3)
    assert s_one["blocks"] == 1
    for strategy, blocks in (("no_blocking", 2), ("Turek", None)):
        cfg = base.replace("[ptmhsampler]", f"[ptmhsampler]\nblocking_strategy={strategy}")
        rows, st = host.run_pt(PRIOR, LIKELIHOOD, cfg, batched=True, seed=3)
        serial, st_serial = host.run_pt(PRIOR, LIKELIHOOD, cfg, batched=False, seed=3)
        assert np.array_equal(rows, serial) and st["evaluations"] == st_serial["evaluations"]
        if blocks is not None:
            assert st["blocks"] == blocks
            # 5 tempered chains x 2 blocks + the prior chain's one draw per mutate move, against 6 evaluations with one block
            assert st["evaluations"] > 1.7 * s_one["evaluations"] and st["batched_calls"] > 1.9 * s_one["batched_calls"]
        else:
            assert st["blocks"] in (1, 2)
        post = rows[rows[:, 0] == 1.0][600:, 3:]
        r = post[:, 1] - (1 + post[:, 0]) ** 2
        assert abs(r.mean()) < 0.2 and abs(r.std() - 1.0) < 0.2
    with pytest.raises(RuntimeError, match="clustered_autoblock"):
        host.run_pt(PRIOR, LIKELIHOOD, base.replace("[ptmhsampler]", "[ptmhsampler]\nblocking_strategy=clustered_autoblock"))
    with pytest.raises(RuntimeError, match="Unknown blocking strategy"):
        host.run_pt(PRIOR, LIKELIHOOD, base.replace("[ptmhsampler]", "[ptmhsampler]\nblocking_strategy=nope"))


def test_sample_handlers_tsv_and_max_a_posteriori(host, tmp_path):
    """The reference's sample sinks on the sampler mirror (Sampler::AddSampleHandler): SampleHandlerTSV writes the posterior chain's
    samples in the reference's text format -- "%.6g", tab-separated, and its line structure: the weight on a line of its own
    (SampleHandlerTSV.cpp:45-47) -- and SampleHandlerStoreMaxAPosteriori keeps the best log-posterior over ALL temperatures."""
    cfg = CONFIG.replace("num_samples=8000", "num_samples=300").replace("adapt_proposal_samples=2000", "adapt_proposal_samples=100")
    path = tmp_path / "samples.tsv"
    rows, stats, best = host.run_pt_with_handlers(PRIOR, LIKELIHOOD, cfg, tsv_file=str(path), seed=11)
    plain, _ = host.run_pt(PRIOR, LIKELIHOOD, cfg, seed=11)
    assert np.array_equal(rows, plain)  # the handlers only listen
    lines = path.read_text().split("\n")
    assert lines[0] == "x1\tx2\tlog prior\tlog likelihood\tweight"
    body = [l for l in lines[1:] if l != ""]
    posterior = rows[rows[:, 0] == 1.0]  # [temperature, lprior, llh, values...]
    assert len(posterior) == 300 and len(body) == 2 * len(posterior)
    for r, (sample_line, weight_line) in zip(posterior, zip(body[0::2], body[1::2])):
        want = "\t".join("%.6g" % v for v in (r[3], r[4], r[1], r[2]))
        assert sample_line == want
        assert weight_line == "1"
    # maximum a posteriori over every emitted sample, the heated chains included
    lpost = rows[:, 1] + rows[:, 2]
    k = int(np.argmax(lpost))
    assert best["lposterior"] == lpost[k] and best["llikelihood"] == rows[k, 2]
    assert np.array_equal(best["values"], rows[k, 3:])
    with pytest.raises(RuntimeError, match="Failed to open output file"):
        host.run_pt_with_handlers(PRIOR, LIKELIHOOD, cfg, tsv_file=str(tmp_path / "no_such_dir" / "x.tsv"), seed=11)
