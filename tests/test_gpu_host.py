"""GPU suite: the C++ host mirror driving the GPU-backed pop_pk_trajectory likelihood through the reference's plugin
surface (LikelihoodFactory type string, likelihood.xml, prior.xml, config.txt): a parallel-tempered run whose every
mutate round is one batched call must equal the run that evaluates chain by chain."""
import math

import numpy as np
import pytest

from bcm3_b200 import synthetic as syn
from bcm3_b200.poppk_data import PK_ONE, PK_TWO

pytestmark = pytest.mark.gpu


def poppk_prior_xml(pk, P):
    two = pk == PK_TWO
    lines = ['<?xml version="1.0" encoding="utf-8"?>', "<prior>",
             '<variable name="mean_absorption" distribution="normal" mu="0.1" sigma="0.2"/>',
             '<variable name="excretion" logspace="true" distribution="uniform" lower="-2.0" upper="-0.5"/>',
             '<variable name="mean_clearance" distribution="normal" mu="0.9" sigma="0.2"/>',
             '<variable name="volume_of_distribution" logspace="true" distribution="uniform" lower="1.2" upper="2.2"/>']
    if two:
        lines += ['<variable name="k_periphery_fwd" logspace="true" distribution="uniform" lower="-1.5" upper="0.5"/>',
                  '<variable name="k_periphery_bwd" logspace="true" distribution="uniform" lower="-2.0" upper="0.0"/>']
    lines += ['<variable name="sigma_absorption" distribution="uniform" lower="0.1" upper="0.6"/>',
              '<variable name="sigma_clearance" distribution="uniform" lower="0.1" upper="0.6"/>',
              f'<variable name="patient" repeat="{2 * P}" distribution="uniform" lower="0.02" upper="0.98"/>',
              '<variable name="standard_deviation" logspace="true" distribution="uniform" lower="-1.0" upper="1.0"/>',
              '<variable name="proportional_standard_deviation" logspace="true" distribution="uniform" lower="-1.5" upper="0.0"/>',
              "</prior>"]
    return "\n".join(lines)


CONFIG = """[sampler]
num_samples=40
use_every_nth=2
[ptmhsampler]
num_chains=5
proposal_type=global_covariance
adapt_proposal_samples=20
adapt_proposal_times=1
"""


@pytest.mark.parametrize("pk", [PK_ONE, PK_TWO])
def test_pt_run_on_gpu_likelihood_batched_equals_serial(built, pk):
    from bcm3_b200 import host_api

    P = 40
    prob = syn.make_poppk_problem(pk, P=P, T=8, t_end=48.0, seed=3)
    prior = poppk_prior_xml(pk, P)
    n, transforms, sdix = host_api.varset_info(prior, "standard_deviation")
    assert n == prob.num_variables and sdix == prob.sd_ix and transforms == prob.transforms.tolist()
    lik = f'<bcm_likelihood type="pop_pk_trajectory"><pk_model drug="lapatinib" type="{"one" if pk == PK_ONE else "two"}" trial="t" pkdata_file="unused.nc"/></bcm_likelihood>'
    a, sa = host_api.run_pt_poppk(prior, lik, CONFIG, prob.trial, batched=True, seed=11)
    b, sb = host_api.run_pt_poppk(prior, lik, CONFIG, prob.trial, batched=False, seed=11)
    assert a.shape == (40 * 5, n + 3)
    assert np.array_equal(a, b)
    assert sa["batched_calls"] >= 80 and sb["batched_calls"] == 0 and sa["evaluations"] == sb["evaluations"]
    assert np.isfinite(a[:, 2]).all()
    # the posterior chain moved and its log-likelihood improved over the start
    post = a[a[:, 0] == 1.0]
    assert post[-1, 2] >= post[0, 2]


def test_single_patient_plugin_pt_run_and_parity(built):
    """likelihood.xml type="pharmacokinetic_trajectory" (LikelihoodFactory.cpp:60): <pk_model patient=> picks one patient of the
    trial; a parallel-tempered run with the no_blocking strategy (one variable, one batched call per block) equals the serial
    run, and the plugin's log-likelihoods are the direct ABI call's on that patient."""
    from bcm3_b200 import host_api
    from bcm3_b200.poppk import PopPKEvaluator
    from bcm3_b200.poppk_data import PopPKProblem, PopPKTrial

    pop = syn.make_poppk_problem(PK_TWO, P=3, T=10, t_end=96.0, seed=9, heterogeneous=True)
    names = ["absorption", "excretion", "clearance", "volume_of_distribution", "k_periphery_fwd", "k_periphery_bwd", "spare_a", "spare_b",
             "standard_deviation", "proportional_standard_deviation"]
    centres = [0.1, -1.3, 0.9, 1.7, -0.5, -1.0, 0.0, 0.0, 0.0, -0.7]
    prior = "\n".join(['<?xml version="1.0" encoding="utf-8"?>', "<prior>"] +
                      [f'<variable name="{n}" logspace="true" distribution="uniform" lower="{c - 0.5}" upper="{c + 0.5}"/>' for n, c in zip(names, centres)] +
                      ["</prior>"])
    lik = '<bcm_likelihood type="pharmacokinetic_trajectory"><pk_model drug="lapatinib" type="two" trial="t" patient="1"/></bcm_likelihood>'
    cfg = CONFIG.replace("[ptmhsampler]", "[ptmhsampler]\nblocking_strategy=no_blocking")
    a, sa = host_api.run_pt_poppk(prior, lik, cfg, pop.trial, batched=True, seed=5)
    b, sb = host_api.run_pt_poppk(prior, lik, cfg, pop.trial, batched=False, seed=5)
    assert np.array_equal(a, b) and sa["evaluations"] == sb["evaluations"] and sa["blocks"] == 10
    assert sa["batched_calls"] >= 80 * 10 and np.isfinite(a[:, 2]).all()
    # the same patient through the ABI directly
    tr = pop.trial
    one = PopPKTrial(drug=tr.drug, time=tr.time, observed_concentration=tr.observed_concentration[1:2], dose=tr.dose[1:2], dosing_interval=tr.dosing_interval[1:2],
                     dose_after_dose_change=tr.dose_after_dose_change[1:2], dose_change_time=tr.dose_change_time[1:2], intermittent=tr.intermittent[1:2],
                     treatment_interruptions=tr.treatment_interruptions[1:2])
    prob = PopPKProblem(pk_type=PK_TWO, trial=one, transforms=np.full(10, 2, dtype=np.int32), sd_ix=8, single=True)
    post = a[a[:, 0] == 1.0]
    ev = PopPKEvaluator(prob)
    logp, _ = ev.evaluate(post[:, 3:])
    ev.close()
    assert np.array_equal(logp, post[:, 2])
    with pytest.raises(RuntimeError, match="Cannot find patient"):
        host_api.run_pt_poppk(prior, lik.replace('patient="1"', 'patient="7"'), cfg, pop.trial)


def test_cell_population_plugin_matches_the_direct_abi_call(built):
    """likelihood.xml -> LikelihoodFactory -> CellPopulationLikelihoodB200::EvaluateLogProbabilityBatch gives what the ABI
    gives for the same problem, batched and chain by chain."""
    from bcm3_b200 import host_api
    from bcm3_b200 import synthetic_cellpop as sc
    from bcm3_b200.cellpop import CellPopEvaluator
    from tests.util import cellpop_xml

    prob = sc.make_cellpop_problem(N=8, num_cells=96, T=10, data_cells=4, seed=9)
    vals = sc.make_chain_values(3, seed=9)
    ev = CellPopEvaluator(prob)
    want, _ = ev.evaluate(vals)
    ev.close()
    prior, lik, species = cellpop_xml(prob)
    batched, _ = host_api.cellpop_evaluate(prior, lik, prob, species, values=vals, batched=True)
    serial, _ = host_api.cellpop_evaluate(prior, lik, prob, species, values=vals, batched=False)
    assert np.array_equal(batched, want) and np.array_equal(serial, want)


def test_pt_run_on_gpu_cell_population_batched_equals_serial(built):
    """The C++ sampler on the GPU-backed cell_population likelihood: one batched call per mutate round reproduces the
    chain-by-chain run sample for sample (results do not depend on the batch a chain is evaluated in)."""
    from bcm3_b200 import host_api
    from bcm3_b200 import synthetic_cellpop as sc
    from tests.util import cellpop_xml

    prob = sc.make_cellpop_problem(N=8, num_cells=96, T=10, data_cells=4, seed=9)
    prior, lik, species = cellpop_xml(prob)
    prior = prior.replace('lower="-5" upper="5"', 'lower="-1.5" upper="0.5"')  # keep the start-up draws in a sane range
    cfg = CONFIG.replace("num_samples=40", "num_samples=24").replace("num_chains=5", "num_chains=4")
    a, sa = host_api.run_pt_cellpop(prior, lik, cfg, prob, species, batched=True, seed=5)
    b, sb = host_api.run_pt_cellpop(prior, lik, cfg, prob, species, batched=False, seed=5)
    assert a.shape == (24 * 4, prob.num_variables + 3)
    assert np.array_equal(a, b, equal_nan=True)
    assert sa["batched_calls"] >= 48 and sb["batched_calls"] == 0 and sa["evaluations"] == sb["evaluations"]


def test_cell_population_plugin_sums_experiments_and_data_sets(built):
    """Two experiments, the first with two data sets (different species, timepoints, error model): the plugin's result is
    the reference's sum over experiments of the sum over data sets (CellPopulationLikelihood.cpp:82-101, Experiment.cpp:346-355),
    each term checked against the CPU checker run with the experiment's common simulation end."""
    import oracle
    from tests.util import cellpop_two_experiment_setup, open_cellpop_session
    from bcm3_b200 import synthetic_cellpop as sc

    prior, lik, species, problems = cellpop_two_experiment_setup()
    vals = sc.make_chain_values(4, seed=9)
    s = open_cellpop_session(prior, lik, species, problems)
    s.post_initialize()
    batched = s.evaluate(vals, batched=True)
    serial = s.evaluate(vals, batched=False)
    s.close()
    assert np.array_equal(batched, serial)
    chk = oracle.load("ref" if oracle.available("ref") else "port")
    terms = [[chk.cellpop_evaluate(p, vals)["logp"] for p in exp] for exp in problems]
    want = sum((sum(exp[1:], 0.0 + exp[0]) for exp in terms), np.zeros(len(vals)))
    assert np.isfinite(want).all()
    # pure relative error of the sum, bounded by 1e-6 or the reference's own reproducibility on these inputs (summed over the terms)
    from tests.util import parity_tolerance, reference_noise_floor_cellpop
    floors = [reference_noise_floor_cellpop(p, vals) for exp in problems for p in exp]
    floor_abs = sum((f[1] * np.abs(f[0]["logp"]) for f in floors), np.zeros(len(vals))) if all(f is not None for f in floors) else None
    tol = parity_tolerance(None if floor_abs is None else floor_abs / np.abs(want))
    assert np.all(np.abs(batched - want) <= tol * np.abs(want)), (batched, want, tol)
    # the shorter data set integrated only to its own last timepoint is NOT the same number: the end time enters CVODE's
    # initial step (cvHin), which is why the descriptor carries simulation_end_time
    import dataclasses
    alone = chk.cellpop_evaluate(dataclasses.replace(problems[0][1], simulation_end_time=None), vals)["logp"]
    assert not np.array_equal(alone, terms[0][1])


def test_shared_integration_gives_the_same_bits_as_one_handle_per_data_set(built):
    """One integration per experiment shared by its data sets (the kernel interpolates at the union of their timepoints and sums
    every data set's own species) against one handle -- one integration of the same cells -- per <data> element: the same
    accepted steps, the same Nordsieck polynomial at every output time, the same sums: bit-identical log-likelihoods. Half the
    integrator launches for the two-data-set experiment."""
    from bcm3_b200 import synthetic_cellpop as sc
    from tests.util import cellpop_two_experiment_setup, open_cellpop_session

    prior, lik, species, problems = cellpop_two_experiment_setup()
    vals = sc.make_chain_values(4, seed=9)
    out = {}
    for share in (True, False):
        s = open_cellpop_session(prior, lik, species, problems)
        s.share_integration(share)
        s.post_initialize()
        out[share] = s.evaluate(vals, batched=True)
        s.close()
    assert np.isfinite(out[True]).all()
    assert np.array_equal(out[True], out[False])


def test_cell_population_experiment_specific_parameter(built):
    """<experiment_specific_parameter>: the second experiment's cells see k_in_second where the model reads k_in
    (Experiment.cpp:515-527, 640-642) -- the same as evaluating that experiment with the column replaced."""
    from tests.util import cellpop_two_experiment_setup, open_cellpop_session
    from bcm3_b200 import synthetic_cellpop as sc
    from bcm3_b200.cellpop import CellPopEvaluator
    import dataclasses

    prior, lik, species, problems = cellpop_two_experiment_setup()
    prior = prior.replace("</variableset>", '<variable name="k_in_second" logspace="true" distribution="uniform" lower="-5" upper="5"/></variableset>')
    at = lik.index(">", lik.index('<experiment name="second"')) + 1
    lik = lik[:at] + '<experiment_specific_parameter parameter_name="k_in" replacement_parameter_name="k_in_second"/>' + lik[at:]
    base = sc.make_chain_values(3, seed=4)
    vals = np.concatenate([base, base[:, :1] + 0.2], axis=1)  # the extra variable: another k_in
    s = open_cellpop_session(prior, lik, species, problems)
    s.post_initialize()
    got = s.evaluate(vals, batched=True)
    serial = s.evaluate(vals, batched=False)
    s.close()
    assert np.array_equal(got, serial)
    want = np.zeros(len(vals))
    for ei, exp in enumerate(problems):
        v = vals.copy()
        if ei == 1:
            v[:, 0] = v[:, 6]
        term = np.zeros(len(vals))
        for p in exp:
            tr = np.concatenate([p.transforms, p.transforms[:1]])
            ev = CellPopEvaluator(dataclasses.replace(p, transforms=tr))
            term = term + ev.evaluate(v)[0]
            ev.close()
        want = want + term
    assert np.isfinite(want).all() and np.array_equal(got, want)
    # and the replacement matters
    s = open_cellpop_session(prior, lik.replace('<experiment_specific_parameter parameter_name="k_in" replacement_parameter_name="k_in_second"/>', ""),
                             species, problems)
    s.post_initialize()
    assert not np.array_equal(s.evaluate(vals), got)
    s.close()


def test_cell_population_plugin_with_real_generator_output_and_non_sampled_parameters(built, tmp_path):
    """The plugin on the fixture whose model text the reference's own SBML code generator emitted, through the three
    bcm3::Likelihood virtuals of Likelihood.h:18-22: AddNonSampledParameters names `basal`, SetNonSampledParameters replaces
    its value between evaluations (bcmopt/main.cpp:231) -- equal to the direct ABI call on a problem that carries that value --
    and OutputEvaluationStatistics writes its report."""
    import dataclasses

    from bcm3_b200 import host_api
    from bcm3_b200.cellpop import CellPopEvaluator
    from tests.util import sbml_cell_cycle_problem, sbml_cell_cycle_values

    prob = sbml_cell_cycle_problem(num_cells=64, T=10)
    vals = sbml_cell_cycle_values(3)
    species = ["Cdc20", "Cdh1", "CycA", "CycB", "CycD", "CycD2", "CycE", "CycEp27", "E2F", "Emi1", "Rb", "p27", "pRb"]
    names = ["k_syn", "k_deg", "k_act", "k_inh", "variability_scale", "stdev"]
    prior = "<variableset>" + "".join(
        f'<variable name="{n}" {"" if n == "variability_scale" else "logspace=" + chr(34) + "true" + chr(34) + " "}distribution="uniform" lower="-5" upper="5"/>'
        for n in names) + "</variableset>"
    lik = (f'<bcm_likelihood type="cell_population"><experiment name="cycle" model_file="cell_cycle.xml" entry_time="0" num_cells="{prob.num_cells}" '
           f'max_cells="{prob.num_cells}" divide_cells="false"><cell_variability distribution="diagonal_gaussian">'
           '<variable model_parameter="k_syn" apply="multiplicative_log" scale="variability_scale"/>'
           '<variable model_parameter="k_deg" apply="multiplicative_log" scale="variability_scale" negate="true"/>'
           f'<variable initial_condition_species="Rb" apply="additive" scale="{float(np.log(0.02))!r}"/></cell_variability>'
           '<data type="time_course_population_average" data_name="readout" species_name="CycE+CycEp27" stdev="stdev"/>'
           '</experiment></bcm_likelihood>')
    s = host_api.CellPopSession(prior, lik)
    s.add_non_sampled_parameters(["basal"])
    s.set_model_with_non_sampled(prob, species)
    s.set_data(0, 0, prob.timepoints, prob.observed)
    s.set_sobol(0, prob.sobol)
    s.post_initialize()
    for basal in (0.01, 0.05):
        s.set_non_sampled_parameters([basal])
        got = s.evaluate(vals)
        ev = CellPopEvaluator(dataclasses.replace(prob, non_sampled_parameters=np.array([basal])))
        want, _ = ev.evaluate(vals)
        ev.close()
        assert np.array_equal(got, want)
    s.output_evaluation_statistics(str(tmp_path))
    s.close()
    report = (tmp_path / "cellpop_evaluation_statistics.txt").read_text().splitlines()
    assert report[0].split("\t") == ["experiment", "data_set", "evaluations", "kernel_launches"] and report[1].split("\t")[:3] == ["cycle", "0", "6"]


def test_cell_population_plugin_with_dividing_cells(built):
    """likelihood.xml with the reference's default divide_cells="true" and a model that has the cell-cycle species: the plugin
    finds them by name (Cell.cpp:40-55, 127-133), the evaluation equals the direct ABI call on the same problem."""
    from bcm3_b200 import host_api
    from bcm3_b200 import synthetic_cellpop as sc
    from bcm3_b200.cellpop import CellPopEvaluator
    from tests.util import cellpop_xml

    M = 5
    prob = sc.make_dividing_problem(M=M, num_cells=16, max_cells=400, t_end=5.5, T=16)
    vals = sc.make_chain_values(3, seed=5)
    ev = CellPopEvaluator(prob)
    want, _ = ev.evaluate(vals)
    ev.close()
    prior, lik, species = cellpop_xml(prob, max_cells="400")
    lik = lik.replace(' divide_cells="false"', "").replace('model_parameter="k_in"', 'model_parameter="k_cascade"')
    assert f'scale="{math.log(0.01)!r}"/>' in lik
    lik = lik.replace(f'scale="{math.log(0.01)!r}"/>', f'scale="{math.log(0.01)!r}" only_initial_cells="true"/>')
    species = species[:M] + list(sc.DIVISION_SPECIES)
    lik = lik.replace(f'species_name="x{prob.num_species - 1}"', f'species_name="x{M - 1}"')
    got, desc = host_api.cellpop_evaluate(prior, lik, prob, species, values=vals, batched=True)
    assert "divide_cells=1" in desc and f"cytokinesis_species={M}" in desc
    assert np.array_equal(got, want)


def test_cell_population_time_course_and_population_average_in_one_experiment(built):
    """<data type="time_course"> (per-cell trajectories + matching) next to a population average over the same cells: one
    integration, the plugin's result is the sum of the two data sets' terms as the reference adds them (Experiment.cpp:346-355),
    each term from the CPU checker."""
    import dataclasses
    import math
    import oracle
    from bcm3_b200 import synthetic_cellpop as sc
    from tests.util import cellpop_xml, open_cellpop_session, parity_tolerance

    tc = sc.make_time_course_problem(N=8, num_cells=32, T=10, seed=61)
    avg = dataclasses.replace(tc, data_kind="time_course_population_average", observed=np.nanmean(tc.observed, axis=0)[None, :] * 1.02,
                              obs_species=[2, 3], stdev=0.05, error_model="student_t4")
    prior, _, species = cellpop_xml(tc)
    obs = lambda p: "+".join(species[s] for s in p.obs_species)
    lik = ('<bcm_likelihood type="cell_population">'
           f'<experiment name="imaging" model_file="cascade.xml" entry_time="0" num_cells="{tc.num_cells}" max_cells="{tc.num_cells}" divide_cells="false">'
           '<cell_variability distribution="diagonal_gaussian">'
           '<variable model_parameter="k_in" apply="multiplicative_log" scale="variability_scale"/>'
           '<variable model_parameter="k_deg" apply="multiplicative_log" scale="variability_scale" negate="true"/>'
           f'<variable initial_condition_species="x1" apply="additive" scale="{math.log(0.01)!r}"/>'
           '</cell_variability>'
           f'<data type="time_course" data_name="cells" species_name="{obs(tc)}" stdev="{tc.stdev!r}"/>'
           f'<data type="time_course_population_average" data_name="bulk" species_name="{obs(avg)}" stdev="0.05" error_model="student_t4"/>'
           '</experiment></bcm_likelihood>')
    vals = sc.make_chain_values(3, seed=61)
    out = {}
    for share in (True, False):
        s = open_cellpop_session(prior, lik, species, [[tc, avg]])
        s.share_integration(share)
        s.post_initialize()
        out[share] = s.evaluate(vals, batched=True)
        s.close()
    chk = oracle.load("ref" if oracle.available("ref") else "port")
    want = chk.cellpop_evaluate(tc, vals)["logp"] + chk.cellpop_evaluate(avg, vals)["logp"]
    assert np.isfinite(want).all()
    for share in (True, False):
        assert np.all(np.abs(out[share] - want) <= parity_tolerance(None) * np.abs(want)), (share, out[share], want)


def test_cell_population_plugin_time_course_with_offset_scale_optimisation(built):
    """<data type="time_course" optimize_offset_scale="true" ...> through the plugin: observations in arbitrary units, every
    (observed, simulated) pair regressed before it is scored (DataLikelihoodTimeCourseBase.cpp:317-322); against the CPU checker."""
    import dataclasses
    import math
    import oracle
    from bcm3_b200 import synthetic_cellpop as sc
    from tests.util import cellpop_xml, open_cellpop_session, parity_tolerance

    tc = sc.make_time_course_problem(N=8, num_cells=40, T=12, seed=63, missing_fraction=0.05)
    tc = dataclasses.replace(tc, observed=0.1 + 3.0 * tc.observed, optimize_offset_scale=True, optimize_offset_range=(-0.5, 0.5),
                             optimize_scale_range=(0.2, 5.0), stdev=0.1)
    prior, _, species = cellpop_xml(tc)
    obs = "+".join(species[s] for s in tc.obs_species)
    lik = ('<bcm_likelihood type="cell_population">'
           f'<experiment name="imaging" model_file="cascade.xml" entry_time="0" num_cells="{tc.num_cells}" max_cells="{tc.num_cells}" divide_cells="false">'
           '<cell_variability distribution="diagonal_gaussian">'
           '<variable model_parameter="k_in" apply="multiplicative_log" scale="variability_scale"/>'
           '<variable model_parameter="k_deg" apply="multiplicative_log" scale="variability_scale" negate="true"/>'
           f'<variable initial_condition_species="x1" apply="additive" scale="{math.log(0.01)!r}"/>'
           '</cell_variability>'
           f'<data type="time_course" data_name="cells" species_name="{obs}" stdev="0.1" optimize_offset_scale="true" optimize_offset_min="-0.5" '
           'optimize_offset_max="0.5" optimize_scale_min="0.2" optimize_scale_max="5.0"/>'
           '</experiment></bcm_likelihood>')
    vals = sc.make_chain_values(3, seed=63)
    s = open_cellpop_session(prior, lik, species, [[tc]])
    s.post_initialize()
    got = s.evaluate(vals, batched=True)
    s.close()
    chk = oracle.load("ref" if oracle.available("ref") else "port")
    want = chk.cellpop_evaluate(tc, vals)["logp"]
    assert np.isfinite(want).all()
    assert np.all(np.abs(got - want) <= parity_tolerance(None) * np.abs(want)), (got, want)


def test_cell_population_plugin_time_course_with_three_markers(built):
    """species_name="a;b+c;d" with ';'-separated stdev / offset / scale lists: three markers per cell in one likelihood (the plugin
    hands every further marker to the library as a data set that names the first as its owner); against the CPU checker."""
    import math
    import oracle
    from bcm3_b200 import synthetic_cellpop as sc
    from tests.util import cellpop_xml, open_cellpop_session, parity_tolerance
    import dataclasses

    tc = sc.make_time_course_problem(N=8, num_cells=28, T=10, seed=65, missing_fraction=0.08, extra_marker_species=((5, 6), (3,)))
    prior, _, species = cellpop_xml(tc)
    names = lambda sp: "+".join(species[s] for s in sp)
    mk = tc.extra_markers
    species_attr = ";".join([names(tc.obs_species)] + [names(m.obs_species) for m in mk])
    lst = lambda f: ";".join(repr(float(v)) for v in f)
    lik = ('<bcm_likelihood type="cell_population">'
           f'<experiment name="imaging" model_file="cascade.xml" entry_time="0" num_cells="{tc.num_cells}" max_cells="{tc.num_cells}" divide_cells="false">'
           '<cell_variability distribution="diagonal_gaussian">'
           '<variable model_parameter="k_in" apply="multiplicative_log" scale="variability_scale"/>'
           '<variable model_parameter="k_deg" apply="multiplicative_log" scale="variability_scale" negate="true"/>'
           f'<variable initial_condition_species="x1" apply="additive" scale="{math.log(0.01)!r}"/>'
           '</cell_variability>'
           f'<data type="time_course" data_name="cells" species_name="{species_attr}" stdev="{lst([tc.stdev] + [m.stdev for m in mk])}" '
           f'offset="{lst([tc.offset] + [m.offset for m in mk])}" scale="{lst([tc.scale] + [m.scale for m in mk])}" error_model="student_t4"/>'
           '</experiment></bcm_likelihood>')
    tc = dataclasses.replace(tc, error_model="student_t4")
    vals = sc.make_chain_values(3, seed=65)
    from bcm3_b200 import host_api

    s = host_api.CellPopSession(prior, lik)
    assert s.num_data_sets == [3]  # the three markers: every one takes its observed block through set_data
    s.set_model(tc, species)
    s.set_sobol(0, tc.sobol)
    s.set_data(0, 0, tc.timepoints, tc.observed)
    for l, m in enumerate(mk, start=1):
        s.set_data(0, l, tc.timepoints, m.observed)
    s.post_initialize()
    got = s.evaluate(vals, batched=True)
    serial = s.evaluate(vals, batched=False)
    s.close()
    assert np.array_equal(got, serial)
    chk = oracle.load("ref" if oracle.available("ref") else "port")
    want = chk.cellpop_evaluate(tc, vals)["logp"]
    assert np.isfinite(want).all()
    assert np.all(np.abs(got - want) <= parity_tolerance(None) * np.abs(want)), (got, want)


def test_cell_population_plugin_time_course_log_ratio_with_two_markers(built):
    """use_log_ratio="true" species_name="x7/x3;x6/x2": two ratiometric markers per cell (DataLikelihoodTimeCourse.cpp:380-397) -- the
    plugin adds the denominators as value-row blocks of the handle; against the CPU checker."""
    import dataclasses
    import math
    import oracle
    from bcm3_b200 import host_api, synthetic_cellpop as sc
    from bcm3_b200.cellpop_data import Marker
    from tests.util import cellpop_xml, parity_tolerance

    tc = sc.make_time_course_problem(N=8, num_cells=24, T=10, seed=67, log_ratio_denominator=3, noise=0.05)
    obs = tc.observed.copy()
    obs[:, 0] = np.nan  # both species still are 0 at the first timepoint
    rng = np.random.default_rng(67)
    second = Marker(obs_species=[6], observed=np.where(np.isnan(obs), np.nan, 0.3 + 0.1 * rng.standard_normal(obs.shape)), stdev=0.4, log_ratio_denominator=2)
    tc = dataclasses.replace(tc, observed=obs, extra_markers=[second])
    prior, _, species = cellpop_xml(tc)
    lik = ('<bcm_likelihood type="cell_population">'
           f'<experiment name="fret" model_file="cascade.xml" entry_time="0" num_cells="{tc.num_cells}" max_cells="{tc.num_cells}" divide_cells="false">'
           '<cell_variability distribution="diagonal_gaussian">'
           '<variable model_parameter="k_in" apply="multiplicative_log" scale="variability_scale"/>'
           '<variable model_parameter="k_deg" apply="multiplicative_log" scale="variability_scale" negate="true"/>'
           f'<variable initial_condition_species="x1" apply="additive" scale="{math.log(0.01)!r}"/>'
           '</cell_variability>'
           f'<data type="time_course" data_name="ratio" use_log_ratio="true" species_name="x7 / x3; x6/x2" stdev="{tc.stdev!r};0.4"/>'
           '</experiment></bcm_likelihood>')
    vals = sc.make_chain_values(3, seed=67)
    s = host_api.CellPopSession(prior, lik)
    assert s.num_data_sets == [2]
    s.set_model(tc, species)
    s.set_sobol(0, tc.sobol)
    s.set_data(0, 0, tc.timepoints, tc.observed)
    s.set_data(0, 1, tc.timepoints, second.observed)
    s.post_initialize()
    got = s.evaluate(vals, batched=True)
    s.close()
    chk = oracle.load("ref" if oracle.available("ref") else "port")
    want = chk.cellpop_evaluate(tc, vals)["logp"]
    assert np.isfinite(want).all()
    assert np.all(np.abs(got - want) <= parity_tolerance(None) * np.abs(want)), (got, want)
