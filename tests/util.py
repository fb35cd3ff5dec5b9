"""Shared helpers of the test-suite."""
import glob
import os

import numpy as np

from bcm3_b200.poppk_data import PopPKProblem, PopPKTrial

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "poppk_*.npz")))


SINGLE_GOLDEN_NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "pksingle_*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    trial = PopPKTrial(
        drug=str(z["drug"]), time=z["time"], observed_concentration=z["observed_concentration"], dose=z["dose"],
        dosing_interval=z["dosing_interval"], dose_after_dose_change=z["dose_after_dose_change"],
        dose_change_time=z["dose_change_time"], intermittent=z["intermittent"],
        treatment_interruptions=z["treatment_interruptions"])
    named = {k: int(z[k]) for k in ("n_transit_ix", "mean_transit_time_ix", "biphasic_uptake_time_ix", "mean_absorption2_ix") if k in z.files}
    if "single" in z.files:  # pharmacokinetic_trajectory fixtures (tests/golden/make_golden_single.py)
        named.update(single=bool(z["single"]), fixed_vod=float(z["fixed_vod"]), fixed_periphery_fwd=float(z["fixed_periphery_fwd"]),
                     fixed_periphery_bwd=float(z["fixed_periphery_bwd"]))
    prob = PopPKProblem(pk_type=int(z["pk_type"]), trial=trial, transforms=z["transforms"], sd_ix=int(z["sd_ix"]), **named)
    return prob, {k: z[k] for k in ("values", "logp", "conc", "patient_ll", "counters", "noise_floor", "noise_floor_counter_match") if k in z.files}


def rel_err(a, b):
    """Relative error that treats equal infinities / NaN patterns as exact."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    out = np.zeros_like(a)
    fin = np.isfinite(a) & np.isfinite(b)
    out[fin] = np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-300)
    bad = ~fin & ~((a == b) | (np.isnan(a) & np.isnan(b)))
    out[bad] = np.inf
    return out


def counter_match_floor(name):
    """Round-off flips a step-size/order decision in roughly 0.5 % of the systems per 300 steps (the same rate is seen
    between the reference and its own restatement), so long-horizon fixtures get a lower floor."""
    return 0.90 if name.endswith("maxsteps") else 0.97


def assert_matches_golden(got_logp, got_conc, got_counters, gold, logp_tol=1e-6, min_counter_match=0.97):
    """The bar of BASELINE.json: per-chain log-likelihood within 1e-6 relative of the reference's CVODE path.
    Step counters must coincide for (almost) all systems: equal counters mean the same step-size/order decisions."""
    assert rel_err(got_logp, gold["logp"]).max() <= logp_tol
    if got_conc is not None:
        assert (np.isnan(got_conc) == np.isnan(gold["conc"])).all()
    if got_counters is not None:
        same = (np.asarray(got_counters, dtype=np.int64) == gold["counters"]).all(axis=-1)
        assert same.mean() >= min_counter_match, f"only {same.mean():.3f} of the systems reproduce CVODE's counters"


def make_nan_inf_case():
    """A trial whose three chains end in -inf, -inf and NaN under the reference's serial loop
    (LikelihoodPopPKTrajectory.cpp:427-440):
      patient 40: dosing every 2 h over 240 h => exceeds max_steps => patient_logllh = -inf for every chain;
      patient 5 (dose 1e4) / patient 50 (dose 1e5): with a NEGATIVE proportional sd the Student-t scale
      sd + sd2*x turns negative for large concentrations => log(negative) = NaN.
    chain 0: sd2 > 0            -> -inf
    chain 1: sd2 = -0.002       -> NaN only at patient 50, after the -inf patient: never reached -> -inf
    chain 2: sd2 = -0.01        -> NaN at patient 5, before the -inf patient -> NaN (the sampler aborts on it)"""
    from bcm3_b200 import synthetic as syn
    from bcm3_b200.poppk_data import PK_ONE, TRANSFORM_NONE

    prob = syn.make_poppk_problem(PK_ONE, P=61, T=6, t_end=240.0, seed=77)
    tr = prob.trial
    tr.dosing_interval[40] = 2.0
    tr.dose[5] = 1e4
    tr.dose[50] = 1e5
    transforms = prob.transforms.copy()
    transforms[prob.sd_ix + 1] = TRANSFORM_NONE
    prob = PopPKProblem(pk_type=PK_ONE, trial=tr, transforms=transforms, sd_ix=prob.sd_ix)
    vals = syn.make_chain_values(prob, 3, seed=77)
    vals[:, prob.sd_ix] = 3.0
    vals[:, prob.sd_ix + 1] = [0.2, -0.002, -0.01]
    return prob, vals


PHARMACO_GOLDEN_NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "pharmaco_*.npz")))


def load_pharmaco_golden(name):
    from bcm3_b200.pharmaco import PharmacoProblem
    from bcm3_b200.poppk_data import PopPKTrial

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    trial = PopPKTrial(drug=str(z["drug"]), time=z["time"], observed_concentration=z["observed_concentration"], dose=z["dose"],
                       dosing_interval=z["dosing_interval"], dose_after_dose_change=z["dose_after_dose_change"], dose_change_time=z["dose_change_time"],
                       intermittent=z["intermittent"], treatment_interruptions=z["treatment_interruptions"])
    prob = PharmacoProblem(trial=trial, variable_names=[str(n) for n in z["variable_names"]], transforms=z["transforms"],
                           peripheral_compartment=bool(z["peripheral_compartment"]), num_transit_compartments=int(z["num_transit_compartments"]),
                           bioavailability=bool(z["bioavailability"]), **{k: bool(z[k]) for k in ("single", "biphasic_absorption", "metabolite") if k in z.files})
    return prob, {k: z[k] for k in ("values", "logp", "conc", "patient_ll")}


CELLPOP_GOLDEN_NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "cellpop_*.npz")))


def load_cellpop_golden(name):
    from bcm3_b200.cellpop_data import APPLY_TYPES, CellPopProblem, Variability

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    inv_apply = {v: k for k, v in APPLY_TYPES.items()}
    variability = []
    for row in z["variability_rows"]:
        kind_flags, target, apply, scale_ix, scale_fixed, negate = row
        kind, only_initial = int(kind_flags) & 3, bool(int(kind_flags) & 4)
        is_ic = kind == 1
        variability.append(Variability(apply=inv_apply[int(apply)], model_parameter=None if (is_ic or kind == 2) else int(target),
                                       initial_condition_species=int(target) if is_ic else None, entry_time=(kind == 2),
                                       scale_ix=None if scale_ix < 0 else int(scale_ix), scale_fixed=float(scale_fixed), negate=bool(negate),
                                       only_initial_cells=only_initial))
    opt_int = lambda k: int(z[k]) if k in z.files else None
    extra = {}
    if "relative_to_time_average" in z.files:
        extra.update(relative_to_time_average=bool(z["relative_to_time_average"]))
    if "simulation_end_time" in z.files:
        extra.update(simulation_end_time=float(z["simulation_end_time"]))
    if "data_kind" in z.files:
        extra.update(data_kind=str(z["data_kind"]))
    if "optimize_offset_scale" in z.files and bool(z["optimize_offset_scale"]):
        extra.update(optimize_offset_scale=True, optimize_offset_range=tuple(float(v) for v in z["optimize_offset_range"]),
                     optimize_scale_range=tuple(float(v) for v in z["optimize_scale_range"]))
    if "marker_observed" in z.files:
        from bcm3_b200.cellpop_data import Marker

        opt = lambda v: None if v < 0 else int(v)
        extra.update(extra_markers=[
            Marker(obs_species=[int(sp) for sp in z["marker_obs_species"][l] if sp >= 0], observed=z["marker_observed"][l],
                   stdev_ix=opt(q[0]), stdev=float(q[1]), proportional_stdev_ix=opt(q[2]), proportional_stdev=float(q[3]),
                   offset_ix=opt(q[4]), offset=float(q[5]), scale_ix=opt(q[6]), scale=float(q[7]))
            for l, q in enumerate(z["marker_parameters"])])
    if "include_only_cells_that_went_through_mitosis" in z.files and bool(z["include_only_cells_that_went_through_mitosis"]):
        extra.update(include_only_cells_that_went_through_mitosis=True)
    if "nuclear_envelope_species" in z.files:
        extra.update(nuclear_envelope_species=int(z["nuclear_envelope_species"]))
    if "use_only_nondivided" in z.files and bool(z["use_only_nondivided"]):
        extra.update(use_only_nondivided=True)
    if "log_ratio_denominator" in z.files:
        extra.update(log_ratio_denominator=int(z["log_ratio_denominator"]))
    if "saturation_scale_ix" in z.files:
        extra.update(saturation_scale_ix=int(z["saturation_scale_ix"]))
    if "value_relative_to_timepoint_ix" in z.files:
        extra.update(value_relative_to_timepoint_ix=int(z["value_relative_to_timepoint_ix"]))
    if "treatment_species" in z.files:
        extra.update(treatment_species=int(z["treatment_species"]), treatment_times=z["treatment_times"])
    if "divide_cells" in z.files and bool(z["divide_cells"]):
        extra.update(divide_cells=True, max_cells=int(z["max_cells"]), cytokinesis_species=opt_int("cytokinesis_species"),
                     division_reset_species=tuple(int(i) for i in z["division_reset_species"]))
    if "apoptosis_species" in z.files:
        extra.update(apoptosis_species=int(z["apoptosis_species"]))
    if "variability_distribution" in z.files:
        extra.update(variability_distribution=str(z["variability_distribution"]),
                     covariance=[int(ix) if ix >= 0 else float(fx) for ix, fx in z["covariance_rows"]],
                     proportional_stdev_ix=opt_int("proportional_stdev_ix"), proportional_stdev=float(z["proportional_stdev"]))
    prob = CellPopProblem(
        derivative_code=str(z["derivative_code"]), num_species=int(z["num_species"]), initial_conditions=z["initial_conditions"],
        transforms=z["transforms"], num_cells=int(z["num_cells"]), timepoints=z["timepoints"], observed=z["observed"],
        obs_species=[int(s) for s in z["obs_species"]], constant_species=z["constant_species"],
        non_sampled_parameters=z["non_sampled_parameters"], sobol=z["sobol"], variability=variability,
        entry_time_ix=opt_int("entry_time_ix"), entry_time=float(z["entry_time"]), error_model=str(z["error_model"]), weight=float(z["weight"]),
        stdev_ix=opt_int("stdev_ix"), stdev=float(z["stdev"]), offset_ix=opt_int("offset_ix"), offset=float(z["offset"]),
        scale_ix=opt_int("scale_ix"), scale=float(z["scale"]), missing_simulation_time_stdev=float(z["missing_simulation_time_stdev"]),
        solver_relative_tolerance=float(z["solver_relative_tolerance"]), solver_absolute_tolerance=float(z["solver_absolute_tolerance"]),
        solver_min_timestep=float(z["solver_min_timestep"]), solver_max_steps=int(z["solver_max_steps"]), **extra)
    return prob, {k: z[k] for k in ("values", "logp", "cell_values", "cell_steps", "population_average", "noise_floor", "noise_floor_solver",
                                    "noise_floor_rhs", "noise_floor_trajectory", "noise_floor_step_match") if k in z.files}


NORTH_STAR_RTOL = 1e-6  # BASELINE.json: relative error of each per-chain log-likelihood against the reference's CVODE path


def parity_tolerance(noise_floor=None):
    """Per-chain relative tolerance of a parity assertion against the compiled reference: the north-star 1e-6 wherever the
    reference itself is reproducible at that level, and twice the reference's own measured irreproducibility elsewhere.

    `noise_floor` (per chain) is how far the reference's result moves between builds of its own sources that differ only in
    compiler flags (FMA contraction of the solver / of the generated right-hand side) and when its inputs move by one unit in
    the last place: tests/golden/measure_noise_floor.py
    stores it in every golden fixture, reference_noise_floor_cellpop() measures it for fresh inputs. It is ONE draw of a
    round-off-driven quantity (a step-size decision flips or it does not), so the bound is twice the draw, not the draw."""
    if noise_floor is None:
        return NORTH_STAR_RTOL
    return np.maximum(NORTH_STAR_RTOL, 2.0 * np.asarray(noise_floor, dtype=np.float64))


def assert_logp_parity(got, want, noise_floor=None, what=""):
    """PURE relative error per chain (no scaling by the size of the sum's terms)."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    err = rel_err(got, want)
    tol = parity_tolerance(noise_floor)
    assert np.all(err <= tol), f"{what} relative error {err} above tolerance {tol} (reference noise floor {noise_floor})"
    return err


def reference_noise_floor_cellpop(prob, vals, threads=4):
    """The same measurement as tests/golden/measure_noise_floor.py for fresh inputs: the compiled reference (A) against its
    -ffp-contract=off build (B), against itself with the generated right-hand side compiled with contraction (C) and against
    itself at inputs one unit in the last place away (D).
    Returns (logp_A result dict, per-chain floor, fraction of cells with identical step counts between the builds), or
    None when the reference builds are not present (container without /root/reference and without a shipped oracle/_ref)."""
    import oracle

    strict = os.path.join(os.path.dirname(oracle.REF_LIB), "libbcm3ref_strict.so")
    if not (oracle.available("ref") and os.path.exists(strict)):
        return None
    a = oracle.load("ref")
    b = _strict_oracle(strict)
    ra = a.cellpop_evaluate(prob, vals, threads=threads, want_steps=True, want_cell_values=True, want_average=True)
    rb = b.cellpop_evaluate(prob, vals, threads=threads, want_steps=True)
    oracle.rhs_build = "contracted"
    try:
        rc = a.cellpop_evaluate(prob, vals, threads=threads, want_steps=True)
    finally:
        oracle.rhs_build = "strict"
    floor = np.maximum(rel_err(rb["logp"], ra["logp"]), rel_err(rc["logp"], ra["logp"]))
    steps = min((ra["cell_steps"] == rb["cell_steps"]).mean(), (ra["cell_steps"] == rc["cell_steps"]).mean())
    # (D) the reference at inputs moved by one unit in the last place, both directions: its own conditioning. Small models give
    # the compiler nothing to contract, so A == B == C there says nothing about how sensitive the step decisions are.
    for direction in (np.inf, -np.inf):
        rd = a.cellpop_evaluate(prob, np.nextafter(np.asarray(vals, dtype=np.float64), direction), threads=threads, want_steps=True)
        floor = np.maximum(floor, rel_err(rd["logp"], ra["logp"]))
        steps = min(steps, (ra["cell_steps"] == rd["cell_steps"]).mean())
    return ra, floor, float(steps)


_strict_cache = {}


def _strict_oracle(path):
    import oracle

    if path not in _strict_cache:
        _strict_cache[path] = oracle.Oracle("ref", path)
    return _strict_cache[path]


def cellpop_step_match_floor(gold_or_fraction, num_solves=None):
    """Fraction of cells whose accepted-step count must equal the reference's. Equal step counts mean the same step-size/order
    decisions, and a round-off-sized difference flips them at the reference's own rate: `noise_floor_step_match` is the
    fraction on which the reference's own builds agree with each other (0.06 on the stiff fixture, 0.12-0.24 with pulsed
    treatments, 0.7-0.98 elsewhere). The expected agreement of an independent implementation is taken as half of that, and the
    assertion allows the binomial scatter of the count over the fixture's solves (two standard deviations): on the stiff
    fixture (48 solves, 3 of which agree between the reference's builds) that leaves no requirement at all -- step counts
    cannot demonstrate control-flow fidelity there; the mean step count (asserted separately, 2 %) and the parity of the
    result are what is checked."""
    if isinstance(gold_or_fraction, dict):
        frac = float(gold_or_fraction["noise_floor_step_match"])
        if num_solves is None and "cell_steps" in gold_or_fraction:
            num_solves = int(np.asarray(gold_or_fraction["cell_steps"]).size)
    else:
        frac = float(gold_or_fraction)
    q = 0.5 * frac
    if not num_solves:
        return q
    return max(0.0, q - 2.0 * np.sqrt(q * (1.0 - q) / num_solves))


# ---- the fixture whose model text came out of the reference's own SBML code generator (tests/golden/make_golden_sbml.py) ----
def sbml_cell_cycle_problem(num_cells=48, T=14, t_end=30.0, data_cells=8, seed=31):
    """13-species cell-cycle network, `derivative_code` = tests/golden/sbml_cell_cycle_generated.txt verbatim (species order,
    printed constants and helper calls as SBMLModel::GenerateCode chose them). The read-out is CycE + CycEp27
    (species_name="CycE+CycEp27"), per-cell variability on k_syn and k_deg and on the initial amount of Rb."""
    import math

    from scipy.integrate import solve_ivp
    from scipy.special import ndtri

    from bcm3_b200 import synthetic_cellpop as sc
    from bcm3_b200.cellpop_data import CellPopProblem, Variability
    from bcm3_b200.poppk_data import TRANSFORM_LOG10, TRANSFORM_NONE

    text = open(os.path.join(GOLDEN_DIR, "sbml_cell_cycle_generated.txt")).read()
    header = {}
    for line in text.splitlines():
        if line.startswith("// ") and ":" in line:
            key, _, rest = line[3:].partition(":")
            header[key.strip()] = rest.split()
    species = header["ode_species"]
    ic = np.array([float(x) for x in header["ode_initial"]])
    const = np.array([float(x) for x in header["constant_initial"]])
    code = text[text.index("EXPORT_PREFIX"):]
    N = len(species)
    names = header["variables"]
    transforms = np.full(len(names), TRANSFORM_LOG10, dtype=np.int32)
    transforms[names.index("variability_scale")] = TRANSFORM_NONE
    timepoints = t_end * np.arange(T) / (T - 1.0)
    variability = [Variability(apply="multiplicative_log", model_parameter=names.index("k_syn"), scale_ix=names.index("variability_scale")),
                   Variability(apply="multiplicative_log", model_parameter=names.index("k_deg"), scale_ix=names.index("variability_scale"), negate=True),
                   Variability(apply="additive", initial_condition_species=species.index("Rb"), scale_fixed=math.log(0.02))]
    sobol = sc.sobol_points(num_cells, len(variability))
    obs = [species.index("CycE"), species.index("CycEp27")]
    non_sampled = np.array([0.01])  # basal
    v = sbml_cell_cycle_values(1)[0]
    tv = np.where(transforms == TRANSFORM_LOG10, 10.0 ** v, v)
    f = sc.python_rhs(code, N)
    acc = np.zeros(T)
    rng = np.random.default_rng(seed)
    u = sc.sobol_points(data_cells, len(variability))
    for ci in range(data_cells):
        p = tv.copy()
        y0 = ic.copy()
        z = ndtri(u[ci]) * math.exp(tv[names.index("variability_scale")])
        p[names.index("k_syn")] *= math.exp(z[0])
        p[names.index("k_deg")] *= math.exp(-z[1])
        y0[species.index("Rb")] += ndtri(u[ci][2]) * 0.02
        sol = solve_ivp(lambda t, y: f(t, y, const, p, non_sampled), (0.0, float(timepoints[-1])), y0, method="LSODA", t_eval=timepoints, rtol=1e-7, atol=1e-9)
        acc += sol.y[obs].sum(axis=0)
    observed = (acc / data_cells)[None, :] + 0.01 * rng.standard_normal((1, T))
    return CellPopProblem(derivative_code=code, num_species=N, initial_conditions=ic, transforms=transforms, num_cells=num_cells, timepoints=timepoints,
                          observed=observed, obs_species=obs, constant_species=const, non_sampled_parameters=non_sampled, sobol=sobol,
                          variability=variability, stdev_ix=names.index("stdev"))


def sbml_cell_cycle_values(C, seed=77):
    """[C][6] sampled values (k_syn, k_deg, k_act, k_inh as log10; variability_scale natural log of the s.d.; stdev log10)."""
    import math

    base = np.array([math.log10(0.6), math.log10(0.4), math.log10(1.2), math.log10(0.8), math.log(0.2), math.log10(0.02)])
    rng = np.random.default_rng(seed)
    return base[None, :] + rng.normal(0.0, 0.05, (C, 6))


# ---- cell_population through the C++ plugin surface: prior.xml / likelihood.xml of the synthetic models ----
CELLPOP_VARIABLE_NAMES = ["k_in", "k_cascade", "k_deg", "k_feedback", "variability_scale", "stdev"]


def cellpop_xml(prob, **experiment_overrides):
    """prior.xml and likelihood.xml (reference schema) describing bcm3_b200.synthetic_cellpop.make_cellpop_problem."""
    import math

    logspace = lambda n: 'logspace="true" ' if n != "variability_scale" else ""
    prior = "<variableset>" + "".join(
        f'<variable name="{n}" {logspace(n)}distribution="uniform" lower="-5" upper="5"/>' for n in CELLPOP_VARIABLE_NAMES) + "</variableset>"
    species = [f"x{i}" for i in range(prob.num_species)]
    exp = dict(name="synthetic", model_file="cascade.xml", entry_time="0", num_cells=str(prob.num_cells), max_cells=str(prob.num_cells),
               divide_cells="false")
    exp.update(experiment_overrides)
    attrs = " ".join(f'{k}="{v}"' for k, v in exp.items())
    obs = "+".join(species[s] for s in prob.obs_species)
    lik = (f'<bcm_likelihood type="cell_population"><experiment {attrs}>'
           '<cell_variability distribution="diagonal_gaussian">'
           '<variable model_parameter="k_in" apply="multiplicative_log" scale="variability_scale"/>'
           '<variable model_parameter="k_deg" apply="multiplicative_log" scale="variability_scale" negate="true"/>'
           f'<variable initial_condition_species="x1" apply="additive" scale="{math.log(0.01)!r}"/>'
           '</cell_variability>'
           f'<data type="time_course_population_average" data_name="readout" species_name="{obs}" stdev="stdev"/>'
           '</experiment></bcm_likelihood>')
    return prior, lik, species


def cellpop_two_experiment_setup(seed=9):
    """A likelihood.xml with two experiments -- the first with two data sets on different species, timepoints and error
    models -- plus, for every <data> element, the single-data-set problem that reproduces it (for the direct ABI call and
    the CPU checkers): data sets of one experiment share the experiment's simulation end (Experiment.cpp:190-214, 655-656).
    Returns (prior, likelihood, species, [[problem per data set] per experiment])."""
    import dataclasses
    import math

    from bcm3_b200 import synthetic_cellpop as sc

    a = sc.make_cellpop_problem(N=8, num_cells=96, T=10, data_cells=4, seed=seed)
    b = sc.make_cellpop_problem(N=8, num_cells=64, T=7, data_cells=3, seed=seed)  # same model (the seed fixes the rate laws), other cells and times
    assert a.derivative_code == b.derivative_code
    prior, _, species = cellpop_xml(a)
    end = float(a.timepoints[-1])
    rng = np.random.default_rng(seed)
    a2 = dataclasses.replace(a, obs_species=[2, 3], timepoints=a.timepoints[:6].copy(),
                             observed=np.abs(a.observed[:, :6] * 1.7 + 0.05 * rng.standard_normal(a.observed[:, :6].shape)),
                             error_model="student_t4", stdev_ix=None, stdev=0.3, weight=0.5, simulation_end_time=end)
    b1 = dataclasses.replace(b, entry_time=0.25 * float(b.timepoints[1]))
    variability = ('<cell_variability distribution="diagonal_gaussian">'
                   '<variable model_parameter="k_in" apply="multiplicative_log" scale="variability_scale"/>'
                   '<variable model_parameter="k_deg" apply="multiplicative_log" scale="variability_scale" negate="true"/>'
                   f'<variable initial_condition_species="x1" apply="additive" scale="{math.log(0.01)!r}"/>'
                   '</cell_variability>')
    obs = lambda p: "+".join(species[s] for s in p.obs_species)
    lik = ('<bcm_likelihood type="cell_population">'
           f'<experiment name="first" model_file="cascade.xml" entry_time="0" num_cells="{a.num_cells}" max_cells="{a.num_cells}" divide_cells="false">'
           + variability +
           f'<data type="time_course_population_average" data_name="readout" species_name="{obs(a)}" stdev="stdev"/>'
           f'<data type="time_course_population_average" data_name="early" species_name="{obs(a2)}" stdev="0.3" error_model="student_t4" weight="0.5"/>'
           '</experiment>'
           f'<experiment name="second" model_file="cascade.xml" entry_time="{b1.entry_time!r}" num_cells="{b.num_cells}" max_cells="{b.num_cells}" divide_cells="false">'
           + variability +
           f'<data type="time_course_population_average" data_name="readout" species_name="{obs(b1)}" stdev="stdev"/>'
           '</experiment></bcm_likelihood>')
    return prior, lik, species, [[a, a2], [b1]]


def open_cellpop_session(prior, lik, species, problems):
    from bcm3_b200 import host_api

    s = host_api.CellPopSession(prior, lik)
    assert s.num_data_sets == [len(e) for e in problems]
    s.set_model(problems[0][0], species)
    for ei, exp in enumerate(problems):
        s.set_sobol(ei, exp[0].sobol)
        for di, p in enumerate(exp):
            s.set_data(ei, di, p.timepoints, p.observed)
    return s
