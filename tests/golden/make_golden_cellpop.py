"""Generates tests/golden/cellpop_*.npz with the reference's own compiled solver stack (oracle/_ref: real ODESolverCVODE with
its difference-quotient Jacobian and zero-skipping LU + vendored CVODE 5.3.0) driving the generated-derivative text compiled
for the host. Run where /root/reference is mounted:  python tests/golden/make_golden_cellpop.py"""
import dataclasses
import zlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from bcm3_b200 import synthetic_cellpop as sc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: kwargs of make_cellpop_problem, C, tweaks
    "cellpop_n12_normal": (dict(N=12, num_cells=48, T=16, data_cells=8, seed=21), 3, {}),
    "cellpop_n12_t4_two_species": (dict(N=12, num_cells=40, T=12, data_cells=8, seed=22, replicates=3, missing_fraction=0.2,
                                        two_species_readout=True), 3, dict(error_model="student_t4", offset=0.01, scale=1.1, weight=0.5)),
    "cellpop_n5_late_entry": (dict(N=5, num_cells=33, T=10, data_cells=8, seed=23), 2, dict(entry_time=1.0)),
    "cellpop_n24_stiff": (dict(N=24, num_cells=24, T=10, data_cells=4, seed=24, rate_decades=4.0), 2, {}),
    # <cell_variability distribution="full_gaussian"> (two fixed angles and one sampled) + additive_proportional_normal error model
    "cellpop_n8_fullgauss_addprop": (dict(N=8, num_cells=64, T=12, data_cells=8, seed=25, replicates=2), 3,
                                     dict(variability_distribution="full_gaussian", covariance=[0.3, sc.VAR_VARIABILITY_SCALE, 0.15],
                                          error_model="additive_proportional_normal", proportional_stdev=0.1)),
    # <treatment_trajectory type="pulses"> driving the input (constant species 0): time-dependent RHS + 8 discontinuities, late entry
    "cellpop_n8_treatment_pulses": (dict(N=8, num_cells=40, T=20, t_end=40.0, data_cells=4, seed=27), 2,
                                    dict(treatment_species=0, treatment_times=np.array([20.0, 1.0]), obs_species=[0, 2], entry_time=4.0)),
    # <data relative_to_time_average="true">: log of the population average over its time average, then scaled
    "cellpop_n6_relative": (dict(N=6, num_cells=32, T=10, data_cells=8, seed=28, replicates=2), 2,
                            dict(relative_to_time_average=True, offset=0.05, scale=0.8)),
    # one data set of an experiment whose other data sets run longer: cells are integrated to the experiment's last requested
    # time (Experiment.cpp:190-214, 655-656), which enters CVODE's initial step -- late entry
    "cellpop_n8_longer_experiment": (dict(N=8, num_cells=48, T=8, data_cells=8, seed=29), 2,
                                     dict(simulation_end_time=13.5, entry_time=0.5)),
    # the same with a pulsed treatment: discontinuities (re-initialisations) keep coming after the data set's last timepoint
    "cellpop_n8_treatment_pulses_longer": (dict(N=8, num_cells=40, T=20, t_end=40.0, data_cells=4, seed=27), 2,
                                           dict(treatment_species=0, treatment_times=np.array([20.0, 1.0, 45.0]), obs_species=[0, 2], entry_time=4.0,
                                                simulation_end_time=55.0)),
    # the model text is the OUTPUT OF THE REFERENCE'S OWN SBML CODE GENERATOR (tests/golden/make_golden_sbml.py): 13 species, 24
    # reactions, hill_function with a non-integer exponent, the fixed-exponent variants 2/4/10/16/100, michaelis_menten_function,
    # tQSSA, safepow, synthcap, exp, a division, a non-sampled parameter, three constant species, stoichiometry 2
    "cellpop_sbml_cell_cycle": (dict(_builder="sbml_cell_cycle", num_cells=48, T=14), 3, {}),
    # <experiment divide_cells="true">: 16 initial cells, two generations of daughters within the experiment (Experiment.cpp:726-782,
    # Cell.cpp:119-148, 463-538); a variability variable with only_initial_cells="true"
    "cellpop_dividing_two_generations": (dict(_builder="dividing", M=5, num_cells=16, max_cells=400, t_end=5.5, T=16), 3, {}),
    # the same with an "apoptosis" species: some cells die before or instead of dividing
    "cellpop_dividing_with_apoptosis": (dict(_builder="dividing", M=5, num_cells=16, max_cells=400, t_end=5.5, T=16, with_apoptosis=True), 3, {}),
    # <data type="time_course">: one observed trajectory per cell, every observed cell matched to one simulated cell by the
    # reference's own Hungarian implementation (DataLikelihoodTimeCourse.cpp:230-365, dependencies/hungarian2 compiled into oracle/_ref)
    "cellpop_time_course_n8_normal": (dict(_builder="time_course", N=8, num_cells=24, T=12, seed=41), 3, {}),
    "cellpop_time_course_n6_t4_missing": (dict(_builder="time_course", N=6, num_cells=40, T=10, seed=42, missing_fraction=0.15,
                                               two_species_readout=True), 3,
                                          dict(error_model="student_t4", offset=0.01, scale=1.1, weight=0.5)),
    "cellpop_time_course_n6_addprop": (dict(_builder="time_course", N=6, num_cells=33, T=10, seed=43), 2,
                                       dict(error_model="additive_proportional_normal", proportional_stdev=0.1, _positive_data=True)),
    # optimize_offset_scale: the observed trajectories are in arbitrary fluorescence units (x 2.7 + 0.4 here); every (observed,
    # simulated) pair is regressed first, offset and scale clamped (the offset clamp of +-0.3 bites for some pairs)
    "cellpop_time_course_n8_optimize": (dict(_builder="time_course", N=8, num_cells=24, T=12, seed=44, missing_fraction=0.1), 3,
                                        dict(optimize_offset_scale=True, optimize_offset_range=(-0.3, 0.3), optimize_scale_range=(0.1, 10.0),
                                             _affine_data=(0.4, 2.7), stdev=0.08)),
    # saturation_scale="k_feedback": the scaled trajectories pass through s / (1 + exp(-x)) - s / 2 (DataLikelihoodTimeCourse.cpp:243-254);
    # the data went through the same curve with s = 0.5
    "cellpop_time_course_n6_saturation": (dict(_builder="time_course", N=6, num_cells=20, T=10, seed=47), 3,
                                          dict(saturation_scale_ix=sc.VAR_K_FEEDBACK, scale=4.0, _saturated_data=(4.0, 0.5), stdev=0.01)),
    # species_name="a;b+c;d": three markers per cell, each with its own scale, offset and stdev (DataLikelihoodBase.cpp:130-233), missing values
    "cellpop_time_course_n8_three_markers": (dict(_builder="time_course", N=8, num_cells=20, T=10, seed=48, missing_fraction=0.1,
                                                  extra_marker_species=((5, 6), (3,))), 3, dict(error_model="student_t4", weight=0.8)),
    # use_log_ratio="true" species_name="x7/x3": a ratiometric reporter, the cell's value is log10 of the ratio (the first timepoint, where
    # the downstream species still are 0, is left out of the data: log10(0 / 0) there)
    "cellpop_time_course_n8_log_ratio": (dict(_builder="time_course", N=8, num_cells=20, T=10, seed=50, log_ratio_denominator=3, noise=0.05), 3,
                                         dict(_drop_first_timepoint=True)),
    "cellpop_time_points_n8_two_markers": (dict(_builder="time_points", N=8, num_cells=24, T=8, seed=49, extra_marker_species=((4,),)), 3, {}),
    # <data type="time_points">: at every timepoint its own set of observed cells, matched to the simulated cells (rectangular
    # Hungarian calls: fewer observed than simulated cells at most timepoints), DataLikelihoodTimePoints.cpp:209-345
    "cellpop_time_points_n8_normal": (dict(_builder="time_points", N=8, num_cells=24, T=8, seed=45), 3, {}),
    # value_relative_to_timepoint_ix (DataLikelihoodBase.cpp:49): simulated values relative to the cell's own value at timepoint 2;
    # Student-t error model, offset / scale / weight
    "cellpop_time_points_n6_t4_relative": (dict(_builder="time_points", N=6, num_cells=30, T=9, seed=46, relative_to=2), 3,
                                           dict(error_model="student_t4", offset=0.01, scale=1.1, weight=0.5, stdev=0.01)),
    # <data type="time_points"> on a DIVIDING population: snapshots with more observed cells at the later timepoints (two generations of
    # daughters); a cell slot that is not filled, or a cell outside its life span, has no value and is left out per timepoint
    "cellpop_time_points_dividing": (dict(_builder="dividing_snapshots", M=5, num_cells=16, max_cells=400, t_end=5.5, T=12), 3, {}),
    # the same with use_only_nondivided="true": only the 16 initial cells are matched (DataLikelihoodTimePoints.cpp:349-351)
    "cellpop_time_points_dividing_nondivided": (dict(_builder="dividing_snapshots", M=5, num_cells=16, max_cells=400, t_end=5.5, T=12, nondivided=True), 3, {}),
    # include_only_cells_that_went_through_mitosis="true": the population average over the cells whose nuclear envelope species fell
    # below 0.5 after some accepted step (Cell::EnteredMitosis), over the number of such cells alive at the timepoint
    "cellpop_dividing_mitotic_only": (dict(_builder="dividing", M=5, num_cells=16, max_cells=400, t_end=5.5, T=16), 3,
                                      dict(include_only_cells_that_went_through_mitosis=True, nuclear_envelope_species=6)),
    "cellpop_n6_proportional": (dict(N=6, num_cells=32, T=10, data_cells=8, seed=26), 2,
                                dict(error_model="proportional_normal", proportional_stdev=0.25, _positive_data=True)),
}


def main():
    ref = oracle.load("ref")
    only = sys.argv[1:]
    for name, (kw, C, tweaks) in CASES.items():
        if only and name not in only:
            continue
        tweaks = dict(tweaks)
        positive = tweaks.pop("_positive_data", False)
        affine = tweaks.pop("_affine_data", None)
        saturated = tweaks.pop("_saturated_data", None)
        drop_first = tweaks.pop("_drop_first_timepoint", False)
        kw = dict(kw)
        builder = kw.pop("_builder", None)
        if builder == "dividing" and tweaks:
            prob = dataclasses.replace(sc.make_dividing_problem(**kw), **tweaks)
            fixed_values = sc.make_chain_values(C, seed=5)
        elif builder == "dividing_snapshots":
            # observations: values of cells that exist at each timepoint in the reference's own simulation at the reference parameters
            nondivided = kw.pop("nondivided", False)
            base = sc.make_dividing_problem(**kw)
            sim = ref.cellpop_evaluate(base, sc.default_values()[None, :], threads=1, want_cell_values=True)["cell_values"][0]  # [T][max_cells]
            rng = np.random.default_rng(77)
            T = base.num_timepoints
            slots = 40
            observed = np.full((slots, T), np.nan)
            for ti in range(T):
                alive = np.flatnonzero(~np.isnan(sim[ti, :base.num_cells] if nondivided else sim[ti]))
                take = rng.permutation(alive)[:max(1, min(len(alive) // 2, slots, int(rng.integers(4, 14))))]
                where = rng.permutation(slots)[:len(take)]
                observed[where, ti] = sim[ti, take] + 0.02 * rng.standard_normal(len(take))
            prob = dataclasses.replace(base, data_kind="time_points", observed=observed, stdev_ix=None, stdev=0.03, use_only_nondivided=nondivided)
            fixed_values = sc.make_chain_values(C, seed=5)
        elif builder == "dividing":
            prob = sc.make_dividing_problem(**kw)
            fixed_values = sc.make_chain_values(C, seed=5)
        elif builder == "time_course":
            prob = dataclasses.replace(sc.make_time_course_problem(**kw), **tweaks)
            fixed_values = None
        elif builder == "time_points":
            prob = dataclasses.replace(sc.make_time_points_problem(**kw), **tweaks)
            fixed_values = None
        elif builder == "sbml_cell_cycle":
            from tests.util import sbml_cell_cycle_problem, sbml_cell_cycle_values

            prob = sbml_cell_cycle_problem(**kw)
            fixed_values = sbml_cell_cycle_values(C)
        else:
            prob = dataclasses.replace(sc.make_cellpop_problem(**kw), **tweaks)
            fixed_values = None
        if drop_first:
            obs = prob.observed.copy()
            obs[:, 0] = np.nan
            prob = dataclasses.replace(prob, observed=obs)
        if saturated is not None:
            prob = dataclasses.replace(prob, observed=saturated[1] / (1.0 + np.exp(-saturated[0] * prob.observed)) - 0.5 * saturated[1])
        if affine is not None:
            prob = dataclasses.replace(prob, observed=affine[0] + affine[1] * prob.observed)
        if positive:  # a proportional error model has sigma = 0 (log-density NaN) at data <= 0
            prob = dataclasses.replace(prob, observed=np.abs(prob.observed) + 0.05)
        vals = fixed_values if fixed_values is not None else sc.make_chain_values(C, seed=zlib.crc32(name.encode()) % 10000)
        r = ref.cellpop_evaluate(prob, vals, threads=1, want_cell_values=True, want_steps=True, want_average=True)
        out = {f.name: getattr(prob, f.name) for f in dataclasses.fields(prob) if f.name not in ("variability", "covariance", "extra_markers")}
        if prob.extra_markers:  # [marker][...]: species lists padded with -1, one row of (stdev_ix, stdev, proportional_stdev_ix, proportional_stdev, offset_ix, offset, scale_ix, scale)
            opt = lambda v: -1.0 if v is None else float(v)
            out["marker_obs_species"] = np.array([list(m.obs_species) + [-1] * (8 - len(m.obs_species)) for m in prob.extra_markers], dtype=np.int64)
            out["marker_observed"] = np.stack([np.asarray(m.observed, dtype=np.float64) for m in prob.extra_markers])
            out["marker_parameters"] = np.array([[opt(m.stdev_ix), m.stdev, opt(m.proportional_stdev_ix), m.proportional_stdev, opt(m.offset_ix), m.offset,
                                                  opt(m.scale_ix), m.scale] for m in prob.extra_markers], dtype=np.float64)
        out["covariance_rows"] = prob.covariance_rows()
        out = {k: (np.array(v) if not isinstance(v, np.ndarray) else v) for k, v in out.items() if v is not None}
        out["variability_rows"] = prob.variability_rows()
        out.update(values=vals, logp=r["logp"], cell_values=r["cell_values"], cell_steps=r["cell_steps"], population_average=r["population_average"])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "logp", r["logp"], "steps mean", r["cell_steps"].mean(), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
