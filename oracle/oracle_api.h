/* TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * C interface shared by the two CPU checkers of the batched-likelihood path:
 *   - oracle/_ref/libbcm3ref.so  : the reference's OWN compiled solver stack
 *     (vendored CVODE 5.3.0 + src/odecommon, unmodified, compiled in place from
 *     /root/reference by oracle/ref/build_ref.sh) plus a thin restatement of
 *     the Boost/NetCDF-bound per-patient glue (oracle/ref/poppk_ref.cpp);
 *   - oracle/liboracle.so        : a plain-C restatement of the whole path
 *     (oracle/cvode_bdf.c + oracle/poppk_oracle.c), pinned against the former
 *     through tests/golden/.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load either of them.
 */
#ifndef BCM3_ORACLE_API_H
#define BCM3_ORACLE_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* pk_type values; reference enum EPKModelType, LikelihoodPopPKTrajectory.h */
#define ORACLE_PK_ONE 0 /* "one": N=2, LikelihoodPopPKTrajectory.cpp:446-467 */
#define ORACLE_PK_TWO 1 /* "two": N=3, LikelihoodPopPKTrajectory.cpp:469-494 */
#define ORACLE_PK_ONE_BIPHASIC 2 /* "one_biphasic_uptake", cpp:496-529 */
#define ORACLE_PK_TWO_BIPHASIC 3 /* "two_biphasic_uptake", cpp:531-571 */
#define ORACLE_PK_ONE_TRANSIT 4  /* "one_transit", cpp:573-604 */
#define ORACLE_PK_TWO_TRANSIT 5  /* "two_transit", cpp:606-642 */
#define ORACLE_PK_IS_TWO(t) ((t) == ORACLE_PK_TWO || (t) == ORACLE_PK_TWO_BIPHASIC || (t) == ORACLE_PK_TWO_TRANSIT)
#define ORACLE_PK_IS_BIPHASIC(t) ((t) == ORACLE_PK_ONE_BIPHASIC || (t) == ORACLE_PK_TWO_BIPHASIC)
#define ORACLE_PK_IS_TRANSIT(t) ((t) == ORACLE_PK_ONE_TRANSIT || (t) == ORACLE_PK_TWO_TRANSIT)

/* transform codes; reference VariableSet::TransformVariable, VariableSet.cpp:97-124 */
#define ORACLE_TRANSFORM_NONE 0
#define ORACLE_TRANSFORM_LOG 1   /* exp(x) */
#define ORACLE_TRANSFORM_LOG10 2 /* fastpow10(x) */
#define ORACLE_TRANSFORM_LOGIT 3

/* What LikelihoodPopPKTrajectory::Initialize (cpp:50-252) leaves in the object
 * after reading likelihood.xml + the NetCDF group. All arrays caller-owned. */
typedef struct {
	int32_t pk_type;
	int32_t num_patients;   /* P */
	int32_t num_timepoints; /* T */
	int32_t num_variables;  /* nvar = npk - nfixed + 2*(P+1) + 2, cpp:127 */
	int32_t sd_ix;          /* varset->GetVariableIndex("standard_deviation"), cpp:263 */
	int32_t max_steps;      /* ODESolverCVODE::max_steps, 2000, ODESolverCVODE.cpp:45 */
	double fixed_vod;       /* NaN when sampled, cpp:65 */
	double fixed_periphery_fwd;
	double fixed_periphery_bwd;
	double mol_weight;      /* MW of the drug, cpp:377-393 */
	double rtol;            /* (double)1e-6f, cpp:238 */
	double atol;            /* minimum_dose * (double)1e-6f, cpp:238 */
	const double* time;                   /* [T] */
	const double* observed_concentration; /* [P][T], NaN = missing */
	const double* dose;                   /* [P] */
	const double* dosing_interval;        /* [P] */
	const double* dose_after_dose_change; /* [P], NaN = no change */
	const double* dose_change_time;       /* [P] */
	const int32_t* intermittent;          /* [P] 0..3 */
	const uint32_t* skipped_days;         /* [P] bit d set = day d skipped (29 days) */
	const int32_t* simulate_until;        /* [P] number of leading timepoints simulated */
	const int32_t* transforms;            /* [nvar] ORACLE_TRANSFORM_* */
	/* variables the variants look up by name (cpp:296-310); -1 when the model type does not use them */
	int32_t n_transit_ix, mean_transit_time_ix, biphasic_uptake_time_ix, mean_absorption2_ix;
	/* 1: LikelihoodPharmacokineticTrajectory (src/likelihoods/LikelihoodPharmacokineticTrajectory.cpp:259-352), the likelihood
	 * of ONE patient (P = 1): absorption and elimination are variables 0 and 2 themselves instead of population quantiles
	 * (its cpp:276-279), the biphasic switching time is not clipped (its cpp:302), the dosing schedule is a bool (its
	 * cpp:184-186: any intermittent value acts as 1), a NaN concentration is not turned into -inf (its cpp:343-350) */
	int32_t single;
} oracle_poppk_problem;

/* per-(chain, patient) solver counters, summed over the restarts of one solve */
enum {
	ORACLE_CNT_STEPS = 0, /* ODESolverCVODE current_step (accepted steps) */
	ORACLE_CNT_NFE,
	ORACLE_CNT_NSETUPS,
	ORACLE_CNT_NJE,
	ORACLE_CNT_NETF,
	ORACLE_CNT_NCFN,
	ORACLE_CNT_NNI,
	ORACLE_CNT_OK, /* 1 = solve succeeded */
	ORACLE_NUM_COUNTERS
};

/* Evaluate num_chains parameter vectors (values[c*nvar + i]) the way
 * LikelihoodPopPKTrajectory::EvaluateLogProbability (cpp:259-444) does, one
 * chain per worker thread as SamplerPTChain.cpp:315-326 schedules them.
 *   logp      [C]            out
 *   conc      [C][P][T]      out, optional: conversion*trajectory(1,i) (NaN where not simulated)
 *   patient_ll[C][P]         out, optional: patient_logllh
 *   counters  [C][P][ORACLE_NUM_COUNTERS] out, optional
 * With conc/patient_ll/counters requested, the early `break` on logp == -inf
 * (cpp:438) is NOT taken so that every patient is reported; logp is unaffected.
 * Returns 0 on success. */
int oracle_poppk_evaluate(const oracle_poppk_problem* prob, size_t num_chains, const double* values,
                          double* logp, double* conc, double* patient_ll, int64_t* counters, int num_threads);

/* ---- cell_population (rows a8-a12 of SURVEY.md section 8): independent non-dividing cells, one cell_variability block
 * (diagonal_gaussian or full_gaussian), one time_course_population_average data set ---- */
typedef void (*oracle_derivative_fn)(double* out, const double* species, const double* constant_species, const double* parameters,
                                     const double* non_sampled_parameters); /* SolverCodeGenerator.h:6 */

/* A further MARKER of a per-cell data set (species_name="a+b;c": the part after a ';', DataLikelihoodTimeCourseBase.cpp:79-87):
 * its own observed block, species sum and stdev / offset / scale entries (DataLikelihoodBase.cpp:130-233: lists separated by ';') */
typedef struct {
	int32_t num_obs_species;
	int32_t obs_species[8];
	int32_t stdev_ix, offset_ix, scale_ix, proportional_stdev_ix; /* -1: fixed */
	double stdev, offset, scale, proportional_stdev;
	const double* observed; /* [observed cells][T] */
	int32_t log_ratio_denominator; /* use_log_ratio: species index of the denominator of "a/b" (obs_species = {a}), -1: none */
} oracle_cellpop_marker;

typedef struct {
	int32_t num_species, num_constant_species, num_variables, num_non_sampled, num_cells, num_timepoints, num_replicates, variability_dim;
	int32_t entry_time_ix; /* -1: fixed */
	int32_t max_steps;
	int32_t error_model; /* 0 normal, 1 student_t4, 2 proportional_normal, 3 additive_proportional_normal */
	int32_t stdev_ix, offset_ix, scale_ix; /* -1: fixed */
	int32_t proportional_stdev_ix;         /* -1: fixed */
	int32_t full_gaussian;                 /* 0: diagonal_gaussian, 1: full_gaussian (spherical Cholesky, VariabilityDescription.cpp:99-131) */
	int32_t num_obs_species;
	int32_t obs_species[8];
	double entry_time, rel_tol, abs_tol, min_dt, weight, stdev, offset, scale, missing_stdev, proportional_stdev;
	const double* covariance;          /* [D (D - 1) / 2][2]: variable index (or -1), fixed value; full_gaussian only */
	const double* initial_conditions;  /* [N] */
	const double* constant_species;    /* [Nc] */
	const double* non_sampled;         /* [Nn] */
	const double* sobol;               /* [cells][D] */
	const double* timepoints;          /* [T] */
	const double* observed;            /* [R][T] */
	const double* variability;         /* [D][6]: kind (0 parameter, 1 initial condition, 2 entry time: dimension only), target, apply, scale_ix, scale_fixed, negate */
	const int32_t* transforms;         /* [nvar] */
	oracle_derivative_fn derivative;   /* generated_derivative compiled for the host from the generated text */
	/* one <treatment_trajectory type="pulses"> (TreatmentTrajectoryPulses.cpp): the constant species it drives (-1: none)
	 * and the sorted pulse times; its value is a function of time and every pulse adds four discontinuities */
	int32_t treatment_species, treatment_num_pulses;
	const double* treatment_times;
	int32_t relative_to_time_average; /* <data relative_to_time_average="true">, DataLikelihoodTimeCoursePopulationAverage.cpp:105-115 */
	/* an experiment with several data sets integrates every cell to the last time ANY of them requests (Experiment.cpp:190-214,
	 * 655-656); evaluating one data set of such an experiment needs that end time. 0: the last of `timepoints` */
	int32_t have_sim_end_time;
	double sim_end_time;
	int32_t stdev_relative_to_scale; /* <data stdev_relative_to_scale="true">: stdev *= data scale, DataLikelihoodBase.cpp:151-153 */
	/* Dividing and dying cells (Experiment.cpp:726-782, CellPopulation.cpp:36-104, Cell.cpp:119-148, 463-538), the path without
	 * stored integration points: after every accepted step a cell whose "cytokinesis" species exceeds 1 ends there and two
	 * daughters start from its state (seven named species reset) with quasi-random rows num_cells + 2 * row(parent) + child;
	 * a cell whose "apoptosis" species exceeds 1 ends there. With divide_cells the per-cell output arrays have max_cells
	 * columns and the quasi-random table has sobol_rows rows (the reference makes 100 * num_cells). */
	int32_t divide_cells, max_cells, sobol_rows;
	int32_t cytokinesis_ix, apoptosis_ix; /* ODE species indices, -1: the model has no such species */
	int32_t reset_ix[7];                  /* cytokinesis, nuclear_envelope, G1S_break, G2_break, spindle_components, assembled_spindle, chromatid_separation */
	double max_dt; /* <experiment solver_max_timestep=> -> SetSolverParameter("max_dt") -> CVodeSetMaxStep (Cell.cpp:73); infinity: none */
	/* 0: <data type="time_course_population_average">; 2: see below; 1: <data type="time_course"> -- per-cell trajectories, one observed cell per
	 * row of `observed` ([num_replicates = observed cells][T]), every observed cell matched to one simulated cell by minimum-cost
	 * perfect matching (DataLikelihoodTimeCourse.cpp:230-365, 431-505; synchronize="none", no parent information, one marker) */
	int32_t data_kind;
	/* data_kind 2: <data type="time_points"> (DataLikelihoodTimePoints.cpp:209-345) -- `observed` [observed cell slots][T], NaN =
	 * no such cell at that timepoint; at every timepoint the observed cells present are matched to simulated cells.
	 * value_relative_to_timepoint_ix (DataLikelihoodBase.cpp:49; -1: none): the simulated value is (x + offset) / x(that
	 * timepoint) * scale instead of x * scale + offset */
	int32_t value_relative_to_timepoint_ix;
	/* time_course: <data optimize_offset_scale="true" optimize_offset_min= optimize_offset_max= optimize_scale_min= optimize_scale_max=>
	 * (DataLikelihoodTimeCourseBase.cpp:43-57, 317-322): for every (observed, simulated) pair the observed trajectory is regressed
	 * on the simulated one (bcm3::linear_regress_columns, Correlation.cpp:158-200), offset and scale clamped to their ranges */
	int32_t optimize_offset_scale;
	double optimize_offset_min, optimize_offset_max, optimize_scale_min, optimize_scale_max;
	/* time_course: <data saturation_scale="variable">: the scaled and shifted trajectories pass through
	 * s / (1 + exp(-x)) - s / 2 with s = that variable (DataLikelihoodTimeCourse.cpp:243-254); -1: none. (A NUMERIC saturation_scale
	 * is parsed by the reference and then overwritten with DBL_MAX in PrepateEvaluation, DataLikelihoodTimeCourseBase.cpp:243-246:
	 * only the variable form is usable there, and only it exists here.) */
	int32_t saturation_scale_ix;
	/* per-cell data kinds: the markers after the first (which is obs_species / observed / stdev ... above) */
	int32_t num_extra_markers;
	const oracle_cellpop_marker* extra_markers;
	/* time_course: <data use_log_ratio="true" species_name="a/b"> (DataLikelihoodTimeCourseBase.cpp:142-147, 171-201;
	 * DataLikelihoodTimeCourse.cpp:380-397): the cell's value is log10(a / b) = 0.4342944819032518 * log(a / b) with b replaced by
	 * 1e-16 when it is smaller; obs_species = {a}, this = b (-1: no ratio). Every marker of the data set is a ratio then. */
	int32_t log_ratio_denominator;
	/* time_points: <data use_only_nondivided="true"> (DataLikelihoodTimePoints.cpp:27, 349-351): the daughters of a dividing
	 * population (cell index >= num_cells) are left out of this data set */
	int32_t use_only_nondivided;
	/* <data include_only_cells_that_went_through_mitosis="true"> (population average, DataLikelihoodTimeCourseBase.cpp:44,
	 * DataLikelihoodTimeCoursePopulationAverage.cpp:171-176): only the cells that entered mitosis -- whose "nuclear_envelope"
	 * species (index nuclear_envelope_ix) was below 0.5 after some accepted step (Cell.cpp:487-492, Cell.h:27) -- are averaged, over
	 * the number of such cells alive at the timepoint (CellPopulation.cpp:106-121) */
	int32_t include_only_mitotic, nuclear_envelope_ix;
} oracle_cellpop_problem;

/* CellPopulationLikelihood::EvaluateLogProbability (CellPopulationLikelihood.cpp:82-101) for num_chains vectors.
 *   logp [C]; cell_values [C][T][cells] (optional); cell_steps [C][cells] (optional); population_average [C][T] (optional) */
int oracle_cellpop_evaluate(const oracle_cellpop_problem* prob, size_t num_chains, const double* values, double* logp,
                            double* cell_values, int32_t* cell_steps, double* population_average, int num_threads);

/* The same evaluation, reporting the solver's counters per cell: counters [C][cells][ORACLE_NUM_COUNTERS] (steps, nfe, nsetups,
 * nje, netf, ncfn, nni, ok; nfe counts the integrator's own right-hand-side evaluations -- the N per difference-quotient
 * Jacobian are nje * N on top). bench.py computes the algorithmic FLOPs of the cell_population workloads from these. */
int oracle_cellpop_evaluate_counters(const oracle_cellpop_problem* prob, size_t num_chains, const double* values, double* logp,
                                     int64_t* counters, int num_threads);

/* ---- pharmaco_population: PharmacoLikelihoodPopulation (src/pharmaco/PharmacoLikelihoodPopulation.cpp:202-340) around the
 * matrix-exponential compartment model PharmacokineticModel (src/pharmaco/PharmacokineticModel.cpp:111-247) ---- */
typedef struct {
	int32_t num_patients, num_timepoints, num_variables;
	int32_t use_peripheral, num_transit, use_bioavailability;
	/* variable indices, -1 = not in the prior (PostInitialize, cpp:102-188) */
	int32_t additive_sd_ix, proportional_sd_ix, mean_absorption_ix, mean_excretion_ix, mean_clearance_ix, mean_vod_ix;
	int32_t sigma_absorption_ix, sigma_excretion_ix, sigma_clearance_ix, sigma_vod_ix, sigma_transit_ix;
	int32_t periph_fwd_ix, periph_bwd_ix, mean_transit_time_ix;
	double mol_weight;
	const int32_t* transforms;              /* [nvar] */
	/* p<i>_<name> variable indices per patient (InitializePatientMarginals, cpp:342-354), NULL where the prior has no sigma */
	const int32_t *p_absorption_ix, *p_excretion_ix, *p_clearance_ix, *p_vod_ix, *p_transit_ix, *p_bioavailability_ix;
	/* the trial as the NetCDF file holds it (Patient::Load, PharmacoPatient.cpp:8-116) */
	const double* time;                     /* [T] */
	const double* observed_concentration;   /* [P][T], NaN = missing */
	const double *dose, *dosing_interval, *dose_after_dose_change, *dose_change_time; /* [P] */
	const int32_t* intermittent;            /* [P] */
	const uint32_t* skipped_days;           /* [P] bit d = day d skipped */
	/* 1: PharmacoLikelihoodSingle (src/pharmaco/PharmacoLikelihoodSingle.cpp:153-221), ONE patient (P = 1): mean_*_ix are the
	 * indices of "absorption", "excretion", "clearance", "volume_of_distribution", used through the variable transform; the
	 * direct-absorption route and the metabolite compartment (elimination fixed at 1, cpp:143) can be switched on */
	int32_t single, use_biphasic, use_metabolite, direct_absorption_ix, metabolite_conversion_ix;
} oracle_pharmaco_problem;

/* logp [C]; conc [C][P][T] optional (conversion * simulated concentration at the observations that have a value, NaN elsewhere);
 * patient_ll [C][P] optional */
int oracle_pharmaco_evaluate(const oracle_pharmaco_problem* prob, size_t num_chains, const double* values, double* logp, double* conc,
                             double* patient_ll, int num_threads);

/* hungarianMinimumWeightPerfectMatching (dependencies/hungarian2/hungarian.cpp) on a complete cost matrix [n][n], the edge list
 * in row order as DataLikelihoodTimeCourse::Evaluate builds it (.cpp:288-323); match [n], -1 everywhere without a matching.
 * "ref": the reference's compiled function; "port": the product's restatement (see oracle/cellpop_port.cpp). */
int oracle_hungarian_match(int n, const double* cost, int32_t* match);

/* "ref" or "port" */
const char* oracle_kind(void);

#ifdef __cplusplus
}
#endif

#endif
