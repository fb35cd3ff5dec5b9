/* bcm3b200.h -- C ABI of the B200-native batched likelihood evaluator.
 *
 * Drop-in boundary for ONE path of NKI-CCB/bcm3: evaluating
 *   bcm3::Likelihood::EvaluateLogProbability(threadix, values, logp)      (src/sampler/Likelihood.h:29)
 * for every parallel-tempered chain's proposal at once. The reference has no FFI for this
 * path; the closest precedent is LikelihoodDLL (src/likelihoods/LikelihoodDLL.h:17-18:
 * `initialize_likelihood` / `evaluate_log_probability`, resolved with dlsym,
 * LikelihoodDLL.cpp:72-75) and the R bridge's all-pointer convention
 * (src/bcmrbridge/interface.cpp:27-101). The entry points below are what a
 * GPU-backed `bcm3::Likelihood` subclass binds (INTEGRATION.md shows the subclass).
 *
 * Conventions: plain pointers and sizes, caller-owned buffers, row = chain.
 * Return 0 = ok, < 0 = unrecoverable (maps to `return false` in the reference's
 * bool convention). logp = -inf is a legal value (failed ODE solve,
 * LikelihoodPopPKTrajectory.cpp:400-408); a NaN log-likelihood is reported through
 * status[c] != 0 (the reference sampler treats NaN as an error, Sampler.cpp:172-178).
 * There is no CPU fallback: every call fails with BCM3B200_ERR_CUDA when no device is usable.
 */
#ifndef BCM3B200_H
#define BCM3B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BCM3B200_OK 0
#define BCM3B200_ERR_ARG (-1)     /* bad argument / unknown name / shape mismatch */
#define BCM3B200_ERR_STATE (-2)   /* data missing or call out of order */
#define BCM3B200_ERR_CUDA (-3)    /* CUDA runtime error (bcm3b200_last_error has the text) */
#define BCM3B200_ERR_UNSUPPORTED (-4)

/* per-chain status codes written by evaluate */
#define BCM3B200_STATUS_OK 0
#define BCM3B200_STATUS_NAN 1 /* log-likelihood is NaN: the reference aborts sampling on this (Sampler.cpp:172-178) */

/* Create an evaluator.
 *   model_kind : the reference's likelihood.xml type string (LikelihoodFactory.cpp:62,66,81):
 *                "pop_pk_trajectory" | "pharmacokinetic_trajectory" | "cell_population" | "pharmaco_population" | "pharmaco_single"
 *                pharmaco_single = PharmacoLikelihoodSingle (LikelihoodFactory.cpp:64, src/pharmaco/PharmacoLikelihoodSingle.cpp): the
 *                matrix-exponential model of pharmaco_population for ONE patient (num_patients=1, trial data as below), whose
 *                rates are the chain's variables themselves: keys drug num_patients num_timepoints num_variables
 *                [peripheral_compartment] [num_transit_compartments] [biphasic_absorption=0|1] [metabolite=0|1] (cpp:39-50) and
 *                <name>_ix for additive_sd, proportional_sd, absorption, excretion, clearance, volume_of_distribution,
 *                peripheral_forward_rate, peripheral_backward_rate, mean_transit_time, direct_absorption,
 *                metabolite_conversion_rate (the variable names of cpp:75-146; the metabolite's elimination is fixed at 1).
 *                pharmacokinetic_trajectory = LikelihoodPharmacokineticTrajectory (LikelihoodFactory.cpp:60,
 *                src/likelihoods/LikelihoodPharmacokineticTrajectory.cpp:259-352), the likelihood of ONE patient on the same
 *                models, solver and dosing logic: the keys and data of pop_pk_trajectory below with num_patients=1 (the host
 *                picks the patient its <pk_model patient=> names out of the trial), and these differences, all the
 *                reference's: variables 0 and 2 are the absorption rate and the clearance themselves (no population level, its
 *                cpp:276-279), the biphasic pair is read at positions 6 and 7 and the switching time is not clipped (its
 *                cpp:302-303), every timepoint is simulated, any "intermittent" value acts as schedule 1 (a bool, its
 *                cpp:184-186), no variable-count check, absolute tolerance dose * 1e-6f, a NaN concentration stays NaN.
 *                A batch is one ODE system per chain.
 *                cell_population = CellPopulationLikelihood (src/likelihoods/cellpop/CellPopulationLikelihood.cpp:27-101) with
 *                one Experiment (Experiment.cpp:404-633) per handle:
 *                  num_species num_constant_species num_variables num_non_sampled num_cells num_timepoints num_replicates
 *                  variability_dim variability_distribution=diagonal_gaussian|full_gaussian  entry_time | entry_time_ix
 *                  solver_relative_tolerance solver_absolute_tolerance solver_min_timestep solver_max_timestep solver_max_steps
 *                  error_model=normal|student_t4|proportional_normal|additive_proportional_normal  weight
 *                  relative_to_time_average stdev_relative_to_scale missing_simulation_time_stdev simulation_end_time
 *                  stdev | stdev_ix, proportional_stdev | proportional_stdev_ix, offset | offset_ix, scale | scale_ix
 *                  obs_species=<i>+<j>+...  treatment_species=<constant species index>
 *                  divide_cells=0|1 max_cells cytokinesis_species apoptosis_species division_reset_species=<7 indices, '+'>
 *                       (Experiment.cpp:726-782, CellPopulation.cpp:36-104, Cell.cpp:119-148: a cell whose cytokinesis species
 *                        passes 1 is replaced by two daughters, one whose apoptosis species passes 1 ends; a chain that
 *                        outgrows max_cells or the quasi-random table evaluates to -inf; such handles cannot be sharded)
 *                  data_kind=time_course_population_average|time_course|time_points (the <data type=>; default the population average).
 *                       time_points = DataLikelihoodTimePoints (src/cellpop/DataLikelihoodTimePoints.cpp:209-345), synchronize="none",
 *                        one marker: "observed" is [num_replicates = observed cell slots <= num_cells][T], NaN = no such cell at
 *                        that timepoint; at every timepoint the observed cells present are matched to the simulated cells that
 *                        have a value there (normal | student_t4); value_relative_to_timepoint_ix=<t> (DataLikelihoodBase.cpp:49):
 *                        the simulated value is (x + offset) / x(timepoint t) * scale instead of x * scale + offset;
 *                        works with divide_cells (num_replicates <= max_cells; use_only_nondivided=1 leaves the daughters out)
 *                       time_course = DataLikelihoodTimeCourse (src/cellpop/DataLikelihoodTimeCourse.cpp:230-365, 431-505,
 *                        566-588) with synchronize="none", one marker and no parent information: "observed" holds one
 *                        trajectory per OBSERVED CELL ([num_replicates = observed cells][T], NaN = missing), there are as
 *                        many observed as simulated cells (the reference refuses anything else, .cpp:178-187), every pair's
 *                        log-likelihood is computed on the device and the observed cells are matched to the simulated ones
 *                        on the host exactly as the reference's Hungarian call does (bcm3b200_match_cells below); such
 *                        handles cannot be sharded or combined with divide_cells.
 *                        optimize_offset_scale=1 [optimize_offset_min=-1 optimize_offset_max=1 optimize_scale_min=0.1
 *                        optimize_scale_max=10] (DataLikelihoodTimeCourseBase.cpp:43-57, 317-322): every pair regresses the
 *                        observed on the simulated trajectory first (normal | student_t4 only);
 *                        saturation_scale_ix=<variable>: the signal saturation of DataLikelihoodTimeCourse.cpp:243-254
 *                       several MARKERS per cell (species_name="a+b;c" with ';'-separated stdev / offset / scale lists,
 *                        DataLikelihoodTimeCourseBase.cpp:79-87, DataLikelihoodBase.cpp:130-233): every marker after the first is
 *                        passed as a further data set of the handle (num_data_sets, suffix @k: num_timepoints, num_replicates,
 *                        obs_species, the stdev / offset / scale keys, "timepoints@k", "observed@k") that carries
 *                        marker_of@k=<index of the per-cell data set it belongs to, 0 = the first>; it has the timepoints and
 *                        observed cells of that data set and no term of its own -- its values enter that data set's cell
 *                        likelihoods (time_course: all markers summed per pair; time_points: a simulated cell counts when
 *                        marker 0 has a value, missing observations of a marker are skipped)
 *                       use_log_ratio (time_course, species_name="a/b", DataLikelihoodTimeCourse.cpp:380-397): obs_species = a; the
 *                        denominator b is one more entry of the handle (suffix @k: num_timepoints, num_replicates=1, obs_species=b,
 *                        "timepoints@k", an all-zero "observed@k") that carries denominator_of@k=<index of the data set or
 *                        marker it divides>: the cell's value is 0.4342944819032518 * log(a / max-guarded b) before scale / offset
 *                  include_only_cells_that_went_through_mitosis=0|1 (population average, DataLikelihoodTimeCoursePopulationAverage.cpp:
 *                       171-176) with nuclear_envelope_species=<index>: only the cells whose nuclear envelope species fell below
 *                       0.5 after some accepted step (Cell.cpp:487-492, Cell.h:27) are averaged, over their own number
 *                  num_data_sets=<D <= 4>: the experiment's further <data> elements share this handle's ONE integration of
 *                       the cells; data set k >= 1 repeats num_timepoints, num_replicates, obs_species, error_model, weight, data_kind,
 *                       the stdev/offset/scale keys and the relative_to/missing keys with the suffix @k ("stdev_ix@1=5")
 *                       and supplies "timepoints@k", "observed@k"; the result is the sum over the data sets in order
 *                  [shard_rank=0] [shard_count=1] [device=0]
 *                Data: initial_conditions[N] constant_species non_sampled_parameters timepoints[T] observed[R][T]
 *                transforms[nvar] sobol[rows][D] (the quasi-random table, VariabilityPseudoRandomIterator.cpp:14-26; rows >=
 *                num_cells, dividing populations draw the daughters' rows from it) variability[D][6] (kind: 0 parameter,
 *                1 initial condition, 2 entry time, +16 = only_initial_cells; target; apply; scale variable; fixed scale; negate)
 *                variability_covariance[D(D-1)/2][2] treatment_times[pulses]. Text: derivative_code (bcm3b200_set_text).
 *                DESIGN.md section 9 has the meaning of every key with the reference line it restates.
 *                pharmaco_population = PharmacoLikelihoodPopulation (src/pharmaco/PharmacoLikelihoodPopulation.cpp:43-340): the
 *                population PK model advanced with the matrix exponential (PharmacokineticModel.cpp:111-247), no ODE solver.
 *                  drug=<name> num_patients=<P> num_timepoints=<T> num_variables=<nvar>
 *                  [peripheral_compartment=0|1] [num_transit_compartments=<k>] [bioavailability=0|1]   (<pk_model> attributes, cpp:51-56)
 *                  <role>_ix=<index of the prior variable>, roles (variable names of cpp:104-186): additive_sd, proportional_sd,
 *                       mean_absorption, mean_excretion, mean_clearance, mean_volume_of_distribution, sigma_absorption,
 *                       sigma_excretion, sigma_clearance, sigma_volume_of_distribution, sigma_transit_time,
 *                       peripheral_forward_rate, peripheral_backward_rate, mean_transit_time -- absent = not in the prior
 *                  [shard_rank=0] [shard_count=1] [device=0]
 *                Data: the trial arrays of pop_pk_trajectory below (the same NetCDF group, Patient::Load PharmacoPatient.cpp:8-116)
 *                plus, for every marginal the prior switches on, the indices of the per-patient variables p<i>_<name>
 *                (InitializePatientMarginals, cpp:342-354) as "patient_absorption_ix" | "patient_excretion_ix" |
 *                "patient_clearance_ix" | "patient_volume_of_distribution_ix" | "patient_transit_time_ix" |
 *                "patient_bioavailability_ix", each [P]. Entry points: finalize, evaluate_batch (with a communicator: the
 *                complete result on every rank), enqueue_batch (d_partial [3][C] as for pop_pk_trajectory, except that a NaN
 *                term anywhere makes the chain NaN), exchange_partials, get_diagnostics (conc, patient_ll; no counters).
 *   model_desc : `key=value;...` text, desc_bytes long (no terminator needed). Keys for pop_pk_trajectory
 *                mirror the <pk_model> attributes (LikelihoodPopPKTrajectory.cpp:58-87) plus sizes:
 *                  type=one|two|one_biphasic_uptake|two_biphasic_uptake|one_transit|two_transit  drug=<name>
 *                       (as in the reference BOTH biphasic strings select the two-compartment biphasic model, cpp:73-76;
 *                        its one-compartment biphasic right-hand side, unreachable from the reference's XML, is
 *                        available as type=one_compartment_biphasic_uptake)
 *                  [volume_of_distribution=<v>] [k_periphery_fwd=<v>] [k_periphery_bwd=<v>]  fixed instead of sampled
 *                       (cpp:64-67): each takes one variable out of the prior (cpp:122-130) while the vector is still read
 *                       at the all-sampled positions, exactly as the reference does
 *                  num_patients=<P>  num_timepoints=<T>
 *                  num_variables=<nvar>  sd_ix=<index of "standard_deviation">  [max_steps=2000]
 *                  transit types:  n_transit_ix= mean_transit_time_ix=   (indices of the variables of those names,
 *                  biphasic types: biphasic_uptake_time_ix= mean_absorption2_ix=    cpp:296-310)
 *                  [shard_rank=0] [shard_count=1]   contiguous slice of patients owned by this handle
 *                  [device=0]                       first CUDA device ordinal
 *   device_count: number of CUDA devices (device .. device+device_count-1) this handle spreads its patients / cells
 *                over inside this process; 1 for the one-process-per-GPU launch.
 * Replaces: LikelihoodFactory::CreateLikelihood -> make_shared<LikelihoodPopPKTrajectory> + Initialize
 *           (LikelihoodFactory.cpp:31-101). */
int bcm3b200_create(const char* model_kind, const void* model_desc, size_t desc_bytes, int device_count, void** handle);

/* Attach one named static input (all as doubles; integer-valued inputs are passed as doubles).
 * pop_pk_trajectory names and shapes = the NetCDF variables read at LikelihoodPopPKTrajectory.cpp:94-161
 * (FULL arrays over all P patients even when the handle owns a shard):
 *   "time"[T]  "observed_concentration"[P][T] (NaN = missing)  "dose"[P]  "dosing_interval"[P]
 *   "dose_after_dose_change"[P] (NaN = none)  "dose_change_time"[P]  "intermittent"[P] (0..3)
 *   "treatment_interruptions"[P][29] (0/1)
 *   "transforms"[nvar] : VariableSet transform per variable, 0 none / 1 log / 2 log10 / 3 logit (VariableSet.cpp:97-124)
 * Replaces: the NetCDFDataFile reads in LikelihoodPopPKTrajectory::Initialize. */
int bcm3b200_set_data(void* handle, const char* name, const double* data, const size_t* shape, int ndim);

/* Attach one named text input. cell_population: "derivative_code" = the C++ text the reference's SBML code generator
 * emits (SBMLModel::GenerateCode, src/sbml/SBMLModel.cpp:291-389; ABI derivative_fn of SolverCodeGenerator.h:6):
 *   EXPORT_PREFIX void generated_derivative(OdeReal* out, const OdeReal* species, const OdeReal* constant_species,
 *                                           const OdeReal* parameters, const OdeReal* non_sampled_parameters) {...}
 * The library compiles it for the device at finalize, the way the reference compiles it for the host
 * (SolverCodeGenerator.cpp:390,407-414). */
int bcm3b200_set_text(void* handle, const char* name, const char* text, size_t text_bytes);

/* Derive simulate_until / tolerances / skipped-day masks (LikelihoodPopPKTrajectory.cpp:163-204,238) and upload
 * the static data to the device(s). Called implicitly by the first evaluate. Replaces: PostInitialize. */
int bcm3b200_finalize(void* handle);

/* Evaluate num_chains parameter vectors at once. HOST buffers:
 *   values [num_chains][num_variables]  (row c = chain c's VectorReal `values`)
 *   logp   [num_chains]  out
 *   status [num_chains]  out, may be NULL
 * With shard_count > 1 the result is this shard's partial (see bcm3b200_evaluate_batch_device for the
 * exact combination rule); otherwise it equals the reference's EvaluateLogProbability per chain.
 * Replaces: C calls of LikelihoodPopPKTrajectory::EvaluateLogProbability (cpp:259-444). */
int bcm3b200_evaluate_batch(void* handle, size_t num_chains, size_t num_variables, const double* values, double* logp,
                            int* status);

/* Same evaluation on DEVICE buffers of the handle's first device, enqueued on `stream` (a cudaStream_t), no sync:
 *   d_values  [num_chains][num_variables] device (8-byte alignment is enough; rows whose per-patient block happens to be
 *             16-byte aligned are read with 128-bit loads)
 *   d_partial [3][num_chains] device, out:
 *       row 0: sum of the finite per-patient log-likelihoods of this shard
 *       row 1: global index of the first patient whose log-likelihood is -inf (+inf if none)
 *       row 2: global index of the first patient whose log-likelihood is NaN  (+inf if none)
 * Combination across shards (NCCL all-reduce: SUM on row 0, MIN on rows 1-2) followed by
 * bcm3b200_combine_partials reproduces the reference's serial `logp += patient_logllh; if (logp == -inf) break;`
 * loop (cpp:427-440) exactly in its -inf / NaN outcome. */
int bcm3b200_evaluate_batch_device(void* handle, size_t num_chains, size_t num_variables, const double* d_values,
                                   double* d_partial, void* stream);

/* HOST values in, DEVICE partial out, enqueued on `stream` without synchronising: copies only this handle's slice of
 * the batch (chain-level entries + its patients' probabilities) host->device, then runs the kernels.
 * `values` should be page-locked (bcm3b200_host_alloc) for the copy to overlap; it must stay untouched until the
 * stream has passed the copy (the library stages nothing of its own, so several batches may be in flight on one
 * handle as long as each has its own `values` and `d_partial`). d_partial as in bcm3b200_evaluate_batch_device. This is the entry the
 * one-process-per-GPU launch uses: enqueue, NCCL all-reduce d_partial, read back 3*C doubles. */
int bcm3b200_enqueue_batch(void* handle, size_t num_chains, size_t num_variables, const double* values, double* d_partial,
                           void* stream);

/* cell_population handles created with shard_count > 1 own a contiguous slice of the simulated cells (the reference
 * walks them in one loop, Experiment.cpp:470-560 / DataLikelihoodTimeCourse.cpp:161-197). For those handles
 * bcm3b200_enqueue_batch writes d_partial [num_chains][2 T + 1]: per timepoint the shard's sum of the existing cells'
 * observed values and their count, then its number of failed cells, all as doubles. One SUM all-reduce of d_partial
 * over the ranks followed by bcm3b200_cellpop_finish on every rank gives the population average (sum / count; the
 * unsharded handle divides every cell by the population size before summing -- the two differ at round-off), the
 * data likelihood and the -inf verdict for failed cells, exactly as the unsharded bcm3b200_evaluate_batch does.
 * bcm3b200_get_stat(h, "partial_doubles_per_chain") = 2 T + 1. Synchronises `stream` before returning. */
int bcm3b200_cellpop_finish(void* handle, size_t num_chains, const double* d_partial, double* logp, int* status, void* stream);

/* ---- the exchange step inside the library (NCCL over NVLink / NVSwitch) ----
 * One process per GPU: every rank creates its handle with shard_rank = its rank and shard_count = the number of ranks,
 * rank 0 obtains an id with bcm3b200_comm_unique_id and hands it to the other ranks by whatever means the host program has
 * (MPI_Bcast, a file, a socket: it is BCM3B200_COMM_ID_BYTES of plain bytes = an ncclUniqueId), then EVERY rank calls
 * bcm3b200_comm_init (collective). From then on
 *   bcm3b200_evaluate_batch        returns the COMPLETE per-chain log-likelihoods on every rank (both model kinds): this
 *                                  rank's slice, one all-gather of the per-rank partial blocks stream-ordered behind the
 *                                  reduction kernel, their combination in rank order on the device -- bit-identical on all
 *                                  ranks and from run to run, which an all-reduce does not promise --, 3 C doubles back;
 *   bcm3b200_exchange_partials     is that exchange alone, in place on a device block produced by bcm3b200_enqueue_batch /
 *                                  bcm3b200_evaluate_batch_device ([3][C] for pop_pk_trajectory: SUM row 0, MIN rows 1-2;
 *                                  [C][2 T + 1] for cell_population: SUM), enqueued on `stream`, no synchronisation.
 * Every rank must make the same calls with the same num_chains (the exchange is a collective).
 * Without a communicator (shard_count == 1, or bcm3b200_comm_init never called) the exchange is a no-op and a sharded
 * handle yields its partial as described above. NCCL is loaded at run time (libnccl.so.2; BCM3B200_NCCL_LIB overrides).
 * Replaces: nothing in the reference (its evaluation threads share one address space, SamplerPT.cpp:438-475); this is the
 * ownership row of the multi-GPU design.
 * cell_population handles created with device_count > 1 spread their cells over the devices of ONE process and combine
 * them the same way over ncclCommInitAll communicators inside bcm3b200_evaluate_batch. */
#define BCM3B200_COMM_ID_BYTES 128
int bcm3b200_comm_unique_id(void* id, size_t id_bytes);
int bcm3b200_comm_init(void* handle, const void* id, size_t id_bytes);
int bcm3b200_exchange_partials(void* handle, size_t num_chains, double* d_partial, void* stream);

/* partial [3][num_chains] (host) -> logp[num_chains], status[num_chains] (may be NULL) */
int bcm3b200_combine_partials(size_t num_chains, const double* partial, double* logp, int* status);

/* Diagnostics of the LAST evaluate on this handle, for parity checks against the reference
 * (LikelihoodPopPKTrajectory::GetSimulatedConcentrations, .h:25; ODESolver::GetNumSteps, ODESolver.h:35).
 * Must be enabled before the evaluate with bcm3b200_set_option(h, "diagnostics", 1).
 *   conc       [num_chains][P_local][T] : conversion * trajectory(1, i), NaN where not simulated
 *   patient_ll [num_chains][P_local]
 *   counters   [num_chains][P_local][8] : steps, nfe, nsetups, nje, netf, ncfn, nni, ok
 * Any pointer may be NULL. */
int bcm3b200_get_diagnostics(void* handle, double* conc, double* patient_ll, int32_t* counters);

/* cell_population diagnostics of the LAST evaluate (any pointer may be NULL), with rows = stat "value_rows" (the timepoints of
 * all data sets of the handle one after the other; T for one data set) and cells = stat "cell_columns" (this handle's cells;
 * max_cells for a dividing population):
 *   cell_values [C][rows][cells] observed-species value of every simulated cell at every data timepoint (NaN = cell absent)
 *   cell_status [C][cells] 1 = solved, 0 = CVODE failure;  cell_steps [C][cells] accepted steps (ODESolver::GetNumSteps)
 *   population_average [C][rows] (DataLikelihoodTimeCoursePopulationAverage::population_average before offset/scale) */
int bcm3b200_get_cell_diagnostics(void* handle, double* cell_values, int32_t* cell_status, int32_t* cell_steps,
                                  double* population_average);

/* options: "diagnostics" (0/1, all kinds);
 *   pop_pk_trajectory: "block_size" (0 = auto, else a multiple of 32 up to 384), "sort_patients" (0/1: patients ranked by
 *     expected cost, results unchanged), "sort_min_systems", "chain_fastest_grid" (0/1: block order of the ranked launch);
 *   cell_population: "cellpop_kernel" (0 auto, 1 one cell per warp, 2 per thread, 3 per lane group), "cellpop_group_lanes"
 *     (0 auto), "cellpop_rhs_lanes" (0 never, 1 where it pays, 2 always: lane-parallel form of the generated right-hand
 *     side, bit-identical to the text as it stands) -- the last two before finalize; "cellpop_steps_report". */
int bcm3b200_set_option(void* handle, const char* name, int64_t value);

/* stats: "num_patients_local", "patient_offset", "last_kernel_launches", "total_kernel_launches",
 * "num_evaluations" (chains evaluated so far = the reference's num_likelihood_evaluations, Sampler.cpp:169),
 * "last_kernel_us" (device time of the last host-buffer evaluate's kernels, microseconds, max over devices);
 * cell_population: "num_cells_local", "cell_columns", "value_rows", "partial_doubles_per_chain" (= 2 value_rows + 1);
 * pharmaco_population: "num_compartments". */
int bcm3b200_get_stat(void* handle, const char* name, int64_t* value);

void bcm3b200_destroy(void* handle);

/* Page-locked host memory for the caller's `values` / `logp` buffers: evaluate_batch copies straight from it
 * with asynchronous DMA instead of staging pageable memory. Plain malloc'ed buffers work too, only slower. */
void* bcm3b200_host_alloc(size_t bytes);
void bcm3b200_host_free(void* p);

/* Measured FP64 FMA throughput of `device` in TFLOP/s (dependent-chain-free DFMA microbenchmark, best of 5 after one
 * warm-up): the denominator of the FP64 roofline fraction reported by bench.py. */
int bcm3b200_measure_fp64_peak(int device, double* tflops);

/* text of the last error raised on the calling thread ("" if none) */
const char* bcm3b200_last_error(void);

/* number of usable CUDA devices (0 when there is no driver / no GPU) */
int bcm3b200_device_count(void);

/* The host-side step of the per-cell time_course likelihood (cell_population, data_kind = time_course), exported for callers
 * and tests that want it on its own: the assignment of n observed cells (rows of cost [n][n]) to n simulated cells (columns)
 * that the reference's hungarianMinimumWeightPerfectMatching returns for the complete edge list
 * (dependencies/hungarian2/hungarian.cpp as DataLikelihoodTimeCourse.cpp:323 calls it -- not always the minimum-cost
 * matching, see bcm3_b200/csrc/matching_host.cuh). match [n]: the column of every row; BCM3B200_ERR_STATE (match = -1) when no
 * perfect matching was found. Pure host code: needs no device. */
int bcm3b200_match_cells(int n, const double* cost, int32_t* match);

#ifdef __cplusplus
}
#endif

#endif /* BCM3B200_H */
