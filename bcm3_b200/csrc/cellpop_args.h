// cellpop_args.h -- launch arguments shared by libbcm3b200.so and the per-model kernel library that is compiled
// at PostInitialize time from the generated RHS text (the reference compiles and dlopens its generated code the
// same way: src/cellpop/SolverCodeGenerator.cpp:390,407-414).
#pragma once

#include <stdint.h>

#define CP_MAX_VARIABILITY 8
#define CP_GROUP_SCALARS 33 /* per-cell scalar slots of cellpop_group.cuh before the parameter overrides (enum SC_*) */

// apply types, VariabilityDescriptionVariable.cpp:172-207
enum {
	CP_APPLY_ADDITIVE = 0,
	CP_APPLY_ADDITIVE_LOG,
	CP_APPLY_ADDITIVE_LOG2,
	CP_APPLY_MULTIPLICATIVE,
	CP_APPLY_MULTIPLICATIVE_LOG,
	CP_APPLY_MULTIPLICATIVE_LOG2,
	CP_APPLY_REPLACE
};

struct CpArgs {
	// batch
	int num_chains, num_cells, cell_offset; // cells of this shard, global index of the first one
	const double* transformed;   // [C][nvar] transformed variables (VariableSet::TransformVariable applied)
	int nvar;
	// model
	const double* initial_conditions; // [N]
	const double* constant_species;   // [Nc]
	const double* non_sampled;        // [Nn]
	// per-cell variability (one diagonal_gaussian block): quasi-random uniforms, row = cell
	const double* sobol; // [num_cells_total][D]
	int D;
	int var_scale_ix[CP_MAX_VARIABILITY];    // variable index of the scale, or -1
	double var_scale_fixed[CP_MAX_VARIABILITY];
	int var_apply[CP_MAX_VARIABILITY];
	int var_negate[CP_MAX_VARIABILITY];
	int var_slot[CP_MAX_VARIABILITY];        // override slot (parameter; -1: a dimension without a target) or species index (initial condition)
	int var_is_ic[CP_MAX_VARIABILITY];
	int var_full;           // 1: full_gaussian -- v = L z with the chain's Cholesky factor (VariabilityDescription.cpp:99-131)
	const double* var_chol; // [C][D][D] row-major lower triangle, written per batch by cellpop_cholesky_kernel
	// time
	int entry_time_ix;     // variable index or -1
	double entry_time_fixed;
	const double* timepoints; // [T] absolute times, sorted
	int T;
	// absolute time every cell is integrated to: the last timepoint of ALL data sets of the experiment (Experiment.cpp:655-656);
	// equals timepoints[T - 1] when this data set is the only or the longest one
	double sim_end_time;
	// solver (Experiment.cpp:411-416, Cell.cpp:70-74)
	double rel_tol, abs_tol, min_dt;
	double max_dt_inv; // CVodeSetMaxStep(solver_max_timestep): 1 / hmax, 0 = no ceiling (Experiment.cpp:413, Cell.cpp:73, cvode_io.c:344-376)
	int max_steps;
	// observed species: sum of these ODE-integrated species per timepoint
	int num_obs_species;
	int obs_species[8];
	// Several data sets of one experiment share its cells' integration (group kernel built with CP_NUM_DATASETS = K > 1,
	// Experiment.cpp:190-214, 298-312): `timepoints` is then the sorted union of their timepoints, tp_rows [T][K] says which row
	// of cell_values the value of data set k at union time u goes to (-1: data set k does not ask for that time), and every data
	// set sums its own observed species. num_rows = rows of cell_values per chain (= T with one data set).
	int num_data_sets, num_rows;
	const int32_t* tp_rows;
	int num_obs_species_more[3];
	int obs_species_more[3][8];
	// outputs
	double* cell_values; // [C][num_rows][cell_stride]  (NaN where the cell does not exist at that time)
	int32_t* cell_status; // [C][num_cells] 1 = ok, 0 = solver failure
	int32_t* cell_steps;  // [C][num_cells] or null
	int debug_report;     // 0: cell_steps = accepted steps; 1: RHS evaluations (nfe); 2: linear setups; 3: Jacobian evaluations
	const int32_t* cell_order; // [num_cells] or null: the order in which the group kernel hands out this shard's cells
	// one <treatment_trajectory type="pulses"> (TreatmentTrajectoryPulses.cpp): constant species it drives (-1: none)
	int treatment_species, treatment_num_pulses;
	const double* treatment_times; // [treatment_num_pulses] sorted
	// ---- dividing and dying cells (group kernel built with CP_DIVISION; Experiment.cpp:726-782, CellPopulation.cpp:36-104,
	// Cell.cpp:119-148, 463-538, the branch without stored integration points). The population is integrated one generation at
	// a time: `items` lists the (chain, slot) pairs of the generation, the per-cell records say when a cell was created, which
	// quasi-random row it takes and which cell it inherits its state from; a cell whose "cytokinesis" / "apoptosis" species
	// exceeds 1 after an accepted step ends there and leaves its state behind for its daughters.
	int var_only_initial[CP_MAX_VARIABILITY]; // <variable only_initial_cells="true">: skipped unless the cell's "initial cell" flag is set
	int initial_flag;     // the flag of the experiment's first cells: 1 when num_cells > 1, 0 for a single cell (Experiment.cpp:662-670); daughters: 0
	int cell_stride;      // cell columns of cell_values / cell_status / cell_steps and of the records below (num_cells when nothing divides)
	int cytokinesis_ix, apoptosis_ix; // ODE species indices, -1: none
	int reset_ix[7];      // species a daughter resets to 0, 1, 1, 1, 0, 0, 0 (Cell.cpp:127-133)
	const int32_t* items; // [num_items][2] (chain, slot), or null: every cell of every chain
	int num_items;
	const double* cell_creation; // [C][stride] absolute creation time
	const int32_t* cell_row;     // [C][stride] quasi-random row
	const int32_t* cell_parent;  // [C][stride] slot of the parent, -1: an initial cell
	double* cell_end_y;          // [C][stride][N] y after the step that ended a dividing cell
	double* cell_end_time;       // [C][stride] absolute time at which the cell's integration ended (division, death or the end of the experiment)
	int32_t* cell_event;         // [C][stride] 0 none, 1 divided, 2 died
	// Cell::EnteredMitosis (Cell.h:27; Cell.cpp:487-492): the cell's "nuclear_envelope" species was below 0.5 after some accepted
	// step. nuclear_envelope_ix = its ODE species index (-1: not tracked), cell_mitotic [C][stride] receives the flag (null: none)
	int nuclear_envelope_ix;
	int32_t* cell_mitotic;
};

#ifdef __CUDACC__
// VariabilityDescription::GetPseudorandomVector (VariabilityDescription.cpp:50-139): component d of the cell's vector
__device__ __forceinline__ double cellpop_variability_value(const CpArgs& a, const double* tv, int c, long long gcell, int d)
{
	if (a.var_full) {
		const double* L = a.var_chol + ((long long)c * a.D + d) * a.D;
		double v = 0.0;
		for (int j = 0; j <= d; j++) v += L[j] * normcdfinv(a.sobol[gcell * a.D + j]);
		return v;
	}
	// diagonal_gaussian: the scale is on log scale (:54-64)
	const double scale = (a.var_scale_ix[d] >= 0) ? tv[a.var_scale_ix[d]] : a.var_scale_fixed[d];
	return normcdfinv(a.sobol[gcell * a.D + d]) * exp(scale);
}
#endif

typedef int (*cellpop_launch_fn)(const CpArgs* args, void* stream);
typedef int (*cellpop_thread_launch_fn)(const CpArgs* args, double* scratch, void* stream);
typedef long long (*cellpop_thread_scratch_fn)(int num_chains, int num_cells);
typedef int (*cellpop_group_launch_fn)(const CpArgs* args, double* scratch, void* stream);
typedef long long (*cellpop_group_scratch_fn)(int num_chains, int num_cells);
typedef int (*cellpop_group_info_fn)(int* lanes_per_cell, int* threads_per_block, int* smem_bytes_per_block);
