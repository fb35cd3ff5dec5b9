// cellpop_host.cuh -- host side of the cell_population evaluator inside libbcm3b200.so:
//   * per-model kernel module: generated RHS text -> translation unit -> nvcc -> dlopen, the same mechanism the
//     reference uses for its CPU code (src/cellpop/SolverCodeGenerator.cpp:32-100 writes code.cpp, :390 runs
//     `cmake . ; make`, :407-414 dlopens the result), cached by content hash like its codegen_<name>/ directory;
//   * model-independent kernels: variable transform (CellPopulationLikelihood.cpp:82-101), population average and
//     data likelihood (DataLikelihoodTimeCoursePopulationAverage.cpp:85-197, DataLikelihoodTimeCourseBase.cpp:229-315).
#pragma once

#include <dlfcn.h>
#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <thread>

#include "cellpop_args.h"
#include "matching_host.cuh"

namespace bcm3b200 {

enum : int { CP_ERR_NORMAL = 0, CP_ERR_STUDENT_T4 = 1, CP_ERR_PROPORTIONAL_NORMAL = 2, CP_ERR_ADDITIVE_PROPORTIONAL_NORMAL = 3 };

struct CellPopState {
	// description
	// <data type=>: 0 = time_course_population_average; 1 = time_course -- one observed trajectory per cell ("observed" is
	// [observed cells][T], num_replicates = observed cells = num_cells), every observed cell matched to one simulated cell
	// (DataLikelihoodTimeCourse.cpp:230-365): the [observed x simulated] log-likelihood block is a kernel, the matching runs
	// on the host (matching_host.cuh)
	// 2 = time_points (DataLikelihoodTimePoints.cpp:209-345): "observed" is [observed cell slots][T], NaN = no such cell at that
	// timepoint; at every timepoint the observed cells present are matched to the simulated cells that have a value there
	int data_kind = 0;
	// time_course: <data optimize_offset_scale=...> (DataLikelihoodTimeCourseBase.cpp:43-57, 317-322)
	bool optimize_offset_scale = false;
	double optimize_offset_min = -1.0, optimize_offset_max = 1.0, optimize_scale_min = 0.1, optimize_scale_max = 10.0;
	int saturation_scale_ix = -1; // time_course: <data saturation_scale="variable">, DataLikelihoodTimeCourse.cpp:243-254
	// <data include_only_cells_that_went_through_mitosis="true"> (population average, DataLikelihoodTimeCoursePopulationAverage.cpp:
	// 171-176) needs Cell::EnteredMitosis: the "nuclear_envelope" species (descriptor key nuclear_envelope_species) below 0.5 after
	// some accepted step (Cell.cpp:487-492) -- tracked by the event code of the kernel (CP_DIVISION builds)
	bool include_only_mitotic = false;
	int nuclear_envelope_ix = -1;
	bool track_mitosis() const
	{
		if (nuclear_envelope_ix < 0) return false;
		if (include_only_mitotic) return true;
		for (const auto& m : more)
			if (m->include_only_mitotic) return true;
		return false;
	}
	bool use_only_nondivided = false; // time_points with dividing cells: daughters are left out (DataLikelihoodTimePoints.cpp:349-351)
	int value_relative_to_timepoint_ix = -1; // time_points: simulated value = (x + offset) / x(that timepoint) * scale (DataLikelihoodBase.cpp:49)
	int N = 0, Nc = 0, nvar = 0, Nn = 0, num_cells = 0, T = 0, R = 1, D = 0;
	int entry_time_ix = -1;
	double entry_time_fixed = 0.0;
	double rel_tol = 4.0 * 1.1920928955078125e-07, abs_tol = 4.0 * 1.1920928955078125e-07; // 4 * FLT_EPSILON, Experiment.cpp:415-416
	double min_dt = 1e-8;                                                                   // Experiment.cpp:412
	double max_dt = std::numeric_limits<double>::infinity();                                // Experiment.cpp:413 solver_max_timestep
	int max_steps = 10000;                                                                  // Experiment.cpp:414
	int error_model = CP_ERR_NORMAL;
	double weight = 1.0;
	int stdev_ix = -1, offset_ix = -1, scale_ix = -1, prop_stdev_ix = -1;
	double stdev_fixed = 1.0, offset_fixed = 0.0, scale_fixed = 1.0, prop_stdev_fixed = 1.0;
	double missing_simulation_time_stdev = 300.0; // DataLikelihoodTimeCourseBase.cpp:22
	bool have_sim_end_time = false; // descriptor key simulation_end_time: the experiment's last requested time over ALL its data sets
	double sim_end_time = 0.0;
	bool full_gaussian = false;                   // <cell_variability distribution="full_gaussian">
	bool relative_to_time_average = false;        // <data relative_to_time_average="true">
	bool stdev_relative_to_scale = false;         // <data stdev_relative_to_scale="true">: stdev *= data scale (DataLikelihoodBase.cpp:151-153)
	int steps_report = 0;                         // option "cellpop_steps_report": what get_cell_diagnostics returns as cell_steps (0 steps, 1 nfe, 2 nsetups, 3 nje)
	int treatment_species = -1;                   // <treatment_trajectory type="pulses" species_name=...>: constant species index
	std::vector<int> obs_species;
	int shard_rank = 0, shard_count = 1, device = 0;
	// <experiment divide_cells="true" max_cells=> and the species the reference's cell events are built around (by index):
	// Experiment.cpp:726-782, Cell.cpp:119-148, 463-538
	bool divide_cells = false;
	int max_cells = 0, cytokinesis_ix = -1, apoptosis_ix = -1, sobol_rows = 0;
	int reset_ix[7] = { -1, -1, -1, -1, -1, -1, -1 };
	bool division() const { return (divide_cells && cytokinesis_ix >= 0) || apoptosis_ix >= 0 || track_mitosis(); } // the model library carries the event code
	// Further data sets of the same experiment (descriptor keys with the suffix @1, @2, @3; data "timepoints@k", "observed@k"):
	// they share the integration of the experiment's cells (Experiment.cpp:190-214, 298-312) -- the kernel interpolates at the
	// union of all timepoints and every data set sums its own species -- and their log-likelihoods are added in order (:346-355)
	struct MoreData {
		int T = 0, R = 1, error_model = CP_ERR_NORMAL, data_kind = 0, value_relative_to_timepoint_ix = -1;
		bool optimize_offset_scale = false;
		double optimize_offset_min = -1.0, optimize_offset_max = 1.0, optimize_scale_min = 0.1, optimize_scale_max = 10.0;
		int saturation_scale_ix = -1;
		bool use_only_nondivided = false, include_only_mitotic = false;
		// >= 0: not a data set of its own but a further MARKER (species_name="a;b": the part after a ';') of the per-cell data set
		// with that index (0 = the handle's first data set, j = more[j - 1]): its rows, observed block and stdev / offset / scale
		// entries enter that data set's cell likelihoods (DataLikelihoodTimeCourse.cpp:449-489, DataLikelihoodTimePoints.cpp:264-289)
		int marker_of = -1;
		// >= 0: the rows of this entry hold the DENOMINATOR of the log ratio (use_log_ratio, species_name="a/b",
		// DataLikelihoodTimeCourse.cpp:380-397) of the time_course data set or marker with that index (same numbering)
		int denominator_of = -1;
		bool rides() const { return marker_of >= 0 || denominator_of >= 0; } // value rows for another entry, no term of its own
		int stdev_ix = -1, offset_ix = -1, scale_ix = -1, prop_stdev_ix = -1;
		double stdev_fixed = 1.0, offset_fixed = 0.0, scale_fixed = 1.0, prop_stdev_fixed = 1.0, weight = 1.0, missing_stdev = 300.0;
		bool relative_to_time_average = false, stdev_relative_to_scale = false;
		std::vector<int> obs_species;
		std::vector<double> timepoints, observed;
		DevBuf<double> d_time, d_obs;
	};
	std::vector<std::unique_ptr<MoreData>> more;
	int rows() const // rows of cell_values / population averages per chain: every data set's timepoints one after the other
	{
		int r = T;
		for (const auto& m : more) r += m->T;
		return r;
	}
	int TU = 0; // timepoints the kernel interpolates at (the sorted union; = T with one data set)
	DevBuf<double> d_union_time;
	DevBuf<int32_t> d_tp_rows;
	int capacity() const { return division() ? (divide_cells ? max_cells : cells_local) : cells_local; }
	DevBuf<double> d_creation, d_end_y, d_end_time;
	DevBuf<int32_t> d_row, d_parent, d_event, d_items, d_wave, d_item_offsets, d_mitotic, d_row_only_mitotic;
	std::vector<int32_t> h_wave;
	std::string derivative_code;
	std::map<std::string, std::vector<double>> data;
	// derived
	bool finalized = false;
	int cell_offset = 0, cells_local = 0;
	void* module = nullptr;
	uint64_t module_key = 0; // hash of the module source the loaded library was built from
	cellpop_launch_fn launch = nullptr;
	cellpop_thread_launch_fn thread_launch = nullptr;
	cellpop_thread_scratch_fn thread_scratch = nullptr;
	cellpop_group_launch_fn group_launch = nullptr;
	cellpop_group_scratch_fn group_scratch = nullptr;
	cellpop_group_info_fn group_info = nullptr;
	int kernel_choice = 0; // 0 auto (lane groups for N <= 96, one cell per warp above), 1 warp, 2 thread, 3 group
	int built_kernel = 0;  // the one kernel the model library was compiled with (1, 2 or 3)
	int group_lanes = 0;   // 0 auto: smallest power of two with ceil(N / G) <= 3
	int rhs_lanes = 1;     // option "cellpop_rhs_lanes": lane-parallel right-hand side (cellpop_lane_rhs): 0 never, 1 where it pays (16 or 32 lanes per cell), 2 wherever the text can be regrouped
	DevBuf<double> d_scratch;
	std::string module_path;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	DevBuf<double> d_ic, d_const, d_nonsampled, d_sobol, d_time, d_obs, d_values, d_transformed, d_cellvals, d_avg, d_logp;
	DevBuf<int32_t> d_transforms, d_status, d_steps, d_count, d_nfail, d_cov_ix, d_cell_order;
	DevBuf<double> d_cov_fixed, d_chol, d_treatment_times, d_partial, d_cell_lik;
	std::vector<double> h_cell_lik; // [C][observed][simulated] of the time_course data set being matched
	bool any_time_course() const
	{
		if (data_kind != 0) return true;
		for (const auto& m : more)
			if (m->data_kind != 0) return true;
		return false;
	}
	bool diagnostics = false;
	int last_C = 0;
	double last_kernel_ms = 0.0;
	int64_t total_launches = 0, last_launches = 0, num_evaluations = 0;
	CpArgs args;

	~CellPopState()
	{
		if (ev0) cudaEventDestroy(ev0);
		if (ev1) cudaEventDestroy(ev1);
		if (stream) cudaStreamDestroy(stream);
		// the module stays loaded: unloading a library that registered CUDA kernels is not safe
	}
};

// ---------------------------------------------------------------------------------------------------------------
// model-independent kernels

// transformed[c][i] = TransformVariable(i, values[c][i])
__global__ void cellpop_transform_kernel(const double* __restrict__ values, const int32_t* __restrict__ transforms, int nvar, int C,
                                         double* __restrict__ out)
{
	const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= (long long)nvar * C) return;
	out[e] = transform_variable(transforms[e % nvar], values[e]);
}

// per (chain, timepoint): population size = cells that exist at that time (non-NaN value), average = sum_i x_i / size in
// a fixed tree order; per chain: number of failed cells (block t == 0 only)
// mitotic [C][num_cells] / row_only_mitotic [T] (both or neither): rows with the flag average the cells that entered mitosis only,
// over the number of such cells that exist at the timepoint (DataLikelihoodTimeCoursePopulationAverage.cpp:171-176)
__global__ void cellpop_average_kernel(const double* __restrict__ cell_values, const int32_t* __restrict__ status, int num_cells, int T,
                                       double* __restrict__ avg, int32_t* __restrict__ count, int32_t* __restrict__ nfail,
                                       const int32_t* __restrict__ mitotic = nullptr, const int32_t* __restrict__ row_only_mitotic = nullptr)
{
	__shared__ double sh[256];
	__shared__ int shi[256];
	const int t = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
	const double* v = cell_values + ((long long)c * T + t) * num_cells;
	const int32_t* mit = (mitotic && row_only_mitotic && row_only_mitotic[t]) ? mitotic + (long long)c * num_cells : nullptr;
	int n = 0;
	for (int i = tid; i < num_cells; i += blockDim.x) n += (isnan(v[i]) || (mit && !mit[i])) ? 0 : 1;
	shi[tid] = n;
	__syncthreads();
	for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
		if (tid < off) shi[tid] += shi[tid + off];
		__syncthreads();
	}
	const int pop = shi[0];
	__syncthreads();
	double s = 0.0;
	for (int i = tid; i < num_cells; i += blockDim.x) {
		const double x = v[i];
		if (!isnan(x) && !(mit && !mit[i])) s += x / (double)pop; // NotifySimulatedValue divides every value by the population size, .cpp:161-197
	}
	sh[tid] = s;
	__syncthreads();
	for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
		if (tid < off) sh[tid] += sh[tid + off];
		__syncthreads();
	}
	if (tid == 0) {
		avg[c * T + t] = sh[0];
		count[c * T + t] = pop;
	}
	if (t == 0) {
		int f = 0;
		const int32_t* st = status + (long long)c * num_cells;
		for (int i = tid; i < num_cells; i += blockDim.x) f += st[i] ? 0 : 1;
		__syncthreads();
		shi[tid] = f;
		__syncthreads();
		for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
			if (tid < off) shi[tid] += shi[tid + off];
			__syncthreads();
		}
		if (tid == 0) nfail[c] = shi[0];
	}
}

// Sharded form of the same reduction (cells split over ranks): per (chain, timepoint) the shard's raw sum and count of
// existing cells, per chain its failed cells -- partial[c][0..T) sums, [T..2T) counts, [2T] failures, all as doubles so
// that one SUM all-reduce combines the shards. Fixed tree order inside the shard.
__global__ void cellpop_partial_kernel(const double* __restrict__ cell_values, const int32_t* __restrict__ status, int num_cells, int T,
                                       double* __restrict__ partial)
{
	__shared__ double sh[256];
	__shared__ int shi[256];
	const int t = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
	const double* v = cell_values + ((long long)c * T + t) * num_cells;
	double s = 0.0;
	int n = 0;
	for (int i = tid; i < num_cells; i += blockDim.x) {
		const double x = v[i];
		if (!isnan(x)) {
			s += x;
			n++;
		}
	}
	sh[tid] = s;
	shi[tid] = n;
	__syncthreads();
	for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
		if (tid < off) {
			sh[tid] += sh[tid + off];
			shi[tid] += shi[tid + off];
		}
		__syncthreads();
	}
	double* out = partial + (long long)c * (2 * T + 1);
	if (tid == 0) {
		out[t] = sh[0];
		out[T + t] = (double)shi[0];
	}
	if (t == 0) {
		__syncthreads();
		int f = 0;
		const int32_t* st = status + (long long)c * num_cells;
		for (int i = tid; i < num_cells; i += blockDim.x) f += st[i] ? 0 : 1;
		shi[tid] = f;
		__syncthreads();
		for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
			if (tid < off) shi[tid] += shi[tid + off];
			__syncthreads();
		}
		if (tid == 0) out[2 * T] = (double)shi[0];
	}
}

// combined partial -> population average and failure count in the layout the data-likelihood kernel reads
__global__ void cellpop_unpack_partial_kernel(const double* __restrict__ partial, int T, int C, double* __restrict__ avg, int32_t* __restrict__ count,
                                              int32_t* __restrict__ nfail)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= C * (T + 1)) return;
	const int c = i / (T + 1), t = i % (T + 1);
	const double* in = partial + (long long)c * (2 * T + 1);
	if (t < T) {
		const double n = in[T + t];
		avg[c * T + t] = (n > 0.0) ? in[t] / n : 0.0;
		count[c * T + t] = (int32_t)n;
	} else {
		nfail[c] = (int32_t)in[2 * T];
	}
}

// ---- dividing populations: one generation at a time (Experiment::ParallelSimulation / SimulateCell, Experiment.cpp:691-782) ----
// wave [C][4]: first slot of the generation being integrated, one past its last slot, overflow flag, unused

// Generation 0: the experiment's initial cells (Experiment.cpp:662-670) and the reset of every per-cell record
__global__ void cellpop_division_init_kernel(int C, int n0, int stride, const double* __restrict__ transformed, int nvar, int entry_time_ix,
                                             double entry_time_fixed, int row0, double* __restrict__ creation, int32_t* __restrict__ row,
                                             int32_t* __restrict__ parent, int32_t* __restrict__ event, int32_t* __restrict__ status,
                                             int32_t* __restrict__ steps, int32_t* __restrict__ items, int32_t* __restrict__ wave)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (long long)C * stride) return;
	const int c = (int)(i / stride), slot = (int)(i % stride);
	event[i] = 0;
	status[i] = 1;
	steps[i] = 0;
	parent[i] = -1;
	row[i] = row0 + slot;
	creation[i] = (entry_time_ix >= 0) ? transformed[(long long)c * nvar + entry_time_ix] : entry_time_fixed;
	if (slot < n0) {
		items[2 * ((long long)c * n0 + slot)] = c;
		items[2 * ((long long)c * n0 + slot) + 1] = slot;
	}
	if (slot == 0) {
		wave[4 * c] = 0;
		wave[4 * c + 1] = n0;
		wave[4 * c + 2] = 0;
		wave[4 * c + 3] = 0;
	}
}

// After a generation: every cell of it that divided before the end of the experiment gets two daughters, in the order of the
// parents -- the order in which the reference's loop over the population appends them (CellPopulation::AddNewCell,
// CellPopulation.cpp:36-104). One block per chain. A daughter without a free cell object or quasi-random row fails the
// evaluation of the chain (AddNewCell returns max, SimulateCell returns false).
__global__ void cellpop_spawn_kernel(int stride, int n0, int sobol_rows, double target_time, const int32_t* __restrict__ event,
                                     const double* __restrict__ end_time, double* __restrict__ creation, int32_t* __restrict__ row,
                                     int32_t* __restrict__ parent, int32_t* __restrict__ wave)
{
	__shared__ int sh[256];
	__shared__ int base, overflow;
	const int c = blockIdx.x, tid = threadIdx.x;
	const int begin = wave[4 * c], end = wave[4 * c + 1];
	const long long o = (long long)c * stride;
	if (tid == 0) {
		base = 0;
		overflow = wave[4 * c + 2];
	}
	__syncthreads();
	for (int s0 = begin; s0 < end; s0 += blockDim.x) {
		const int slot = s0 + tid;
		const bool divides = slot < end && (event[o + slot] & 1) && end_time[o + slot] < target_time;
		sh[tid] = divides ? 1 : 0;
		__syncthreads();
		for (int off = 1; off < blockDim.x; off <<= 1) { // inclusive scan
			const int v = (tid >= off) ? sh[tid - off] : 0;
			__syncthreads();
			sh[tid] += v;
			__syncthreads();
		}
		const int rank = base + sh[tid] - (divides ? 1 : 0);
		if (divides) {
			for (int child = 0; child < 2; child++) {
				const int d = end + 2 * rank + child;
				const long long r = (long long)n0 + 2ll * row[o + slot] + child; // CellPopulation.cpp:75
				if (d >= stride || r >= sobol_rows) {
					overflow = 1;
				} else {
					creation[o + d] = end_time[o + slot];
					row[o + d] = (int32_t)r;
					parent[o + d] = slot;
				}
			}
		}
		__syncthreads();
		if (tid == 0) base += sh[blockDim.x - 1];
		__syncthreads();
	}
	if (tid == 0) {
		int next_end = end + 2 * base;
		if (next_end > stride) next_end = stride;
		wave[4 * c] = end;
		wave[4 * c + 1] = overflow ? end : next_end; // a failed chain integrates nothing more
		wave[4 * c + 2] = overflow;
	}
}

// the (chain, slot) list of the next generation; offsets [C] = exclusive prefix of the generations' sizes (host)
__global__ void cellpop_items_kernel(int C, const int32_t* __restrict__ wave, const int32_t* __restrict__ offsets, int32_t* __restrict__ items)
{
	const int c = blockIdx.x;
	const int begin = wave[4 * c], n = wave[4 * c + 1] - begin;
	for (int k = threadIdx.x; k < n; k += blockDim.x) {
		items[2 * ((long long)offsets[c] + k)] = c;
		items[2 * ((long long)offsets[c] + k) + 1] = begin + k;
	}
}

// a chain whose population outgrew max_cells (or the quasi-random table) counts as failed: logp = -inf (Experiment.cpp:356-358)
__global__ void cellpop_overflow_kernel(int C, const int32_t* __restrict__ wave, int32_t* __restrict__ nfail)
{
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c < C && wave[4 * c + 2]) nfail[c] += 1;
}

// Per chain: the Cholesky factor of the cell-variability covariance in the spherical parametrisation of
// VariabilityDescription.cpp:99-131 -- L(i, j) = exp(scale_i) * prod_{k < j} sin(pi c_ik) * [j < i] cos(pi c_ij), c_ik =
// covariance value (i - 1) i / 2 + k. One thread per chain; out [C][D][D] row-major, zeros above the diagonal.
struct CpCholArgs {
	int D, nvar;
	int scale_ix[CP_MAX_VARIABILITY];
	double scale_fixed[CP_MAX_VARIABILITY];
	const int32_t* cov_ix;   // [D (D - 1) / 2] variable index or -1
	const double* cov_fixed; // [D (D - 1) / 2]
};
__global__ void cellpop_cholesky_kernel(const CpCholArgs a, const double* __restrict__ transformed, int C, double* __restrict__ out)
{
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= C) return;
	const double* tv = transformed + (long long)c * a.nvar;
	double* L = out + (long long)c * a.D * a.D;
	for (int i = 0; i < a.D; i++) {
		const double exp_scale = exp((a.scale_ix[i] >= 0) ? tv[a.scale_ix[i]] : a.scale_fixed[i]);
		for (int j = 0; j < a.D; j++) {
			double l = 0.0;
			if (j <= i) {
				l = exp_scale;
				for (int k = 0; k < i; k++) {
					if (k <= j) {
						const int e = (i - 1) * i / 2 + k;
						const double cov_value = ((a.cov_ix[e] >= 0) ? tv[a.cov_ix[e]] : a.cov_fixed[e]) * 3.14159265358979323846;
						if (k == j) l *= cos(cov_value);
						else l *= sin(cov_value);
					}
				}
			}
			L[i * a.D + j] = l;
		}
	}
}

struct CpLikArgs {
	int avg_stride, avg_row0, accumulate; // the data set's averages are avg[c * avg_stride + avg_row0 + i]; accumulate: logp[c] += instead of =
	const double* avg;       // [C][avg_stride]
	const int32_t* nfail;    // [C]
	const double* transformed; // [C][nvar]
	const double* timepoints;  // [T]
	const double* observed;    // [R][T]
	int T, R, nvar, error_model, relative_to_time_average, stdev_relative_to_scale;
	int stdev_ix, offset_ix, scale_ix, prop_stdev_ix;
	double stdev_fixed, offset_fixed, scale_fixed, prop_stdev_fixed, weight, missing_stdev;
	double* logp; // [C]
	int kind;     // 1 / 2: a per-cell time_course / time_points data set -- its term is added on the host after the matching, here it is 0
};

// DataLikelihoodTimeCoursePopulationAverage::Evaluate (.cpp:85-159) for one species column; one thread per chain
__global__ void cellpop_datalik_kernel(const CpLikArgs a, int C)
{
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= C) return;
	if (a.nfail[c] > 0) { // Simulate() failed for some cell: Experiment.cpp:356-358
		a.logp[c] = -INFINITY;
		return;
	}
	if (a.kind != 0) {
		if (!a.accumulate) a.logp[c] = 0.0;
		return;
	}
	const double* avg = a.avg + (long long)c * a.avg_stride + a.avg_row0;
	const double* tv = a.transformed + (long long)c * a.nvar;
	double stdev = (a.stdev_ix >= 0) ? tv[a.stdev_ix] : a.stdev_fixed;
	const double offset = (a.offset_ix >= 0) ? tv[a.offset_ix] : a.offset_fixed;
	const double scale = (a.scale_ix >= 0) ? tv[a.scale_ix] : a.scale_fixed;
	if (a.stdev_relative_to_scale) stdev *= scale; // GetCurrentSTDev, DataLikelihoodBase.cpp:151-153
	const double prop_stdev = (a.prop_stdev_ix >= 0) ? tv[a.prop_stdev_ix] : a.prop_stdev_fixed;
	const double minus_log_sigma = -log(stdev);
	const double inv_two_sigma_sq = 1.0 / (2.0 * stdev * stdev);
	// offset / scale of the population average (.cpp:105-115); with relative_to_time_average the logarithm of every value
	// relative to the average over the timepoints
	double time_average = 0.0;
	if (a.relative_to_time_average) {
		for (int i = 0; i < a.T; i++) time_average += avg[i] + offset;
		time_average /= (double)a.T;
	}
	auto transformed_average = [&](int i) {
		double x = avg[i];
		if (a.relative_to_time_average) {
			x += offset;
			x = log(x / time_average);
			x *= scale;
		} else {
			x *= scale;
			x += offset;
		}
		return x;
	};
	double logp = 0.0;
	for (int i = 0; i < a.T; i++) {
		const double x = transformed_average(i);
		if (isnan(x)) {
			// missing-value penalty, .cpp:121-144
			double first_ok = a.timepoints[a.T - 1], last_ok = a.timepoints[0];
			for (int m = 0; m < a.T; m++)
				if (!isnan(transformed_average(m))) {
					first_ok = a.timepoints[m];
					break;
				}
			for (int m = a.T - 1; m >= 0; m--)
				if (!isnan(transformed_average(m))) {
					last_ok = a.timepoints[m];
					break;
				}
			const double time_offset = fmin(fabs(a.timepoints[i] - first_ok), fabs(a.timepoints[i] - last_ok));
			double pen;
			if (a.error_model == CP_ERR_STUDENT_T4) {
				pen = logpdf_tnu4(time_offset, 0.0, a.missing_stdev);
			} else {
				pen = -log(a.missing_stdev) - 0.91893853320467274178032973640562 - time_offset * time_offset / (2.0 * a.missing_stdev * a.missing_stdev);
			}
			for (int j = 0; j < a.R; j++)
				if (!isnan(a.observed[j * a.T + i])) logp += pen;
		} else {
			for (int j = 0; j < a.R; j++) {
				const double obs = a.observed[j * a.T + i];
				if (isnan(obs)) continue;
				// EvaluateValue(observed_data, x, 0): called with the arguments swapped: simulated := obs, observed := x
				if (a.error_model == CP_ERR_STUDENT_T4) {
					logp += logpdf_tnu4(x, obs, stdev);
				} else if (a.error_model == CP_ERR_PROPORTIONAL_NORMAL || a.error_model == CP_ERR_ADDITIVE_PROPORTIONAL_NORMAL) {
					// DataLikelihoodTimeCourseBase.cpp:281-287 with the swapped arguments: the proportional part scales with the datum
					double sigma = prop_stdev * fmax(obs, 0.0);
					if (a.error_model == CP_ERR_ADDITIVE_PROPORTIONAL_NORMAL) sigma = stdev + sigma;
					const double d = x - obs;
					logp += -log(sigma) - 0.91893853320467274178032973640562 - d * d / (2.0 * sigma * sigma);
				} else {
					const double d = x - obs;
					logp += minus_log_sigma - 0.91893853320467274178032973640562 - d * d * inv_two_sigma_sq;
				}
			}
		}
	}
	// Experiment.cpp:346-355: the data sets' log-likelihoods are added in order
	a.logp[c] = a.accumulate ? a.logp[c] + logp * a.weight : logp * a.weight;
}

// DataLikelihoodTimeCourse::CalculateCellLikelihood (.cpp:431-505) for every (observed cell i, simulated cell j) pair of a chain:
// lik[c][i][j] = sum over the markers and over the timepoints with an observation of the log-density of the observed value
// around the simulated cell's (scaled, shifted) trajectory; a simulated value that is missing pays
// CalculateMissingValueLikelihood (.cpp:566-588). One thread per pair, j fastest: the trajectory reads of a warp are one
// coalesced row segment, the observed value is a broadcast. cell_values is the kernel's [C][rows][cell_stride] block; every
// marker (species_name="a+b;c": the entries between the ';') has its own rows, observed block and stdev / offset / scale.
struct CpMarkerArgs {
	int row0, stdev_ix, offset_ix, scale_ix, prop_stdev_ix;
	int den_row0; // >= 0: use_log_ratio -- the value is 0.4342944819032518 * log(rows(row0) / rows(den_row0)), denominator >= 1e-16
	double stdev_fixed, offset_fixed, scale_fixed, prop_stdev_fixed;
	const double* observed; // [n_obs][T]
};
struct CpCellLikArgs {
	const double* cell_values;
	int rows, cell_stride, T, n_obs, n_sim, nvar, error_model, stdev_relative_to_scale;
	const double* transformed;
	const double* timepoints; // [T]
	int L;                    // markers, 1..4
	CpMarkerArgs mk[4];
	double missing_stdev;
	double* lik; // [C][n_obs][n_sim]
	// time_points (DataLikelihoodTimePoints.cpp:255-290): only timepoint `only_k` (>= 0), the simulated value optionally relative
	// to the cell's own value at timepoint `rel_k`, LogPdfNormal's division form; a simulated value that is missing in marker 0
	// gives NaN (the host leaves such cells out of the matching)
	int only_k, rel_k;
	// time_course with optimize_offset_scale: the observed trajectory regressed on the simulated one per pair and marker
	// (OptimizeOffsetScale, DataLikelihoodTimeCourseBase.cpp:317-322 = bcm3::linear_regress_columns, Correlation.cpp:158-200, + the clamps)
	int optimize;
	double opt_offset_min, opt_offset_max, opt_scale_min, opt_scale_max;
	int saturation_ix; // >= 0: signal saturation s / (1 + exp(-x)) - s / 2 with s = that transformed variable (.cpp:243-254)
	int c0; // first chain of this launch: grid z = chains c0 .. c0 + gridDim.z - 1, lik holds those chains only
};
__global__ void cellpop_cell_likelihood_kernel(const CpCellLikArgs a)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, cz = blockIdx.z, c = a.c0 + cz;
	if (j >= a.n_sim) return;
	const double* tv = a.transformed + (long long)c * a.nvar;
	double cell_logp = 0.0;
	bool no_value = false; // time_points: the simulated cell has no value in marker 0
	for (int l = 0; l < a.L; l++) {
		const CpMarkerArgs& mk = a.mk[l];
		double stdev = (mk.stdev_ix >= 0) ? tv[mk.stdev_ix] : mk.stdev_fixed;
		const double offset = (mk.offset_ix >= 0) ? tv[mk.offset_ix] : mk.offset_fixed;
		const double scale = (mk.scale_ix >= 0) ? tv[mk.scale_ix] : mk.scale_fixed;
		if (a.stdev_relative_to_scale) stdev *= scale;
		const double prop_stdev = (mk.prop_stdev_ix >= 0) ? tv[mk.prop_stdev_ix] : mk.prop_stdev_fixed;
		const double* traj = a.cell_values + ((long long)c * a.rows + mk.row0) * a.cell_stride + j;
		const double* obs = mk.observed + (long long)i * a.T;
		if (a.only_k >= 0) {
			const int k = a.only_k;
			double x = traj[(long long)k * a.cell_stride];
			if (a.rel_k >= 0) {
				x += offset;
				x /= traj[(long long)a.rel_k * a.cell_stride];
				x *= scale;
			} else {
				x *= scale;
				x += offset;
			}
			if (l == 0 && (isnan(traj[(long long)k * a.cell_stride]) || (a.rel_k >= 0 && isnan(traj[(long long)a.rel_k * a.cell_stride])))) no_value = true;
			const double y = obs[k];
			if (isnan(y)) continue; // DataLikelihoodTimePoints.cpp:275-279
			if (a.error_model == CP_ERR_NORMAL) {
				const double two_sigma_sq = 2.0 * stdev * stdev, d = y - x;
				cell_logp += -log(stdev) - 0.91893853320467274178032973640562 - d * d / two_sigma_sq;
			} else {
				cell_logp += logpdf_tnu4(y, x, stdev);
			}
			continue;
		}
		const double minus_log_sigma = -log(stdev);
		const double inv_two_sigma_sq = 1.0 / (2.0 * stdev * stdev);
		const double* dtraj = (mk.den_row0 >= 0) ? a.cell_values + ((long long)c * a.rows + mk.den_row0) * a.cell_stride + j : nullptr;
		auto value = [&](int k) { // NotifySimulatedValue's log ratio (.cpp:380-397), then .cpp:236-254: *= data scale, += data offset, the signal saturation
			double v = traj[(long long)k * a.cell_stride];
			if (dtraj) {
				const double den = dtraj[(long long)k * a.cell_stride];
				v = 0.4342944819032518276511289189166 * ((den < 1e-16) ? log(v / 1e-16) : log(v / den));
			}
			v *= scale;
			v += offset;
			if (a.saturation_ix >= 0) {
				const double s = tv[a.saturation_ix];
				v *= -1.0;
				v = exp(v);
				v += 1.0;
				v = 1.0 / v;
				v *= s;
				v -= 0.5 * s;
			}
			return v;
		};
		double opt_offset = 0.0, opt_scale = 1.0;
		if (a.optimize) {
			double mu_x = 0.0, mu_y = 0.0, xvar_calc = 0.0, cov_calc = 0.0, empirical_n = 0.0;
			for (int k = 0; k < a.T; k++) {
				const double xv = value(k), yv = obs[k];
				if (isnan(xv) || isnan(yv)) continue;
				empirical_n += 1.0;
				const double invN = 1.0 / empirical_n, mu_x_nm1 = mu_x, mu_y_nm1 = mu_y;
				mu_x += (xv - mu_x) * invN;
				mu_y += (yv - mu_y) * invN;
				if (empirical_n > 1) {
					const double ratio = (empirical_n - 1) / empirical_n, dx = xv - mu_x_nm1, dy = yv - mu_y_nm1;
					xvar_calc += dx * dx * ratio;
					cov_calc += dx * dy * ratio;
				}
			}
			if (empirical_n >= 2) {
				opt_scale = cov_calc / xvar_calc;
				opt_offset = mu_y - mu_x * opt_scale;
			}
			// std::min(std::max(v, lo), hi): a NaN slope (no variance in the simulated trajectory) stays NaN
			opt_scale = (opt_scale < a.opt_scale_min) ? a.opt_scale_min : opt_scale;
			opt_scale = (a.opt_scale_max < opt_scale) ? a.opt_scale_max : opt_scale;
			opt_offset = (opt_offset < a.opt_offset_min) ? a.opt_offset_min : opt_offset;
			opt_offset = (a.opt_offset_max < opt_offset) ? a.opt_offset_max : opt_offset;
		}
		for (int k = 0; k < a.T; k++) {
			const double y = obs[k];
			if (isnan(y)) continue;
			const double x = opt_offset + opt_scale * value(k);
			if (isnan(x)) {
				double first_ok = a.timepoints[a.T - 1], last_ok = a.timepoints[0];
				for (int m = 0; m < a.T; m++)
					if (!isnan(value(m))) {
						first_ok = a.timepoints[m];
						break;
					}
				for (int m = a.T - 1; m >= 0; m--)
					if (!isnan(value(m))) {
						last_ok = a.timepoints[m];
						break;
					}
				const double time_offset = fmin(fabs(a.timepoints[k] - first_ok), fabs(a.timepoints[k] - last_ok));
				if (a.error_model == CP_ERR_STUDENT_T4) cell_logp += logpdf_tnu4(time_offset, 0.0, a.missing_stdev);
				else cell_logp += -log(a.missing_stdev) - 0.91893853320467274178032973640562 - time_offset * time_offset / (2.0 * a.missing_stdev * a.missing_stdev);
			} else if (a.error_model == CP_ERR_NORMAL) {
				const double d = y - x;
				cell_logp += minus_log_sigma - 0.91893853320467274178032973640562 - d * d * inv_two_sigma_sq;
			} else if (a.error_model == CP_ERR_STUDENT_T4) {
				cell_logp += logpdf_tnu4(y, x, stdev);
			} else { // .cpp:272-283: sigma from the simulated value
				double sigma = prop_stdev * fmax(x, 0.0);
				if (a.error_model == CP_ERR_ADDITIVE_PROPORTIONAL_NORMAL) sigma += stdev;
				const double d = y - x;
				cell_logp += -log(sigma) - 0.91893853320467274178032973640562 - d * d * (1.0 / (2.0 * (sigma * sigma)));
			}
		}
		if (cell_logp == -INFINITY) break; // .cpp:485-488
	}
	if (no_value) cell_logp = NAN;
	a.lik[((long long)cz * a.n_obs + i) * a.n_sim + j] = cell_logp;
}

// ---------------------------------------------------------------------------------------------------------------
// Lane-parallel right-hand side.
//
// The reference's generator emits one statement per reaction, `ratelaws[r] = <expr>;`, then one per species,
// `out[i] = +ratelaws[a]-2.000000*ratelaws[b]...;` (SBMLModel.cpp:291-365). Evaluated as it stands, that text is SCALAR
// code: the lanes that share a cell all execute every rate law. Here the statements are regrouped by SHAPE -- the
// expression with its array indices and numeric literals replaced by placeholders -- and every shape becomes one loop in
// which the lanes of the cell's group evaluate different reactions of that shape at the same time, reading their indices and
// literals from a table; the species sums are then assembled by the lane that owns the species, term by term in the order
// of the text. Every rate law and every sum is the same sequence of IEEE operations on the same operands as in the
// original text (the module is compiled with -fmad=false), so the result is bit-identical to the scalar evaluation --
// tests/test_gpu_cellpop.py::test_lane_parallel_rhs_is_bit_identical checks exactly that.
// Anything the parser does not recognise makes it give up, and the model runs the scalar text unchanged.
struct LaneRhs {
	bool ok = false;
	int num_ratelaws = 0;
	int num_shapes = 0;
	int table_doubles = 0; // size of the tables when they are copied into shared memory
	std::string code; // tables + generated_ratelaws_lanes<G>() + generated_assemble()
};

namespace lane_rhs_detail {

inline std::string trim(const std::string& s)
{
	size_t a = s.find_first_not_of(" \t\r\n");
	if (a == std::string::npos) return "";
	size_t b = s.find_last_not_of(" \t\r\n");
	return s.substr(a, b - a + 1);
}
inline bool is_ident_start(char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_'; }
inline bool is_ident(char c) { return is_ident_start(c) || (c >= '0' && c <= '9'); }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

struct Law {
	int target = -1;
	std::string shape;             // expression with \x01 (index) / \x02 (literal) placeholders
	std::vector<int> idx;          // array indices in order of appearance
	std::vector<std::string> lit;  // literal texts in order of appearance
};

// expression -> shape; false on anything unexpected
inline bool scan(const std::string& e, Law& law)
{
	size_t i = 0;
	while (i < e.size()) {
		const char c = e[i];
		if (is_ident_start(c)) {
			size_t j = i;
			while (j < e.size() && is_ident(e[j])) j++;
			const std::string id = e.substr(i, j - i);
			if (id == "ratelaws" || id == "out") return false; // a rate law that depends on another statement: keep the text order
			if (id == "species" || id == "constant_species" || id == "parameters" || id == "non_sampled_parameters") {
				if (j >= e.size() || e[j] != '[') return false;
				size_t k = j + 1;
				while (k < e.size() && is_digit(e[k])) k++;
				if (k == j + 1 || k >= e.size() || e[k] != ']') return false;
				law.idx.push_back(atoi(e.substr(j + 1, k - j - 1).c_str()));
				law.shape += id + "[\x01]";
				i = k + 1;
			} else {
				law.shape += id; // a helper or a math function
				i = j;
			}
		} else if (is_digit(c)) {
			size_t j = i;
			while (j < e.size() && is_digit(e[j])) j++;
			bool real = false;
			if (j < e.size() && e[j] == '.') {
				real = true;
				j++;
				while (j < e.size() && is_digit(e[j])) j++;
			}
			if (j < e.size() && (e[j] == 'e' || e[j] == 'E')) {
				size_t k = j + 1;
				if (k < e.size() && (e[k] == '+' || e[k] == '-')) k++;
				if (k < e.size() && is_digit(e[k])) {
					real = true;
					while (k < e.size() && is_digit(e[k])) k++;
					j = k;
				}
			}
			if (j < e.size() && e[j] == 'f') return false; // single-precision literals: ODE_SINGLE_PRECISION builds are not supported
			if (real) {
				law.lit.push_back(e.substr(i, j - i));
				law.shape += "\x02";
			} else {
				law.shape += e.substr(i, j - i); // an integer stays part of the shape
			}
			i = j;
		} else {
			if (c == '[' || c == ']' || c == ';' || c == '{' || c == '}' || c == '=' || c == '"') return false;
			if (c != ' ' && c != '\t' && c != '\n' && c != '\r') law.shape += c;
			i++;
		}
	}
	return !law.shape.empty();
}

struct Term {
	int law;
	std::string coef; // signed literal text
};

// `+ratelaws[3]-2.000000*ratelaws[5]` | `0.0`
inline bool parse_sum(const std::string& e, std::vector<Term>& terms)
{
	const std::string t = trim(e);
	if (t == "0.0" || t == "0" || t == "0.000000") return true;
	size_t i = 0;
	while (i < t.size()) {
		char sign = '+';
		if (t[i] == '+' || t[i] == '-') sign = t[i++];
		else if (i != 0) return false;
		std::string coef = "1.0";
		if (i < t.size() && is_digit(t[i])) {
			size_t j = i;
			while (j < t.size() && (is_digit(t[j]) || t[j] == '.')) j++;
			if (j >= t.size() || t[j] != '*') return false;
			coef = t.substr(i, j - i);
			i = j + 1;
		}
		if (t.compare(i, 9, "ratelaws[") != 0) return false;
		size_t j = i + 9;
		size_t k = j;
		while (k < t.size() && is_digit(t[k])) k++;
		if (k == j || k >= t.size() || t[k] != ']') return false;
		terms.push_back(Term{ atoi(t.substr(j, k - j).c_str()), std::string(1, sign) + coef });
		i = k + 1;
	}
	return true;
}

} // namespace lane_rhs_detail

inline LaneRhs cellpop_lane_rhs(const std::string& code_without_jacobian, int N)
{
	using namespace lane_rhs_detail;
	LaneRhs out;
	const size_t open = code_without_jacobian.find('{');
	const size_t close = code_without_jacobian.rfind('}');
	if (open == std::string::npos || close == std::string::npos || close <= open) return out;
	const std::string body = code_without_jacobian.substr(open + 1, close - open - 1);
	int NR = -1;
	std::vector<Law> laws;
	std::vector<std::vector<Term>> sums(N);
	std::vector<bool> have_sum(N, false);
	size_t pos = 0;
	while (pos < body.size()) {
		size_t end = body.find(';', pos);
		if (end == std::string::npos) end = body.size();
		const std::string st = trim(body.substr(pos, end - pos));
		pos = end + 1;
		if (st.empty()) continue;
		if (st.compare(0, 17, "OdeReal ratelaws[") == 0) {
			NR = atoi(st.c_str() + 17);
			continue;
		}
		const bool is_law = st.compare(0, 9, "ratelaws[") == 0, is_out = st.compare(0, 4, "out[") == 0;
		if (!is_law && !is_out) return out;
		const size_t b0 = is_law ? 9 : 4;
		size_t b1 = b0;
		while (b1 < st.size() && is_digit(st[b1])) b1++;
		if (b1 == b0 || b1 >= st.size() || st[b1] != ']') return out;
		const int index = atoi(st.substr(b0, b1 - b0).c_str());
		size_t eq = st.find('=', b1);
		if (eq == std::string::npos || trim(st.substr(b1 + 1, eq - b1 - 1)) != "") return out;
		const std::string expr = trim(st.substr(eq + 1));
		if (is_law) {
			Law law;
			law.target = index;
			if (!scan(expr, law)) return out;
			laws.push_back(law);
		} else {
			if (index < 0 || index >= N || have_sum[index]) return out;
			if (!parse_sum(expr, sums[index])) return out;
			have_sum[index] = true;
		}
	}
	if (NR <= 0 || (int)laws.size() != NR) return out;
	{ // every rate law defined exactly once, every species assembled, every term refers to a rate law
		std::vector<int> seen(NR, 0);
		for (const Law& l : laws) {
			if (l.target < 0 || l.target >= NR || seen[l.target]++) return out;
		}
		for (int i = 0; i < N; i++) {
			if (!have_sum[i]) return out;
			for (const Term& t : sums[i])
				if (t.law < 0 || t.law >= NR) return out;
		}
	}
	// group by shape, in order of first appearance
	std::vector<std::string> shape_names;
	std::vector<std::vector<int>> members;
	for (size_t li = 0; li < laws.size(); li++) {
		size_t sidx = 0;
		while (sidx < shape_names.size() && shape_names[sidx] != laws[li].shape) sidx++;
		if (sidx == shape_names.size()) {
			shape_names.push_back(laws[li].shape);
			members.emplace_back();
		}
		members[sidx].push_back((int)li);
	}
	std::ostringstream tab_idx, tab_lit, tab_target, fn;
	int off_idx = 0, off_lit = 0, off_target = 0;
	fn << "template <int G_, class SP, class CS, class PP, class NS>\n"
	      "__device__ __forceinline__ void generated_ratelaws_lanes(int lg, double* ratelaws, const SP& species, const CS& constant_species, const PP& parameters, "
	      "const NS& non_sampled_parameters)\n{\n";
	for (size_t sidx = 0; sidx < shape_names.size(); sidx++) {
		const std::vector<int>& mem = members[sidx];
		const Law& first = laws[mem[0]];
		const size_t ni = first.idx.size(), nk = first.lit.size();
		// a placeholder whose value is the same in every member stays a constant of the code
		std::vector<int> idx_slot(ni, -1), lit_slot(nk, -1);
		int vi = 0, vk = 0;
		for (size_t k = 0; k < ni; k++) {
			bool same = true;
			for (int m : mem) same = same && laws[m].idx[k] == first.idx[k];
			if (!same) idx_slot[k] = vi++;
		}
		for (size_t k = 0; k < nk; k++) {
			bool same = true;
			for (int m : mem) same = same && laws[m].lit[k] == first.lit[k];
			if (!same) lit_slot[k] = vk++;
		}
		std::string expr;
		size_t ci = 0, ck = 0;
		for (char c : first.shape) {
			if (c == '\x01') {
				expr += (idx_slot[ci] < 0) ? std::to_string(first.idx[ci]) : ("CP_IDX(I + " + std::to_string(idx_slot[ci]) + ")");
				ci++;
			} else if (c == '\x02') {
				expr += (lit_slot[ck] < 0) ? first.lit[ck] : ("CP_LIT(K + " + std::to_string(lit_slot[ck]) + ")");
				ck++;
			} else {
				expr += c;
			}
		}
		fn << "\t// shape " << sidx << ": " << mem.size() << " reaction(s)\n";
		fn << "\tfor (int m = lg; m < " << mem.size() << "; m += G_) {\n";
		if (vi) fn << "\t\tconst int I = " << off_idx << " + m * " << vi << ";\n";
		if (vk) fn << "\t\tconst int K = " << off_lit << " + m * " << vk << ";\n";
		fn << "\t\tratelaws[CP_TARGET(" << off_target << " + m)] = " << expr << ";\n\t}\n";
		for (int m : mem) {
			for (size_t k = 0; k < ni; k++)
				if (idx_slot[k] >= 0) tab_idx << laws[m].idx[k] << ", ";
			for (size_t k = 0; k < nk; k++)
				if (lit_slot[k] >= 0) tab_lit << laws[m].lit[k] << ", ";
			tab_target << laws[m].target << ", ";
		}
		off_idx += vi * (int)mem.size();
		off_lit += vk * (int)mem.size();
		off_target += (int)mem.size();
	}
	fn << "}\n";
	std::ostringstream ob, ol, oc;
	int nterms = 0;
	for (int i = 0; i < N; i++) {
		ob << nterms << ", ";
		for (const Term& t : sums[i]) {
			ol << t.law << ", ";
			oc << t.coef << ", ";
			nterms++;
		}
	}
	ob << nterms;
	std::ostringstream o;
	o << "// ---- lane-parallel form of generated_derivative (" << shape_names.size() << " shapes, " << NR << " reactions), made by cellpop_lane_rhs ----\n";
	o << "#define CP_RHS_LANES 1\n#define CP_NUM_RATELAWS " << NR << "\n";
	// Tables: literals and coefficients (doubles), then indices, targets, sum boundaries and sum terms (ints). With
	// CP_TABLES_SHARED (set by cellpop_module_source when they fit behind the cells' blocks) the kernel copies them into shared
	// memory once per block and the accessors read that copy; otherwise they are read through the read-only cache.
	const int n_lit = off_lit + 1, n_coef = nterms + 1, n_idx = off_idx + 1, n_target = off_target + 1, n_begin = N + 1, n_law = nterms + 1;
	o << "#define CP_TAB_NLIT " << n_lit << "\n#define CP_TAB_NCOEF " << n_coef << "\n#define CP_TAB_NIDX " << n_idx << "\n#define CP_TAB_NTARGET " << n_target
	  << "\n#define CP_TAB_NBEGIN " << n_begin << "\n#define CP_TAB_NLAW " << n_law << "\n";
	o << "__device__ const int cp_rl_idx[] = { " << tab_idx.str() << "0 };\n";
	o << "__device__ const double cp_rl_lit[] = { " << tab_lit.str() << "0.0 };\n";
	o << "__device__ const int cp_rl_target[] = { " << tab_target.str() << "0 };\n";
	o << "__device__ const int cp_out_begin[] = { " << ob.str() << " };\n";
	o << "__device__ const int cp_out_law[] = { " << ol.str() << "0 };\n";
	o << "__device__ const double cp_out_coef[] = { " << oc.str() << "0.0 };\n";
	o << "#ifndef CP_TABLES_SHARED\n#define CP_TABLES_SHARED 0\n#endif\n"
	     "#if CP_TABLES_SHARED\n"
	     "extern __shared__ double smem_d[];\n"
	     "#define CP_TAB_INTS (reinterpret_cast<const int*>(smem_d + CP_TAB_BASE + CP_TAB_NLIT + CP_TAB_NCOEF))\n"
	     "#define CP_LIT(k) (smem_d[CP_TAB_BASE + (k)])\n#define CP_COEF(k) (smem_d[CP_TAB_BASE + CP_TAB_NLIT + (k)])\n"
	     "#define CP_IDX(k) (CP_TAB_INTS[(k)])\n#define CP_TARGET(k) (CP_TAB_INTS[CP_TAB_NIDX + (k)])\n"
	     "#define CP_BEGIN(k) (CP_TAB_INTS[CP_TAB_NIDX + CP_TAB_NTARGET + (k)])\n#define CP_LAW(k) (CP_TAB_INTS[CP_TAB_NIDX + CP_TAB_NTARGET + CP_TAB_NBEGIN + (k)])\n"
	     "#else\n"
	     "#define CP_LIT(k) __ldg(cp_rl_lit + (k))\n#define CP_COEF(k) __ldg(cp_out_coef + (k))\n#define CP_IDX(k) __ldg(cp_rl_idx + (k))\n"
	     "#define CP_TARGET(k) __ldg(cp_rl_target + (k))\n#define CP_BEGIN(k) __ldg(cp_out_begin + (k))\n#define CP_LAW(k) __ldg(cp_out_law + (k))\n"
	     "#endif\n";
	o << fn.str();
	// out[i] = the signed terms of the text, left to right (a coefficient of 1 multiplies exactly)
	o << "__device__ __forceinline__ double generated_assemble(int i, const double* ratelaws)\n{\n"
	     "\tint t = CP_BEGIN(i);\n\tconst int t1 = CP_BEGIN(i + 1);\n\tif (t == t1) return 0.0;\n"
	     "\tdouble acc = CP_COEF(t) * ratelaws[CP_LAW(t)];\n"
	     "\tfor (t++; t < t1; t++) acc = acc + CP_COEF(t) * ratelaws[CP_LAW(t)];\n\treturn acc;\n}\n";
	out.table_doubles = n_lit + n_coef + (n_idx + n_target + n_begin + n_law + 1) / 2;
	out.ok = true;
	out.num_ratelaws = NR;
	out.num_shapes = (int)shape_names.size();
	out.code = o.str();
	return out;
}

// ---------------------------------------------------------------------------------------------------------------
// per-model module

inline uint64_t fnv1a(const std::string& s)
{
	uint64_t h = 1469598103934665603ull;
	for (unsigned char ch : s) {
		h ^= ch;
		h *= 1099511628211ull;
	}
	return h;
}

inline std::string library_dir()
{
	Dl_info info;
	if (dladdr((void*)&fnv1a, &info) && info.dli_fname) {
		std::string p(info.dli_fname);
		size_t sl = p.find_last_of('/');
		return sl == std::string::npos ? std::string(".") : p.substr(0, sl);
	}
	return ".";
}

// Turns the reference generator's text into a device function template and wraps it into a translation unit.
// which of the three mappings the model library is built with; fixed at finalize (option "cellpop_kernel" or the
// BCM3B200_CELLPOP_KERNEL environment variable, else by size)
inline int cellpop_resolve_kernel(const CellPopState& cp)
{
	int choice = cp.kernel_choice;
	const char* kenv = getenv("BCM3B200_CELLPOP_KERNEL");
	if (kenv && !strcmp(kenv, "warp")) choice = 1;
	if (kenv && !strcmp(kenv, "thread")) choice = 2;
	if (kenv && !strcmp(kenv, "group")) choice = 3;
	if (choice < 1 || choice > 3) choice = (cp.N <= 96) ? 3 : 1;
	return choice;
}

inline int cellpop_module_source(const CellPopState& cp, const std::vector<int>& override_vars, std::string& src)
{
	std::string code = cp.derivative_code;
	// the (never installed) generated Jacobian uses host-only types: cut it off (Cell.cpp:57-76 never calls SetJacobianFunction)
	size_t jac = code.find("EXPORT_PREFIX void generated_jacobian");
	if (jac != std::string::npos) code = code.substr(0, jac);
	const std::string sig = "EXPORT_PREFIX void generated_derivative(OdeReal* out, const OdeReal* species, const OdeReal* constant_species, "
	                        "const OdeReal* parameters, const OdeReal* non_sampled_parameters)";
	size_t at = code.find(sig);
	if (at == std::string::npos)
		return fail(BCM3B200_ERR_ARG, "derivative_code does not contain the reference generator's generated_derivative signature (SBMLModel.cpp:295)");
	code.replace(at, sig.size(),
	             "template <class OUT, class SP, class CS, class PP, class NS>\n__device__ __forceinline__ void generated_derivative(OUT out, const SP& species, "
	             "const CS& constant_species, const PP& parameters, const NS& non_sampled_parameters)");
	// std::numeric_limits in generated text (none emitted by the rate-law printer today, kept for safety)
	std::ostringstream o;
	o << "// generated by libbcm3b200 for a cell_population model -- do not edit\n";
	o << "#include <cuda_runtime.h>\n#include <limits>\n";
	o << "#define CP_N " << cp.N << "\n";
	o << "#define CP_NUM_OVERRIDES " << override_vars.size() << "\n";
	o << "#define CP_PARAM_OVERRIDE_BODY";
	for (size_t s = 0; s < override_vars.size(); s++) o << " if (k == " << override_vars[s] << ") return ov[" << s << "];";
	o << "\n#define CP_OVERRIDE_INIT";
	for (size_t s = 0; s < override_vars.size(); s++) o << " S.params.ov[" << s << "] = tv[" << override_vars[s] << "];";
	o << "\n";
	// shared memory per warp grows with N^2: keep a block below ~96 KB
	const size_t per_warp = sizeof(double) * ((size_t)13 * cp.N + 2 * (size_t)cp.N * (cp.N + 1) + 24) + sizeof(int) * cp.N;
	int warps = 4;
	while (warps > 1 && per_warp * warps > 96 * 1024) warps >>= 1;
	o << "#define CP_WARPS_PER_BLOCK " << warps << "\n";
	// one cell per thread (small N): 12 N doubles of shared memory per thread
	o << "#define CP_THREADS_PER_BLOCK " << ((size_t)12 * cp.N * 64 * sizeof(double) <= 112 * 1024 ? 64 : 32) << "\n";
	// lane groups: G lanes per cell, every lane owns at most 3 components; warps per block sized for >= 2 blocks per SM
	int G = cp.group_lanes;
	if (const char* genv = getenv("BCM3B200_CELLPOP_GROUP")) G = atoi(genv);
	if (G != 2 && G != 4 && G != 8 && G != 16 && G != 32) {
		G = 2;
		while (G < 32 && (cp.N + G - 1) / G > 3) G <<= 1;
	}
	// block shape: ONE block per SM holding as many cells as its shared memory allows (the per-cell block of
	// cellpop_group.cuh), at most 12 warps -- above that the register budget (64K / threads) starts to force spills. One
	// large lock-step block shares instruction fetches best: measured 130.6 ms (1 x 12 warps) vs 136.7 (2 x 6) vs 146.0
	// (4 x 3) at config 3.
	const int Eg = (cp.N + G - 1) / G;
	// lane-parallel right-hand side (see cellpop_lane_rhs): the cell's shared block grows by one double per reaction
	LaneRhs lanes;
	{
		const char* renv = getenv("BCM3B200_CELLPOP_RHS_LANES");
		// default: models with 16 or 32 lanes per cell (measured: 50 species 1 933 -> 1 464 ms; with 4 lanes per cell and 8 cells
		// per warp the per-shape loops diverge between the groups and the scalar text is faster, 143 vs 178 ms at 12 species)
		const bool want = renv ? atoi(renv) != 0 : (cp.rhs_lanes == 2 || (cp.rhs_lanes == 1 && G >= 16));
		if (want && cellpop_resolve_kernel(cp) == 3) lanes = cellpop_lane_rhs(code.substr(at), cp.N);
	}
	// Newton matrices in global memory instead of the cells' shared blocks (CP_M_GLOBAL of cellpop_group.cuh): an experiment
	// switch, see DESIGN.md section 9
	bool m_global = false;
	if (const char* menv = getenv("BCM3B200_CELLPOP_M_GLOBAL")) m_global = atoi(menv) != 0;
	size_t per_cell;
	{ // the constants of cellpop_group.cuh: RS, OFF_SCAL, SC_COUNT, OFF_ZNH, OFF_RL, CS
		const int RS = cp.N | 1;
		const int off_scal = (m_global ? 0 : cp.N * RS) + 2 * cp.N + (cp.N + 1) / 2;
		const int sc_count = CP_GROUP_SCALARS + (override_vars.empty() ? 1 : (int)override_vars.size());
		int cs = off_scal + sc_count + 4 * Eg * G + (lanes.ok ? lanes.num_ratelaws : 0);
		while (cs % 16 != (RS * G) % 16) cs++;
		per_cell = sizeof(double) * (size_t)cs;
	}
	const int cells_per_warp = 32 / G;
	int warps_max = (int)((220 * 1024) / (per_cell * cells_per_warp));
	if (warps_max > 12) warps_max = 12;
	if (warps_max < 1) warps_max = 1;
	int gwarps = warps_max, gblocks = 1;
	if (const char* wenv = getenv("BCM3B200_CELLPOP_GROUP_WARPS")) gwarps = atoi(wenv) > 0 ? atoi(wenv) : gwarps;
	if (const char* benv = getenv("BCM3B200_CELLPOP_GROUP_MIN_BLOCKS")) gblocks = atoi(benv) > 0 ? atoi(benv) : gblocks;
	o << "#define CP_GROUP " << G << "\n";
	if (m_global) o << "#define CP_M_GLOBAL 1\n";
	o << "#define CP_GROUP_WARPS " << gwarps << "\n";
	if (lanes.ok) {
		// The tables of the lane-parallel right-hand side can be copied behind the cells' blocks when the 227 KB of a block have
		// room for them (BCM3B200_CELLPOP_TABLES_SHARED=1). Measured on B200, 50 species x 6 000 cells x 16 chains: 1 476 ms
		// with the copy in shared memory vs 1 459 ms through the read-only cache (long-scoreboard stalls 3.3 -> 2.4 cycles per
		// issue, instruction-fetch stalls 3.0 -> 4.3: nothing gained), so the default stays the read-only cache.
		const size_t cells_bytes = per_cell * cells_per_warp * gwarps, table_bytes = sizeof(double) * (size_t)lanes.table_doubles;
		bool shared_tables = false;
		if (const char* tenv = getenv("BCM3B200_CELLPOP_TABLES_SHARED")) shared_tables = atoi(tenv) != 0 && cells_bytes + table_bytes <= 227 * 1024;
		if (shared_tables) o << "#define CP_TABLES_SHARED 1\n#define CP_TAB_BASE " << (cells_bytes / sizeof(double)) << "\n";
	}
	o << "#define CP_GROUP_MIN_BLOCKS " << gblocks << "\n";
	// Block lock-step shares instruction fetches between the warps of a block but makes every trip as long as the slowest
	// warp's. With 8 or more cells per warp the warps are statistically alike and sharing wins (N = 12: 130 vs 150 ms,
	// N = 6: 13.5 vs 16.5 ms); with few cells per warp one linear setup (2 N^3 / 3 flops + N right-hand sides) stalls the
	// whole block (N = 24: 217 vs 183 ms, N = 50: 299 vs 201 ms without it).
	int lockstep = (G <= 4) ? 1 : 0;
	if (const char* lenv = getenv("BCM3B200_CELLPOP_GROUP_LOCKSTEP")) lockstep = atoi(lenv);
	o << "#define CP_GROUP_LOCKSTEP " << lockstep << "\n";
	// rate-law helpers: inlined for the small models that run in lock-step (N = 12: 131 vs 140 ms), real functions for the
	// large ones, whose right-hand side alone would outgrow the instruction cache (N = 50: 161 vs 199 ms)
	int helper_inline = lockstep;
	if (const char* henv = getenv("BCM3B200_CELLPOP_HELPER_INLINE")) helper_inline = atoi(henv);
	o << "#define CP_HELPER_INLINE " << helper_inline << "\n";
	if (cp.division()) o << "#define CP_DIVISION 1\n";
	if (!cp.more.empty()) o << "#define CP_NUM_DATASETS " << (1 + cp.more.size()) << "\n";
	if (const char* benv2 = getenv("BCM3B200_CELLPOP_GROUP_BATCHED")) o << "#define CP_GROUP_BATCHED " << atoi(benv2) << "\n";
	if (const char* senv = getenv("BCM3B200_CELLPOP_GROUP_STATIC_LU_MAX")) o << "#define CP_GROUP_STATIC_LU_MAX " << atoi(senv) << "\n";
	if (const char* senv = getenv("BCM3B200_CELLPOP_LU_SKIP_ZEROS")) o << "#define CP_LU_SKIP_ZEROS " << atoi(senv) << "\n";
	if (const char* senv = getenv("BCM3B200_CELLPOP_SOLVE_SLOTTED")) o << "#define CP_SOLVE_SLOTTED " << atoi(senv) << "\n";
	if (const char* senv = getenv("BCM3B200_CELLPOP_SOLVE_SKIP")) o << "#define CP_SOLVE_SKIP " << atoi(senv) << "\n";
	if (const char* senv = getenv("BCM3B200_CELLPOP_PIVOT_REDUX")) o << "#define CP_PIVOT_REDUX " << atoi(senv) << "\n";
	if (const char* senv = getenv("BCM3B200_CELLPOP_LOCKSTEP_TEAM")) o << "#define CP_GROUP_LOCKSTEP_TEAM " << atoi(senv) << "\n";
	if (const char* senv = getenv("BCM3B200_CELLPOP_LOCKSTEP_EVERY")) o << "#define CP_GROUP_LOCKSTEP_EVERY " << atoi(senv) << "\n";
	if (const char* senv = getenv("BCM3B200_CELLPOP_SJ_UNROLL")) o << "#define CP_SJ_UNROLL " << atoi(senv) << "\n";
	o << "#include \"cellpop_prelude.cuh\"\n";
	o << code << "\n";
	const int which = cellpop_resolve_kernel(cp);
	if (lanes.ok) o << lanes.code << "\n";
	o << (which == 1 ? "#include \"cellpop_warp.cuh\"\n" : which == 2 ? "#include \"cellpop_thread.cuh\"\n" : "#include \"cellpop_group.cuh\"\n");
	src = o.str();
	return BCM3B200_OK;
}

inline int cellpop_build_module(CellPopState& cp, const std::vector<int>& override_vars)
{
	cp.built_kernel = cellpop_resolve_kernel(cp);
	std::string src;
	int rc = cellpop_module_source(cp, override_vars, src);
	if (rc != BCM3B200_OK) return rc;
	const std::string csrc = library_dir() + "/csrc";
	// -fmad=false: the generated rate laws are +,-,*,/,sqrt expressions; without FMA contraction every one of those is
	// correctly rounded on the GPU exactly as in a host build of the same text with -ffp-contract=off, so the RHS -- the
	// part whose rounding the stiff, switch-like models amplify most (two host builds of the reference's own generated
	// code, with and without contraction, already differ by 3.5e-5 in single trajectories) -- is bit-identical to it.
	// content hash over everything that determines the binary
	std::string keyed = src + "|fmad=false";
	for (const char* f : { "/cellpop_warp.cuh", "/cellpop_thread.cuh", "/cellpop_group.cuh", "/bdf_thread.cuh", "/cellpop_prelude.cuh", "/cellpop_args.h" }) {
		std::ifstream in(csrc + f);
		std::stringstream ss;
		ss << in.rdbuf();
		keyed += ss.str();
	}
	char hash[32];
	snprintf(hash, sizeof(hash), "%016llx", (unsigned long long)fnv1a(keyed));
	const char* env = getenv("BCM3B200_CACHE");
	const std::string cache = env ? std::string(env) : (library_dir() + "/codegen_cache");
	mkdir(cache.c_str(), 0755);
	const std::string dir = cache + "/cellpop_" + hash;
	mkdir(dir.c_str(), 0755);
	const std::string so = dir + "/libcellpop_model.so";
	if (access(so.c_str(), R_OK) != 0) {
		// Several ranks may build the same model at once: every process writes its own source, log and output names and
		// renames the finished library into place. nvcc is started with an argument vector (fork + execv), never through a
		// shell: $NVCC, $BCM3B200_CACHE and the install path may contain anything.
		const std::string pid = std::to_string((long)getpid());
		const std::string cu = dir + "/model." + pid + ".cu", log = dir + "/build." + pid + ".log", tmp = so + ".tmp." + pid;
		{
			std::ofstream f(cu);
			f << src;
		}
		const char* nvcc_env = getenv("NVCC");
		const std::string nvcc = nvcc_env ? nvcc_env : "/usr/local/cuda/bin/nvcc";
		const std::string inc = "-I" + csrc;
		std::vector<std::string> args = { nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17", "--shared",
			                              "-Xcompiler", "-fPIC", inc, "-o", tmp, cu };
		std::vector<char*> argv;
		for (std::string& a : args) argv.push_back(&a[0]);
		argv.push_back(nullptr);
		int status = -1;
		const pid_t child = fork();
		if (child == 0) {
			const int fd = open(log.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
			if (fd >= 0) {
				dup2(fd, 1);
				dup2(fd, 2);
				close(fd);
			}
			execv(argv[0], argv.data());
			_exit(127);
		}
		if (child > 0) {
			while (waitpid(child, &status, 0) < 0 && errno == EINTR) {
			}
		}
		if (child < 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0) {
			std::ifstream lf(log);
			std::stringstream ss;
			ss << lf.rdbuf();
			std::string text = ss.str();
			if (text.size() > 700) text = text.substr(0, 700);
			return fail(BCM3B200_ERR_STATE, "compiling the generated model failed (%s, log %s): %s", nvcc.c_str(), log.c_str(), text.c_str());
		}
		rename(tmp.c_str(), so.c_str());
		rename(cu.c_str(), (dir + "/model.cu").c_str());
		rename(log.c_str(), (dir + "/build.log").c_str());
	}
	cp.module = dlopen(so.c_str(), RTLD_NOW | RTLD_LOCAL);
	if (!cp.module) return fail(BCM3B200_ERR_STATE, "dlopen(%s) failed: %s", so.c_str(), dlerror());
	if (cp.built_kernel == 1) {
		cp.launch = (cellpop_launch_fn)dlsym(cp.module, "cellpop_launch");
		if (!cp.launch) return fail(BCM3B200_ERR_STATE, "generated module lacks cellpop_launch");
	} else if (cp.built_kernel == 2) {
		cp.thread_launch = (cellpop_thread_launch_fn)dlsym(cp.module, "cellpop_thread_launch");
		cp.thread_scratch = (cellpop_thread_scratch_fn)dlsym(cp.module, "cellpop_thread_scratch_doubles");
		if (!cp.thread_launch || !cp.thread_scratch) return fail(BCM3B200_ERR_STATE, "generated module lacks cellpop_thread_launch");
	} else {
		cp.group_launch = (cellpop_group_launch_fn)dlsym(cp.module, "cellpop_group_launch");
		cp.group_scratch = (cellpop_group_scratch_fn)dlsym(cp.module, "cellpop_group_scratch_doubles");
		cp.group_info = (cellpop_group_info_fn)dlsym(cp.module, "cellpop_group_info");
		if (!cp.group_launch || !cp.group_scratch || !cp.group_info) return fail(BCM3B200_ERR_STATE, "generated module lacks cellpop_group_launch");
	}
	cp.module_path = so;
	return BCM3B200_OK;
}

inline int cellpop_finalize(CellPopState& cp, bool need_device)
{
	if (cp.finalized) return BCM3B200_OK;
	static const char* required[] = { "initial_conditions", "timepoints", "observed", "transforms" };
	for (const char* n : required)
		if (!cp.data.count(n)) return fail(BCM3B200_ERR_STATE, "missing data \"%s\"", n);
	if (cp.derivative_code.empty()) return fail(BCM3B200_ERR_STATE, "missing text \"derivative_code\"");
	if (cp.D > 0 && (!cp.data.count("sobol") || !cp.data.count("variability"))) return fail(BCM3B200_ERR_STATE, "variability needs \"sobol\" and \"variability\"");
	if (cp.D > CP_MAX_VARIABILITY) return fail(BCM3B200_ERR_UNSUPPORTED, "more than %d variability dimensions", CP_MAX_VARIABILITY);
	if (cp.obs_species.empty() || cp.obs_species.size() > 8) return fail(BCM3B200_ERR_ARG, "obs_species must name 1..8 species");
	if (cp.have_sim_end_time && !(cp.sim_end_time >= cp.data["timepoints"].back())) return fail(BCM3B200_ERR_ARG, "simulation_end_time lies before the last timepoint");
	for (int s : cp.obs_species)
		if (s < 0 || s >= cp.N) return fail(BCM3B200_ERR_ARG, "obs_species index out of range");
	// the data likelihoods read transformed[c][ix] on the device: an index past the variables would be an out-of-bounds read
	{
		auto bad = [&](int ix) { return ix >= cp.nvar; };
		if (bad(cp.stdev_ix) || bad(cp.offset_ix) || bad(cp.scale_ix) || bad(cp.prop_stdev_ix) || bad(cp.entry_time_ix))
			return fail(BCM3B200_ERR_ARG, "stdev_ix / offset_ix / scale_ix / proportional_stdev_ix / entry_time_ix out of range");
		for (size_t k = 0; k < cp.more.size(); k++)
			if (bad(cp.more[k]->stdev_ix) || bad(cp.more[k]->offset_ix) || bad(cp.more[k]->scale_ix) || bad(cp.more[k]->prop_stdev_ix))
				return fail(BCM3B200_ERR_ARG, "stdev_ix@%zu / offset_ix@%zu / scale_ix@%zu / proportional_stdev_ix@%zu out of range", k + 1, k + 1, k + 1, k + 1);
	}
	if (!cp.more.empty()) {
		if (cp.more.size() > 3) return fail(BCM3B200_ERR_UNSUPPORTED, "more than four data sets per handle");
		if (cellpop_resolve_kernel(cp) != 3) return fail(BCM3B200_ERR_UNSUPPORTED, "several data sets per handle need the lane-group kernel (cellpop_kernel = auto, N <= 96)");
		for (size_t k = 0; k < cp.more.size(); k++) {
			const CellPopState::MoreData& m = *cp.more[k];
			if (m.T < 1 || (int)m.timepoints.size() != m.T || (int)m.observed.size() != m.R * m.T)
				return fail(BCM3B200_ERR_STATE, "data set %zu: \"timepoints@%zu\" / \"observed@%zu\" missing or of the wrong shape", k + 1, k + 1, k + 1);
			if (m.obs_species.empty() || m.obs_species.size() > 8) return fail(BCM3B200_ERR_ARG, "obs_species@%zu must name 1..8 species", k + 1);
			for (int sp : m.obs_species)
				if (sp < 0 || sp >= cp.N) return fail(BCM3B200_ERR_ARG, "obs_species@%zu index out of range", k + 1);
		}
	}
	{
		bool wants = cp.include_only_mitotic;
		for (const auto& mp : cp.more) wants = wants || mp->include_only_mitotic;
		if (wants) {
			if (cp.nuclear_envelope_ix < 0 || cp.nuclear_envelope_ix >= cp.N)
				return fail(BCM3B200_ERR_ARG, "include_only_cells_that_went_through_mitosis needs nuclear_envelope_species = the index of the model's \"nuclear_envelope\" species");
			if (cp.data_kind != 0 && cp.include_only_mitotic) return fail(BCM3B200_ERR_ARG, "include_only_cells_that_went_through_mitosis acts on population averages only");
			if (cp.shard_count != 1) return fail(BCM3B200_ERR_UNSUPPORTED, "include_only_cells_that_went_through_mitosis is not split over ranks");
			if (cellpop_resolve_kernel(cp) != 3) return fail(BCM3B200_ERR_UNSUPPORTED, "include_only_cells_that_went_through_mitosis needs the lane-group kernel (cellpop_kernel = auto, N <= 96)");
		}
	}
	if (cp.any_time_course()) {
		// what the per-cell likelihood is built for: no parent information (non-dividing cells), all cells on one
		// device, as many observed as simulated cells (the reference refuses anything else, DataLikelihoodTimeCourse.cpp:178-187)
		if (cp.division()) { // snapshots of a dividing population are fine; per-cell trajectories would need the parent information
			bool only_snapshots = (cp.data_kind != 1);
			for (const auto& mp : cp.more) only_snapshots = only_snapshots && (mp->data_kind != 1);
			if (!only_snapshots) return fail(BCM3B200_ERR_UNSUPPORTED, "data_kind time_course with dividing / dying cells (parent information) is not built");
		}
		if (cp.shard_count != 1) return fail(BCM3B200_ERR_UNSUPPORTED, "data_kind time_course / time_points is not split over ranks (every observed cell is compared with every simulated cell)");
		if ((cp.divide_cells ? cp.max_cells : cp.num_cells) > 4096) return fail(BCM3B200_ERR_UNSUPPORTED, "data_kind time_course / time_points with more than 4096 cells (the matching is O(n^3) on the host)");
		if (cp.saturation_scale_ix >= cp.nvar) return fail(BCM3B200_ERR_ARG, "saturation_scale_ix out of range");
		for (size_t k = 0; k < cp.more.size(); k++)
			if (cp.more[k]->saturation_scale_ix >= cp.nvar) return fail(BCM3B200_ERR_ARG, "saturation_scale_ix@%zu out of range", k + 1);
		for (size_t k = 0; k < cp.more.size(); k++) {
			const CellPopState::MoreData& dn = *cp.more[k];
			if (dn.denominator_of < 0) continue;
			if (dn.marker_of >= 0 || dn.denominator_of > (int)k) return fail(BCM3B200_ERR_ARG, "denominator_of@%zu must name an earlier data set or marker", k + 1);
			const bool of_first = (dn.denominator_of == 0);
			const CellPopState::MoreData* owner = of_first ? nullptr : cp.more[(size_t)dn.denominator_of - 1].get();
			if (owner && owner->denominator_of >= 0) return fail(BCM3B200_ERR_ARG, "denominator_of@%zu names a denominator", k + 1);
			const int okind = of_first ? cp.data_kind : (owner->marker_of >= 0 ? (owner->marker_of == 0 ? cp.data_kind : cp.more[(size_t)owner->marker_of - 1]->data_kind) : owner->data_kind);
			if (okind != 1) return fail(BCM3B200_ERR_ARG, "denominator_of@%zu: the log ratio exists for data_kind time_course only (DataLikelihoodTimeCourse.cpp:380-397)", k + 1);
			const std::vector<double>& otime = of_first ? cp.data["timepoints"] : owner->timepoints;
			if (dn.timepoints != otime) return fail(BCM3B200_ERR_ARG, "denominator_of@%zu: a denominator shares the timepoints of its numerator", k + 1);
			if (dn.obs_species.size() != 1) return fail(BCM3B200_ERR_ARG, "denominator_of@%zu: the ratio of exactly two species (DataLikelihoodTimeCourseBase.cpp:178-181)", k + 1);
		}
		for (size_t k = 0; k < cp.more.size(); k++) {
			const CellPopState::MoreData& mk = *cp.more[k];
			if (mk.marker_of < 0) continue;
			if (mk.marker_of > (int)k) return fail(BCM3B200_ERR_ARG, "marker_of@%zu must name an earlier data set", k + 1);
			const bool of_first = (mk.marker_of == 0);
			const CellPopState::MoreData* parent = of_first ? nullptr : cp.more[(size_t)mk.marker_of - 1].get();
			if (parent && parent->marker_of >= 0) return fail(BCM3B200_ERR_ARG, "marker_of@%zu names a marker, not a data set", k + 1);
			const int pkind = of_first ? cp.data_kind : parent->data_kind, pT = of_first ? cp.T : parent->T, pR = of_first ? cp.R : parent->R;
			const std::vector<double>& ptime = of_first ? cp.data["timepoints"] : parent->timepoints;
			if (pkind == 0) return fail(BCM3B200_ERR_ARG, "marker_of@%zu: several markers exist for the per-cell data kinds only", k + 1);
			if (mk.T != pT || mk.R != pR || mk.timepoints != ptime)
				return fail(BCM3B200_ERR_ARG, "marker_of@%zu: a marker shares the timepoints and the observed cells of its data set", k + 1);
		}
		auto check_optimize = [&](int kind, bool optimize, int error_model) -> int {
			if (!optimize) return BCM3B200_OK;
			if (kind != 1) return fail(BCM3B200_ERR_ARG, "optimize_offset_scale belongs to data_kind time_course");
			// the reference's proportional models read per-cell sigma tables that it only fills WITHOUT optimize_offset_scale
			// (DataLikelihoodTimeCourse.cpp:221-224, 272-283, 473-481): undefined there, refused here
			if (error_model == CP_ERR_PROPORTIONAL_NORMAL || error_model == CP_ERR_ADDITIVE_PROPORTIONAL_NORMAL)
				return fail(BCM3B200_ERR_UNSUPPORTED, "optimize_offset_scale with a proportional error model is undefined in the reference");
			return BCM3B200_OK;
		};
		int rco = check_optimize(cp.data_kind, cp.optimize_offset_scale, cp.error_model);
		if (rco != BCM3B200_OK) return rco;
		for (size_t k = 0; k < cp.more.size(); k++) {
			rco = check_optimize(cp.more[k]->data_kind, cp.more[k]->optimize_offset_scale, cp.more[k]->error_model);
			if (rco != BCM3B200_OK) return rco;
		}
		auto check_kind = [&](int kind, int R, int T, bool relative, int error_model, int rel_ix) -> int {
			if (kind == 1 && (R != cp.num_cells || relative))
				return fail(BCM3B200_ERR_ARG, "data_kind time_course needs num_replicates (observed cells) = num_cells and no relative_to_time_average");
			if (kind == 2) {
				if (R < 1 || R > (cp.divide_cells ? cp.max_cells : cp.num_cells) || relative) // DataLikelihoodTimePoints.cpp:137-140
					return fail(BCM3B200_ERR_ARG, "data_kind time_points needs 1 <= num_replicates (observed cell slots) <= num_cells (max_cells when dividing) and no relative_to_time_average");
				if (error_model != CP_ERR_NORMAL && error_model != CP_ERR_STUDENT_T4)
					return fail(BCM3B200_ERR_UNSUPPORTED, "data_kind time_points knows the normal and student_t4 error models (DataLikelihoodTimePoints.cpp:280-287)");
				if (rel_ix >= T) return fail(BCM3B200_ERR_ARG, "value_relative_to_timepoint_ix out of range");
			}
			return BCM3B200_OK;
		};
		int rck = check_kind(cp.data_kind, cp.R, cp.T, cp.relative_to_time_average, cp.error_model, cp.value_relative_to_timepoint_ix);
		if (rck != BCM3B200_OK) return rck;
		for (size_t k = 0; k < cp.more.size(); k++) {
			const CellPopState::MoreData& mk = *cp.more[k];
			if (mk.rides()) continue;
			rck = check_kind(mk.data_kind, mk.R, mk.T, mk.relative_to_time_average, mk.error_model, mk.value_relative_to_timepoint_ix);
			if (rck != BCM3B200_OK) return rck;
		}
	}
	if (cp.division()) {
		if (cp.cytokinesis_ix >= cp.N || cp.apoptosis_ix >= cp.N) return fail(BCM3B200_ERR_ARG, "cytokinesis_species / apoptosis_species index out of range");
		if (cp.divide_cells && cp.cytokinesis_ix >= 0) {
			if (cp.max_cells < cp.num_cells) return fail(BCM3B200_ERR_ARG, "max_cells is smaller than num_cells");
			for (int k = 0; k < 7; k++)
				if (cp.reset_ix[k] < 0 || cp.reset_ix[k] >= cp.N)
					return fail(BCM3B200_ERR_ARG, "divide_cells needs division_reset_species = the seven species a daughter resets (Cell.cpp:127-133)");
		}
		if (cellpop_resolve_kernel(cp) != 3) return fail(BCM3B200_ERR_UNSUPPORTED, "dividing / dying cells need the lane-group kernel (cellpop_kernel = auto, N <= 96)");
		if (cp.shard_count != 1) return fail(BCM3B200_ERR_UNSUPPORTED, "a dividing population is not split over ranks (the max_cells cap is population-wide)");
	}

	// variability rows: is_ic, target index, apply, scale_ix, scale_fixed, negate
	CpArgs& a = cp.args;
	memset(&a, 0, sizeof(a));
	std::vector<int> override_vars;
	for (int d = 0; d < cp.D; d++) {
		const double* row = cp.data["variability"].data() + (size_t)d * 6;
		// row[0]: 0 = parameter, 1 = initial condition, 2 = entry time. The reference never applies an entry-time variable
		// (VariabilityDescription::ApplyVariabilityEntryTime has no caller), but it takes its quasi-random dimension.
		const int kind = (int)row[0] & 3;
		a.var_only_initial[d] = ((int)row[0] & 4) ? 1 : 0; // <variable only_initial_cells="true">
		const bool no_target = (kind == 2);
		a.var_is_ic[d] = (kind == 1);
		const int target = (int)row[1];
		a.var_apply[d] = (int)row[2];
		a.var_scale_ix[d] = (int)row[3];
		a.var_scale_fixed[d] = row[4];
		a.var_negate[d] = row[5] != 0.0;
		if (a.var_apply[d] < 0 || a.var_apply[d] > CP_APPLY_REPLACE) return fail(BCM3B200_ERR_ARG, "bad variability apply type");
		if (no_target) {
			a.var_slot[d] = -1;
		} else if (a.var_is_ic[d]) {
			if (target < 0 || target >= cp.N) return fail(BCM3B200_ERR_ARG, "variability species index out of range");
			a.var_slot[d] = target;
		} else {
			if (target < 0 || target >= cp.nvar) return fail(BCM3B200_ERR_ARG, "variability parameter index out of range");
			int slot = -1;
			for (size_t s = 0; s < override_vars.size(); s++)
				if (override_vars[s] == target) slot = (int)s;
			if (slot < 0) {
				slot = (int)override_vars.size();
				override_vars.push_back(target);
			}
			a.var_slot[d] = slot;
		}
	}
	// set_data after a finalize clears `finalized`: the module (keyed by the model text and the override list) is loaded
	// once per distinct source, streams and events are created once
	{
		std::string src;
		int rc = cellpop_module_source(cp, override_vars, src);
		if (rc != BCM3B200_OK) return rc;
		const uint64_t key = fnv1a(src);
		if (!cp.module || key != cp.module_key) {
			rc = cellpop_build_module(cp, override_vars);
			if (rc != BCM3B200_OK) return rc;
			cp.module_key = key;
		}
	}
	if (!need_device) return BCM3B200_OK; // compile-only (CPU container)

	const long long lo = (long long)cp.num_cells * cp.shard_rank / cp.shard_count;
	const long long hi = (long long)cp.num_cells * (cp.shard_rank + 1) / cp.shard_count;
	cp.cell_offset = (int)lo;
	cp.cells_local = (int)(hi - lo);
	CUDA_TRY(cudaSetDevice(cp.device));
	if (!cp.stream) CUDA_TRY(cudaStreamCreateWithFlags(&cp.stream, cudaStreamNonBlocking));
	if (!cp.ev0) CUDA_TRY(cudaEventCreate(&cp.ev0));
	if (!cp.ev1) CUDA_TRY(cudaEventCreate(&cp.ev1));
	auto up = [&](DevBuf<double>& b, const std::vector<double>& v) -> cudaError_t {
		cudaError_t e = b.ensure(v.size() ? v.size() : 1);
		if (e != cudaSuccess || v.empty()) return e;
		return cudaMemcpy(b.p, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice);
	};
	CUDA_TRY(up(cp.d_ic, cp.data["initial_conditions"]));
	CUDA_TRY(up(cp.d_const, cp.data["constant_species"]));
	CUDA_TRY(up(cp.d_nonsampled, cp.data["non_sampled_parameters"]));
	CUDA_TRY(up(cp.d_sobol, cp.data["sobol"]));
	CUDA_TRY(up(cp.d_time, cp.data["timepoints"]));
	CUDA_TRY(up(cp.d_obs, cp.data["observed"]));
	if (cp.full_gaussian) {
		const size_t ncov = (size_t)cp.D * (cp.D - 1) / 2;
		const std::vector<double>& cov = cp.data["variability_covariance"];
		if (cov.size() != 2 * ncov) return fail(BCM3B200_ERR_STATE, "full_gaussian with %d variables needs \"variability_covariance\" of shape [%zu][2]", cp.D, ncov);
		std::vector<int32_t> cix(ncov ? ncov : 1, -1);
		std::vector<double> cfx(ncov ? ncov : 1, 0.0);
		for (size_t e = 0; e < ncov; e++) {
			cix[e] = (int32_t)cov[2 * e];
			cfx[e] = cov[2 * e + 1];
			if (cix[e] >= cp.nvar) return fail(BCM3B200_ERR_ARG, "covariance variable index out of range");
		}
		CUDA_TRY(cp.d_cov_ix.ensure(cix.size()));
		CUDA_TRY(cp.d_cov_fixed.ensure(cfx.size()));
		CUDA_TRY(cudaMemcpy(cp.d_cov_ix.p, cix.data(), sizeof(int32_t) * cix.size(), cudaMemcpyHostToDevice));
		CUDA_TRY(cudaMemcpy(cp.d_cov_fixed.p, cfx.data(), sizeof(double) * cfx.size(), cudaMemcpyHostToDevice));
	}
	std::vector<int32_t> tr(cp.nvar);
	for (int i = 0; i < cp.nvar; i++) tr[i] = (int32_t)cp.data["transforms"][i];
	CUDA_TRY(cp.d_transforms.ensure(cp.nvar ? cp.nvar : 1));
	CUDA_TRY(cudaMemcpy(cp.d_transforms.p, tr.data(), sizeof(int32_t) * cp.nvar, cudaMemcpyHostToDevice));

	// Order in which the group kernel hands out this shard's cells: along a Morton (Z-order) curve through the first three
	// coordinates of their quasi-random variability vectors, so that consecutive cells -- the ones that share a warp --
	// have nearly the same per-cell parameters. Results do not depend on the order (every cell writes its own column of
	// cell_values, which is reduced in cell order afterwards).
	a.cell_order = nullptr;
	if (cp.D > 0 && cp.cells_local > 1 && !getenv("BCM3B200_CELLPOP_NO_ORDER")) {
		const std::vector<double>& sob = cp.data["sobol"];
		const int dims = cp.D < 3 ? cp.D : 3;
		std::vector<std::pair<uint32_t, int32_t>> keyed(cp.cells_local);
		for (int i = 0; i < cp.cells_local; i++) {
			uint32_t key = 0;
			for (int d = 0; d < dims; d++) {
				double u = sob[(size_t)(cp.cell_offset + i) * cp.D + d];
				uint32_t qv = (uint32_t)(u < 0.0 ? 0.0 : (u >= 1.0 ? 1023.0 : u * 1024.0));
				if (qv > 1023u) qv = 1023u;
				for (int b = 0; b < 10; b++) key |= ((qv >> b) & 1u) << (b * dims + d);
			}
			keyed[i] = { key, (int32_t)i };
		}
		std::sort(keyed.begin(), keyed.end());
		std::vector<int32_t> order(cp.cells_local);
		for (int i = 0; i < cp.cells_local; i++) order[i] = keyed[i].second;
		CUDA_TRY(cp.d_cell_order.ensure(order.size()));
		CUDA_TRY(cudaMemcpy(cp.d_cell_order.p, order.data(), sizeof(int32_t) * order.size(), cudaMemcpyHostToDevice));
		a.cell_order = cp.d_cell_order.p;
	}
	a.treatment_species = -1;
	a.treatment_num_pulses = 0;
	a.treatment_times = nullptr;
	if (cp.treatment_species >= 0) {
		if (cp.treatment_species >= cp.Nc) return fail(BCM3B200_ERR_ARG, "treatment_species is not a constant species of the model");
		if (cp.built_kernel != 3) return fail(BCM3B200_ERR_UNSUPPORTED, "treatment trajectories need the lane-group kernel (cellpop_kernel = auto for N <= 96)");
		std::vector<double> times = cp.data["treatment_times"];
		std::sort(times.begin(), times.end()); // TreatmentTrajectoryPulses::Load, .cpp:17
		CUDA_TRY(cp.d_treatment_times.ensure(times.size() ? times.size() : 1));
		if (!times.empty()) CUDA_TRY(cudaMemcpy(cp.d_treatment_times.p, times.data(), sizeof(double) * times.size(), cudaMemcpyHostToDevice));
		a.treatment_species = cp.treatment_species;
		a.treatment_num_pulses = (int)times.size();
		a.treatment_times = cp.d_treatment_times.p;
	}
	a.num_cells = cp.cells_local;
	a.cell_stride = cp.capacity();
	a.initial_flag = cp.num_cells > 1 ? 1 : 0; // Experiment.cpp:662-670
	a.cytokinesis_ix = (cp.divide_cells && cp.division()) ? cp.cytokinesis_ix : -1;
	a.apoptosis_ix = cp.division() ? cp.apoptosis_ix : -1;
	for (int k = 0; k < 7; k++) a.reset_ix[k] = cp.reset_ix[k];
	a.items = nullptr;
	a.num_items = 0;
	a.cell_offset = cp.cell_offset;
	a.nvar = cp.nvar;
	a.initial_conditions = cp.d_ic.p;
	a.constant_species = cp.d_const.p;
	a.non_sampled = cp.d_nonsampled.p;
	a.sobol = cp.d_sobol.p;
	a.D = cp.D;
	a.entry_time_ix = cp.entry_time_ix;
	a.entry_time_fixed = cp.entry_time_fixed;
	a.timepoints = cp.d_time.p;
	a.T = cp.T;
	cp.TU = cp.T;
	a.num_data_sets = 1 + (int)cp.more.size();
	a.num_rows = cp.rows();
	a.tp_rows = nullptr;
	double last_requested = cp.data["timepoints"].back();
	if (!cp.more.empty()) {
		// the union of the data sets' timepoints (exact comparisons: equal times are one interpolation) and, per union time and data
		// set, the row of cell_values the value goes to
		const int K = 1 + (int)cp.more.size();
		std::vector<double> all(cp.data["timepoints"]);
		for (const auto& m : cp.more) {
			all.insert(all.end(), m->timepoints.begin(), m->timepoints.end());
			last_requested = std::max(last_requested, m->timepoints.back());
		}
		std::sort(all.begin(), all.end());
		all.erase(std::unique(all.begin(), all.end()), all.end());
		std::vector<int32_t> rows(all.size() * K, -1);
		int row0 = 0;
		for (int k = 0; k < K; k++) {
			const std::vector<double>& tp = k == 0 ? cp.data["timepoints"] : cp.more[k - 1]->timepoints;
			for (size_t i = 0; i < tp.size(); i++) {
				if (i > 0 && !(tp[i] > tp[i - 1])) return fail(BCM3B200_ERR_ARG, "several data sets per handle need strictly increasing timepoints in every one of them");
				const size_t u = (size_t)(std::lower_bound(all.begin(), all.end(), tp[i]) - all.begin());
				rows[u * K + k] = row0 + (int)i;
			}
			row0 += (int)tp.size();
		}
		CUDA_TRY(cp.d_union_time.ensure(all.size()));
		CUDA_TRY(cudaMemcpy(cp.d_union_time.p, all.data(), sizeof(double) * all.size(), cudaMemcpyHostToDevice));
		CUDA_TRY(cp.d_tp_rows.ensure(rows.size()));
		CUDA_TRY(cudaMemcpy(cp.d_tp_rows.p, rows.data(), sizeof(int32_t) * rows.size(), cudaMemcpyHostToDevice));
		a.timepoints = cp.d_union_time.p;
		a.T = cp.TU = (int)all.size();
		a.tp_rows = cp.d_tp_rows.p;
		for (size_t k = 0; k < cp.more.size(); k++) {
			CellPopState::MoreData& m = *cp.more[k];
			a.num_obs_species_more[k] = (int)m.obs_species.size();
			for (size_t i = 0; i < m.obs_species.size(); i++) a.obs_species_more[k][i] = m.obs_species[i];
			CUDA_TRY(m.d_time.ensure(m.timepoints.size()));
			CUDA_TRY(cudaMemcpy(m.d_time.p, m.timepoints.data(), sizeof(double) * m.timepoints.size(), cudaMemcpyHostToDevice));
			CUDA_TRY(m.d_obs.ensure(m.observed.size()));
			CUDA_TRY(cudaMemcpy(m.d_obs.p, m.observed.data(), sizeof(double) * m.observed.size(), cudaMemcpyHostToDevice));
		}
	}
	// the cells are integrated to the last time any data set of the experiment requests (Experiment.cpp:655-656)
	a.sim_end_time = cp.have_sim_end_time ? std::max(cp.sim_end_time, last_requested) : last_requested;
	a.rel_tol = cp.rel_tol;
	a.abs_tol = cp.abs_tol;
	a.min_dt = cp.min_dt;
	// CVodeSetMaxStep (cvode_io.c:344-376): 0 and infinity mean no ceiling; a ceiling below the minimum step is an error
	a.max_dt_inv = (cp.max_dt > 0.0 && std::isfinite(cp.max_dt)) ? 1.0 / cp.max_dt : 0.0;
	if (cp.max_dt < 0.0 || a.max_dt_inv * cp.min_dt > 1.0) return fail(BCM3B200_ERR_ARG, "solver_max_timestep is negative or below solver_min_timestep");
	if (a.max_dt_inv > 0.0 && cp.built_kernel != 3) return fail(BCM3B200_ERR_UNSUPPORTED, "solver_max_timestep needs the lane-group kernel (cellpop_kernel = auto, N <= 96)");
	a.max_steps = cp.max_steps;
	a.num_obs_species = (int)cp.obs_species.size();
	for (size_t k = 0; k < cp.obs_species.size(); k++) a.obs_species[k] = cp.obs_species[k];
	cp.finalized = true;
	return BCM3B200_OK;
}

// stage 1: upload the batch, transform, integrate every (chain, cell) of this shard -> cell_values / status on the device
inline int cellpop_run_cells(CellPopState& cp, size_t C, size_t nvar, const double* values, cudaStream_t st)
{
	if ((int)nvar != cp.nvar) return fail(BCM3B200_ERR_ARG, "num_variables %zu != %d", nvar, cp.nvar);
	int rc = cellpop_finalize(cp, true);
	if (rc != BCM3B200_OK) return rc;
	CUDA_TRY(cudaSetDevice(cp.device));
	const int T = cp.rows(), nc = cp.cells_local; // rows of per-cell values per chain
	const size_t cols = (size_t)(cp.capacity() ? cp.capacity() : 1); // cell columns of the per-cell outputs
	CUDA_TRY(cp.d_values.ensure(C * nvar));
	CUDA_TRY(cp.d_transformed.ensure(C * nvar));
	CUDA_TRY(cp.d_cellvals.ensure(C * (size_t)T * cols));
	CUDA_TRY(cp.d_status.ensure(C * cols));
	CUDA_TRY(cp.d_steps.ensure(C * cols));
	CUDA_TRY(cp.d_avg.ensure(C * (size_t)T));
	CUDA_TRY(cp.d_count.ensure(C * (size_t)T));
	CUDA_TRY(cp.d_nfail.ensure(C));
	CUDA_TRY(cp.d_logp.ensure(C));
	CUDA_TRY(cudaMemcpyAsync(cp.d_values.p, values, sizeof(double) * C * nvar, cudaMemcpyHostToDevice, st));
	CUDA_TRY(cudaEventRecord(cp.ev0, st));
	const long long ne = (long long)C * nvar;
	cellpop_transform_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(cp.d_values.p, cp.d_transforms.p, (int)nvar, (int)C, cp.d_transformed.p);
	CUDA_TRY(cudaGetLastError());
	CpArgs a = cp.args;
	a.var_full = 0;
	a.var_chol = nullptr;
	if (cp.full_gaussian && cp.D > 0) {
		CUDA_TRY(cp.d_chol.ensure(C * (size_t)cp.D * cp.D));
		CpCholArgs ca;
		ca.D = cp.D;
		ca.nvar = cp.nvar;
		for (int d = 0; d < cp.D; d++) {
			ca.scale_ix[d] = a.var_scale_ix[d];
			ca.scale_fixed[d] = a.var_scale_fixed[d];
		}
		ca.cov_ix = cp.d_cov_ix.p;
		ca.cov_fixed = cp.d_cov_fixed.p;
		cellpop_cholesky_kernel<<<(unsigned)((C + 63) / 64), 64, 0, st>>>(ca, cp.d_transformed.p, (int)C, cp.d_chol.p);
		CUDA_TRY(cudaGetLastError());
		a.var_full = 1;
		a.var_chol = cp.d_chol.p;
	}
	a.num_chains = (int)C;
	a.transformed = cp.d_transformed.p;
	a.cell_values = cp.d_cellvals.p;
	a.cell_status = cp.d_status.p;
	a.cell_steps = cp.d_steps.p;
	a.debug_report = cp.steps_report;
	cp.last_launches = 1;
	const bool use_group = (cp.built_kernel == 3);
	const bool use_thread = (cp.built_kernel == 2);
	if (nc > 0 && use_group && cp.division()) {
		// One generation at a time (Experiment::ParallelSimulation, Experiment.cpp:691-724): the cells of a generation are
		// independent work items of one launch; the daughters of those that divided form the next generation.
		const int stride = cp.capacity();
		const size_t recs = C * (size_t)stride;
		CUDA_TRY(cp.d_creation.ensure(recs));
		CUDA_TRY(cp.d_end_time.ensure(recs));
		CUDA_TRY(cp.d_end_y.ensure(recs * (size_t)cp.N));
		CUDA_TRY(cp.d_row.ensure(recs));
		CUDA_TRY(cp.d_parent.ensure(recs));
		CUDA_TRY(cp.d_event.ensure(recs));
		CUDA_TRY(cp.d_items.ensure(2 * recs));
		CUDA_TRY(cp.d_wave.ensure(4 * C));
		CUDA_TRY(cp.d_item_offsets.ensure(C));
		cp.h_wave.resize(4 * C);
		CUDA_TRY(cudaMemsetAsync(cp.d_cellvals.p, 0xFF, sizeof(double) * C * (size_t)T * stride, st)); // all-ones = NaN: the cell does not exist
		a.nuclear_envelope_ix = -1;
		a.cell_mitotic = nullptr;
		if (cp.track_mitosis()) {
			CUDA_TRY(cp.d_mitotic.ensure(recs));
			CUDA_TRY(cudaMemsetAsync(cp.d_mitotic.p, 0, sizeof(int32_t) * recs, st));
			a.nuclear_envelope_ix = cp.nuclear_envelope_ix;
			a.cell_mitotic = cp.d_mitotic.p;
			// which rows of the value block belong to data sets that average the mitotic cells only
			std::vector<int32_t> only((size_t)T, 0);
			int r = 0;
			for (int k = -1; k < (int)cp.more.size(); k++) {
				const int Tk = (k < 0) ? cp.T : cp.more[(size_t)k]->T;
				const bool flag = (k < 0) ? cp.include_only_mitotic : cp.more[(size_t)k]->include_only_mitotic;
				for (int i = 0; i < Tk; i++) only[(size_t)(r + i)] = flag ? 1 : 0;
				r += Tk;
			}
			CUDA_TRY(cp.d_row_only_mitotic.ensure((size_t)T));
			CUDA_TRY(cudaMemcpyAsync(cp.d_row_only_mitotic.p, only.data(), sizeof(int32_t) * (size_t)T, cudaMemcpyHostToDevice, st));
			CUDA_TRY(cudaStreamSynchronize(st)); // `only` is a local
		}
		cellpop_division_init_kernel<<<(unsigned)((recs + 255) / 256), 256, 0, st>>>((int)C, nc, stride, cp.d_transformed.p, (int)nvar, cp.entry_time_ix,
		                                                                            cp.entry_time_fixed, cp.cell_offset, cp.d_creation.p, cp.d_row.p,
		                                                                            cp.d_parent.p, cp.d_event.p, cp.d_status.p, cp.d_steps.p, cp.d_items.p,
		                                                                            cp.d_wave.p);
		CUDA_TRY(cudaGetLastError());
		cp.last_launches++;
		a.cell_stride = stride;
		a.items = cp.d_items.p;
		a.cell_creation = cp.d_creation.p;
		a.cell_row = cp.d_row.p;
		a.cell_parent = cp.d_parent.p;
		a.cell_end_y = cp.d_end_y.p;
		a.cell_end_time = cp.d_end_time.p;
		a.cell_event = cp.d_event.p;
		const long long need = cp.group_scratch((int)C, nc);
		if (need < 0) return fail(BCM3B200_ERR_CUDA, "cellpop group kernel does not fit on this device: %s", cudaGetErrorString((cudaError_t)(-need)));
		CUDA_TRY(cp.d_scratch.ensure((size_t)need));
		long long items = (long long)C * nc;
		for (int generation = 0; items > 0; generation++) {
			if (generation > 64) return fail(BCM3B200_ERR_STATE, "more than 64 generations of dividing cells");
			a.num_items = (int)items;
			int lrc = cp.group_launch(&a, cp.d_scratch.p, (void*)st);
			if (lrc != 0) return fail(BCM3B200_ERR_CUDA, "cellpop group kernel launch failed: %s", cudaGetErrorString((cudaError_t)lrc));
			cp.last_launches++;
			if (!cp.divide_cells || cp.cytokinesis_ix < 0) break; // cells only die: one generation
			cellpop_spawn_kernel<<<(unsigned)C, 256, 0, st>>>(stride, cp.num_cells, cp.sobol_rows, a.sim_end_time, cp.d_event.p, cp.d_end_time.p, cp.d_creation.p,
			                                                 cp.d_row.p, cp.d_parent.p, cp.d_wave.p);
			CUDA_TRY(cudaGetLastError());
			CUDA_TRY(cudaMemcpyAsync(cp.h_wave.data(), cp.d_wave.p, sizeof(int32_t) * 4 * C, cudaMemcpyDeviceToHost, st));
			CUDA_TRY(cudaStreamSynchronize(st));
			std::vector<int32_t> offsets(C);
			items = 0;
			for (size_t c = 0; c < C; c++) {
				offsets[c] = (int32_t)items;
				items += cp.h_wave[4 * c + 1] - cp.h_wave[4 * c];
			}
			cp.last_launches++;
			if (items == 0) break;
			CUDA_TRY(cudaMemcpyAsync(cp.d_item_offsets.p, offsets.data(), sizeof(int32_t) * C, cudaMemcpyHostToDevice, st));
			cellpop_items_kernel<<<(unsigned)C, 256, 0, st>>>((int)C, cp.d_wave.p, cp.d_item_offsets.p, cp.d_items.p);
			CUDA_TRY(cudaGetLastError());
			CUDA_TRY(cudaStreamSynchronize(st)); // `offsets` is a stack buffer
			cp.last_launches++;
		}
	} else if (nc > 0 && use_group) {
		const long long need = cp.group_scratch((int)C, nc);
		if (need < 0) return fail(BCM3B200_ERR_CUDA, "cellpop group kernel does not fit on this device: %s", cudaGetErrorString((cudaError_t)(-need)));
		CUDA_TRY(cp.d_scratch.ensure((size_t)need));
		int lrc = cp.group_launch(&a, cp.d_scratch.p, (void*)st);
		if (lrc != 0) return fail(BCM3B200_ERR_CUDA, "cellpop group kernel launch failed: %s", cudaGetErrorString((cudaError_t)lrc));
		cp.last_launches++;
	} else if (nc > 0 && use_thread) {
		CUDA_TRY(cp.d_scratch.ensure((size_t)cp.thread_scratch((int)C, nc)));
		int lrc = cp.thread_launch(&a, cp.d_scratch.p, (void*)st);
		if (lrc != 0) return fail(BCM3B200_ERR_CUDA, "cellpop thread kernel launch failed: %s", cudaGetErrorString((cudaError_t)lrc));
		cp.last_launches++;
	} else if (nc > 0) {
		int lrc = cp.launch(&a, (void*)st);
		if (lrc != 0) return fail(BCM3B200_ERR_CUDA, "cellpop kernel launch failed: %s", cudaGetErrorString((cudaError_t)lrc));
		cp.last_launches++;
	}
	cp.last_C = (int)C;
	return BCM3B200_OK;
}

// stage 3: data likelihood of every chain from the population averages in d_avg / d_nfail -> d_logp
inline int cellpop_data_likelihood(CellPopState& cp, size_t C, cudaStream_t st)
{
	CpLikArgs la;
	la.avg = cp.d_avg.p;
	la.avg_stride = cp.rows();
	la.avg_row0 = 0;
	la.accumulate = 0;
	la.nfail = cp.d_nfail.p;
	la.transformed = cp.d_transformed.p;
	la.timepoints = cp.d_time.p;
	la.observed = cp.d_obs.p;
	la.T = cp.T;
	la.R = cp.R;
	la.nvar = cp.nvar;
	la.error_model = cp.error_model;
	la.stdev_ix = cp.stdev_ix;
	la.offset_ix = cp.offset_ix;
	la.scale_ix = cp.scale_ix;
	la.stdev_fixed = cp.stdev_fixed;
	la.relative_to_time_average = cp.relative_to_time_average ? 1 : 0;
	la.stdev_relative_to_scale = cp.stdev_relative_to_scale ? 1 : 0;
	la.prop_stdev_ix = cp.prop_stdev_ix;
	la.prop_stdev_fixed = cp.prop_stdev_fixed;
	la.offset_fixed = cp.offset_fixed;
	la.scale_fixed = cp.scale_fixed;
	la.weight = cp.weight;
	la.missing_stdev = cp.missing_simulation_time_stdev;
	la.logp = cp.d_logp.p;
	la.kind = cp.data_kind;
	cellpop_datalik_kernel<<<(unsigned)((C + 63) / 64), 64, 0, st>>>(la, (int)C);
	CUDA_TRY(cudaGetLastError());
	// the further data sets of the experiment, from their own rows of the averages, added in order
	int row0 = cp.T;
	for (const auto& mp : cp.more) {
		const CellPopState::MoreData& m = *mp;
		if (m.rides()) { // a further marker (or a ratio's denominator) of a per-cell data set: no term of its own
			row0 += m.T;
			continue;
		}
		la.avg_row0 = row0;
		la.accumulate = 1;
		la.timepoints = m.d_time.p;
		la.observed = m.d_obs.p;
		la.T = m.T;
		la.R = m.R;
		la.error_model = m.error_model;
		la.stdev_ix = m.stdev_ix;
		la.offset_ix = m.offset_ix;
		la.scale_ix = m.scale_ix;
		la.stdev_fixed = m.stdev_fixed;
		la.relative_to_time_average = m.relative_to_time_average ? 1 : 0;
		la.stdev_relative_to_scale = m.stdev_relative_to_scale ? 1 : 0;
		la.prop_stdev_ix = m.prop_stdev_ix;
		la.prop_stdev_fixed = m.prop_stdev_fixed;
		la.offset_fixed = m.offset_fixed;
		la.scale_fixed = m.scale_fixed;
		la.weight = m.weight;
		la.missing_stdev = m.missing_stdev;
		la.kind = m.data_kind;
		cellpop_datalik_kernel<<<(unsigned)((C + 63) / 64), 64, 0, st>>>(la, (int)C);
		CUDA_TRY(cudaGetLastError());
		cp.last_launches++;
		row0 += m.T;
	}
	return BCM3B200_OK;
}

// The per-cell data sets of the handle. time_course (DataLikelihoodTimeCourse::Evaluate, .cpp:230-365): the [observed x
// simulated] block of cell log-likelihoods on the device, then per chain on the host the reference's admission rules (a NaN
// anywhere or an observed cell without enough finite entries: -inf), the matching (matching_host.cuh) and the sum of the matched
// entries in observed-cell order. time_points (DataLikelihoodTimePoints::Evaluate, DataLikelihoodTimePoints.cpp:209-345): the same
// per TIMEPOINT -- the observed cells present at that time against the simulated cells that have a value there, one block,
// one matching and one partial sum per timepoint, accumulated in time order. Times the data set's weight; the terms are added
// to logp [C] (host) after the population-average data sets' terms, data set by data set. Chains go to the host's cores.
template <class F>
inline void cellpop_for_each_chain(size_t C, bool parallel, F&& f)
{
	size_t workers = std::thread::hardware_concurrency();
	if (workers < 1) workers = 1;
	if (workers > C) workers = C;
	if (!parallel || workers <= 1) {
		for (size_t c = 0; c < C; c++) f(c);
		return;
	}
	std::atomic<size_t> next(0);
	std::vector<std::thread> th;
	for (size_t w = 0; w < workers; w++)
		th.emplace_back([&]() {
			for (;;) {
				const size_t c = next.fetch_add(1);
				if (c >= C) break;
				f(c);
			}
		});
	for (auto& t : th) t.join();
}

inline int cellpop_time_course_terms(CellPopState& cp, size_t C, cudaStream_t st, double* logp)
{
	const double ninf = -std::numeric_limits<double>::infinity();
	int row0 = 0;
	for (int k = -1; k < (int)cp.more.size(); k++) {
		const CellPopState::MoreData* m = (k >= 0) ? cp.more[(size_t)k].get() : nullptr;
		const int T = m ? m->T : cp.T;
		const int kind = m ? m->data_kind : cp.data_kind;
		if (m && m->rides()) { // rows of a further marker / of a denominator: used by the data set they belong to
			row0 += T;
			continue;
		}
		if (kind != 0) {
			// simulated cell columns: the cells of the handle; with a dividing population (time_points only) every cell slot up to
			// max_cells -- a slot that was never filled, or a cell outside its life span, holds NaN and is left out per timepoint
			const int n = (kind == 2) ? cp.capacity() : cp.cells_local;
			const int n_obs = m ? m->R : cp.R; // time_course: = n; time_points: the observed cell slots (<= n)
			const int newborn_from = (kind == 2 && (m ? m->use_only_nondivided : cp.use_only_nondivided)) ? cp.num_cells : n;
			CpCellLikArgs a;
			a.cell_values = cp.d_cellvals.p;
			a.rows = cp.rows();
			a.cell_stride = cp.capacity();
			a.T = T;
			a.n_obs = n_obs;
			a.n_sim = n;
			a.nvar = cp.nvar;
			a.error_model = m ? m->error_model : cp.error_model;
			a.stdev_relative_to_scale = (m ? m->stdev_relative_to_scale : cp.stdev_relative_to_scale) ? 1 : 0;
			a.transformed = cp.d_transformed.p;
			a.timepoints = m ? m->d_time.p : cp.d_time.p;
			// marker 0: the data set itself; then the entries of `more` that name it as their data set
			std::vector<const std::vector<double>*> observed_of_marker;
			auto denominator_rows = [&](int entry) { // first row of the entry of `more` that is the denominator of `entry`, or -1
				int r = cp.T;
				for (size_t q = 0; q < cp.more.size(); q++) {
					if (cp.more[q]->denominator_of == entry) return r;
					r += cp.more[q]->T;
				}
				return -1;
			};
			a.L = 1;
			a.mk[0].row0 = row0;
			a.mk[0].den_row0 = (kind == 1) ? denominator_rows(k + 1) : -1;
			a.mk[0].observed = m ? m->d_obs.p : cp.d_obs.p;
			a.mk[0].stdev_ix = m ? m->stdev_ix : cp.stdev_ix;
			a.mk[0].offset_ix = m ? m->offset_ix : cp.offset_ix;
			a.mk[0].scale_ix = m ? m->scale_ix : cp.scale_ix;
			a.mk[0].prop_stdev_ix = m ? m->prop_stdev_ix : cp.prop_stdev_ix;
			a.mk[0].stdev_fixed = m ? m->stdev_fixed : cp.stdev_fixed;
			a.mk[0].offset_fixed = m ? m->offset_fixed : cp.offset_fixed;
			a.mk[0].scale_fixed = m ? m->scale_fixed : cp.scale_fixed;
			a.mk[0].prop_stdev_fixed = m ? m->prop_stdev_fixed : cp.prop_stdev_fixed;
			observed_of_marker.push_back(m ? &m->observed : &cp.data["observed"]);
			{
				int r = cp.T;
				for (size_t q = 0; q < cp.more.size(); q++) {
					const CellPopState::MoreData& f = *cp.more[q];
					if (f.marker_of == k + 1 && a.L < 4) {
						CpMarkerArgs& mk = a.mk[a.L++];
						mk.row0 = r;
						mk.den_row0 = (kind == 1) ? denominator_rows((int)q + 1) : -1;
						mk.observed = f.d_obs.p;
						mk.stdev_ix = f.stdev_ix;
						mk.offset_ix = f.offset_ix;
						mk.scale_ix = f.scale_ix;
						mk.prop_stdev_ix = f.prop_stdev_ix;
						mk.stdev_fixed = f.stdev_fixed;
						mk.offset_fixed = f.offset_fixed;
						mk.scale_fixed = f.scale_fixed;
						mk.prop_stdev_fixed = f.prop_stdev_fixed;
						observed_of_marker.push_back(&f.observed);
					}
					r += f.T;
				}
			}
			a.missing_stdev = m ? m->missing_stdev : cp.missing_simulation_time_stdev;
			a.optimize = (kind == 1 && (m ? m->optimize_offset_scale : cp.optimize_offset_scale)) ? 1 : 0;
			a.opt_offset_min = m ? m->optimize_offset_min : cp.optimize_offset_min;
			a.opt_offset_max = m ? m->optimize_offset_max : cp.optimize_offset_max;
			a.opt_scale_min = m ? m->optimize_scale_min : cp.optimize_scale_min;
			a.opt_scale_max = m ? m->optimize_scale_max : cp.optimize_scale_max;
			a.saturation_ix = (kind == 1) ? (m ? m->saturation_scale_ix : cp.saturation_scale_ix) : -1;
			a.only_k = -1;
			a.rel_k = (kind == 2) ? (m ? m->value_relative_to_timepoint_ix : cp.value_relative_to_timepoint_ix) : -1;
			const double weight = m ? m->weight : cp.weight;
			const size_t block = (size_t)n_obs * n;
			// the blocks of at most ~1 GiB worth of chains at a time (4 096 cells: 128 MiB per chain)
			size_t budget = (size_t)1 << 27; // doubles
			if (const char* benv = getenv("BCM3B200_CELL_LIKELIHOOD_DOUBLES")) budget = (size_t)strtoull(benv, nullptr, 10); // tests: force small chunks
			size_t chunk = block ? budget / block : C;
			if (chunk < 1) chunk = 1;
			if (chunk > C) chunk = C;
			CUDA_TRY(cp.d_cell_lik.ensure(chunk * block ? chunk * block : 1));
			a.lik = cp.d_cell_lik.p;
			cp.h_cell_lik.resize(chunk * block);
			for (size_t c0 = 0; c0 < C; c0 += chunk) {
			const size_t Cc = std::min(chunk, C - c0);
			a.c0 = (int)c0;
			auto run_block = [&]() -> int { // the block for a.only_k (or all timepoints) of chains c0 .. c0 + Cc - 1 to the host
				if (block == 0) return BCM3B200_OK;
				cellpop_cell_likelihood_kernel<<<dim3((unsigned)((n + 127) / 128), (unsigned)n_obs, (unsigned)Cc), 128, 0, st>>>(a);
				CUDA_TRY(cudaGetLastError());
				cp.last_launches++;
				cp.total_launches++;
				CUDA_TRY(cudaMemcpyAsync(cp.h_cell_lik.data(), cp.d_cell_lik.p, sizeof(double) * Cc * block, cudaMemcpyDeviceToHost, st));
				CUDA_TRY(cudaStreamSynchronize(st));
				return BCM3B200_OK;
			};
			if (kind == 1) {
				int rc = run_block();
				if (rc != BCM3B200_OK) return rc;
				cellpop_for_each_chain(Cc, n >= 32, [&](size_t cc) {
					const size_t c = c0 + cc;
					if (!(logp[c] > ninf)) return; // already -inf (a failed cell, an earlier data set) or NaN: nothing to add to
					const double* L = cp.h_cell_lik.data() + cc * block;
					for (int i = 0; i < n; i++) {
						int finite_count = 0;
						for (int j = 0; j < n; j++) {
							const double v = L[(size_t)i * n + j];
							if (v != v) { // .cpp:301-304
								logp[c] = ninf;
								return;
							}
							if (v > ninf) finite_count++;
						}
						if (finite_count < n) { // .cpp:316-320
							logp[c] = ninf;
							return;
						}
					}
					std::vector<double> cost(block);
					for (size_t e = 0; e < block; e++) cost[e] = -L[e];
					const std::vector<int> match = payor_matching_complete(n, cost.data());
					if ((int)match.size() != n) {
						logp[c] = ninf;
						return;
					}
					double term = 0.0;
					for (int i = 0; i < n; i++) {
						if (match[i] < 0) {
							logp[c] = ninf;
							return;
						}
						term += L[(size_t)i * n + match[i]];
					}
					logp[c] += term * weight;
				});
			} else {
				std::vector<double> term(Cc, 0.0);
				std::vector<char> dead(Cc, 0);
				for (int ti = 0; ti < T; ti++) {
					std::vector<int> rows; // the observed cells with a finite value in any marker, .cpp:222-227
					for (int i = 0; i < n_obs; i++) {
						bool any = false;
						for (const std::vector<double>* ob : observed_of_marker) any = any || std::isfinite((*ob)[(size_t)i * T + ti]);
						if (any) rows.push_back(i);
					}
					if (rows.empty()) continue; // .cpp:229-231
					a.only_k = ti;
					int rc = run_block();
					if (rc != BCM3B200_OK) return rc;
					const int fd = (int)rows.size();
					cellpop_for_each_chain(Cc, fd >= 32, [&](size_t c) {
						if (dead[c] || !(logp[c0 + c] > ninf)) return;
						const double* L = cp.h_cell_lik.data() + c * block;
						// the simulated cells with a value at this timepoint (and at the reference timepoint), .cpp:234-239: the block holds
						// NaN for the others in every row
						std::vector<int> cols;
						const double* first = L + (size_t)rows[0] * n;
						for (int j = 0; j < newborn_from; j++)
							if (first[j] == first[j]) cols.push_back(j);
						if ((int)cols.size() < fd) { // .cpp:241-245
							dead[c] = 1;
							return;
						}
						// the reference's Hungarian call keeps the edges to the first fd right nodes (hungarian.cpp:81)
						std::vector<double> cost((size_t)fd * fd);
						for (int p = 0; p < fd; p++)
							for (int q = 0; q < fd; q++) cost[(size_t)p * fd + q] = -L[(size_t)rows[p] * n + cols[q]];
						const std::vector<int> match = payor_matching_complete(fd, cost.data());
						if ((int)match.size() != fd) {
							dead[c] = 1;
							return;
						}
						for (int p = 0; p < fd; p++) {
							if (match[p] < 0) {
								dead[c] = 1;
								return;
							}
							term[c] += L[(size_t)rows[p] * n + cols[match[p]]];
						}
					});
				}
				for (size_t c = 0; c < Cc; c++) {
					if (!(logp[c0 + c] > ninf)) continue;
					if (dead[c]) logp[c0 + c] = ninf;
					else logp[c0 + c] += term[c] * weight;
				}
			}
			} // chain chunks
		}
		row0 += T;
	}
	return BCM3B200_OK;
}

inline int cellpop_evaluate(CellPopState& cp, size_t C, size_t nvar, const double* values, double* logp, int* status)
{
	if (cp.shard_count != 1)
		return fail(BCM3B200_ERR_UNSUPPORTED, "a sharded cell_population handle yields partials: use bcm3b200_enqueue_batch + bcm3b200_cellpop_finish");
	if (C == 0) return BCM3B200_OK;
	cudaStream_t st = cp.stream;
	int rc = cellpop_run_cells(cp, C, nvar, values, st);
	if (rc != BCM3B200_OK) return rc;
	const int T = cp.rows(), nc = cp.cells_local;
	cellpop_average_kernel<<<dim3(T, (unsigned)C), 256, 0, st>>>(cp.d_cellvals.p, cp.d_status.p, cp.capacity(), T, cp.d_avg.p, cp.d_count.p, cp.d_nfail.p,
	                                                             cp.track_mitosis() ? cp.d_mitotic.p : nullptr, cp.track_mitosis() ? cp.d_row_only_mitotic.p : nullptr);
	CUDA_TRY(cudaGetLastError());
	if (cp.division() && cp.divide_cells && nc > 0) {
		cellpop_overflow_kernel<<<(unsigned)((C + 63) / 64), 64, 0, st>>>((int)C, cp.d_wave.p, cp.d_nfail.p);
		CUDA_TRY(cudaGetLastError());
		cp.last_launches++;
	}
	rc = cellpop_data_likelihood(cp, C, st);
	if (rc != BCM3B200_OK) return rc;
	cp.last_launches += 2;
	cp.total_launches += cp.last_launches;
	CUDA_TRY(cudaEventRecord(cp.ev1, st));
	CUDA_TRY(cudaMemcpyAsync(logp, cp.d_logp.p, sizeof(double) * C, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	float ms = 0.f;
	if (cudaEventElapsedTime(&ms, cp.ev0, cp.ev1) == cudaSuccess) cp.last_kernel_ms = ms;
	if (cp.any_time_course()) {
		rc = cellpop_time_course_terms(cp, C, st, logp);
		if (rc != BCM3B200_OK) return rc;
	}
	if (status)
		for (size_t c = 0; c < C; c++) status[c] = std::isnan(logp[c]) ? BCM3B200_STATUS_NAN : BCM3B200_STATUS_OK;
	cp.num_evaluations += (int64_t)C;
	return BCM3B200_OK;
}

// sharded stage 1+2: this shard's partial [C][2 T + 1] on the device, enqueued on `stream`
inline int cellpop_enqueue_partial(CellPopState& cp, size_t C, size_t nvar, const double* values, double* d_partial, cudaStream_t st)
{
	if (C == 0) return BCM3B200_OK;
	int rc = cellpop_run_cells(cp, C, nvar, values, st);
	if (rc != BCM3B200_OK) return rc;
	cellpop_partial_kernel<<<dim3(cp.rows(), (unsigned)C), 256, 0, st>>>(cp.d_cellvals.p, cp.d_status.p, cp.cells_local, cp.rows(), d_partial);
	CUDA_TRY(cudaGetLastError());
	cp.last_launches += 1;
	cp.total_launches += cp.last_launches;
	return BCM3B200_OK;
}

// sharded stage 3: the combined (summed over shards) partial -> logp, on every rank
inline int cellpop_finish(CellPopState& cp, size_t C, const double* d_partial, double* logp, int* status, cudaStream_t st)
{
	if (C == 0) return BCM3B200_OK;
	if (!cp.finalized || cp.last_C != (int)C) return fail(BCM3B200_ERR_STATE, "bcm3b200_cellpop_finish without a matching bcm3b200_enqueue_batch");
	CUDA_TRY(cudaSetDevice(cp.device));
	const int n = (int)C * (cp.rows() + 1);
	cellpop_unpack_partial_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_partial, cp.rows(), (int)C, cp.d_avg.p, cp.d_count.p, cp.d_nfail.p);
	CUDA_TRY(cudaGetLastError());
	int rc = cellpop_data_likelihood(cp, C, st);
	if (rc != BCM3B200_OK) return rc;
	cp.total_launches += 2;
	CUDA_TRY(cudaEventRecord(cp.ev1, st));
	CUDA_TRY(cudaMemcpyAsync(logp, cp.d_logp.p, sizeof(double) * C, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	float ms = 0.f;
	if (cudaEventElapsedTime(&ms, cp.ev0, cp.ev1) == cudaSuccess) cp.last_kernel_ms = ms;
	if (status)
		for (size_t c = 0; c < C; c++) status[c] = std::isnan(logp[c]) ? BCM3B200_STATUS_NAN : BCM3B200_STATUS_OK;
	cp.num_evaluations += (int64_t)C;
	return BCM3B200_OK;
}

} // namespace bcm3b200
