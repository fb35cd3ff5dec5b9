// poppk_kernel.cuh -- K0+K1+K3 of the PopPK path: one (chain, patient) ODE system per thread.
//
// Replaces, for all chains at once, the per-patient loop of
//   LikelihoodPopPKTrajectory::EvaluateLogProbability   src/likelihoods/LikelihoodPopPKTrajectory.cpp:259-444
// and under it ODESolver::SolveReturnSolution / ODESolverCVODE::Solve (src/odecommon/ODESolver.cpp:93-134,
// ODESolverCVODE.cpp:322-463) with the CVODE-faithful integrator of bdf_thread.cuh.
//
// Launch shape: grid = (ceil(P_local / BLOCK), C); thread = patient, blockIdx.y = chain. A warp holds 32
// consecutive patients of ONE chain, so parameter reads (double2 per lane) and observation reads
// (time-major obs_t[T][P_pad]) are coalesced and the per-chain reduction is warp-local.
// Control flow is organised so that a warp stays converged where it matters: all lanes run the
// restart path (CVodeReInit + cvHin) together at the start of each dosing segment and then iterate
// "one step attempt per loop trip" until every lane has reached its next discontinuity.
#pragma once

#include <cstdint>

#include "bdf_thread.cuh"

// Loop votes. With BCM3_BLOCK_LOCKSTEP the segment and attempt loops are uniform over the whole thread block
// (__syncthreads_or): the warps of a block then walk the large, straight-line integrator code in lock-step and
// share the instruction-cache lines they fetch, at the price of waiting for the slowest warp each trip.
#ifndef BCM3_BLOCK_LOCKSTEP
#define BCM3_BLOCK_LOCKSTEP 1 // measured on B200: 24.2 -> 19.8 ms on 320k two-compartment systems (profiles/r01_variants.log)
#endif
#if BCM3_BLOCK_LOCKSTEP
#define BCM3_LOOP_VOTE(pred) (__syncthreads_or((pred) ? 1 : 0) != 0)
#else
#define BCM3_LOOP_VOTE(pred) (__any_sync(0xffffffffu, (pred)) != 0)
#endif

namespace bcm3b200 {

// pk_model type=, LikelihoodPopPKTrajectory.cpp:69-83
enum : int { PK_ONE = 0, PK_TWO = 1, PK_ONE_BIPHASIC = 2, PK_TWO_BIPHASIC = 3, PK_ONE_TRANSIT = 4, PK_TWO_TRANSIT = 5 };
enum : int { PKV_PLAIN = 0, PKV_BIPHASIC = 1, PKV_TRANSIT = 2 };
enum : int { TR_NONE = 0, TR_LOG = 1, TR_LOG10 = 2, TR_LOGIT = 3 };

// indices into PkArgs::ix / PkArgs::tr (the chain-level entries of the variable vector)
enum : int {
	SV_MEAN_ABSORPTION = 0, // values[0]            mean log10 k_absorption
	SV_EXCRETION,           // values[1]
	SV_MEAN_CLEARANCE,      // values[2]            mean log10 clearance
	SV_VOD,                 // values[3]
	SV_PERIPHERY_FWD,       // values[4]  (two-compartment)
	SV_PERIPHERY_BWD,       // values[5]
	SV_SIGMA_ABSORPTION,    // values[npk + 0]
	SV_SIGMA_CLEARANCE,     // values[npk + 1]
	SV_SD,                  // values[sd_ix]
	SV_SD2,                 // values[sd_ix + 1]
	SV_NAMED_A,             // by name: n_transit (transit types) / biphasic_uptake_time (biphasic types), cpp:296-310
	SV_NAMED_B,             // by name: mean_transit_time / mean_absorption2
	SV_COUNT
};

struct PkArgs {
	// static data of this shard (device pointers)
	const double* time;            // [T]
	const double* obs_t;           // [T][P_pad]  time-major observed concentrations, NaN = missing
	const double* dose;            // [P_local]
	const double* dosing_interval; // [P_local]
	const double* dose_after_dose_change;
	const double* dose_change_time;
	const int32_t* intermittent;
	const uint32_t* skipped_days;
	const int32_t* simulate_until;
	int P_local, P_pad, patient_offset, T, max_steps;
	double rtol, atol, conv_base; // conv_base = 1e6 / MW
	// <pk_model volume_of_distribution= k_periphery_fwd= k_periphery_bwd=> (cpp:64-67): NaN = sampled. As in the reference the
	// positional indices of the variable vector do not move when a parameter is fixed (cpp:267-272, 285-294), and the
	// peripheral pair is taken from the attributes only when k_periphery_fwd is given (cpp:288).
	double fixed_vod, fixed_periphery_fwd, fixed_periphery_bwd;
	// the batch
	const double* values; // [C][row_stride]
	long long row_stride;
	long long col_patient0; // column of patient_offset's first probability
	int ix[SV_COUNT];
	int tr[SV_COUNT];
	const int* order; // [C][P_local] or null: patient handled by thread r of chain c (patients ranked by absorption rate)
	// 1: grid = (C, blocks per chain) -- the block index that runs fastest is the CHAIN, so with ranked patients (most work
	// first within every chain) the grid as a whole runs from the most expensive blocks of all chains to the cheapest ones and
	// its tail is made of the shortest blocks; 0: grid = (blocks per chain, C)
	int chain_fastest;
	int single; // 1: pharmacokinetic_trajectory -- the one patient's parameters are the chain's variables themselves (LikelihoodPharmacokineticTrajectory.cpp:259-330)
	// outputs
	double* patient_ll;    // [C][P_local] log-likelihood of every (chain, patient), in patient order whatever the launch shape
	double* diag_conc;     // [C][P_local][T] or null
	int32_t* diag_counters; // [C][P_local][8] or null
};

// VariableSet::TransformVariable, src/sampler/VariableSet.cpp:97-124; bcm3::fastpow10, MathFunctions.h:13
__device__ __forceinline__ double fastpow10(double x) { return exp(x * 2.3025850929940459); }
__device__ __forceinline__ double transform_variable(int tr, double x)
{
	switch (tr) {
	case TR_LOG:
		return exp(x);
	case TR_LOG10:
		return fastpow10(x);
	case TR_LOGIT:
		if (x > 0) {
			double z = exp(-x);
			return 1.0 / (1.0 + z);
		} else {
			double z = exp(x);
			return z / (1.0 + z);
		}
	default:
		return x;
	}
}

// bcm3::QuantileNormal(p, mu, sigma) = boost::math::quantile(normal(mu, sigma), p)
// (ProbabilityDistributions.cpp:359-363): mu + sigma * Phi^-1(p)
__device__ __forceinline__ double quantile_normal(double p, double mu, double sigma)
{
	double r = normcdfinv(p);
	r *= sigma;
	r += mu;
	return r;
}

// bcm3::LogPdfTnu4, ProbabilityDistributions.cpp:216-224
__device__ __forceinline__ double logpdf_tnu4(double x, double mu, double sigma)
{
	double xn = (x - mu) / sigma;
	return -0.9808292530117262 - 2.5 * log1p(0.25 * xn * xn) - log(sigma);
}

// RHS / Jacobian of the six model types: LikelihoodPopPKTrajectory.cpp:446-467 (one), :469-494 (two), :496-571 (biphasic
// uptake: the absorption rate switches between k_absorption and k_absorption2 at the discontinuities), :573-642 (transit:
// no bolus, a gamma-shaped input k_tr * (k_tr s)^n e^(-k_tr s) / n! * dose into the depot, s = time since the last dose,
// n! by Stirling's series exactly as the reference writes it).
template <int NN, int VARIANT_>
struct PkModel {
	static constexpr int N = NN;
	static constexpr int VARIANT = VARIANT_;
	double ka, kex, kel, kf, kb;
	// biphasic: ka2, switch flag; transit: k_transit, n_transit, log n!, last treatment time, the patient's doses
	double ka2, k_transit, n_transit, log_n_factorial, last_treatment, dose, dose_after, dose_change_time;
	bool biphasic_switch;

	__device__ __forceinline__ double current_ka() const
	{
		if (VARIANT == PKV_BIPHASIC) return biphasic_switch ? ka : ka2;
		return ka;
	}
	__device__ __forceinline__ void rhs(double t, const double (&y)[NN], double (&f)[NN]) const
	{
		const double k = current_ka();
		if (VARIANT == PKV_TRANSIT) {
			double d = dose;
			if (t >= dose_change_time) d = dose_after;
			const double s = t - last_treatment;
			double transit = exp((n_transit * log(k_transit * s) - k_transit * s) - log_n_factorial);
			transit = k_transit * transit * d;
			f[0] = transit - (k + kex) * y[0];
		} else {
			f[0] = -(k + kex) * y[0];
		}
		if (NN == 2) {
			f[1] = k * y[0] - kel * y[1];
		} else {
			f[1] = k * y[0] - kel * y[1] - kf * y[1] + kb * y[2];
			f[2] = kf * y[1] - kb * y[2];
		}
	}
	__device__ __forceinline__ void jac(double (&A)[NN * NN]) const
	{
		const double k = current_ka();
		A[0 * NN + 0] = -(k + kex);
		A[1 * NN + 0] = k;
		if (NN == 2) {
			A[1 * NN + 1] = -kel;
		} else {
			A[1 * NN + 1] = -(kel + kf);
			A[1 * NN + 2] = kb;
			A[2 * NN + 1] = kf;
			A[2 * NN + 2] = -kb;
		}
	}
};
typedef PkModel<2, PKV_PLAIN> PkOneModel;
typedef PkModel<3, PKV_PLAIN> PkTwoModel;
typedef PkModel<2, PKV_BIPHASIC> PkOneBiphasicModel;
typedef PkModel<3, PKV_BIPHASIC> PkTwoBiphasicModel;
typedef PkModel<2, PKV_TRANSIT> PkOneTransitModel;
typedef PkModel<3, PKV_TRANSIT> PkTwoTransitModel;

// LikelihoodPopPKTrajectory::CheckGiveTreatment, cpp:644-671
__device__ __forceinline__ bool check_give_treatment(double t, uint32_t skipped_days, int intermittent)
{
	bool give = true;
	int day = static_cast<int>(floor(t / 24.0));
	if (day >= 0 && day < 29 && ((skipped_days >> day) & 1u)) give = false;
	if (intermittent == 1) {
		double tw = t - 7.0 * 24.0 * floor(t / (7.0 * 24.0));
		if (tw >= 5.0 * 24.0) give = false;
	} else if (intermittent == 2) {
		double tc = t - 28.0 * 24.0 * floor(t / (28.0 * 24.0));
		if (tc >= 21.0 * 24.0) give = false;
	} else if (intermittent == 3) {
		double tw = t - 7.0 * 24.0 * floor(t / (7.0 * 24.0));
		if (tw >= 4.0 * 24.0) give = false;
	}
	return give;
}

// STRIDE = thread stride of the per-thread shared-memory state columns = the largest block this instantiation runs with

// One translation unit per model family instantiates the integrator (poppk_inst_plain.cu, _biphasic.cu, _transit.cu: 8 kernels
// each = one/two compartments x diagnostics x state-column stride) so that they compile in parallel and deterministically.
// Returns a cudaError_t value.
// the two instantiations of the state-column stride (= the largest block size each serves) and the blocks per SM each is
// compiled for: registers per thread follow from threads x blocks (168 at 3 x 128 and 1 x 384)
#ifndef BCM3_POPPK_STRIDE_SMALL
#define BCM3_POPPK_STRIDE_SMALL 128
#endif
#ifndef BCM3_POPPK_SMALL_BLOCKS
#define BCM3_POPPK_SMALL_BLOCKS 3
#endif
#ifndef BCM3_POPPK_STRIDE_BIG
#define BCM3_POPPK_STRIDE_BIG 384
#endif
#ifndef BCM3_POPPK_BIG_BLOCKS
#define BCM3_POPPK_BIG_BLOCKS 1
#endif
#ifndef BCM3_POPPK_KERNEL_ATTR /* experiments: -DBCM3_POPPK_KERNEL_ATTR=__maxnreg__(152) */
#define BCM3_POPPK_KERNEL_ATTR __launch_bounds__(STRIDE, (STRIDE <= BCM3_POPPK_STRIDE_SMALL) ? BCM3_POPPK_SMALL_BLOCKS : BCM3_POPPK_BIG_BLOCKS)
#endif
int launch_poppk_plain(bool two, bool diagnostics, int stride, dim3 grid, int block, size_t smem_bytes, cudaStream_t stream, const PkArgs& a);
int launch_poppk_biphasic(bool two, bool diagnostics, int stride, dim3 grid, int block, size_t smem_bytes, cudaStream_t stream, const PkArgs& a);
int launch_poppk_transit(bool two, bool diagnostics, int stride, dim3 grid, int block, size_t smem_bytes, cudaStream_t stream, const PkArgs& a);

#ifdef BCM3_POPPK_AUX_KERNELS /* compiled once, in bcm3b200.cu; the integrator instances live in poppk_inst_*.cu */
// Sort key of (chain, patient): chain in the high word; in the low word the bits of a positive float that orders the
// patients by expected work: every dose restarts the integrator (about as many steps per dosing interval whatever its
// length), so the number of dosing intervals inside the simulated window comes first, and within it the absorption rate
// (ka as in K0 of poppk_kernel, squashed into [0, 1)).
__global__ void poppk_rank_kernel(const PkArgs a, int C, unsigned long long* __restrict__ keys, int* __restrict__ patients)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
	if (j >= a.P_local || c >= C) return;
	const double* vrow = a.values + (long long)c * a.row_stride;
	const double p = vrow[a.col_patient0 + 2ll * j];
	const double ka = fastpow10(quantile_normal(p, vrow[a.ix[SV_MEAN_ABSORPTION]], vrow[a.ix[SV_SIGMA_ABSORPTION]]));
	const int ntp = a.simulate_until[j];
	const double end_time = (ntp > 0) ? a.time[ntp - 1] : 0.0;
	const double interval = a.dosing_interval[j];
	double intervals = (interval > 0.0) ? floor(end_time / interval) + 1.0 : 1.0;
	if (!(intervals >= 1.0)) intervals = 1.0;
	if (intervals > 1e6) intervals = 1e6;
	double frac = (log10(ka) + 4.0) * 0.125; // ka in 1e-4 .. 1e4 per hour -> [0, 1)
	if (!(frac >= 0.0)) frac = 0.0;
	if (frac > 0.999) frac = 0.999;
	float kf = (float)(intervals + frac);
	if (!(kf >= 0.0f)) kf = INFINITY; // keep the order total whatever the inputs
	const long long e = (long long)c * a.P_local + j;
	// descending work within the chain: the longest-running blocks are scheduled first (shorter tail at the end of the grid)
	keys[e] = ((unsigned long long)c << 32) | (unsigned long long)(0xffffffffu - __float_as_uint(kf));
	patients[e] = j;
}

#endif

template <class Model, bool DIAG, int STRIDE>
__global__ void BCM3_POPPK_KERNEL_ATTR poppk_kernel(const PkArgs a)
{
	constexpr int N = Model::N;
	constexpr unsigned FULL = 0xffffffffu;
	// Dynamic shared memory: the integrator's per-thread state columns FIRST, at compile-time offsets (slot * STRIDE + tid:
	// the hot loop addresses them with nothing but the thread index), then the output times and the simulated values,
	// whose offsets depend on T and are only needed at output times.
	extern __shared__ double smem[];
	double* s_time = smem + BdfSlots<N>::COUNT * STRIDE; // [T]
	double* s_sim = s_time + a.T;                         // [T][blockDim.x]  simulated central-compartment amounts at the output times

	const int tid = threadIdx.x;
	const int c = a.chain_fastest ? blockIdx.x : blockIdx.y;
	int jl = (a.chain_fastest ? blockIdx.y : blockIdx.x) * blockDim.x + tid; // patient index inside this shard
	bool valid = jl < a.P_local;
	// Patients ranked by absorption rate for this chain (poppk_rank_kernel + radix sort): the number of steps a solve takes
	// is ~96 % determined by ka (correlation 0.98 on the config-5 workload: the absorption transient sets the step sizes
	// after every dose), a warp runs until its slowest lane is done and a block until its slowest warp is. In arrival
	// order a warp's lanes are busy 84 % of its trips; ranked by ka, 97 %, and the warps of a block finish together.
	if (a.order && valid) jl = a.order[(long long)c * a.P_local + jl];
	const int T = a.T;

	for (int i = tid; i < T; i += blockDim.x) s_time[i] = a.time[i];
	__syncthreads();

	// ---- K0: parameters of this (chain, patient); cpp:263-295 ----
	const double* vrow = a.values + (long long)c * a.row_stride;
	Model model;
	double sd, sd2, conversion;
	{
		double k_vod = isnan(a.fixed_vod) ? transform_variable(a.tr[SV_VOD], vrow[a.ix[SV_VOD]]) : a.fixed_vod;
		model.kex = transform_variable(a.tr[SV_EXCRETION], vrow[a.ix[SV_EXCRETION]]);
		if (N == 3) {
			if (isnan(a.fixed_periphery_fwd)) {
				model.kf = transform_variable(a.tr[SV_PERIPHERY_FWD], vrow[a.ix[SV_PERIPHERY_FWD]]);
				model.kb = transform_variable(a.tr[SV_PERIPHERY_BWD], vrow[a.ix[SV_PERIPHERY_BWD]]);
			} else {
				model.kf = a.fixed_periphery_fwd;
				model.kb = a.fixed_periphery_bwd;
			}
		} else {
			model.kf = 0.0;
			model.kb = 0.0;
		}
		sd = transform_variable(a.tr[SV_SD], vrow[a.ix[SV_SD]]);
		sd2 = transform_variable(a.tr[SV_SD2], vrow[a.ix[SV_SD2]]);
		double2 pp = make_double2(0.5, 0.5);
		if (valid && !a.single) {
			// the patient's probability pair: one 16-byte load where the pair is 16-byte aligned (always in the compact layout
			// of the host entries; in a caller's [C][nvar] device block only when c * nvar + npk + 2 is even and the block
			// itself is 16-byte aligned -- uniform over the thread block), two 8-byte loads otherwise
			const double* pair = vrow + a.col_patient0 + 2ll * jl;
			if ((reinterpret_cast<uintptr_t>(pair) & 15u) == 0) {
				pp = *reinterpret_cast<const double2*>(pair);
			} else {
				pp.x = pair[0];
				pp.y = pair[1];
			}
		}
		if (a.single) { // LikelihoodPharmacokineticTrajectory.cpp:276-279: no population level, the variables are the rates
			model.ka = transform_variable(a.tr[SV_MEAN_ABSORPTION], vrow[a.ix[SV_MEAN_ABSORPTION]]);
			model.kel = transform_variable(a.tr[SV_MEAN_CLEARANCE], vrow[a.ix[SV_MEAN_CLEARANCE]]) / k_vod;
		} else {
			model.ka = fastpow10(quantile_normal(pp.x, vrow[a.ix[SV_MEAN_ABSORPTION]], vrow[a.ix[SV_SIGMA_ABSORPTION]]));
			model.kel = fastpow10(quantile_normal(pp.y, vrow[a.ix[SV_MEAN_CLEARANCE]], vrow[a.ix[SV_SIGMA_CLEARANCE]])) / k_vod;
		}
		conversion = a.conv_base / k_vod;
		model.ka2 = model.k_transit = model.n_transit = model.log_n_factorial = model.last_treatment = 0.0;
		model.biphasic_switch = true;
		if (Model::VARIANT == PKV_TRANSIT) { // cpp:296-301
			model.n_transit = transform_variable(a.tr[SV_NAMED_A], vrow[a.ix[SV_NAMED_A]]);
			model.k_transit = (model.n_transit + 1) / transform_variable(a.tr[SV_NAMED_B], vrow[a.ix[SV_NAMED_B]]);
			model.log_n_factorial = 0.9189385332046727 + (model.n_transit + 0.5) * log(model.n_transit) - model.n_transit + log(1 + 1 / (12.0 * model.n_transit));
		}
		if (Model::VARIANT == PKV_BIPHASIC) model.ka2 = transform_variable(a.tr[SV_NAMED_B], vrow[a.ix[SV_NAMED_B]]); // cpp:308-309
	}


	double dose = 0.0, dosing_interval = 1.0, dose_after = 0.0, dose_change_time = 0.0;
	int intermittent = 0, ntp = 0;
	uint32_t skipped = 0;
	if (valid) {
		dose = a.dose[jl];
		dosing_interval = a.dosing_interval[jl];
		dose_after = a.dose_after_dose_change[jl];
		dose_change_time = a.dose_change_time[jl];
		intermittent = a.intermittent[jl];
		skipped = a.skipped_days[jl];
		ntp = a.simulate_until[jl];
	}
	model.dose = dose;
	model.dose_after = dose_after;
	model.dose_change_time = dose_change_time;
	// biphasic: the switching time is clipped just below the dosing interval (cpp:304-306)
	double biphasic_switch_time = 0.0;
	if (Model::VARIANT == PKV_BIPHASIC) {
		biphasic_switch_time = transform_variable(a.tr[SV_NAMED_A], vrow[a.ix[SV_NAMED_A]]);
		if (!a.single) biphasic_switch_time = fmin(biphasic_switch_time, dosing_interval - 1e-2); // the single-patient likelihood does not clip (its cpp:302)
	}

	// ---- K1: ODESolver::SolveReturnSolution + ODESolverCVODE::Solve ----
	BdfThread<N, Model, DIAG, STRIDE> S;
	S.sh = smem + tid;
	S.create();

	bool done = !valid || ntp <= 0;
	bool failed = false;
	int tpi = 0;
	int current_step = 0;
	double y[N];
	y[0] = (Model::VARIANT == PKV_TRANSIT) ? 0.0 : dose; // initial_conditions, cpp:366-373
#pragma unroll
	for (int i = 1; i < N; i++) y[i] = 0.0;

	// ODESolver.cpp:109-118: output times at t ~ 0 take the initial condition
	if (!done) {
		while (s_time[tpi] < DBL_EPSILON) {
			s_sim[tpi * blockDim.x + tid] = y[1];
			tpi++;
			if (tpi == ntp) {
				done = true;
				break;
			}
		}
	}
	const double end_time = (ntp > 0) ? s_time[ntp - 1] : 0.0;
	double t = 0.0;
	// SetDiscontinuity(dosing_interval, TreatmentCallback), cpp:362-363 (ignored for time <= 0, ODESolver.cpp:62-71)
	double current_dose_time = dosing_interval;
	double next_disc = (dosing_interval > 0.0) ? dosing_interval : NAN;
	if (Model::VARIANT == PKV_BIPHASIC) { // cpp:357-360: first discontinuity = the switch to the second absorption phase
		current_dose_time = 0.0;
		next_disc = (biphasic_switch_time > 0.0) ? biphasic_switch_time : NAN;
	}
	S.tstopset = !isnan(next_disc);
	S.tstop() = next_disc;

#pragma unroll 1
	for (;;) {
		if (!BCM3_LOOP_VOTE(!done)) break;

		// CVodeReInit + first-call block, all lanes of the warp together
		if (!done) {
			if (!S.restart(t, y, end_time, model, a.rtol, a.atol)) {
				failed = true;
				done = true;
			}
		}
		bool seg_end = false;
		bool newstep = true;

#pragma unroll 1
		for (;;) {
			const bool active = !done && !seg_end;
			if (!BCM3_LOOP_VOTE(active)) break;
			if (active && newstep) {
				newstep = false;
				if (!S.begin_step(a.rtol, a.atol)) { // CV_TOO_MUCH_ACC
					failed = true;
					done = true;
				}
			}
			// the lanes that attempt a step in this trip; the votes inside attempt() range over exactly these
			const bool go = active && !done;
			const unsigned mask = __ballot_sync(FULL, go);
			if (go) {
				const int r = S.attempt(model, mask);
				if (r == BDF_ATTEMPT_FAILED) {
					failed = true;
					done = true;
				} else if (r == BDF_ATTEMPT_DONE) {
					double yout[N], tret;
					const bool tstop_return = S.after_step(yout, tret);
					t = tret;
					current_step++;
					// dense output at every requested time passed by this step, ODESolverCVODE.cpp:406-427
					while (tpi < ntp && tret >= s_time[tpi]) {
						double v[N];
						if (!S.dky(s_time[tpi], v)) {
							failed = true;
							done = true;
							break;
						}
						s_sim[tpi * blockDim.x + tid] = v[1];
						tpi++;
					}
					if (!failed) {
						if (t >= end_time) {
							done = true; // ODESolverCVODE.cpp:436
						} else if (current_step == a.max_steps) {
							failed = true; // :440-446
							done = true;
						} else if (!isnan(next_disc) && (tstop_return || next_disc == t)) {
							seg_end = true; // :448
#pragma unroll
							for (int i = 0; i < N; i++) y[i] = yout[i];
						}
					}
					newstep = true;
				}
			}
		}

		// TreatmentCallback (cpp:673-690), then CVodeReInit(t, y) + CVodeSetStopTime at the top of the loop
		if (!done && seg_end) {
			if (Model::VARIANT != PKV_BIPHASIC) {
				current_dose_time += dosing_interval;
				if (check_give_treatment(t, skipped, intermittent)) {
					double d = dose;
					if (t >= dose_change_time) d = dose_after;
					if (Model::VARIANT == PKV_TRANSIT) model.last_treatment = t;
					else y[0] = y[0] + d;
				}
				next_disc = current_dose_time;
			} else if (model.biphasic_switch) { // TreatmentCallbackBiphasic, cpp:692-718
				model.biphasic_switch = false;
				current_dose_time += dosing_interval;
				next_disc = current_dose_time;
			} else if (check_give_treatment(t, skipped, intermittent)) {
				double d = dose;
				if (t >= dose_change_time) d = dose_after;
				y[0] = y[0] + d;
				model.biphasic_switch = true;
				next_disc = current_dose_time + biphasic_switch_time;
			} else {
				current_dose_time += dosing_interval;
				next_disc = current_dose_time;
			}
			if (!isnan(next_disc) && next_disc < INFINITY) {
				S.tstop() = next_disc;
				S.tstopset = true;
			}
		}
	}

	// ---- observation model: Student-t4 log-pdf, cpp:409-424 ----
	double ll = 0.0;
	if (valid && ntp > 0) {
		if (failed) {
			ll = -INFINITY;
		} else {
			bool broken = false;
			for (int i = 0; i < T; i++) {
				if (i < ntp && !broken) {
					double x = conversion * s_sim[i * blockDim.x + tid];
					double yobs = a.obs_t[(long long)i * a.P_pad + jl];
					if (!isnan(yobs)) ll += logpdf_tnu4(x, yobs, sd + sd2 * ((x < 0.0) ? 0.0 : x));
					if (isnan(x) && !a.single) { // cpp:419-422; the single-patient likelihood has no such test (NaN stays NaN)
						ll = -INFINITY;
						broken = true;
					}
				}
			}
		}
	}

	// ---- K3 (first level): the patient's term goes to its own slot [chain][patient]; poppk_chain_reduce sums them in patient
	// order, so a chain's log-likelihood has the same bits whatever the block size, the ranking or the batch it is part of ----
	if (valid) a.patient_ll[(long long)c * a.P_local + jl] = ll;

	if (DIAG && valid) {
		const long long cj = (long long)c * a.P_local + jl;
		if (a.diag_conc) {
			for (int i = 0; i < T; i++) {
				double x = NAN;
				if (i < ntp && !failed) x = conversion * s_sim[i * blockDim.x + tid];
				a.diag_conc[cj * T + i] = x;
			}
		}
		if (a.diag_counters) {
			int32_t* k = a.diag_counters + cj * 8;
			k[0] = current_step;
			k[1] = S.cnt.nfe;
			k[2] = S.cnt.nsetups;
			k[3] = S.cnt.nje;
			k[4] = S.cnt.netf;
			k[5] = S.cnt.ncfn;
			k[6] = S.cnt.nni;
			k[7] = (ntp > 0 && !failed) ? 1 : 0;
		}
	}
}

#ifdef BCM3_POPPK_AUX_KERNELS
// K3 (second level): patient_ll [C][P_local] -> partial[3][C] = (sum of the finite terms, global index of the first -inf
// patient, global index of the first NaN patient). One block per chain; thread t adds patients t, t + 256, ... in order, then
// a fixed tree: the order depends on P_local only (run-to-run and batch-to-batch identical bits).
__global__ void poppk_chain_reduce(const double* __restrict__ patient_ll, int P_local, int patient_offset, int C, double* __restrict__ partial)
{
	__shared__ double sh[3][256];
	const int c = blockIdx.x;
	const int tid = threadIdx.x;
	const double* v = patient_ll + (long long)c * P_local;
	double s = 0.0, mi = INFINITY, mn = INFINITY;
	for (int j = tid; j < P_local; j += blockDim.x) {
		const double ll = v[j];
		const double gidx = (double)(patient_offset + j);
		if (isnan(ll) || ll == INFINITY) mn = fmin(mn, gidx);
		else if (ll == -INFINITY) mi = fmin(mi, gidx);
		else s += ll;
	}
	sh[0][tid] = s;
	sh[1][tid] = mi;
	sh[2][tid] = mn;
	__syncthreads();
	for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
		if (tid < off) {
			sh[0][tid] += sh[0][tid + off];
			sh[1][tid] = fmin(sh[1][tid], sh[1][tid + off]);
			sh[2][tid] = fmin(sh[2][tid], sh[2][tid + off]);
		}
		__syncthreads();
	}
	if (tid == 0) {
		partial[0 * C + c] = sh[0][0];
		partial[1 * C + c] = sh[1][0];
		partial[2 * C + c] = sh[2][0];
	}
}
#endif

} // namespace bcm3b200
