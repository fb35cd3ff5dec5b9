// cellpop_group.cuh -- K0+K2 of the cellpop path: one ODE system (one simulated cell of one chain) per GROUP of G lanes
// (G = 2..32, a power of two chosen per model so that every lane owns E = ceil(N / G) <= 3-4 components).
//
// Same algorithm and reference line map as bdf_thread.cuh / cellpop_warp.cuh (CVODE 5.3.0 BDF + modified Newton with the
// difference-quotient Jacobian of ODESolverCVODE.cpp:496-537 and BCM3's partial-pivot LU,
// EigenPartialPivLUSomewhatSparse.h:38-105). What is different is where things live:
//
//   * component i of every integrator vector belongs to lane (i mod G) of the group, slot (i div G): the Nordsieck array,
//     weights and corrections are REGISTER arrays of E doubles indexed only by compile-time constants (the static_for
//     loops of bdf_thread.cuh), so the vector work of a step costs no address arithmetic and no memory traffic;
//   * norms are xor-butterfly reductions over the G lanes (every lane ends with the same bits, so all scalar
//     bookkeeping -- step size, order, counters -- is replicated per lane and stays group-uniform without broadcasts);
//   * the Newton matrix (LU factors) sits in SHARED memory, one padded row-major block per cell whose strides make the
//     row-per-lane and column-per-lane access patterns of all groups of a half-warp bank-conflict free; the saved
//     Jacobian sits in a global scratch block per resident group (touched only by linear setups);
//   * the right-hand side is evaluated by every lane of the group from a shared copy of y (lanes keep their own
//     components of the result); the N perturbed evaluations of the difference-quotient Jacobian are spread over the
//     lanes, one column each; LU factorisation, permutation and the triangular solves are cooperative over the group;
//   * cells are handed out from a global work queue to resident groups (grid = what fits on the GPU), so a group that
//     finishes a cell picks up the next one instead of idling until the slowest cell of its block is done.
//
// The model library is compiled with -fmad=false so that the generated right-hand side rounds exactly like the checker's
// (see cellpop_host.cuh); the integrator's own multiply-adds are written as explicit fma() -- the reference's host build
// contracts them too.
//
// Control flow is kept converged the way poppk_kernel does it: one step attempt per loop trip, the Newton loops made
// warp-uniform with votes over the participating lanes, the warps of a block in lock-step.
#pragma once

#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "bdf_thread.cuh"
#include "cellpop_args.h"

#ifndef CP_GROUP
#define CP_GROUP 4
#endif
#ifndef CP_GROUP_WARPS
#define CP_GROUP_WARPS 4
#endif
#ifndef CP_GROUP_MIN_BLOCKS
#define CP_GROUP_MIN_BLOCKS 1
#endif

namespace cellpop_group {

using bcm3b200::static_for;
using bcm3b200::static_rfor;

constexpr int N = CP_N;
constexpr int G = CP_GROUP;
constexpr int E = (N + G - 1) / G;
constexpr bool PADDED = (E * G != N);
constexpr int CPW = 32 / G; // cells per warp
constexpr int WPB = CP_GROUP_WARPS;
constexpr int BS = 32 * WPB;
constexpr int QMAX = 5;
#ifndef CP_SJ_UNROLL
#define CP_SJ_UNROLL 1
#endif
constexpr int SJ_UNROLL = CP_SJ_UNROLL; // rows of the saved Jacobian in flight when the Newton matrix is rebuilt from it
constexpr unsigned FULL = 0xffffffffu;
constexpr double UROUND = DBL_EPSILON;

// Shared-memory block of one cell: M (N rows, stride RS), y (N), f (N), row permutation (N ints), cold scalars.
// RS odd: the G lanes of a group touching G consecutive rows of one column hit distinct bank pairs. CS = RS * G (mod 16):
// then bank pair (CS * g + RS * lg) mod 16 = RS * (G * g + lg) is a bijection over the 16 lanes of a half-warp.
constexpr int RS = N | 1;
// CP_M_GLOBAL = 1: the Newton matrix of a cell lives in a global-memory block per resident group (served by L1 / L2) instead
// of the cell's shared block -- 20 KB of the 24.6 KB per cell at 50 species, the difference between 9 and 12 warps per SM
#ifndef CP_M_GLOBAL
#define CP_M_GLOBAL 0
#endif
constexpr int M_SHARED = CP_M_GLOBAL ? 0 : N * RS;
constexpr int OFF_Y = M_SHARED;
constexpr int OFF_F = OFF_Y + N;
constexpr int OFF_PERM = OFF_F + N;
constexpr int OFF_SCAL = OFF_PERM + (N + 1) / 2;
// per-cell scalars that are touched a few times per step at most: kept out of the register file (every lane of the
// group would hold a copy). Lanes of a group always store identical values, so no synchronisation is involved.
enum { SC_TAU = 0 /* [1..5] */, SC_HU = 6, SC_SAVED_TQ5, SC_SAVED_T, SC_HSCALE, SC_ETAMAX, SC_CREATION, SC_END, SC_HPRIME, SC_ETA, SC_GAMMAP, SC_CRATE, SC_DELP, SC_ACNRM,
       SC_TSTOP, SC_NEXT_DISC,
       SC_L /* [0..5] */, SC_TQ = SC_L + 6 /* [1..5] */,
       SC_OV = SC_TQ + 6 /* per-cell parameter overrides */, SC_COUNT = SC_OV + (CP_NUM_OVERRIDES > 0 ? CP_NUM_OVERRIDES : 1) };
// Nordsieck columns 2..5, lane-private: element (j, e) of lane lg at OFF_ZNH + ((j - 2) * E + e) * G + lg
constexpr int OFF_ZNH = OFF_SCAL + SC_COUNT;
static_assert(SC_OV == CP_GROUP_SCALARS, "cellpop_host.cuh sizes the shared block with CP_GROUP_SCALARS");
// rate-law values of the lane-parallel right-hand side (one double per reaction), when the model has one
#ifndef CP_RHS_LANES
#define CP_RHS_LANES 0
#define CP_NUM_RATELAWS 0
#endif
constexpr int OFF_RL = OFF_ZNH + 4 * E * G;
constexpr int REGION_MIN = OFF_RL + CP_NUM_RATELAWS;
constexpr int cell_stride()
{
	const int want = (RS * G) % 16;
	int cs = REGION_MIN;
	while (cs % 16 != want) cs++;
	return cs;
}
constexpr int CS = cell_stride();

enum { T_RETRY = 0, T_DONE = 1, T_FAILED = -1 };

// ---- the generated right-hand side: ONE instance for the whole kernel ----
// Every use (Newton residual, the perturbed evaluations of the difference-quotient Jacobian, cvHin, the order-1 restart)
// goes through rhs_eval(): y is read from shared memory with component j replaced by yj (j = -1: none), the result is
// written to shared memory with a stride (1: the f buffer, RS: a column of the Newton matrix). Inlining the generated
// code at each use costs ~750 instructions per copy, and the hot loop has to stay small.
struct CellParameters {
	const double* base; // transformed variables of the chain
	const double* ov;   // per-cell overrides (shared memory)
	__device__ __forceinline__ double operator[](int k) const
	{
		CP_PARAM_OVERRIDE_BODY
		return __ldg(base + k);
	}
};
struct ConstVector {
	const double* p;
	__device__ __forceinline__ double operator[](int k) const { return __ldg(p + k); }
};
// constant species with the one a pulsed treatment drives replaced by its current value (Cell::SetTreatmentConcentration)
struct ConstSpecies {
	const double* p;
	int treat_ix;
	double treat_value;
	__device__ __forceinline__ double operator[](int k) const { return (k == treat_ix) ? treat_value : __ldg(p + k); }
};
struct SpeciesAt {
	const double* y;
	int j;
	double yj;
	__device__ __forceinline__ double operator[](int i) const { return (i == j) ? yj : y[i]; }
};
struct OutStrided {
	double* p;
	int stride;
	__device__ __forceinline__ double& operator[](int i) const { return p[i * stride]; }
};
struct SpeciesPlain {
	const double* y;
	__device__ __forceinline__ double operator[](int i) const { return y[i]; }
};
// Two instances: the plain one (every Newton residual: hot) reads y as it is; the perturbed one (difference-quotient
// Jacobian columns: one evaluation in ~25) pays a compare-and-select per species read.
__device__ __noinline__ void rhs_eval(unsigned y_off, unsigned out_off, unsigned ov_off, const double* tv, const double* constant_species,
                                      const double* non_sampled, int treat_ix, double treat_value)
{
	extern __shared__ double smem_d[];
	generated_derivative(OutStrided{ smem_d + out_off, 1 }, SpeciesPlain{ smem_d + y_off }, ConstSpecies{ constant_species, treat_ix, treat_value },
	                     CellParameters{ tv, smem_d + ov_off }, ConstVector{ non_sampled });
}
__device__ __noinline__ void rhs_eval_perturbed(unsigned y_off, int j, double yj, double* out, int out_stride, unsigned ov_off, const double* tv,
                                                const double* constant_species, const double* non_sampled, int treat_ix, double treat_value)
{
	extern __shared__ double smem_d[];
	generated_derivative(OutStrided{ out, out_stride }, SpeciesAt{ smem_d + y_off, j, yj },
	                     ConstSpecies{ constant_species, treat_ix, treat_value }, CellParameters{ tv, smem_d + ov_off }, ConstVector{ non_sampled });
}

#if CP_RHS_LANES
// The lane-parallel form (cellpop_host.cuh::cellpop_lane_rhs): the lanes of the cell's group evaluate different reactions at the
// same time into the cell's rate-law buffer, then every lane assembles the species it owns and leaves them in the f
// buffer. ONE instance for all uses, like rhs_eval. The caller has published y (with its group barriers).
__device__ __noinline__ void rhs_eval_lanes(unsigned y_off, unsigned f_off, unsigned rl_off, unsigned ov_off, const double* tv, const double* constant_species,
                                            const double* non_sampled, int treat_ix, double treat_value, int lg, unsigned gmask)
{
	extern __shared__ double smem_d[];
	double* const rl = smem_d + rl_off;
	generated_ratelaws_lanes<G>(lg, rl, SpeciesPlain{ smem_d + y_off }, ConstSpecies{ constant_species, treat_ix, treat_value }, CellParameters{ tv, smem_d + ov_off },
	                            ConstVector{ non_sampled });
	__syncwarp(gmask);
#pragma unroll
	for (int e = 0; e < E; e++) {
		const int i = lg + G * e;
		if (!PADDED || i < N) smem_d[f_off + i] = generated_assemble(i, rl);
	}
}
#endif

#ifndef CP_GROUP_LOCKSTEP_EVERY
#define CP_GROUP_LOCKSTEP_EVERY 1
#endif
#ifndef CP_GROUP_LOCKSTEP
#define CP_GROUP_LOCKSTEP 1
#endif
// Lock-step in TEAMS of this many warps instead of the whole block (0 = the whole block): every team meets at its own named
// barrier (bar.red.or with the team's thread count) -- fewer warps wait for the slowest one, fewer warps share a fetch
#ifndef CP_GROUP_LOCKSTEP_TEAM
#define CP_GROUP_LOCKSTEP_TEAM 0
#endif
// 1: dividing / dying cells (CpArgs: items, per-cell records, event species). The model library of an experiment without
// divide_cells and without an "apoptosis" species is built with 0 and contains none of that code.
#ifndef CP_DIVISION
#define CP_DIVISION 0
#endif
// number of data sets that share the cells' integration (CpArgs::tp_rows); 1 = the plain layout, no indirection
#ifndef CP_NUM_DATASETS
#define CP_NUM_DATASETS 1
#endif
#ifndef CP_GROUP_BATCHED
#define CP_GROUP_BATCHED 1
#endif
#ifndef CP_GROUP_STATIC_LU_MAX
#define CP_GROUP_STATIC_LU_MAX 0
#endif
// Large models (16 or 32 lanes per cell, one or two cells per warp): LU column updates visit only the columns whose pivot-row
// entry is non-zero (as the reference's LU does), the triangular solves run slot by slot with compile-time slots, the pivot
// search uses three warp-reduce instructions instead of log2(G) shuffle stages. Measured (B200, kernel ms): 50 species x
// 6 000 cells x 16 chains 2 126 -> 1 464 with all of them and the lane-parallel right-hand side (each worth 4-24 %); with 4
// lanes per cell (12 species, 8 cells per warp) every one of them LOSES (128 -> 178 ms: the per-group loops over the non-zero
// columns and over the reactions of a shape diverge between the 8 groups of a warp), so small models keep the dense forms.
#ifndef CP_LU_SKIP_ZEROS
#define CP_LU_SKIP_ZEROS (CP_GROUP >= 16)
#endif
#ifndef CP_SOLVE_SLOTTED
#define CP_SOLVE_SLOTTED (CP_GROUP >= 16)
#endif
#ifndef CP_PIVOT_REDUX
#define CP_PIVOT_REDUX (CP_GROUP >= 16)
#endif
// Triangular solves that leave out the columns of L (below the diagonal) and of U (above it) that hold nothing but zeros:
// the factorisation's own votes say which columns those are (the multipliers of a pivot step, the non-zero entries of a
// finished pivot row), kept as one bit per column. A skipped column would have subtracted x_k * 0 from every component, so
// the result has the same bits. In a signalling network most columns of U are empty above the diagonal (a cascade: all
// but the fill-in column), and every skipped column is one shuffle + multiply-add link less in the solve's dependent chain.
// Measured (B200, 50 species x 6 000 cells x 16 chains, same bits): 1 536 ms with it, 1 500 ms without -- the run-time loops
// over the set bits cost more than the skipped columns save (the pivoting of a stiff I - gamma J fills U's superdiagonal,
// and the compile-time slotted loops are unrolled four columns deep); N = 24: 377 vs 377 ms. An experiment switch, off.
// profiles/r02_cellpop50_variants_solve_skip.log
#ifndef CP_SOLVE_SKIP
#define CP_SOLVE_SKIP 0
#endif
#ifndef CP_SJ_UNROLL
#define CP_SJ_UNROLL 1
#endif

__device__ __forceinline__ double step_root(double base, int k) { return bcm3b200::bdf_root_halley(base, k); }

// second half of N_VWrmsNorm: butterfly over the group, mean, square root -- one copy of the code for all call sites
__device__ __noinline__ double norm_finish(double s, unsigned gmask)
{
#pragma unroll
	for (int d = G / 2; d >= 1; d >>= 1) s += __shfl_xor_sync(gmask, s, d);
	return sqrt(s / N);
}

struct GroupBdf {
	// ---- distributed vectors (slot e = component lg + G * e) ----
	double zn01[2][E]; // Nordsieck columns 0 and 1; columns 2..5 are lane-private words of the cell's shared block
	double* znh;
	double ewt[E], acor[E];
	// ---- replicated scalars ----
	double tn, h;
	double gamma, gamrat, rl1;
	double* sc; // shared: tau[1..5], hu, saved_tq5, saved_t, hscale, etamax, creation/end time, l[] and tq[] across the Newton loop
	int q, qprime, L, qwait, nst, nstlp, nstlj, nflag, ncf, nef;
	bool nls_jcur;
	int nfe, nsetups, nje;
#if CP_SOLVE_SKIP
	unsigned lmask[E], umask[E]; // bit kk of slot s: column G * s + kk of L / U has a non-zero entry off the diagonal
#endif
	// ---- placement ----
	double* M;    // shared (global with CP_M_GLOBAL)
	double* ybuf; // shared, N
	double* fbuf; // shared, N
	int* perm;    // shared, N
	double* SJ;   // global, N * N row-major
	unsigned region_off; // offset of the cell's shared block in doubles
	int lg;
	unsigned gmask;
	int gbase; // lane index (within the warp) of the group's lane 0
	// ---- model ----
	const double* constant_species;
	const double* non_sampled;
	const double* tv; // transformed variables of the cell's chain
	double reltol, abstol, hmin, hmax_inv;
	// pulsed treatment (TreatmentTrajectoryPulses.cpp): constant species it drives (-1: none), sorted pulse times
	int treat_ix, treat_n;
	const double* treat_times;
	bool tstopset;

	__device__ __forceinline__ double& tau(int j) const { return sc[SC_TAU + j]; }
	__device__ __forceinline__ double& hu() const { return sc[SC_HU]; }
	__device__ __forceinline__ double& saved_tq5() const { return sc[SC_SAVED_TQ5]; }
	__device__ __forceinline__ double& saved_t() const { return sc[SC_SAVED_T]; }
	__device__ __forceinline__ double& hscale() const { return sc[SC_HSCALE]; }
	__device__ __forceinline__ double& etamax() const { return sc[SC_ETAMAX]; }
	__device__ __forceinline__ double& hprime() const { return sc[SC_HPRIME]; }
	__device__ __forceinline__ double& eta() const { return sc[SC_ETA]; }
	__device__ __forceinline__ double& gammap() const { return sc[SC_GAMMAP]; }
	__device__ __forceinline__ double& crate() const { return sc[SC_CRATE]; }
	__device__ __forceinline__ double& delp() const { return sc[SC_DELP]; }
	__device__ __forceinline__ double& acnrm() const { return sc[SC_ACNRM]; }
	__device__ __forceinline__ double& tstop() const { return sc[SC_TSTOP]; }

	// TreatmentTrajectoryPulses::GetConcentration(t, creation_time), .cpp:21-41
	__device__ __forceinline__ double treatment_value(double t) const
	{
		if (treat_ix < 0) return 0.0;
		const double global_time = t + sc[SC_CREATION];
		for (int i = 0; i < treat_n; i++) {
			const double t_in_pulse = global_time - __ldg(treat_times + i) - 2.0;
			if (t_in_pulse >= 14.0) continue;
			else if (t_in_pulse <= 0.0) return 0.0;
			else if (t_in_pulse < 2.0) return t_in_pulse * 0.5;
			else if (t_in_pulse < 10.0) return 1.0;
			else return 1 - (t_in_pulse - 10.0) * 0.25;
		}
		return 0.0;
	}
	// TreatmentTrajectoryPulses::NextDiscontinuity, .cpp:52-71 (exact comparisons: `time` is a stop time this function returned)
	__device__ __forceinline__ double next_discontinuity(double time) const
	{
		const double creation = sc[SC_CREATION];
		for (int i = 0; i < treat_n; i++) {
			const double tp = __ldg(treat_times + i);
			if (time == tp - creation + 2.0) return tp - creation + 4.0;
			else if (time == tp - creation + 4.0) return tp - creation + 10.0;
			else if (time == tp - creation + 10.0) return tp - creation + 14.0;
			else if (time == tp - creation + 14.0) {
				if (i < treat_n - 1) return __ldg(treat_times + i + 1) - creation + 2.0;
				return __longlong_as_double(0x7ff8000000000000ll);
			}
		}
		return __longlong_as_double(0x7ff8000000000000ll);
	}
	// Cell::Simulate, Cell.cpp:212-229: the first discontinuity ahead of the cell, NaN if none
	__device__ __forceinline__ double first_discontinuity() const
	{
		const double nan = __longlong_as_double(0x7ff8000000000000ll);
		if (treat_ix < 0 || treat_n <= 0) return nan;
		double d = __ldg(treat_times) - sc[SC_CREATION] + 2.0;
		while (d < 0.0) d = next_discontinuity(d);
		return (d > 0.0) ? d : nan; // NaN compares false
	}
	template <int J>
	__device__ __forceinline__ double& Z(int e)
	{
		if constexpr (J < 2) return zn01[J][e];
		else return znh[((J - 2) * E + e) * G];
	}

	__device__ __forceinline__ int idx(int e) const { return lg + G * e; }
	__device__ __forceinline__ bool own(int e) const { return !PADDED || (lg + G * e < N); }
	__device__ __forceinline__ void gsync() const { __syncwarp(gmask); }

	__device__ __forceinline__ double gsum(double v) const
	{
#pragma unroll
		for (int d = G / 2; d >= 1; d >>= 1) v += __shfl_xor_sync(gmask, v, d);
		return v;
	}
	__device__ __forceinline__ double gmax(double v) const
	{
#pragma unroll
		for (int d = G / 2; d >= 1; d >>= 1) {
			const double o = __shfl_xor_sync(gmask, v, d);
			v = (o > v) ? o : v;
		}
		return v;
	}
	// N_VWrmsNorm (nvector_serial_eigen.cpp:386-396); the sum runs lane-major instead of left to right
	__device__ __forceinline__ double wrms(const double (&x)[E]) const
	{
		double s = 0.0;
#pragma unroll
		for (int e = 0; e < E; e++) {
			const double p = x[e] * ewt[e];
			s = fma(p, p, s);
		}
		return norm_finish(s, gmask);
	}
	__device__ __forceinline__ void set_ewt()
	{
#pragma unroll
		for (int e = 0; e < E; e++) ewt[e] = own(e) ? 1.0 / (reltol * fabs(Z<0>(e)) + abstol) : 0.0;
	}
	// publish a distributed vector into the cell's shared y buffer
	__device__ __forceinline__ void publish(double* buf, const double (&x)[E]) const
	{
		gsync(); // earlier readers of the buffer are done
#pragma unroll
		for (int e = 0; e < E; e++)
			if (own(e)) buf[idx(e)] = x[e];
		gsync();
	}
	// Cell::solver_rhs_fn (Cell.cpp:423-433) at the y held in ybuf: every lane of the group evaluates the whole vector into
	// fbuf (identical stores) and reads back the components it owns
	__device__ __forceinline__ void rhs_shared(double t, double (&f)[E])
	{
#if CP_RHS_LANES
		rhs_eval_lanes(region_off + OFF_Y, region_off + OFF_F, region_off + OFF_RL, region_off + OFF_SCAL + SC_OV, tv, constant_species, non_sampled, treat_ix,
		               treatment_value(t), lg, gmask);
#else
		rhs_eval(region_off + OFF_Y, region_off + OFF_F, region_off + OFF_SCAL + SC_OV, tv, constant_species, non_sampled, treat_ix, treatment_value(t));
#endif
#pragma unroll
		for (int e = 0; e < E; e++) f[e] = own(e) ? fbuf[idx(e)] : 0.0; // its own stores (lanes form) or identical stores of all lanes
		nfe++;
	}

	// persistent members that CVodeCreate zero-fills once and CVodeReInit never touches
	__device__ __forceinline__ void create()
	{
		static_for<0, 6>([&](auto J) {
			constexpr int j = decltype(J)::value;
			tau(j) = 0.0;
#pragma unroll
			for (int e = 0; e < E; e++) Z<j>(e) = 0.0;
		});
#pragma unroll
		for (int e = 0; e < E; e++) acor[e] = 0.0;
		gammap() = 0.0; crate() = 1.0; delp() = 0.0; acnrm() = 0.0; saved_tq5() = 0.0;
		eta() = gamma = gamrat = rl1 = 0.0;
		hu() = 0.0;
		nls_jcur = false; nstlj = 0; nfe = 0; nsetups = 0; nje = 0;
		tn = 0.0; q = 1; L = 2; qwait = 2; etamax() = BDF_ETAMX1; nst = 0; nstlp = 0; qprime = 1;
		saved_t() = 0.0; ncf = nef = 0; nflag = bcm3b200::BDF_FIRST_CALL;
		h = hprime() = 0.0;
		hscale() = 0.0;
		tstopset = false;
		tstop() = 0.0;
	}

	// CVodeReInit(t0, y0) + the first-call block of CVode(tout, CV_ONE_STEP) (cvode.c:586-665, 1068-1155); a stop time, if
	// any, has been set by the caller (CVodeSetStopTime). False where CVode returns < 0.
	__device__ __forceinline__ bool restart(double t0, const double (&y0)[E], double tout_in)
	{
		tn = t0; q = 1; L = 2; qwait = 2; etamax() = BDF_ETAMX1; hu() = 0.0; nst = 0; nstlp = 0;
		nstlj = 0; nls_jcur = false;
#pragma unroll
		for (int e = 0; e < E; e++) Z<0>(e) = own(e) ? y0[e] : 0.0;
		set_ewt();
		publish(ybuf, zn01[0]);
		rhs_shared(tn, zn01[1]);
		if (tstopset) {
			if ((tstop() - tn) * (tout_in - tn) <= 0.0) return false; // CV_ILL_INPUT
		}
		double tout = tout_in;
		if (tstopset && (tout - tn) * (tout - tstop()) > 0.0) tout = tstop();
		// cvHin (cvode.c:1884-1984)
		const double tdiff = tout - tn;
		if (tdiff == 0.0) return false;
		const double sign = (tdiff > 0.0) ? 1.0 : -1.0;
		const double tdist = fabs(tdiff);
		const double tround = UROUND * fmax(fabs(tn), fabs(tout));
		if (tdist < 2.0 * tround) return false;
		const double hlb = BDF_HLB_FACTOR * tround;
		double hub_inv = -INFINITY;
#pragma unroll
		for (int e = 0; e < E; e++) {
			if (own(e)) {
				double t2 = fabs(Z<0>(e));
				double t1 = 1.0 / ewt[e];
				t1 = BDF_HUB_FACTOR * t2 + t1;
				t2 = fabs(Z<1>(e));
				t1 = t2 / t1;
				hub_inv = (t1 > hub_inv) ? t1 : hub_inv;
			}
		}
		hub_inv = gmax(hub_inv);
		double hub = BDF_HUB_FACTOR * tdist;
		if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;
		double hg = sqrt(hlb * hub);
		if (hub < hlb) {
			h = (sign < 0.0) ? -hg : hg;
		} else {
			double hnew = hg;
#pragma unroll 1
			for (int count1 = 1; count1 <= BDF_MAX_ITERS; count1++) {
				const double hgs = hg * sign;
				double ytmp[E], ftmp[E];
#pragma unroll
				for (int e = 0; e < E; e++) ytmp[e] = fma(hgs, Z<1>(e), Z<0>(e));
				publish(ybuf, ytmp);
				rhs_shared(tn + hgs, ftmp);
				const double c = 1.0 / hgs;
#pragma unroll
				for (int e = 0; e < E; e++) ftmp[e] = c * (ftmp[e] - Z<1>(e));
				const double yddnrm = wrms(ftmp);
				hnew = (yddnrm * hub * hub > 2.0) ? sqrt(2.0 / yddnrm) : sqrt(hg * hub);
				if (count1 == BDF_MAX_ITERS) break;
				const double hrat = hnew / hg;
				if ((hrat > 0.5) && (hrat < 2.0)) break;
				if ((count1 > 1) && (hrat > 2.0)) {
					hnew = hg;
					break;
				}
				hg = hnew;
			}
			double h0 = BDF_H_BIAS * hnew;
			if (h0 < hlb) h0 = hlb;
			if (h0 > hub) h0 = hub;
			if (sign < 0.0) h0 = -h0;
			h = h0;
		}
		if (hmax_inv > 0.0) { // cvode.c:1121-1122 (a step-size ceiling is rare: the division stays out of the common path)
			const double rh = fabs(h) * hmax_inv;
			if (rh > 1.0) h /= rh;
		}
		if (fabs(h) < hmin) h *= hmin / fabs(h);
		if (tstopset) {
			if ((tn + h - tstop()) * h > 0.0) h = (tstop() - tn) * (1.0 - 4.0 * UROUND);
		}
		hscale() = h;
		hprime() = h;
#pragma unroll
		for (int e = 0; e < E; e++) Z<1>(e) *= h;
		return true;
	}

	// cvRescale, cvode.c:2384-2400
	__device__ __forceinline__ void rescale()
	{
		double c = eta();
		static_for<1, QMAX + 1>([&](auto J) {
			constexpr int j = decltype(J)::value;
			if (j <= q) {
#pragma unroll
				for (int e = 0; e < E; e++) Z<j>(e) *= c;
				c = eta() * c;
			}
		});
		h = hscale() * eta();
		hscale() = h;
	}

	// cvIncreaseBDF, cvode.c:2310-2340
	__device__ __forceinline__ void increase_bdf()
	{
		double ll[6];
#pragma unroll
		for (int i = 0; i < 6; i++) ll[i] = 0.0;
		double alpha1 = 1.0, prod = 1.0, xiold = 1.0, alpha0 = -1.0;
		ll[2] = 1.0;
		const double hs = hscale();
		double hsum = hs;
		static_for<1, QMAX - 1>([&](auto J) {
			constexpr int j = decltype(J)::value;
			if (j < q) {
				hsum += tau(j + 1);
				const double xi = hsum / hs;
				prod *= xi;
				alpha0 -= 1.0 / (j + 1);
				alpha1 += 1.0 / xi;
				static_rfor<2, j + 3>([&](auto I) {
					constexpr int i = decltype(I)::value;
					ll[i] = ll[i] * xiold + ll[i - 1];
				});
				xiold = xi;
			}
		});
		const double A1 = (-alpha0 - alpha1) / prod;
		double znL[E];
#pragma unroll
		for (int e = 0; e < E; e++) znL[e] = A1 * Z<QMAX>(e);
		static_for<2, QMAX + 1>([&](auto J) {
			constexpr int j = decltype(J)::value;
			if (j <= q) {
#pragma unroll
				for (int e = 0; e < E; e++) Z<j>(e) = fma(ll[j], znL[e], Z<j>(e));
			} else if (j <= L) {
#pragma unroll
				for (int e = 0; e < E; e++) Z<j>(e) = znL[e];
			}
		});
	}

	// cvDecreaseBDF, cvode.c:2352-2374
	__device__ __forceinline__ void decrease_bdf()
	{
		double ll[6];
#pragma unroll
		for (int i = 0; i < 6; i++) ll[i] = 0.0;
		ll[2] = 1.0;
		double hsum = 0.0;
		const double hs = hscale();
		static_for<1, QMAX - 1>([&](auto J) {
			constexpr int j = decltype(J)::value;
			if (j <= q - 2) {
				hsum += tau(j);
				const double xi = hsum / hs;
				static_rfor<2, j + 3>([&](auto I) {
					constexpr int i = decltype(I)::value;
					ll[i] = ll[i] * xi + ll[i - 1];
				});
			}
		});
		double znq[E];
#pragma unroll
		for (int e = 0; e < E; e++) znq[e] = Z<2>(e);
		static_for<3, QMAX + 1>([&](auto J) {
			constexpr int j = decltype(J)::value;
			if (j <= q) {
#pragma unroll
				for (int e = 0; e < E; e++) znq[e] = Z<j>(e);
			}
		});
		if (q > 2) {
			static_for<2, QMAX>([&](auto J) {
				constexpr int j = decltype(J)::value;
				if (j < q) {
#pragma unroll
					for (int e = 0; e < E; e++) Z<j>(e) = fma(-ll[j], znq[e], Z<j>(e));
				}
			});
		}
	}

	__device__ __forceinline__ void adjust_order(int deltaq)
	{
		if ((q == 2) && (deltaq != 1)) return;
		if (deltaq == 1) increase_bdf();
		else if (deltaq == -1) decrease_bdf();
	}

	// CVode loop head + cvStep head (cvode.c:1294-1337, 2094-2102). False on CV_TOO_MUCH_ACC.
	__device__ __forceinline__ bool begin_step()
	{
		if (nst > 0) set_ewt();
		if (UROUND * wrms(zn01[0]) > 1.0) return false;
		saved_t() = tn;
		ncf = 0;
		nef = 0;
		nflag = bcm3b200::BDF_FIRST_CALL;
		if ((nst > 0) && (hprime() != h)) {
			if (qprime != q) {
				adjust_order(qprime - q);
				q = qprime;
				L = q + 1;
				qwait = L;
			}
			rescale();
		}
		return true;
	}

	// cvRestore, cvode.c:2918-2927
	__device__ __forceinline__ void restore()
	{
		tn = saved_t();
		static_for<1, QMAX + 1>([&](auto K) {
			constexpr int k = decltype(K)::value;
			static_rfor<k, QMAX + 1>([&](auto J) {
				constexpr int j = decltype(J)::value;
				if (j <= q) {
#pragma unroll
					for (int e = 0; e < E; e++) Z<j - 1>(e) = Z<j - 1>(e) - Z<j>(e);
				}
			});
		});
	}

	// cvLsSetup: Newton matrix M = I - gamma J from a fresh difference-quotient Jacobian (jb) or the saved one, then LU
	__device__ __forceinline__ void linear_setup(bool jb, const double (&y)[E], const double (&fy)[E])
	{
		const double neg_gamma = -gamma;
		if (jb) {
			// ODESolverCVODE::DifferenceQuotientJacobian (ODESolverCVODE.cpp:496-537); y is in ybuf and f(y) in fbuf (residual)
			gsync();
			const double srur = sqrt(UROUND);
			const double treat_now = treatment_value(tn);
			const double fnorm = wrms(fy);
			const double minInc = (fnorm != 0.0) ? (1000.0 * fabs(h) * UROUND * N * fnorm) : 1.0;
			// one instance of the right-hand side code for all slots: the slot's y and weight are picked with selects
#pragma unroll 1
			for (int e = 0; e < E; e++) {
				double ye = y[0], we = ewt[0];
				static_for<1, E>([&](auto J) {
					constexpr int jj = decltype(J)::value;
					if (e >= jj) {
						ye = y[jj];
						we = ewt[jj];
					}
				});
				const int j = lg + G * e;
				if (!PADDED || j < N) {
					const double inc = fmax(srur * fabs(ye), minInc / we);
					// f(y + inc e_j) into column j of M, then the difference quotient into the saved Jacobian and, scaled
					// (SUNMatScaleAddI(-gamma, A): A = -gamma * J, then the unit diagonal added), back into M
					rhs_eval_perturbed(region_off + OFF_Y, j, ye + inc, M + j, RS, region_off + OFF_SCAL + SC_OV, tv, constant_species, non_sampled,
					                   treat_ix, treat_now);
					const double inc_inv = 1.0 / inc;
#pragma unroll 1
					for (int i = 0; i < N; i++) {
						const double Jij = inc_inv * (M[i * RS + j] - fbuf[i]);
						SJ[i * N + j] = Jij;
						double m = neg_gamma * Jij;
						if (i == j) m += 1.0;
						M[i * RS + j] = m;
					}
				}
			}
			nstlj = nst;
			nje++;
		} else {
			gsync();
			// M = I - gamma * (saved Jacobian), row by row; the saved Jacobian is in global memory: with CP_SJ_UNROLL > 1 several
			// rows' loads are in flight at once instead of one round trip per row
#pragma unroll SJ_UNROLL
			for (int r = 0; r < N; r++) {
#pragma unroll
				for (int e = 0; e < E; e++) {
					if (own(e)) {
						const int c = idx(e);
						double m = neg_gamma * SJ[r * N + c];
						if (r == c) m += 1.0;
						M[r * RS + c] = m;
					}
				}
			}
		}
		if constexpr (N <= 3) {
			// SUNLinSolSetup_Dense_Eigen for 2 and 3 states (sunlinsol_dense_eigen.cpp:111-145): an EXPLICIT inverse -- closed
			// form for 2 x 2, cofactors for 3 x 3 (Eigen's compute_inverse), no pivoting -- and x = inverse * b in the solve.
			// On a stiff I - gamma J that inverse is less accurate than a pivoted LU, and the Newton iterates inherit the
			// difference (measured: population averages 1e-7 away from the reference with LU, 1e-10 with this), so it is
			// reproduced: every lane computes the same inverse, lane 0 stores it over M. Products and differences are fused the
			// way the thread integrator's build fuses them (bdf_thread.cuh::linear_setup, which this follows).
			gsync();
			double A[N * N], Ainv[N * N];
#pragma unroll
			for (int i = 0; i < N; i++)
#pragma unroll
				for (int j = 0; j < N; j++) A[i * N + j] = (N == 1) ? 1.0 : M[i * RS + j];
			if constexpr (N == 1) {
				Ainv[0] = 1.0 / M[0];
			} else if constexpr (N == 2) {
				const double invdet = 1.0 / fma(A[0], A[3], -(A[1] * A[2]));
				Ainv[0] = A[3] * invdet;
				Ainv[1] = -A[1] * invdet;
				Ainv[2] = -A[2] * invdet;
				Ainv[3] = A[0] * invdet;
			} else {
#define CPG_A(i, j) A[(i) * N + (j)]
#define CPG_COF(i, j) \
	fma(CPG_A(((i) + 1) % 3, ((j) + 1) % 3), CPG_A(((i) + 2) % 3, ((j) + 2) % 3), -(CPG_A(((i) + 1) % 3, ((j) + 2) % 3) * CPG_A(((i) + 2) % 3, ((j) + 1) % 3)))
				const double c0 = CPG_COF(0, 0), c1 = CPG_COF(1, 0), c2 = CPG_COF(2, 0);
				const double det = fma(c2, CPG_A(2, 0), fma(c1, CPG_A(1, 0), c0 * CPG_A(0, 0)));
				const double invdet = 1.0 / det;
				Ainv[0 * N + 0] = c0 * invdet;
				Ainv[0 * N + 1] = c1 * invdet;
				Ainv[0 * N + 2] = c2 * invdet;
				Ainv[1 * N + 0] = CPG_COF(0, 1) * invdet;
				Ainv[1 * N + 1] = CPG_COF(1, 1) * invdet;
				Ainv[1 * N + 2] = CPG_COF(2, 1) * invdet;
				Ainv[2 * N + 0] = CPG_COF(0, 2) * invdet;
				Ainv[2 * N + 1] = CPG_COF(1, 2) * invdet;
				Ainv[2 * N + 2] = CPG_COF(2, 2) * invdet;
#undef CPG_COF
#undef CPG_A
			}
			gsync();
			if (lg == 0) {
#pragma unroll
				for (int i = 0; i < N; i++)
#pragma unroll
					for (int j = 0; j < N; j++) M[i * RS + j] = Ainv[i * N + j];
			}
			gsync();
			nsetups++;
			return;
		}
		if (lg == 0) {
#pragma unroll
			for (int i = 0; i < N; i++) perm[i] = i;
		}
		gsync();
		// PartialPivLUExtended::compute_optimized (EigenPartialPivLUSomewhatSparse.h:38-105): lane = rows lg, lg + G, ...
		if constexpr (N <= CP_GROUP_STATIC_LU_MAX) {
#if CP_SOLVE_SKIP
#pragma unroll
			for (int e = 0; e < E; e++) lmask[e] = umask[e] = 0xffffffffu;
#endif
			// fully unrolled: which slots still have rows below the pivot is known per (k, slot) at compile time, and every
			// shared-memory access is [lane row base + constant]
			double* const rowbase = M + lg * RS;
			static_for<0, N>([&](auto K) {
				constexpr int k = decltype(K)::value;
				double best = -1.0;
				int bi = N;
				static_for<0, E>([&](auto EE) {
					constexpr int e = decltype(EE)::value;
					if constexpr (G * e + G - 1 >= k) {
						const int i = lg + G * e;
						if ((G * e >= k || i >= k) && own(e)) {
							const double v = fabs(rowbase[G * e * RS + k]);
							if (v > best) {
								best = v;
								bi = i;
							}
						}
					}
				});
#pragma unroll
				for (int d = G / 2; d >= 1; d >>= 1) {
					const double ob = __shfl_xor_sync(gmask, best, d);
					const int oi = __shfl_xor_sync(gmask, bi, d);
					if (ob > best || (ob == best && oi < bi)) {
						best = ob;
						bi = oi;
					}
				}
				double lik[E];
#pragma unroll
				for (int e = 0; e < E; e++) lik[e] = 0.0;
				if (best > 0.0) { // group-uniform; a column of zeros (or of NaN: best stays -1) is left alone
					if (bi != k) {
#pragma unroll
						for (int e = 0; e < E; e++) {
							if (own(e)) {
								const int c = idx(e);
								const double t = M[k * RS + c];
								M[k * RS + c] = M[bi * RS + c];
								M[bi * RS + c] = t;
							}
						}
						if (lg == 0) {
							const int t = perm[k];
							perm[k] = perm[bi];
							perm[bi] = t;
						}
						gsync();
					}
					const double inv_coeff = 1.0 / M[k * RS + k];
					static_for<0, E>([&](auto EE) {
						constexpr int e = decltype(EE)::value;
						if constexpr (G * e + G - 1 > k) {
							const int i = lg + G * e;
							if ((G * e > k || i > k) && own(e)) {
								lik[e] = rowbase[G * e * RS + k] * inv_coeff;
								rowbase[G * e * RS + k] = lik[e];
							}
						}
					});
				} else {
					static_for<0, E>([&](auto EE) {
						constexpr int e = decltype(EE)::value;
						if constexpr (G * e + G - 1 > k) {
							const int i = lg + G * e;
							if ((G * e > k || i > k) && own(e)) lik[e] = rowbase[G * e * RS + k];
						}
					});
				}
				static_for<k + 1, N>([&](auto CC) {
					constexpr int c = decltype(CC)::value;
					const double a_kc = M[k * RS + c];
					static_for<0, E>([&](auto EE) {
						constexpr int e = decltype(EE)::value;
						if constexpr (G * e + G - 1 > k) {
							const int i = lg + G * e;
							if ((G * e > k || i > k) && own(e)) rowbase[G * e * RS + c] = fma(-a_kc, lik[e], rowbase[G * e * RS + c]);
						}
					});
				});
				gsync();
			});
		} else {
			double* const rowbase = M + lg * RS;
#if CP_SOLVE_SKIP
#pragma unroll
			for (int e = 0; e < E; e++) lmask[e] = umask[e] = 0u;
#endif
#pragma unroll 1
			for (int k = 0; k < N; k++) {
				double best = -1.0;
				int bi = N;
#pragma unroll
				for (int e = 0; e < E; e++) {
					const int i = lg + G * e;
					if (i >= k && own(e)) {
						const double v = fabs(rowbase[G * e * RS + k]);
						if (v > best) {
							best = v;
							bi = i;
						}
					}
				}
				// largest |a_ik| of the group, the lowest row among equals (Eigen's maxCoeff visitor keeps the first)
				bool have_pivot;
				if constexpr (CP_PIVOT_REDUX) {
					// three warp-reduce instructions instead of log2(G) shuffle stages: the bit pattern of a non-negative double
					// orders like the number, so the maximum is found on the high word, then on the low word among the lanes
					// that hold that high word, then the smallest row index among the lanes that hold both. A lane without a
					// candidate (best = -1) and a NaN column (best stays -1) carry key 0, like a column of zeros.
					const unsigned hi = (best > 0.0) ? (unsigned)__double2hiint(best) : 0u;
					const unsigned lo = (best > 0.0) ? (unsigned)__double2loint(best) : 0u;
					const unsigned mh = __reduce_max_sync(gmask, hi);
					const unsigned ml = __reduce_max_sync(gmask, (hi == mh) ? lo : 0u);
					const bool cand = (hi == mh) && (lo == ml);
					bi = (int)__reduce_min_sync(gmask, cand ? (unsigned)bi : (unsigned)N);
					have_pivot = (mh | ml) != 0u;
				} else {
#pragma unroll
					for (int d = G / 2; d >= 1; d >>= 1) {
						const double ob = __shfl_xor_sync(gmask, best, d);
						const int oi = __shfl_xor_sync(gmask, bi, d);
						if (ob > best || (ob == best && oi < bi)) {
							best = ob;
							bi = oi;
						}
					}
					have_pivot = best > 0.0;
				}
				double inv_coeff = 1.0;
				if (have_pivot) { // group-uniform; a column of zeros (or of NaN) is left alone
					if (bi != k) {
#pragma unroll
						for (int e = 0; e < E; e++) {
							if (own(e)) {
								const int c = idx(e);
								const double t = M[k * RS + c];
								M[k * RS + c] = M[bi * RS + c];
								M[bi * RS + c] = t;
							}
						}
						if (lg == 0) {
							const int t = perm[k];
							perm[k] = perm[bi];
							perm[bi] = t;
						}
						gsync();
					}
					inv_coeff = 1.0 / M[k * RS + k];
				}
				// multipliers of the rows this lane owns below the pivot
				double lik[E];
#pragma unroll
				for (int e = 0; e < E; e++) {
					const int i = lg + G * e;
					lik[e] = 0.0;
					if (i > k && own(e)) {
						lik[e] = rowbase[G * e * RS + k] * inv_coeff;
						rowbase[G * e * RS + k] = lik[e];
					}
				}
#if CP_SOLVE_SKIP
				{
					bool any = false;
#pragma unroll
					for (int e = 0; e < E; e++) any = any || (lik[e] != 0.0);
					if (__ballot_sync(gmask, any) & gmask) {
						const unsigned bit = 1u << (k % G);
						static_for<0, E>([&](auto S) {
							constexpr int s = decltype(S)::value;
							if (k / G == s) lmask[s] |= bit;
						});
					}
				}
#endif
#if CP_LU_SKIP_ZEROS
				// Column updates, skipping the columns whose pivot-row entry is zero exactly as the reference's LU does
				// (EigenPartialPivLUSomewhatSparse.h:88-93: `if (a_kj != 0.0)`): the lanes look at the pivot row together, one
				// column per lane and slot, a vote gives the set of non-zero columns, and only those are visited. Signalling
				// networks have a handful of non-zeros per row, so this turns the N^3 / 3 multiply-adds into ~N^2.
#pragma unroll
				for (int e = 0; e < E; e++) {
					const int c0 = lg + G * e;
					const double rk = (c0 > k && (!PADDED || c0 < N)) ? M[k * RS + c0] : 0.0;
					unsigned nz = (__ballot_sync(gmask, rk != 0.0) & gmask) >> gbase;
#if CP_SOLVE_SKIP
					umask[e] |= nz; // row k of U is final: its non-zero columns
#endif
					while (nz) {
						const int j = __ffs(nz) - 1;
						nz &= nz - 1;
						const double a_kc = __shfl_sync(gmask, rk, gbase + j);
						const int c = j + G * e;
#pragma unroll
						for (int e2 = 0; e2 < E; e2++) {
							if (lik[e2] != 0.0) rowbase[G * e2 * RS + c] = fma(-a_kc, lik[e2], rowbase[G * e2 * RS + c]);
						}
					}
				}
#else
#pragma unroll 1
				for (int c = k + 1; c < N; c++) {
					const double a_kc = M[k * RS + c];
#pragma unroll
					for (int e = 0; e < E; e++) {
						const int i = lg + G * e;
						if (i > k && own(e)) rowbase[G * e * RS + c] = fma(-a_kc, lik[e], rowbase[G * e * RS + c]);
					}
				}
#endif
				gsync();
			}
		}
		// the triangular solves multiply by the reciprocal pivots
#pragma unroll
		for (int e = 0; e < E; e++) {
			if (own(e)) {
				const int i = idx(e);
				M[i * RS + i] = 1.0 / M[i * RS + i];
			}
		}
		gsync();
		nsetups++;
	}

	// x <- (P L U)^-1 x, cooperative: the pivot component is broadcast, every lane updates the components it owns.
	// Real loops over the lanes inside compile-time slots: ~7 instructions per column (one shuffle pair, one load and one
	// multiply-add per slot) where a single loop over k with run-time slot selection needed ~18.
	__device__ __forceinline__ void lu_solve(double (&b)[E])
	{
		if constexpr (N <= 3) { // x = inverse * b, every row summed left to right (see linear_setup)
			publish(ybuf, b);
#pragma unroll
			for (int e = 0; e < E; e++) {
				if (own(e)) {
					const int i = idx(e);
					double sx = M[i * RS + 0] * ybuf[0];
#pragma unroll
					for (int j = 1; j < N; j++) sx = fma(M[i * RS + j], ybuf[j], sx);
					b[e] = sx;
				} else {
					b[e] = 0.0;
				}
			}
			return;
		}
		publish(ybuf, b);
#pragma unroll
		for (int e = 0; e < E; e++) b[e] = own(e) ? ybuf[perm[idx(e)]] : 0.0;
		const double* const rowbase = M + lg * RS;
#if CP_SOLVE_SKIP
		// As the slotted form below, visiting only the columns with an entry off the diagonal (see CP_SOLVE_SKIP).
		static_for<0, E>([&](auto S) {
			constexpr int s = decltype(S)::value;
			constexpr int KK = (G * (s + 1) <= N) ? G : (N - G * s);
			unsigned m = lmask[s] & (KK >= 32 ? 0xffffffffu : ((1u << (KK & 31)) - 1u));
#pragma unroll 1
			while (m) {
				const int kk = __ffs(m) - 1;
				m &= m - 1;
				const int k = G * s + kk;
				const double xk = __shfl_sync(gmask, b[s], gbase + kk);
				if (lg > kk && own(s)) b[s] = fma(-xk, rowbase[G * s * RS + k], b[s]);
				static_for<s + 1, E>([&](auto EE) {
					constexpr int e2 = decltype(EE)::value;
					if (own(e2)) b[e2] = fma(-xk, rowbase[G * e2 * RS + k], b[e2]);
				});
			}
		});
		static_rfor<0, E>([&](auto S) {
			constexpr int s = decltype(S)::value;
			constexpr int KK = (G * (s + 1) <= N) ? G : (N - G * s);
			const unsigned cols = umask[s] & (KK >= 32 ? 0xffffffffu : ((1u << (KK & 31)) - 1u));
			unsigned m = cols;
#pragma unroll 1
			while (m) {
				const int kk = 31 - __clz(m);
				m &= ~(1u << kk);
				const int k = G * s + kk;
				const double m_k = own(s) ? rowbase[G * s * RS + k] : 0.0;
				const double xk = __shfl_sync(gmask, b[s] * m_k, gbase + kk);
				if (lg < kk) b[s] = fma(-xk, m_k, b[s]);
				else if (lg == kk) b[s] = xk;
				static_for<0, s>([&](auto EE) {
					constexpr int e2 = decltype(EE)::value;
					b[e2] = fma(-xk, rowbase[G * e2 * RS + k], b[e2]);
				});
			}
			// a column without entries above the diagonal: nobody else needs x_k, the owner scales its component
			if (own(s) && !((cols >> lg) & 1u)) b[s] = b[s] * rowbase[G * s * RS + G * s + lg];
		});
#elif CP_SOLVE_SLOTTED
		// Forward substitution with the unit lower factor, column by column: slot by slot (compile time) and lane by lane
		// inside a slot, so that the pivot component is a static register read by a shuffle and every lane touches its
		// rows through [lane row base + constant]. Rows of the pivot's own slot take part only on the lanes behind it.
		static_for<0, E>([&](auto S) {
			constexpr int s = decltype(S)::value;
			constexpr int KK = (G * (s + 1) <= N) ? G : (N - G * s);
#pragma unroll 4
			for (int kk = 0; kk < KK; kk++) {
				const int k = G * s + kk;
				const double xk = __shfl_sync(gmask, b[s], gbase + kk);
				if (lg > kk && own(s)) b[s] = fma(-xk, rowbase[G * s * RS + k], b[s]);
				static_for<s + 1, E>([&](auto EE) {
					constexpr int e2 = decltype(EE)::value;
					if (own(e2)) b[e2] = fma(-xk, rowbase[G * e2 * RS + k], b[e2]);
				});
			}
		});
		// Back substitution with the upper factor (reciprocal pivots on the diagonal): the lane that owns row k reads its
		// diagonal entry with the same load the lanes above it use for their entry of column k.
		static_rfor<0, E>([&](auto S) {
			constexpr int s = decltype(S)::value;
			constexpr int KK = (G * (s + 1) <= N) ? G : (N - G * s);
#pragma unroll 4
			for (int kk = KK - 1; kk >= 0; kk--) {
				const int k = G * s + kk;
				const double m_k = own(s) ? rowbase[G * s * RS + k] : 0.0;
				const double xk = __shfl_sync(gmask, b[s] * m_k, gbase + kk);
				if (lg < kk) b[s] = fma(-xk, m_k, b[s]);
				else if (lg == kk) b[s] = xk;
				static_for<0, s>([&](auto EE) {
					constexpr int e2 = decltype(EE)::value;
					b[e2] = fma(-xk, rowbase[G * e2 * RS + k], b[e2]);
				});
			}
		});
#else
#pragma unroll 1
		for (int k = 0; k < N; k++) {
			const int slot = k / G;
			double bk = b[0];
			static_for<1, E>([&](auto J) {
				constexpr int j = decltype(J)::value;
				if (slot >= j) bk = b[j];
			});
			const double xk = __shfl_sync(gmask, bk, gbase + (k % G));
#pragma unroll
			for (int e = 0; e < E; e++) {
				const int i = lg + G * e;
				if (i > k && own(e)) b[e] = fma(-xk, rowbase[G * e * RS + k], b[e]);
			}
		}
#pragma unroll 1
		for (int k = N - 1; k >= 0; k--) {
			const int slot = k / G;
			double bk = b[0];
			static_for<1, E>([&](auto J) {
				constexpr int j = decltype(J)::value;
				if (slot >= j) bk = b[j];
			});
			const double xk = __shfl_sync(gmask, bk * M[k * RS + k], gbase + (k % G));
#pragma unroll
			for (int e = 0; e < E; e++) {
				const int i = lg + G * e;
				if (i < k) b[e] = fma(-xk, rowbase[G * e * RS + k], b[e]);
				else if (i == k) b[e] = xk;
			}
		}
#endif
	}

	// cvNlsResidual at y = zn[0] + acor: y published to ybuf, f(y) in fy, delta = rl1 zn[1] + acor - gamma f
	__device__ __forceinline__ void residual(double (&y)[E], double (&fy)[E], double (&delta)[E])
	{
#pragma unroll
		for (int e = 0; e < E; e++) y[e] = Z<0>(e) + acor[e];
		publish(ybuf, y);
		rhs_shared(tn, fy);
#pragma unroll
		for (int e = 0; e < E; e++) {
			double r = fma(rl1, Z<1>(e), acor[e]);
			r = fma(-gamma, fy[e], r);
			delta[e] = r;
		}
	}

	// One pass of cvStep's attempt loop (structure of BdfThread::attempt); `mask` = lanes of the warp in this call.
	__device__ __forceinline__ int attempt(unsigned mask)
	{
		double l[6], tq[6];
#pragma unroll
		for (int i = 0; i < 6; i++) {
			l[i] = 0.0;
			tq[i] = 0.0;
		}
		// ---- cvPredict ----
		tn += h;
		if (tstopset) {
			if ((tn - tstop()) * h > 0.0) tn = tstop();
		}
		static_for<1, QMAX + 1>([&](auto K) {
			constexpr int k = decltype(K)::value;
			static_rfor<k, QMAX + 1>([&](auto J) {
				constexpr int j = decltype(J)::value;
				if (j <= q) {
#pragma unroll
					for (int e = 0; e < E; e++) Z<j - 1>(e) += Z<j>(e);
				}
			});
		});
		// ---- cvSetBDF + cvSetTqBDF (cvode.c:2611-2686) ----
		{
			double xi_inv = 1.0, xistar_inv = 1.0, alpha0 = -1.0, alpha0_hat = -1.0;
			double hsum = h;
			l[0] = 1.0;
			l[1] = 1.0;
			if (q > 1) {
				static_for<2, QMAX>([&](auto J) {
					constexpr int j = decltype(J)::value;
					if (j < q) {
						hsum += tau(j - 1);
						xi_inv = h / hsum;
						alpha0 -= 1.0 / j;
						static_rfor<1, j + 1>([&](auto I) {
							constexpr int i = decltype(I)::value;
							l[i] += l[i - 1] * xi_inv;
						});
					}
				});
				alpha0 -= 1.0 / q;
				xistar_inv = -l[1] - alpha0;
				hsum += tau(q - 1);
				xi_inv = h / hsum;
				alpha0_hat = -l[1] - xi_inv;
				static_rfor<1, QMAX + 1>([&](auto I) {
					constexpr int i = decltype(I)::value;
					if (i <= q) l[i] += l[i - 1] * xistar_inv;
				});
			}
			double lq = l[1];
			static_for<2, QMAX + 1>([&](auto J) {
				constexpr int j = decltype(J)::value;
				if (j <= q) lq = l[j];
			});
			const double tauq = tau(q);
			const double A1 = 1.0 - alpha0_hat + alpha0;
			const double A2 = 1.0 + q * A1;
			tq[2] = fabs(A1 / (alpha0 * A2));
			tq[5] = fabs(A2 * xistar_inv / (lq * xi_inv));
			if (qwait == 1) {
				if (q > 1) {
					const double C = xistar_inv / lq;
					const double A3 = alpha0 + 1.0 / q;
					const double A4 = alpha0_hat + xi_inv;
					const double Cpinv = (1.0 - A4 + A3) / A3;
					tq[1] = fabs(C * Cpinv);
				} else {
					tq[1] = 1.0;
				}
				hsum += tauq;
				xi_inv = h / hsum;
				const double A5 = alpha0 - (1.0 / (q + 1));
				const double A6 = alpha0_hat - xi_inv;
				const double Cppinv = (1.0 - A6 + A5) / A2;
				tq[3] = fabs(Cppinv / (xi_inv * (q + 2) * A5));
			}
			tq[4] = BDF_CORTES / tq[2];
			rl1 = 1.0 / l[1];
			gamma = h * rl1;
			if (nst == 0) gammap() = gamma;
			gamrat = (nst > 0) ? gamma / gammap() : 1.0;
		}

		// l[] and tq[1, 2, 3, 5] are not needed until the step is accepted: parked in shared memory over the Newton loop
#pragma unroll
		for (int i = 0; i < 6; i++) {
			sc[SC_L + i] = l[i];
			sc[SC_TQ + i] = tq[i];
		}
		// ---- cvNls + Newton: both loops warp-uniform via votes ----
		int nls_ret = 1;
		{
			const double tol = tq[4];
			int convfail = ((nflag == bcm3b200::BDF_FIRST_CALL) || (nflag == bcm3b200::BDF_PREV_ERR_FAIL)) ? bcm3b200::BDF_NO_FAILURES : bcm3b200::BDF_FAIL_OTHER;
			bool callSetup = (nflag == bcm3b200::BDF_PREV_CONV_FAIL) || (nflag == bcm3b200::BDF_PREV_ERR_FAIL) || (nst == 0) ||
			                 (nst >= nstlp + BDF_MSBP) || (fabs(gamrat - 1.0) > BDF_DGMAX);
#pragma unroll
			for (int e = 0; e < E; e++) acor[e] = 0.0;
			// One trip = one Newton iteration of SUNNonlinSolSolve_Newton (sunnonlinsol_newton.c:183-318): residual, the linear
			// setup if this is the first iteration of a pass that asked for one, solve, convergence test. A failed pass with
			// stale Jacobian data restarts at m = 0 with a forced setup (:301-312). One copy of the residual code serves all.
			bool jbad = false, active = true;
			int m = 0;
#pragma unroll 1
			for (;;) {
				if (!__any_sync(mask, active)) break;
				double y[E], fy[E], delta[E];
#pragma unroll
				for (int e = 0; e < E; e++) y[e] = fy[e] = delta[e] = 0.0;
				if (active) residual(y, fy, delta);
				const bool do_setup = active && (m == 0) && callSetup;
				if (__any_sync(mask, do_setup)) {
					if (do_setup) {
						if (jbad) convfail = bcm3b200::BDF_FAIL_BAD_J;
						const double dgamma = fabs((gamma / gammap()) - 1.0);
						const bool jb = (nst == 0) || (nst > nstlj + BDF_MSBJ) || ((convfail == bcm3b200::BDF_FAIL_BAD_J) && (dgamma < BDF_LS_DGMAX)) ||
						                (convfail == bcm3b200::BDF_FAIL_OTHER);
						linear_setup(jb, y, fy);
						nls_jcur = jb;
						gamrat = 1.0;
						gammap() = gamma;
						crate() = 1.0;
						nstlp = nst;
						callSetup = false;
					}
				}
				if (active) {
#pragma unroll
					for (int e = 0; e < E; e++) delta[e] = -delta[e];
					lu_solve(delta);
					if (gamrat != 1.0) {
						const double sc = 2.0 / (1.0 + gamrat);
#pragma unroll
						for (int e = 0; e < E; e++) delta[e] *= sc;
					}
#pragma unroll
					for (int e = 0; e < E; e++) acor[e] += delta[e];
					const double del = wrms(delta);
					if (m > 0) crate() = fmax(BDF_CRDOWN * crate(), del / delp());
					const double dcon = del * fmin(1.0, crate()) / tol;
					if (dcon <= 1.0) {
						acnrm() = (m == 0) ? del : wrms(acor);
						nls_jcur = false;
						nls_ret = 0;
						active = false;
					} else if (((m >= 1) && (del > BDF_RDIV * delp())) || (m + 1 >= BDF_NLS_MAXCOR)) {
						// this pass failed
						if (nls_jcur) {
							active = false;
						} else {
							callSetup = true;
							jbad = true;
							m = 0;
#pragma unroll
							for (int e = 0; e < E; e++) acor[e] = 0.0;
						}
					} else {
						delp() = del;
						m++;
					}
				}
			}
		}

#pragma unroll
		for (int i = 0; i < 6; i++) {
			l[i] = sc[SC_L + i];
			tq[i] = sc[SC_TQ + i];
		}
		int result;
		// Both failure paths (cvHandleNFlag, cvDoErrorTest) share ONE restore() and ONE rescale() site: per warp a failure
		// happens in a large share of the trips, so this is hot code, and the kernel is bound by its instruction footprint.
		const double dsm = acnrm() * tq[2];
		const bool conv_fail = (nls_ret != 0);
		const bool err_fail = !conv_fail && !(dsm <= 1.0);
		if (conv_fail || err_fail) {
			bool do_rescale = false;
			if (conv_fail) {
				ncf++;
			} else {
				nef++;
				nflag = bcm3b200::BDF_PREV_ERR_FAIL;
			}
			restore();
			if (conv_fail) {
				// ---- cvHandleNFlag ----
				etamax() = 1.0;
				if ((fabs(h) <= hmin * BDF_ONEPSM) || (ncf == BDF_MXNCF)) {
					result = T_FAILED;
				} else {
					eta() = fmax(BDF_ETACF, hmin / fabs(h));
					nflag = bcm3b200::BDF_PREV_CONV_FAIL;
					do_rescale = true;
					result = T_RETRY;
				}
			} else if ((fabs(h) <= hmin * BDF_ONEPSM) || (nef == BDF_MXNEF)) {
				// ---- cvDoErrorTest, failure ----
				result = T_FAILED;
			} else {
				result = T_RETRY;
				etamax() = 1.0;
				if (nef <= BDF_MXNEF1) {
					eta() = 1.0 / (step_root(BDF_BIAS2 * dsm, L) + BDF_ADDON);
					eta() = fmax(BDF_ETAMIN, fmax(eta(), hmin / fabs(h)));
					if (nef >= BDF_SMALL_NEF) eta() = fmin(eta(), BDF_ETAMXF);
					do_rescale = true;
				} else if (q > 1) {
					eta() = fmax(BDF_ETAMIN, hmin / fabs(h));
					adjust_order(-1);
					L = q;
					q--;
					qwait = L;
					do_rescale = true;
				} else {
					eta() = fmax(BDF_ETAMIN, hmin / fabs(h));
					h *= eta();
					hscale() = h;
					qwait = BDF_LONG_WAIT;
					double f[E];
					publish(ybuf, zn01[0]);
					rhs_shared(tn, f);
#pragma unroll
					for (int e = 0; e < E; e++) Z<1>(e) = h * f[e];
				}
			}
			if (do_rescale) rescale();
		} else {
			{
				result = T_DONE;
				// ---- cvCompleteStep ----
				nst++;
				hu() = h;
				for (int i = q; i >= 2; i--) tau(i) = tau(i - 1);
				if ((q == 1) && (nst > 1)) tau(2) = tau(1);
				tau(1) = h;
				static_for<0, QMAX + 1>([&](auto J) {
					constexpr int j = decltype(J)::value;
					if (j <= q) {
#pragma unroll
						for (int e = 0; e < E; e++) Z<j>(e) = fma(l[j], acor[e], Z<j>(e));
					}
				});
				qwait--;
				if ((qwait == 1) && (q != QMAX)) {
#pragma unroll
					for (int e = 0; e < E; e++) Z<QMAX>(e) = acor[e];
					saved_tq5() = tq[5];
				}
				// ---- cvPrepareNextStep ----
				const double etamax_now = etamax();
				if (etamax_now == 1.0) {
					qwait = (qwait > 2) ? qwait : 2;
					qprime = q;
					hprime() = h;
					eta() = 1.0;
				} else {
					const double etaq = 1.0 / (step_root(BDF_BIAS2 * dsm, L) + BDF_ADDON);
					eta() = etaq;
					qprime = q;
					if (qwait == 0) {
						qwait = 2;
						double etaqm1 = 0.0;
						if (q > 1) {
							double znq[E];
#pragma unroll
							for (int e = 0; e < E; e++) znq[e] = Z<2>(e);
							static_for<3, QMAX + 1>([&](auto J) {
								constexpr int j = decltype(J)::value;
								if (j <= q) {
#pragma unroll
									for (int e = 0; e < E; e++) znq[e] = Z<j>(e);
								}
							});
							const double ddn = wrms(znq) * tq[1];
							etaqm1 = 1.0 / (step_root(BDF_BIAS1 * ddn, q) + BDF_ADDON);
						}
						double etaqp1 = 0.0;
						if (q != QMAX) {
							const double stq5 = saved_tq5();
							if (stq5 != 0.0) {
								const double base = h / tau(2);
								double pw = 1.0;
								static_for<1, QMAX + 1>([&](auto I) {
									if (decltype(I)::value <= L) pw *= base;
								});
								const double cquot = (tq[5] / stq5) * pw;
								double tmp[E];
#pragma unroll
								for (int e = 0; e < E; e++) tmp[e] = fma(-cquot, Z<QMAX>(e), acor[e]);
								const double dup = wrms(tmp) * tq[3];
								etaqp1 = 1.0 / (step_root(BDF_BIAS3 * dup, L + 1) + BDF_ADDON);
							}
						}
						const double etam = fmax(etaqm1, fmax(etaq, etaqp1));
						if (etam < BDF_THRESH) {
							eta() = 1.0;
							qprime = q;
						} else if (etam == etaq) {
							eta() = etaq;
							qprime = q;
						} else if (etam == etaqm1) {
							eta() = etaqm1;
							qprime = q - 1;
						} else {
							eta() = etaqp1;
							qprime = q + 1;
#pragma unroll
							for (int e = 0; e < E; e++) Z<QMAX>(e) = acor[e];
						}
					}
					if (eta() < BDF_THRESH) {
						eta() = 1.0;
						hprime() = h;
					} else {
						eta() = fmin(eta(), etamax_now);
						if (hmax_inv > 0.0) eta() /= fmax(1.0, fabs(h) * hmax_inv * eta()); // cvSetEta, cvode.c:3142-3143 (dividing by 1 changes nothing)
						hprime() = h * eta();
					}
				}
				etamax() = BDF_ETAMX3;
#pragma unroll
				for (int e = 0; e < E; e++) acor[e] *= tq[2];
			}
		}
		return result;
	}

	// CVodeGetDky(t, 0), cvode.c:1467-1524
	__device__ __forceinline__ bool dky_ok(double t) const
	{
		const double hu_ = hu();
		double tfuzz = BDF_FUZZ_FACTOR * UROUND * (fabs(tn) + fabs(hu_));
		if (hu_ < 0.0) tfuzz = -tfuzz;
		const double tp = tn - hu_ - tfuzz, tn1 = tn + tfuzz;
		return !((t - tp) * (t - tn1) > 0.0);
	}
	// CVodeGetDky(t, 0) for the components this lane owns
	__device__ __forceinline__ void dky_own(double t, double (&out)[E])
	{
		const double s = (t - tn) / h;
#pragma unroll
		for (int e = 0; e < E; e++) out[e] = 0.0;
		bool first = true;
		static_rfor<0, QMAX + 1>([&](auto J) {
			constexpr int j = decltype(J)::value;
			if (j <= q) {
				double c = 1.0;
#pragma unroll
				for (int i = 0; i < j; i++) c *= s;
				if (first) {
#pragma unroll
					for (int e = 0; e < E; e++) out[e] = c * Z<j>(e);
					first = false;
				} else {
#pragma unroll
					for (int e = 0; e < E; e++) out[e] = fma(c, Z<j>(e), out[e]);
				}
			}
		});
	}
	// tstop handling after an accepted step (cvode.c:1410-1438). True for CV_TSTOP_RETURN: yout = Dky(tstop), tret = tstop;
	// otherwise yout = zn[0], tret = tn.
	__device__ __forceinline__ bool after_step(double (&yout)[E], double& tret)
	{
		if (tstopset) {
			const double troundoff = BDF_FUZZ_FACTOR * UROUND * (fabs(tn) + fabs(h));
			if (fabs(tn - tstop()) <= troundoff) {
				dky_own(tstop(), yout);
				tret = tstop();
				tstopset = false;
				return true;
			}
			if ((tn + hprime() - tstop()) * h > 0.0) {
				hprime() = (tstop() - tn) * (1.0 - 4.0 * UROUND);
				eta() = hprime() / h;
			}
		}
		tret = tn;
#pragma unroll
		for (int e = 0; e < E; e++) yout[e] = Z<0>(e);
		return false;
	}
	// sum over the group of weight[e] * Dky component (weights = multiplicity of the component in the observed list)
	__device__ __forceinline__ double dky_weighted(double t, const double (&weight)[E])
	{
		const double s = (t - tn) / h;
		double acc[E];
#pragma unroll
		for (int e = 0; e < E; e++) acc[e] = 0.0;
		bool first = true;
		static_rfor<0, QMAX + 1>([&](auto J) {
			constexpr int j = decltype(J)::value;
			if (j <= q) {
				double c = 1.0;
#pragma unroll
				for (int i = 0; i < j; i++) c *= s;
				if (first) {
#pragma unroll
					for (int e = 0; e < E; e++) acc[e] = c * Z<j>(e);
					first = false;
				} else {
#pragma unroll
					for (int e = 0; e < E; e++) acc[e] = fma(c, Z<j>(e), acc[e]);
				}
			}
		});
		double sv = 0.0;
#pragma unroll
		for (int e = 0; e < E; e++) sv += weight[e] * acc[e];
		return gsum(sv);
	}
};

__device__ __forceinline__ void apply_variability(double& x, double value, int apply)
{
	switch (apply) {
	case CP_APPLY_ADDITIVE: x += value; break;
	case CP_APPLY_ADDITIVE_LOG: x += exp(value); break;
	case CP_APPLY_ADDITIVE_LOG2: x += pow(2.0, value); break;
	case CP_APPLY_MULTIPLICATIVE: x *= value; break;
	case CP_APPLY_MULTIPLICATIVE_LOG: x *= exp(value); break;
	case CP_APPLY_MULTIPLICATIVE_LOG2: x *= pow(2.0, value); break;
	case CP_APPLY_REPLACE: x = value; break;
	default: break;
	}
}

// Persistent grid: every group of G lanes takes (chain, cell) items from the queue until it is empty.
__global__ void __launch_bounds__(BS, CP_GROUP_MIN_BLOCKS) cellpop_group_kernel(const CpArgs a, double* __restrict__ saved_jacobians, unsigned long long* __restrict__ queue)
{
	extern __shared__ double smem_d[];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int gw = lane / G; // group within the warp
#if CP_RHS_LANES && CP_TABLES_SHARED
	{ // the tables of the lane-parallel right-hand side: one copy per block, behind the cells' blocks
		double* td = smem_d + CP_TAB_BASE;
		for (int i = tid; i < CP_TAB_NLIT; i += BS) td[i] = cp_rl_lit[i];
		for (int i = tid; i < CP_TAB_NCOEF; i += BS) td[CP_TAB_NLIT + i] = cp_out_coef[i];
		int* ti = reinterpret_cast<int*>(td + CP_TAB_NLIT + CP_TAB_NCOEF);
		for (int i = tid; i < CP_TAB_NIDX; i += BS) ti[i] = cp_rl_idx[i];
		for (int i = tid; i < CP_TAB_NTARGET; i += BS) ti[CP_TAB_NIDX + i] = cp_rl_target[i];
		for (int i = tid; i < CP_TAB_NBEGIN; i += BS) ti[CP_TAB_NIDX + CP_TAB_NTARGET + i] = cp_out_begin[i];
		for (int i = tid; i < CP_TAB_NLAW; i += BS) ti[CP_TAB_NIDX + CP_TAB_NTARGET + CP_TAB_NBEGIN + i] = cp_out_law[i];
		__syncthreads();
	}
#endif

	GroupBdf B;
	B.lg = lane % G;
	B.gbase = gw * G;
	B.gmask = (G == 32) ? FULL : (((1u << G) - 1u) << B.gbase);
	double* region = smem_d + (size_t)(warp * CPW + gw) * CS;
	const long long group_id = (long long)blockIdx.x * (WPB * CPW) + warp * CPW + gw;
#if CP_M_GLOBAL
	B.M = saved_jacobians + (long long)gridDim.x * (WPB * CPW) * (N * N) + group_id * (N * RS);
#else
	B.M = region;
#endif
	B.ybuf = region + OFF_Y;
	B.fbuf = region + OFF_F;
	B.perm = reinterpret_cast<int*>(region + OFF_PERM);
	B.sc = region + OFF_SCAL;
	B.znh = region + OFF_ZNH + B.lg;
	B.region_off = (unsigned)((warp * CPW + gw) * CS);
	B.SJ = saved_jacobians + group_id * (N * N);
	B.constant_species = a.constant_species;
	B.non_sampled = a.non_sampled;
	B.treat_ix = a.treatment_species;
	B.treat_n = a.treatment_num_pulses;
	B.treat_times = a.treatment_times;
	B.reltol = a.rel_tol;
	B.abstol = a.abs_tol;
	B.hmin = a.min_dt;
	B.hmax_inv = a.max_dt_inv;

	const int T = a.T;
#if CP_DIVISION
	const long long total = a.items ? (long long)a.num_items : (long long)a.num_chains * a.num_cells;
#else
	const long long total = (long long)a.num_chains * a.num_cells;
#endif
	const long long stride = a.cell_stride; // cell columns of the per-cell outputs
	const double nan = __longlong_as_double(0x7ff8000000000000ll);

	// per-lane multiplicity of owned components in the observed-species list, 4 bits per slot
	unsigned obs_count = 0;
#pragma unroll
	for (int e = 0; e < E; e++) {
		int cnt = 0;
		for (int k = 0; k < a.num_obs_species; k++) cnt += (a.obs_species[k] == B.idx(e)) ? 1 : 0;
		obs_count |= (unsigned)cnt << (4 * e);
	}
	static_assert(E <= 8, "observed-species multiplicities are packed 4 bits per slot");
	constexpr int ND = CP_NUM_DATASETS;
	static_assert(ND >= 1 && ND <= 4, "up to four data sets share one integration");
	unsigned obs_words[ND];
	obs_words[0] = obs_count;
#pragma unroll
	for (int d = 1; d < ND; d++) {
		unsigned w = 0;
#pragma unroll
		for (int e = 0; e < E; e++) {
			int cnt = 0;
			for (int k = 0; k < a.num_obs_species_more[d - 1]; k++) cnt += (a.obs_species_more[d - 1][k] == B.idx(e)) ? 1 : 0;
			w |= (unsigned)cnt << (4 * e);
		}
		obs_words[d] = w;
	}
	// every union timepoint from `from` on: the cell has no value there (lane 0 of the group writes)
	auto fill_nan = [&](double* out, int from) {
		for (int u = from; u < a.T; u++) {
			if constexpr (ND == 1) {
				out[(long long)u * stride] = nan;
			} else {
#pragma unroll
				for (int d = 0; d < ND; d++) {
					const int r = a.tp_rows[u * ND + d];
					if (r >= 0) out[(long long)r * stride] = nan;
				}
			}
		}
	};

	bool have = false, exhausted = false, ok = true, newstep = true;
	int steps = 0, tpi = 0, c = 0, cell = 0;
#if CP_DIVISION
	bool mitotic = false; // Cell::EnteredMitosis of the cell being integrated
#endif
	unsigned trip = 0;
	double* out = nullptr;

#pragma unroll 1
	for (;;) {
#if CP_GROUP_BATCHED
		// A warp takes its next CPW items together, when all of its groups are idle. The items of a chain are handed out in
		// `cell_order` (cells sorted along a space-filling curve through their quasi-random variability vectors), so the
		// cells that share a warp have nearly the same parameters, start in the same trip and take nearly the same
		// step-size, order, Newton and setup decisions in the same trips: the lanes of the warp diverge far less.
		unsigned long long w = ~0ull;
		bool fetched = false;
		if (!exhausted && !__any_sync(FULL, have)) {
			unsigned long long base = 0;
			if (lane == 0) base = atomicAdd(queue, (unsigned long long)CPW);
			base = __shfl_sync(FULL, base, 0);
			w = base + (unsigned long long)gw;
			fetched = true;
			if ((long long)(base + CPW) >= total) exhausted = true; // the queue only grows: nothing left after this batch
		}
		if (fetched && (long long)w < total) {
			{
#else
		if (!have && !exhausted) {
			// ---- next item + K0: Cell::Initialize (Cell.cpp:150-191) ----
			unsigned long long w = 0;
			if (B.lg == 0) w = atomicAdd(queue, 1ull);
			w = __shfl_sync(B.gmask, w, B.gbase);
			if ((long long)w >= total) {
				exhausted = true;
			} else {
#endif
#if CP_DIVISION
				int parent = -1;
				if (a.items) { // one generation of a dividing population: (chain, slot) pairs
					c = a.items[2 * w];
					cell = a.items[2 * w + 1];
					parent = a.cell_parent[(long long)c * stride + cell];
				} else
#endif
				{
					c = (int)(w / (unsigned long long)a.num_cells);
					cell = (int)(w % (unsigned long long)a.num_cells);
					if (a.cell_order) cell = a.cell_order[cell];
				}
				const double* tv = a.transformed + (long long)c * a.nvar;
				B.tv = tv;
				// per-cell parameter overrides: start from the chain's values (CP_OVERRIDE_INIT fills S.params.ov[]), apply the
				// cell's variability, park them in the cell's shared block for CellParameters
				struct {
					struct {
						double ov[CP_NUM_OVERRIDES > 0 ? CP_NUM_OVERRIDES : 1];
					} params;
				} S;
				S.params.ov[0] = 0.0;
				CP_OVERRIDE_INIT
				double* const ovl = S.params.ov;
				double y0[E];
#pragma unroll
				for (int e = 0; e < E; e++) y0[e] = B.own(e) ? a.initial_conditions[B.idx(e)] : 0.0;
				long long gcell = (long long)a.cell_offset + cell;
#if CP_DIVISION
				bool initial_flag = a.initial_flag != 0;
				if (a.items) {
					gcell = a.cell_row[(long long)c * stride + cell];
					if (parent >= 0) {
						// Cell::SetInitialConditionsFromOtherCell (Cell.cpp:119-148): the parent's state at its division, seven species reset
						initial_flag = false;
						const double* py = a.cell_end_y + ((long long)c * stride + parent) * N;
#pragma unroll
						for (int e = 0; e < E; e++) {
							if (B.own(e)) {
								double v = py[B.idx(e)];
#pragma unroll
								for (int k = 0; k < 7; k++)
									if (B.idx(e) == a.reset_ix[k]) v = (k >= 1 && k <= 3) ? 1.0 : 0.0;
								y0[e] = v;
							}
						}
					}
				}
#endif
				for (int d = 0; d < a.D; d++) {
#if CP_DIVISION
					if (a.var_only_initial[d] && !initial_flag) continue; // VariabilityDescriptionVariable.cpp:66-110
#endif
					double v = cellpop_variability_value(a, tv, c, gcell, d);
					if (a.var_negate[d]) v = -v;
					if (a.var_is_ic[d]) {
#pragma unroll
						for (int e = 0; e < E; e++)
							if (B.idx(e) == a.var_slot[d]) apply_variability(y0[e], v, a.var_apply[d]);
					} else {
						if (CP_NUM_OVERRIDES > 0 && a.var_slot[d] >= 0) apply_variability(ovl[a.var_slot[d]], v, a.var_apply[d]);
					}
				}
				B.gsync();
#pragma unroll
				for (int s = 0; s < (CP_NUM_OVERRIDES > 0 ? CP_NUM_OVERRIDES : 1); s++) B.sc[SC_OV + s] = ovl[s];
				// ---- Cell::Simulate + ODESolver::SolveReturnSolution + ODESolverCVODE::Solve: the part before the first step ----
				double creation_time = (a.entry_time_ix >= 0) ? tv[a.entry_time_ix] : a.entry_time_fixed;
#if CP_DIVISION
				if (a.items) creation_time = a.cell_creation[(long long)c * stride + cell];
#endif
				B.sc[SC_CREATION] = creation_time;
				out = a.cell_values + ((long long)c * a.num_rows) * stride + cell;
				ok = true;
				steps = 0;
#if CP_DIVISION
				mitotic = false;
#endif
				tpi = 0;
				newstep = true;
				bool finished = false;
				double sv0[ND];
#pragma unroll
				for (int d = 0; d < ND; d++) {
					double v = 0.0;
#pragma unroll
					for (int e = 0; e < E; e++) v += (double)((obs_words[d] >> (4 * e)) & 15u) * y0[e];
					sv0[d] = B.gsum(v);
				}
				while (tpi < T && (a.timepoints[tpi] - creation_time) < DBL_EPSILON) {
					const double cell_time = a.timepoints[tpi] - creation_time;
					if (B.lg == 0) {
						if constexpr (ND == 1) {
							out[(long long)tpi * stride] = (cell_time < 0.0) ? nan : sv0[0];
						} else {
#pragma unroll
							for (int d = 0; d < ND; d++) {
								const int r = a.tp_rows[tpi * ND + d];
								if (r >= 0) out[(long long)r * stride] = (cell_time < 0.0) ? nan : sv0[d];
							}
						}
					}
					tpi++;
				}
				const double end_time = a.sim_end_time - creation_time;
				if (end_time < DBL_EPSILON) finished = true; // nothing to integrate: every requested time is at or before the cell's creation
				B.sc[SC_END] = end_time;
				if (!finished) {
					// Cell::Simulate (Cell.cpp:212-229): SetDiscontinuity(first discontinuity ahead of the cell) = CVodeSetStopTime
					B.create();
					const double first_disc = B.first_discontinuity();
					B.sc[SC_NEXT_DISC] = first_disc;
					if (first_disc == first_disc) {
						B.tstop() = first_disc;
						B.tstopset = true;
					}
					if (!B.restart(0.0, y0, end_time)) {
						ok = false;
						finished = true;
					}
				}
				if (finished) {
					if (B.lg == 0) {
						if (!ok) fill_nan(out, tpi);
						a.cell_status[(long long)c * stride + cell] = ok ? 1 : 0;
						if (a.cell_steps) a.cell_steps[(long long)c * stride + cell] = 0;
#if CP_DIVISION
						if (a.items) {
							a.cell_event[(long long)c * stride + cell] = 0;
							a.cell_end_time[(long long)c * stride + cell] = a.sim_end_time;
						}
#endif
					}
				} else {
					have = true;
				}
			}
		}
#if CP_GROUP_LOCKSTEP
		// block lock-step: the warps of the block meet before every CP_GROUP_LOCKSTEP_EVERY-th trip (and leave the loop only
		// there, together), so that they run the same code at nearly the same time and share instruction fetches
		if (CP_GROUP_LOCKSTEP_EVERY == 1 || (trip++ % CP_GROUP_LOCKSTEP_EVERY) == 0) {
#if CP_GROUP_LOCKSTEP_TEAM > 0 && (CP_GROUP_WARPS % CP_GROUP_LOCKSTEP_TEAM) == 0 && CP_GROUP_LOCKSTEP_TEAM < CP_GROUP_WARPS
			{
				int any;
				const int team = 1 + (tid / 32) / CP_GROUP_LOCKSTEP_TEAM, pred = (have || !exhausted) ? 1 : 0;
				asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 q, %3, 0;\n\tbar.red.or.pred p, %1, %2, q;\n\tselp.s32 %0, 1, 0, p;\n\t}"
				             : "=r"(any)
				             : "r"(team), "r"(32 * CP_GROUP_LOCKSTEP_TEAM), "r"(pred)
				             : "memory");
				if (any == 0) break;
			}
#else
			if (__syncthreads_or((have || !exhausted) ? 1 : 0) == 0) break;
#endif
		}
#else
		if (!__any_sync(FULL, have || !exhausted)) break;
#endif

		bool done = !have;
		if (!done && newstep) {
			newstep = false;
			if (!B.begin_step()) {
				ok = false;
				done = true;
			}
		}
		const bool go = !done;
		const unsigned mask = __ballot_sync(FULL, go);
		if (go) {
			const int r = B.attempt(mask);
			if (r == T_FAILED) {
				ok = false;
				done = true;
			} else if (r == T_DONE) {
				steps++;
				double yout[E], tret;
				const bool tstop_return = B.after_step(yout, tret);
				const double creation_time = B.sc[SC_CREATION];
				while (tpi < T && tret >= (a.timepoints[tpi] - creation_time)) {
					const double tq = a.timepoints[tpi] - creation_time;
					if (!B.dky_ok(tq)) {
						ok = false;
						done = true;
						break;
					}
#pragma unroll
					for (int d = 0; d < ND; d++) {
						int r = tpi;
						if constexpr (ND > 1) r = a.tp_rows[tpi * ND + d];
						if (r < 0) continue; // group-uniform: this data set does not ask for this time
						double obs_weight[E];
#pragma unroll
						for (int e = 0; e < E; e++) obs_weight[e] = (double)((obs_words[d] >> (4 * e)) & 15u);
						const double sv = B.dky_weighted(tq, obs_weight);
						if (B.lg == 0) out[(long long)r * stride] = sv;
					}
					tpi++;
				}
#if CP_DIVISION
				// Cell::integration_step_cb (Cell.cpp:463-538, the branch without stored integration points), called by the solver
				// after the outputs of the step and before its own end test (ODESolverCVODE.cpp:431-441): a cell whose
				// "cytokinesis" (divide_cells) or "apoptosis" species has passed 1 ends at this step, with this state
				if (ok && a.nuclear_envelope_ix >= 0 && !mitotic) { // Cell.cpp:487-492
					double yne = 1.0;
#pragma unroll
					for (int e = 0; e < E; e++)
						if (B.idx(e) == a.nuclear_envelope_ix) yne = yout[e];
					yne = __shfl_sync(B.gmask, yne, B.gbase + a.nuclear_envelope_ix % G);
					if (yne < 0.5) mitotic = true;
				}
				if (ok) {
					int ev = 0;
					if (a.cytokinesis_ix >= 0 || a.apoptosis_ix >= 0) {
						double ycyt = 0.0, yapo = 0.0;
#pragma unroll
						for (int e = 0; e < E; e++) {
							if (B.idx(e) == a.cytokinesis_ix) ycyt = yout[e];
							if (B.idx(e) == a.apoptosis_ix) yapo = yout[e];
						}
						if (a.cytokinesis_ix >= 0) ycyt = __shfl_sync(B.gmask, ycyt, B.gbase + a.cytokinesis_ix % G);
						if (a.apoptosis_ix >= 0) yapo = __shfl_sync(B.gmask, yapo, B.gbase + a.apoptosis_ix % G);
						if (a.cytokinesis_ix >= 0 && ycyt > 1.0) ev |= 1;
						if (a.apoptosis_ix >= 0 && yapo > 1.0) ev |= 2;
					}
					if (ev) {
						const long long rec = (long long)c * stride + cell;
						if (ev & 1) {
#pragma unroll
							for (int e = 0; e < E; e++)
								if (B.own(e)) a.cell_end_y[rec * N + B.idx(e)] = yout[e];
						}
						if (B.lg == 0) {
							a.cell_event[rec] = ev;
							a.cell_end_time[rec] = tret + creation_time;
							fill_nan(out, tpi); // the cell does not exist after the event
						}
						tpi = T;
						done = true;
					}
				}
#endif
				if (ok && !done) {
					if (tret >= B.sc[SC_END]) done = true;
					else if (steps == a.max_steps) {
						ok = false;
						done = true;
					} else {
						// ODESolverCVODE.cpp:448-461: at a discontinuity ask for the next one and CVodeReInit(t, y)
						const double next_disc = B.sc[SC_NEXT_DISC];
						if (next_disc == next_disc && (tstop_return || next_disc == tret)) {
							const double nd = B.next_discontinuity(tret);
							B.sc[SC_NEXT_DISC] = nd;
							if (nd == nd && nd < INFINITY) {
								B.tstop() = nd;
								B.tstopset = true;
							}
							if (!B.restart(tret, yout, B.sc[SC_END])) {
								ok = false;
								done = true;
							}
						}
					}
				}
				newstep = true;
			}
		}
		if (have && done) {
			if (B.lg == 0) {
				if (!ok) fill_nan(out, tpi);
				a.cell_status[(long long)c * stride + cell] = ok ? 1 : 0;
#if CP_DIVISION
				if (a.cell_mitotic) a.cell_mitotic[(long long)c * stride + cell] = mitotic ? 1 : 0;
#endif
				if (a.cell_steps)
					a.cell_steps[(long long)c * stride + cell] = (a.debug_report == 1) ? B.nfe : (a.debug_report == 2) ? B.nsetups : (a.debug_report == 3) ? B.nje : steps;
			}
			have = false;
		}
	}
}

// internal linkage on purpose: several model libraries live in one process, and a function-local static of an `inline`
// function is a process-wide unique symbol -- the second model would reuse the first one's cached launch configuration
#if CP_RHS_LANES && CP_TABLES_SHARED
constexpr int TABLE_DOUBLES = CP_TAB_NLIT + CP_TAB_NCOEF + (CP_TAB_NIDX + CP_TAB_NTARGET + CP_TAB_NBEGIN + CP_TAB_NLAW + 1) / 2;
static_assert(CP_TAB_BASE == CS * CPW * WPB, "cellpop_host.cuh places the tables right behind the cells' blocks");
#else
constexpr int TABLE_DOUBLES = 0;
#endif
static size_t smem_bytes() { return sizeof(double) * ((size_t)CS * CPW * WPB + TABLE_DOUBLES); }

// launch configuration per DEVICE: the shared-memory opt-in and the occupancy result are per-device state, and one process
// may hold handles of the same model on several devices (the multi-device handle, several ranks' handles in tests)
static int resident_blocks(int* err)
{
	static int blocks_of_device[64] = { 0 };
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev >= 0 && dev < 64 && blocks_of_device[dev] > 0) return blocks_of_device[dev];
	const size_t smem = smem_bytes();
	cudaError_t e = cudaFuncSetAttribute(cellpop_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) { *err = (int)e; return 0; }
	int per_sm = 0, sms = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cellpop_group_kernel, BS, smem);
	if (e != cudaSuccess) { *err = (int)e; return 0; }
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (per_sm < 1) { *err = (int)cudaErrorLaunchOutOfResources; return 0; }
	if (dev >= 0 && dev < 64) blocks_of_device[dev] = per_sm * sms;
	return per_sm * sms;
}

} // namespace cellpop_group

// scratch doubles: 2 for the queue counter + one saved Jacobian per resident group
extern "C" long long cellpop_group_scratch_doubles(int num_chains, int num_cells)
{
	(void)num_chains;
	(void)num_cells;
	int err = 0;
	const int blocks = cellpop_group::resident_blocks(&err);
	if (blocks <= 0) return -(long long)(err ? err : 1);
	return 2ll + (long long)blocks * cellpop_group::WPB * cellpop_group::CPW * (CP_N * CP_N + (CP_M_GLOBAL ? CP_N * cellpop_group::RS : 0));
}

extern "C" int cellpop_group_launch(const CpArgs* args, double* scratch, void* stream)
{
	int err = 0;
	int blocks = cellpop_group::resident_blocks(&err);
	if (blocks <= 0) return err ? err : 1;
	const long long total = args->items ? (long long)args->num_items : (long long)args->num_chains * args->num_cells;
	const long long per_block = cellpop_group::WPB * cellpop_group::CPW;
	const long long needed = (total + per_block - 1) / per_block;
	if (needed < blocks) blocks = (int)(needed > 0 ? needed : 1);
	(void)cudaGetLastError();
	cudaError_t e = cudaMemsetAsync(scratch, 0, 2 * sizeof(double), (cudaStream_t)stream);
	if (e != cudaSuccess) return (int)e + 10000;
	cellpop_group::cellpop_group_kernel<<<blocks, cellpop_group::BS, cellpop_group::smem_bytes(), (cudaStream_t)stream>>>(
	    *args, scratch + 2, reinterpret_cast<unsigned long long*>(scratch));
	e = cudaGetLastError();
	if (e != cudaSuccess && getenv("BCM3B200_DEBUG"))
		fprintf(stderr, "cellpop_group_launch: %s (blocks %d threads %d smem %zu)\n", cudaGetErrorString(e), blocks, cellpop_group::BS, cellpop_group::smem_bytes());
	return (int)e;
}

extern "C" int cellpop_group_info(int* lanes_per_cell, int* threads_per_block, int* smem_bytes_per_block)
{
	*lanes_per_cell = cellpop_group::G;
	*threads_per_block = cellpop_group::BS;
	*smem_bytes_per_block = (int)cellpop_group::smem_bytes();
	return 0;
}
