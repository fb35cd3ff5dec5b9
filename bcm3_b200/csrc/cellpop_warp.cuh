// cellpop_warp.cuh -- K0+K2 of the cellpop path: ONE ODE system (one simulated cell of one chain) PER WARP.
//
// Included at the end of the per-model translation unit that the host generates at PostInitialize time
// (bcm3b200.cu: build_cellpop_module). That unit defines, before including this file:
//   CP_N                 number of ODE-integrated species
//   CP_NUM_OVERRIDES     number of model parameters with per-cell variability; CP_PARAM_OVERRIDE_BODY / CP_OVERRIDE_INIT
//                        the generated accessor / initialiser statements for them
//   generated_derivative the reference generator's RHS text (src/sbml/SBMLModel.cpp:291-365) with its signature made
//                        generic in the `species` and `parameters` argument types
//
// Replaces Cell::Initialize / Cell::Simulate (src/cellpop/Cell.cpp:150-273,423-433) and under them
// ODESolverCVODE::Solve with the difference-quotient Jacobian (src/odecommon/ODESolverCVODE.cpp:322-463,496-537) and
// CVODE 5.3.0 (same control flow as bdf_thread.cuh / the line map there), with BCM3's zero-skipping partial-pivot LU
// (src/utils/EigenPartialPivLUSomewhatSparse.h:38-105).
//
// Mapping: the N state components are spread over the 32 lanes (component e lives in lane e % 32); every vector of
// the integrator (Nordsieck array, weights, corrections, ...) and the N x N Newton matrix live in shared memory.
// Vector operations are lane-parallel, norms are butterfly shuffle reductions (so every lane holds the bit-identical
// value), and ALL scalar bookkeeping (step size, order, coefficients, counters) is computed redundantly by every lane
// from identical inputs -- control flow is therefore warp-uniform by construction: no divergence, no votes needed.
// The N RHS evaluations of a difference-quotient Jacobian run in parallel, one perturbed column per lane.
#pragma once

#include <cstdint>

#include "cellpop_args.h"

#ifndef CP_WARPS_PER_BLOCK
#define CP_WARPS_PER_BLOCK 4
#endif

namespace cellpop {

constexpr int N = CP_N;
constexpr int LD = (N % 2 == 0) ? N + 1 : N; // odd leading dimension: conflict-free column-per-lane writes
constexpr unsigned FULL = 0xffffffffu;

// cvode.c:142-172, cvode_nls.c:29-31, cvode_ls_impl.h:40-42
#define CPC_FUZZ_FACTOR 100.0
#define CPC_HLB_FACTOR 100.0
#define CPC_HUB_FACTOR 0.1
#define CPC_H_BIAS 0.5
#define CPC_MAX_ITERS 4
#define CPC_CORTES 0.1
#define CPC_THRESH 1.5
#define CPC_ETAMX1 10000.0
#define CPC_ETAMX2 10.0
#define CPC_ETAMX3 10.0
#define CPC_ETAMXF 0.2
#define CPC_ETAMIN 0.1
#define CPC_ETACF 0.25
#define CPC_ADDON 0.000001
#define CPC_BIAS1 6.0
#define CPC_BIAS2 6.0
#define CPC_BIAS3 10.0
#define CPC_ONEPSM 1.000001
#define CPC_SMALL_NST 10
#define CPC_MXNCF 10
#define CPC_MXNEF 7
#define CPC_MXNEF1 3
#define CPC_SMALL_NEF 2
#define CPC_LONG_WAIT 10
#define CPC_DGMAX 0.3
#define CPC_MSBP 20
#define CPC_NLS_MAXCOR 3
#define CPC_CRDOWN 0.3
#define CPC_RDIV 2.0
#define CPC_MSBJ 50
#define CPC_LS_DGMAX 0.2
#define CPC_UROUND DBL_EPSILON

enum { CP_FIRST_CALL = 101, CP_PREV_CONV_FAIL = 102, CP_PREV_ERR_FAIL = 103 };
enum { CP_NO_FAILURES = 0, CP_FAIL_BAD_J = 1, CP_FAIL_OTHER = 2 };
enum { CP_STEP_OK = 0, CP_STEP_TSTOP = 1, CP_STEP_FAIL = -1 };

// per-cell view of the variable vector: `parameters[k]` in the generated text. k is a literal there, so the
// comparison chain against the compile-time override indices folds away.
struct CellParameters {
	const double* base;
	double ov[CP_NUM_OVERRIDES > 0 ? CP_NUM_OVERRIDES : 1];
	__device__ __forceinline__ double operator[](int k) const
	{
		CP_PARAM_OVERRIDE_BODY // generated: `if (k == <variable index>) return ov[<slot>];` per overridden parameter
		return base[k];
	}
};

// `species[i]` views
struct SpeciesPlain {
	const double* y;
	__device__ __forceinline__ double operator[](int i) const { return y[i]; }
};
struct SpeciesPerturbed { // y + inc * e_j, without materialising a copy per lane
	const double* y;
	int j;
	double yj;
	__device__ __forceinline__ double operator[](int i) const { return (i == j) ? yj : y[i]; }
};

// shared memory of one warp
struct WarpMem {
	double zn[6][N];
	double ewt[N], acor[N], y[N], ftemp[N], tempv[N], delta[N], yout[N];
	double A[N * LD];      // I - gamma J, then its LU factors (column-major, leading dimension LD)
	double savedJ[N * LD];
	int piv[N];
	double tau[8], l[8], tq[8];
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
	return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, off));
	return v;
}

struct WarpBdf {
	WarpMem* m;
	int lane;
	const double* constant_species;
	const double* non_sampled;
	CellParameters params;
	double reltol, abstol, hmin;

	// integrator scalars, identical in every lane
	double tn, h, hprime, hscale, eta, etamax, hu;
	double gamma, gammap, gamrat, rl1, crate, delp, acnrm, saved_tq5, tstop, tretlast;
	int q, qprime, L, qwait, nst, nstlp, nstlj;
	bool tstopset, nls_jcur;
	int nfe, nsetups, nje;

	// ---- vector helpers (lane-parallel over components) ----
	__device__ __forceinline__ double wrms(const double* x) const
	{
		double s = 0.0;
		for (int i = lane; i < N; i += 32) {
			double p = x[i] * m->ewt[i];
			s += p * p;
		}
		return sqrt(warp_sum(s) * (1.0 / N));
	}

	__device__ __forceinline__ void rhs(double t, const double* ysrc, double* out)
	{
		// every lane evaluates the whole generated function on the same inputs (same cost as one lane doing it under
		// SIMT) and stores identical values; Cell::solver_rhs_fn, Cell.cpp:423-433
		(void)t;
		__syncwarp();
		SpeciesPlain sp{ ysrc };
		generated_derivative(out, sp, constant_species, params, non_sampled);
		__syncwarp();
		nfe++;
	}

	__device__ __forceinline__ void set_ewt(const double* ycur, double* w)
	{
		for (int i = lane; i < N; i += 32) w[i] = 1.0 / (reltol * fabs(ycur[i]) + abstol);
		__syncwarp();
	}

	__device__ __forceinline__ void create()
	{
		for (int j = 0; j < 6; j++)
			for (int i = lane; i < N; i += 32) m->zn[j][i] = 0.0;
		for (int i = lane; i < N; i += 32) m->acor[i] = 0.0;
		if (lane < 8) {
			m->tau[lane] = 0.0;
			m->l[lane] = 0.0;
			m->tq[lane] = 0.0;
		}
		gammap = 0.0;
		crate = 1.0;
		delp = 0.0;
		acnrm = 0.0;
		saved_tq5 = 0.0;
		tstopset = false;
		tstop = 0.0;
		h = hprime = hscale = eta = hu = gamma = gamrat = rl1 = 0.0;
		nls_jcur = false;
		nstlj = 0;
		nfe = 0;
		nsetups = 0;
		nje = 0;
		q = 1;
		L = 2;
		qwait = 2;
		nst = 0;
		nstlp = 0;
		etamax = CPC_ETAMX1;
		__syncwarp();
	}

	// CVodeReInit (cvode.c:586-665)
	__device__ __forceinline__ void reinit(double t0, const double* y0)
	{
		tn = t0;
		q = 1;
		L = 2;
		qwait = 2;
		etamax = CPC_ETAMX1;
		hu = 0.0;
		nst = 0;
		nstlp = 0;
		for (int i = lane; i < N; i += 32) m->zn[0][i] = y0[i];
		__syncwarp();
	}

	// CVodeGetDky(t, 0), cvode.c:1467-1524
	__device__ __forceinline__ bool dky(double t, double* out)
	{
		double tfuzz = CPC_FUZZ_FACTOR * CPC_UROUND * (fabs(tn) + fabs(hu));
		if (hu < 0.0) tfuzz = -tfuzz;
		const double tp = tn - hu - tfuzz, tn1 = tn + tfuzz;
		if ((t - tp) * (t - tn1) > 0.0) return false;
		const double s = (t - tn) / h;
		for (int i = lane; i < N; i += 32) {
			double acc = 0.0;
			for (int j = q; j >= 0; j--) {
				double c = 1.0;
				for (int k = 0; k < j; k++) c *= s;
				acc = (j == q) ? c * m->zn[j][i] : acc + c * m->zn[j][i];
			}
			out[i] = acc;
		}
		__syncwarp();
		return true;
	}

	// ---- linear algebra ----
	// ODESolverCVODE::DifferenceQuotientJacobian (ODESolverCVODE.cpp:496-537): column j on lane j
	__device__ __forceinline__ void dq_jacobian(const double* yc, const double* fy)
	{
		const double srur = sqrt(CPC_UROUND);
		const double fnorm = wrms(fy);
		const double minInc = (fnorm != 0.0) ? (1000.0 * fabs(h) * CPC_UROUND * N * fnorm) : 1.0;
		__syncwarp();
		for (int j = lane; j < N; j += 32) {
			const double inc = fmax(srur * fabs(yc[j]), minInc / m->ewt[j]);
			SpeciesPerturbed sp{ yc, j, yc[j] + inc };
			double* col = m->A + j * LD;
			generated_derivative(col, sp, constant_species, params, non_sampled);
			const double inc_inv = 1.0 / inc;
			for (int i = 0; i < N; i++) col[i] = inc_inv * (col[i] - fy[i]);
		}
		__syncwarp();
	}

	// PartialPivLUExtended::compute_optimized (EigenPartialPivLUSomewhatSparse.h:38-105)
	__device__ __forceinline__ void lu_factor()
	{
		double* A = m->A;
		for (int k = 0; k < N; k++) {
			// pivot: first row with the largest |A(i,k)|, i >= k
			double best = -1.0;
			int bi = k;
			for (int i = k + lane; i < N; i += 32) {
				double v = fabs(A[i + k * LD]);
				if (v > best) {
					best = v;
					bi = i;
				}
			}
#pragma unroll
			for (int off = 16; off > 0; off >>= 1) {
				double ob = __shfl_xor_sync(FULL, best, off);
				int oi = __shfl_xor_sync(FULL, bi, off);
				if (ob > best || (ob == best && oi < bi)) {
					best = ob;
					bi = oi;
				}
			}
			if (lane == 0) m->piv[k] = bi;
			if (best != 0.0) {
				if (bi != k) {
					for (int j = lane; j < N; j += 32) {
						double tmp = A[k + j * LD];
						A[k + j * LD] = A[bi + j * LD];
						A[bi + j * LD] = tmp;
					}
					__syncwarp();
				}
				const double inv_coeff = 1.0 / A[k + k * LD];
				for (int i = k + 1 + lane; i < N; i += 32) A[i + k * LD] *= inv_coeff;
			}
			__syncwarp();
			for (int j = k + 1; j < N; j++) {
				const double a_kj = A[k + j * LD];
				if (a_kj != 0.0) {
					for (int i = k + 1 + lane; i < N; i += 32) A[i + j * LD] -= a_kj * A[i + k * LD];
				}
			}
			__syncwarp();
		}
	}

	// PartialPivLU::solve: b <- P b, unit-lower forward substitution, upper back substitution
	__device__ __forceinline__ void lu_solve(double* b)
	{
		const double* A = m->A;
		if (lane == 0) {
			for (int k = 0; k < N; k++) {
				int p = m->piv[k];
				if (p != k) {
					double tmp = b[k];
					b[k] = b[p];
					b[p] = tmp;
				}
			}
		}
		__syncwarp();
		for (int k = 0; k < N; k++) {
			const double xk = b[k];
			__syncwarp();
			for (int i = k + 1 + lane; i < N; i += 32) b[i] -= xk * A[i + k * LD];
			__syncwarp();
		}
		for (int k = N - 1; k >= 0; k--) {
			if (lane == 0) b[k] /= A[k + k * LD];
			__syncwarp();
			const double xk = b[k];
			for (int i = lane; i < k; i += 32) b[i] -= xk * A[i + k * LD];
			__syncwarp();
		}
	}

	// cvLsSetup + cvLsLinSys (cvode_ls.c:1415-1507,1201-1286)
	__device__ __forceinline__ void ls_setup(int convfail)
	{
		const double dgamma = fabs((gamma / gammap) - 1.0);
		const bool jbad = (nst == 0) || (nst > nstlj + CPC_MSBJ) || ((convfail == CP_FAIL_BAD_J) && (dgamma < CPC_LS_DGMAX)) ||
		                  (convfail == CP_FAIL_OTHER);
		if (!jbad) {
			for (int e = lane; e < N * LD; e += 32) m->A[e] = m->savedJ[e];
		} else {
			dq_jacobian(m->y, m->ftemp);
			for (int e = lane; e < N * LD; e += 32) m->savedJ[e] = m->A[e];
			nstlj = nst;
			nje++;
		}
		nsetups++;
		nls_jcur = jbad;
		__syncwarp();
		// SUNMatScaleAddI(-gamma, A)
		for (int e = lane; e < N * LD; e += 32) m->A[e] *= -gamma;
		__syncwarp();
		for (int i = lane; i < N; i += 32) m->A[i + i * LD] += 1.0;
		__syncwarp();
		lu_factor();
	}

	// cvNlsResidual (cvode_nls.c:281-315): delta = rl1*zn[1] + acor - gamma*f(tn, zn[0] + acor)
	__device__ __forceinline__ void residual()
	{
		for (int i = lane; i < N; i += 32) m->y[i] = m->zn[0][i] + m->acor[i];
		rhs(tn, m->y, m->ftemp);
		for (int i = lane; i < N; i += 32) {
			double r = rl1 * m->zn[1][i] + m->acor[i];
			r += -gamma * m->ftemp[i];
			m->delta[i] = r;
		}
		__syncwarp();
	}

	// cvRescale (cvode.c:2384-2400)
	__device__ __forceinline__ void rescale()
	{
		double c = eta;
		for (int j = 1; j <= q; j++) {
			for (int i = lane; i < N; i += 32) m->zn[j][i] *= c;
			c = eta * c;
		}
		h = hscale * eta;
		hscale = h;
		__syncwarp();
	}

	// cvIncreaseBDF / cvDecreaseBDF / cvAdjustOrder (cvode.c:2213-2374)
	__device__ __forceinline__ void adjust_order(int deltaq)
	{
		if ((q == 2) && (deltaq != 1)) return;
		double ll[6];
		for (int i = 0; i < 6; i++) ll[i] = 0.0;
		ll[2] = 1.0;
		if (deltaq == 1) {
			double alpha1 = 1.0, prod = 1.0, xiold = 1.0, alpha0 = -1.0, hsum = hscale;
			for (int j = 1; j < q; j++) {
				hsum += m->tau[j + 1];
				const double xi = hsum / hscale;
				prod *= xi;
				alpha0 -= 1.0 / (j + 1);
				alpha1 += 1.0 / xi;
				for (int i = j + 2; i >= 2; i--) ll[i] = ll[i] * xiold + ll[i - 1];
				xiold = xi;
			}
			const double A1 = (-alpha0 - alpha1) / prod;
			for (int i = lane; i < N; i += 32) {
				const double znL = A1 * m->zn[5][i];
				m->zn[L][i] = znL;
				for (int j = 2; j <= q; j++) m->zn[j][i] += ll[j] * znL;
			}
		} else if (deltaq == -1) {
			double hsum = 0.0;
			for (int j = 1; j <= q - 2; j++) {
				hsum += m->tau[j];
				const double xi = hsum / hscale;
				for (int i = j + 2; i >= 2; i--) ll[i] = ll[i] * xi + ll[i - 1];
			}
			if (q > 2) {
				for (int i = lane; i < N; i += 32) {
					const double znq = m->zn[q][i];
					for (int j = 2; j < q; j++) m->zn[j][i] += (-ll[j]) * znq;
				}
			}
		}
		__syncwarp();
	}

	// cvPredict / cvRestore (cvode.c:2412-2425, 2918-2927)
	__device__ __forceinline__ void predict()
	{
		tn += h;
		if (tstopset) {
			if ((tn - tstop) * h > 0.0) tn = tstop;
		}
		for (int i = lane; i < N; i += 32) {
			for (int k = 1; k <= q; k++)
				for (int j = q; j >= k; j--) m->zn[j - 1][i] += m->zn[j][i];
		}
		__syncwarp();
	}
	__device__ __forceinline__ void restore(double saved_t)
	{
		tn = saved_t;
		for (int i = lane; i < N; i += 32) {
			for (int k = 1; k <= q; k++)
				for (int j = q; j >= k; j--) m->zn[j - 1][i] = m->zn[j - 1][i] - m->zn[j][i];
		}
		__syncwarp();
	}

	// cvSet + cvSetBDF + cvSetTqBDF (cvode.c:2445-2460, 2611-2686); l/tq kept in per-lane registers via local arrays
	__device__ __forceinline__ void set_bdf(double (&l)[6], double (&tq)[6])
	{
		double alpha0, alpha0_hat, xi_inv, xistar_inv, hsum;
		l[0] = l[1] = xi_inv = xistar_inv = 1.0;
		for (int i = 2; i <= q; i++) l[i] = 0.0;
		alpha0 = alpha0_hat = -1.0;
		hsum = h;
		if (q > 1) {
			for (int j = 2; j < q; j++) {
				hsum += m->tau[j - 1];
				xi_inv = h / hsum;
				alpha0 -= 1.0 / j;
				for (int i = j; i >= 1; i--) l[i] += l[i - 1] * xi_inv;
			}
			alpha0 -= 1.0 / q;
			xistar_inv = -l[1] - alpha0;
			hsum += m->tau[q - 1];
			xi_inv = h / hsum;
			alpha0_hat = -l[1] - xi_inv;
			for (int i = q; i >= 1; i--) l[i] += l[i - 1] * xistar_inv;
		}
		const double A1 = 1.0 - alpha0_hat + alpha0;
		const double A2 = 1.0 + q * A1;
		tq[2] = fabs(A1 / (alpha0 * A2));
		tq[5] = fabs(A2 * xistar_inv / (l[q] * xi_inv));
		if (qwait == 1) {
			if (q > 1) {
				const double C = xistar_inv / l[q];
				const double A3 = alpha0 + 1.0 / q;
				const double A4 = alpha0_hat + xi_inv;
				const double Cpinv = (1.0 - A4 + A3) / A3;
				tq[1] = fabs(C * Cpinv);
			} else {
				tq[1] = 1.0;
			}
			hsum += m->tau[q];
			xi_inv = h / hsum;
			const double A5 = alpha0 - (1.0 / (q + 1));
			const double A6 = alpha0_hat - xi_inv;
			const double Cppinv = (1.0 - A6 + A5) / A2;
			tq[3] = fabs(Cppinv / (xi_inv * (q + 2) * A5));
		}
		tq[4] = CPC_CORTES / tq[2];
		rl1 = 1.0 / l[1];
		gamma = h * rl1;
		if (nst == 0) gammap = gamma;
		gamrat = (nst > 0) ? gamma / gammap : 1.0;
	}

	__device__ __forceinline__ double root(double base, double inv_k) const
	{
		// SUNRpowerR(base, 1/k)
		if (base <= 0.0) return 0.0;
		return pow(base, inv_k);
	}

	// cvStep (cvode.c:2082-2174). Returns false where CVode returns a negative flag.
	__device__ __forceinline__ bool take_step()
	{
		const double saved_t = tn;
		double dsm = 0.0;
		int ncf = 0, nef = 0;
		int nflag = CP_FIRST_CALL;
		// l[] / tq[] are indexed with the run-time order: they live in per-warp shared memory (identical in all lanes)
		double l[6], tq[6];
		for (int i = 0; i < 6; i++) {
			l[i] = 0.0;
			tq[i] = 0.0;
		}
		// tq[1], tq[3] persist across steps only between the qwait == 1 step and ... the same step (see bdf_thread.cuh)

		if ((nst > 0) && (hprime != h)) {
			if (qprime != q) {
				adjust_order(qprime - q);
				q = qprime;
				L = q + 1;
				qwait = L;
			}
			rescale();
		}

		for (;;) {
			predict();
			set_bdf(l, tq);

			// ---- cvNls + Newton ----
			int nls_ret = 1;
			{
				int convfail = ((nflag == CP_FIRST_CALL) || (nflag == CP_PREV_ERR_FAIL)) ? CP_NO_FAILURES : CP_FAIL_OTHER;
				bool callSetup = (nflag == CP_PREV_CONV_FAIL) || (nflag == CP_PREV_ERR_FAIL) || (nst == 0) || (nst >= nstlp + CPC_MSBP) ||
				                 (fabs(gamrat - 1.0) > CPC_DGMAX);
				for (int i = lane; i < N; i += 32) m->acor[i] = 0.0;
				__syncwarp();
				const double tol = tq[4];
				bool jbad = false;
				for (;;) {
					residual();
					if (callSetup) {
						if (jbad) convfail = CP_FAIL_BAD_J;
						ls_setup(convfail);
						gamrat = 1.0;
						gammap = gamma;
						crate = 1.0;
						nstlp = nst;
					}
					bool failed_pass = false;
					for (int mi = 0; mi < CPC_NLS_MAXCOR; mi++) {
						for (int i = lane; i < N; i += 32) m->delta[i] = -m->delta[i];
						__syncwarp();
						lu_solve(m->delta);
						if (gamrat != 1.0) {
							const double sc = 2.0 / (1.0 + gamrat);
							for (int i = lane; i < N; i += 32) m->delta[i] *= sc;
						}
						for (int i = lane; i < N; i += 32) m->acor[i] += m->delta[i];
						__syncwarp();
						const double del = wrms(m->delta);
						if (mi > 0) crate = fmax(CPC_CRDOWN * crate, del / delp);
						const double dcon = del * fmin(1.0, crate) / tol;
						if (dcon <= 1.0) {
							acnrm = (mi == 0) ? del : wrms(m->acor);
							nls_jcur = false;
							nls_ret = 0;
							break;
						}
						if ((mi >= 1) && (del > CPC_RDIV * delp)) {
							failed_pass = true;
							break;
						}
						delp = del;
						if (mi + 1 >= CPC_NLS_MAXCOR) {
							failed_pass = true;
							break;
						}
						residual();
					}
					(void)failed_pass;
					if (nls_ret == 0) break;
					if (!nls_jcur) {
						callSetup = true;
						jbad = true;
						for (int i = lane; i < N; i += 32) m->acor[i] = 0.0;
						__syncwarp();
						continue;
					}
					break;
				}
			}

			// ---- cvHandleNFlag ----
			if (nls_ret != 0) {
				restore(saved_t);
				ncf++;
				etamax = 1.0;
				if ((fabs(h) <= hmin * CPC_ONEPSM) || (ncf == CPC_MXNCF)) return false;
				eta = fmax(CPC_ETACF, hmin / fabs(h));
				nflag = CP_PREV_CONV_FAIL;
				rescale();
				continue;
			}

			// ---- cvDoErrorTest ----
			dsm = acnrm * tq[2];
			if (dsm <= 1.0) break;
			nef++;
			nflag = CP_PREV_ERR_FAIL;
			restore(saved_t);
			if ((fabs(h) <= hmin * CPC_ONEPSM) || (nef == CPC_MXNEF)) return false;
			etamax = 1.0;
			if (nef <= CPC_MXNEF1) {
				eta = 1.0 / (root(CPC_BIAS2 * dsm, 1.0 / L) + CPC_ADDON);
				eta = fmax(CPC_ETAMIN, fmax(eta, hmin / fabs(h)));
				if (nef >= CPC_SMALL_NEF) eta = fmin(eta, CPC_ETAMXF);
				rescale();
				continue;
			}
			if (q > 1) {
				eta = fmax(CPC_ETAMIN, hmin / fabs(h));
				adjust_order(-1);
				L = q;
				q--;
				qwait = L;
				rescale();
				continue;
			}
			eta = fmax(CPC_ETAMIN, hmin / fabs(h));
			h *= eta;
			hscale = h;
			qwait = CPC_LONG_WAIT;
			rhs(tn, m->zn[0], m->tempv);
			for (int i = lane; i < N; i += 32) m->zn[1][i] = h * m->tempv[i];
			__syncwarp();
		}

		// ---- cvCompleteStep ----
		nst++;
		hu = h;
		{
			double tau[6];
			for (int i = 0; i < 6; i++) tau[i] = m->tau[i];
			__syncwarp();
			for (int i = q; i >= 2; i--) tau[i] = tau[i - 1];
			if ((q == 1) && (nst > 1)) tau[2] = tau[1];
			tau[1] = h;
			if (lane == 0)
				for (int i = 0; i < 6; i++) m->tau[i] = tau[i];
			__syncwarp();
		}
		for (int i = lane; i < N; i += 32) {
			const double ac = m->acor[i];
			for (int j = 0; j <= q; j++) m->zn[j][i] += l[j] * ac;
		}
		qwait--;
		if ((qwait == 1) && (q != 5)) {
			for (int i = lane; i < N; i += 32) m->zn[5][i] = m->acor[i];
			saved_tq5 = tq[5];
		}
		__syncwarp();

		// ---- cvPrepareNextStep ----
		if (etamax == 1.0) {
			qwait = (qwait > 2) ? qwait : 2;
			qprime = q;
			hprime = h;
			eta = 1.0;
		} else {
			const double etaq = 1.0 / (root(CPC_BIAS2 * dsm, 1.0 / L) + CPC_ADDON);
			eta = etaq;
			qprime = q;
			if (qwait == 0) {
				qwait = 2;
				double etaqm1 = 0.0;
				if (q > 1) {
					const double ddn = wrms(m->zn[q]) * tq[1];
					etaqm1 = 1.0 / (root(CPC_BIAS1 * ddn, 1.0 / q) + CPC_ADDON);
				}
				double etaqp1 = 0.0;
				if (q != 5) {
					if (saved_tq5 != 0.0) {
						const double base = h / m->tau[2];
						double pw = 1.0;
						for (int i = 1; i <= L; i++) pw *= base;
						const double cquot = (tq[5] / saved_tq5) * pw;
						for (int i = lane; i < N; i += 32) m->tempv[i] = -cquot * m->zn[5][i] + m->acor[i];
						__syncwarp();
						const double dup = wrms(m->tempv) * tq[3];
						etaqp1 = 1.0 / (root(CPC_BIAS3 * dup, 1.0 / (L + 1)) + CPC_ADDON);
					}
				}
				const double etam = fmax(etaqm1, fmax(etaq, etaqp1));
				if (etam < CPC_THRESH) {
					eta = 1.0;
					qprime = q;
				} else if (etam == etaq) {
					eta = etaq;
					qprime = q;
				} else if (etam == etaqm1) {
					eta = etaqm1;
					qprime = q - 1;
				} else {
					eta = etaqp1;
					qprime = q + 1;
					for (int i = lane; i < N; i += 32) m->zn[5][i] = m->acor[i];
					__syncwarp();
				}
			}
			if (eta < CPC_THRESH) {
				eta = 1.0;
				hprime = h;
			} else {
				eta = fmin(eta, etamax);
				hprime = h * eta; // hmax_inv = 0 (solver_max_timestep = inf, Cell.cpp:73)
			}
		}
		etamax = (nst <= CPC_SMALL_NST) ? CPC_ETAMX2 : CPC_ETAMX3;
		for (int i = lane; i < N; i += 32) m->acor[i] *= tq[2];
		__syncwarp();
		return true;
	}

	// cvHin (cvode.c:1884-1984) with cvUpperBoundH0 / cvYddNorm
	__device__ __forceinline__ bool hin(double tout)
	{
		const double tdiff = tout - tn;
		if (tdiff == 0.0) return false;
		const double sign = (tdiff > 0.0) ? 1.0 : -1.0;
		const double tdist = fabs(tdiff);
		const double tround = CPC_UROUND * fmax(fabs(tn), fabs(tout));
		if (tdist < 2.0 * tround) return false;
		const double hlb = CPC_HLB_FACTOR * tround;
		double hub_inv = -INFINITY;
		for (int i = lane; i < N; i += 32) {
			double t2 = fabs(m->zn[0][i]);
			double t1 = 1.0 / m->ewt[i];
			t1 = CPC_HUB_FACTOR * t2 + t1;
			t2 = fabs(m->zn[1][i]);
			t1 = t2 / t1;
			hub_inv = (t1 > hub_inv) ? t1 : hub_inv;
		}
		hub_inv = warp_max(hub_inv);
		double hub = CPC_HUB_FACTOR * tdist;
		if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;
		double hg = sqrt(hlb * hub);
		if (hub < hlb) {
			h = (sign < 0.0) ? -hg : hg;
			return true;
		}
		double hnew = hg;
		for (int count1 = 1; count1 <= CPC_MAX_ITERS; count1++) {
			const double hgs = hg * sign;
			for (int i = lane; i < N; i += 32) m->y[i] = hgs * m->zn[1][i] + m->zn[0][i];
			rhs(tn + hgs, m->y, m->tempv);
			const double c = 1.0 / hgs;
			for (int i = lane; i < N; i += 32) m->tempv[i] = c * (m->tempv[i] - m->zn[1][i]);
			__syncwarp();
			const double yddnrm = wrms(m->tempv);
			hnew = (yddnrm * hub * hub > 2.0) ? sqrt(2.0 / yddnrm) : sqrt(hg * hub);
			if (count1 == CPC_MAX_ITERS) break;
			const double hrat = hnew / hg;
			if ((hrat > 0.5) && (hrat < 2.0)) break;
			if ((count1 > 1) && (hrat > 2.0)) {
				hnew = hg;
				break;
			}
			hg = hnew;
		}
		double h0 = CPC_H_BIAS * hnew;
		if (h0 < hlb) h0 = hlb;
		if (h0 > hub) h0 = hub;
		if (sign < 0.0) h0 = -h0;
		h = h0;
		return true;
	}

	// CVode(tout, CV_ONE_STEP), cvode.c:1006-1443. yout receives y at tret.
	__device__ __forceinline__ int step(double tout, double& tret)
	{
		if (nst == 0) {
			tretlast = tret = tn;
			set_ewt(m->zn[0], m->ewt);
			nstlj = 0;
			nls_jcur = false;
			rhs(tn, m->zn[0], m->zn[1]);
			if (tstopset) {
				if ((tstop - tn) * (tout - tn) <= 0.0) return CP_STEP_FAIL;
			}
			double tout_hin = tout;
			if (tstopset && (tout - tn) * (tout - tstop) > 0.0) tout_hin = tstop;
			if (!hin(tout_hin)) return CP_STEP_FAIL;
			if (fabs(h) < hmin) h *= hmin / fabs(h);
			if (tstopset) {
				if ((tn + h - tstop) * h > 0.0) h = (tstop - tn) * (1.0 - 4.0 * CPC_UROUND);
			}
			hscale = h;
			hprime = h;
			for (int i = lane; i < N; i += 32) m->zn[1][i] *= h;
			__syncwarp();
		}
		if (nst > 0) set_ewt(m->zn[0], m->ewt);
		const double tolsf = CPC_UROUND * wrms(m->zn[0]);
		if (tolsf > 1.0) return CP_STEP_FAIL;
		if (!take_step()) return CP_STEP_FAIL;
		if (tstopset) {
			const double troundoff = CPC_FUZZ_FACTOR * CPC_UROUND * (fabs(tn) + fabs(h));
			if (fabs(tn - tstop) <= troundoff) {
				(void)dky(tstop, m->yout);
				tretlast = tret = tstop;
				tstopset = false;
				return CP_STEP_TSTOP;
			}
			if ((tn + hprime - tstop) * h > 0.0) {
				hprime = (tstop - tn) * (1.0 - 4.0 * CPC_UROUND);
				eta = hprime / h;
			}
		}
		tretlast = tret = tn;
		for (int i = lane; i < N; i += 32) m->yout[i] = m->zn[0][i];
		__syncwarp();
		return CP_STEP_OK;
	}
};

__device__ __forceinline__ void apply_variability(double& x, double value, int apply)
{
	// VariabilityDescriptionVariable::Apply, VariabilityDescriptionVariable.cpp:172-207
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

// grid = (ceil(num_cells / WARPS), C); warp = one cell of chain blockIdx.y
__global__ void __launch_bounds__(32 * CP_WARPS_PER_BLOCK) cellpop_kernel(const CpArgs a)
{
	extern __shared__ unsigned char smem_raw[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int c = blockIdx.y;
	const int cell = blockIdx.x * CP_WARPS_PER_BLOCK + warp; // index inside this shard
	if (cell >= a.num_cells) return; // whole warp exits together
	WarpMem* m = reinterpret_cast<WarpMem*>(smem_raw) + warp;

	WarpBdf S;
	S.m = m;
	S.lane = lane;
	S.constant_species = a.constant_species;
	S.non_sampled = a.non_sampled;
	S.reltol = a.rel_tol;
	S.abstol = a.abs_tol;
	S.hmin = a.min_dt;
	const double* tv = a.transformed + (long long)c * a.nvar;
	S.params.base = tv;

	// ---- K0: Cell::Initialize (Cell.cpp:150-191): per-cell parameters and initial conditions ----
	CP_OVERRIDE_INIT // generated: `S.params.ov[<slot>] = tv[<variable index>];` per overridden parameter
	for (int i = lane; i < N; i += 32) m->y[i] = a.initial_conditions[i];
	__syncwarp();
	const long long gcell = (long long)a.cell_offset + cell;
	for (int d = 0; d < a.D; d++) {
		double v = cellpop_variability_value(a, tv, c, gcell, d);
		if (a.var_negate[d]) v = -v;
		if (a.var_is_ic[d]) {
			if (lane == 0) {
				double x = m->y[a.var_slot[d]];
				apply_variability(x, v, a.var_apply[d]);
				m->y[a.var_slot[d]] = x;
			}
			__syncwarp();
		} else {
#pragma unroll
			for (int s = 0; s < CP_NUM_OVERRIDES; s++) {
				if (a.var_slot[d] == s) apply_variability(S.params.ov[s], v, a.var_apply[d]);
			}
		}
	}

	// ---- Cell::Simulate (Cell.cpp:193-273) + ODESolver::SolveReturnSolution + ODESolverCVODE::Solve ----
	const double creation_time = (a.entry_time_ix >= 0) ? tv[a.entry_time_ix] : a.entry_time_fixed;
	const int T = a.T;
	double* out = a.cell_values + ((long long)c * T) * a.num_cells + cell; // stride num_cells between timepoints
	const double nan = __longlong_as_double(0x7ff8000000000000ll);
	bool ok = true;
	int steps = 0;

	S.create();
	auto observe = [&](const double* yv) {
		double s = 0.0;
		for (int k = 0; k < a.num_obs_species; k++) s += yv[a.obs_species[k]];
		return s;
	};

	// cell-relative output times; times before creation never become available (GetInterpolatedSpeciesValue returns NaN
	// for cell_time < 0, Cell.cpp:319-321); ODESolver.cpp:109-118 hands out the initial condition for cell_time < eps
	int ti = 0;
	while (ti < T && (a.timepoints[ti] - creation_time) < DBL_EPSILON) {
		const double cell_time = a.timepoints[ti] - creation_time;
		if (lane == 0) out[(long long)ti * a.num_cells] = (cell_time < 0.0) ? nan : observe(m->y);
		ti++;
	}
	if ((a.sim_end_time - creation_time) >= DBL_EPSILON) {
		const double end_time = a.sim_end_time - creation_time;
		S.reinit(0.0, m->y);
		S.tstopset = false; // no treatment trajectories: no discontinuities
		double t = 0.0;
		int tpi = ti;
		for (;;) {
			double tret;
			const int r = S.step(end_time, tret);
			if (r == CP_STEP_FAIL) {
				ok = false;
				break;
			}
			t = tret;
			steps++;
			bool bad = false;
			while (tpi < T && tret >= (a.timepoints[tpi] - creation_time)) {
				if (!S.dky(a.timepoints[tpi] - creation_time, m->tempv)) {
					bad = true;
					break;
				}
				if (lane == 0) out[(long long)tpi * a.num_cells] = observe(m->tempv);
				__syncwarp();
				tpi++;
			}
			if (bad) {
				ok = false;
				break;
			}
			if (t >= end_time) break;
			if (steps == a.max_steps) {
				ok = false;
				break;
			}
		}
		if (!ok) {
			for (int k = tpi; k < T; k++)
				if (lane == 0) out[(long long)k * a.num_cells] = nan;
		}
	}
	if (lane == 0) {
		a.cell_status[(long long)c * a.num_cells + cell] = ok ? 1 : 0;
		if (a.cell_steps) a.cell_steps[(long long)c * a.num_cells + cell] = (a.debug_report == 1) ? S.nfe : (a.debug_report == 2) ? S.nsetups : (a.debug_report == 3) ? S.nje : steps;
	}
}

} // namespace cellpop

extern "C" int cellpop_launch(const CpArgs* args, void* stream)
{
	const size_t smem = sizeof(cellpop::WarpMem) * CP_WARPS_PER_BLOCK;
	static bool attr_set = false;
	if (!attr_set && smem > 48 * 1024) {
		cudaError_t e = cudaFuncSetAttribute(cellpop::cellpop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) return (int)e;
		attr_set = true;
	}
	dim3 grid((args->num_cells + CP_WARPS_PER_BLOCK - 1) / CP_WARPS_PER_BLOCK, args->num_chains);
	cellpop::cellpop_kernel<<<grid, 32 * CP_WARPS_PER_BLOCK, smem, (cudaStream_t)stream>>>(*args);
	return (int)cudaGetLastError();
}

extern "C" int cellpop_num_species(void) { return CP_N; }
