// cellpop_thread.cuh -- K0+K2 of the cellpop path for SMALL networks: one ODE system (one simulated cell of one chain) PER
// THREAD. Same algorithm and reference line map as cellpop_warp.cuh / bdf_thread.cuh (CVODE 5.3.0 BDF + Newton with the
// difference-quotient Jacobian of ODESolverCVODE.cpp:496-537 and BCM3's zero-skipping partial-pivot LU), different mapping:
//
//   * the integrator's vectors (Nordsieck array, weights, corrections, work vectors: 12 N doubles) live in SHARED memory,
//     laid out [element][thread] so that the lanes of a warp touch consecutive words (no bank conflicts) and the loops over
//     the order and over N are real loops with run-time bounds -- compact code, any N;
//   * the Newton matrix and the saved Jacobian (2 N^2 doubles) live in a GLOBAL scratch buffer laid out [element][thread]
//     (coalesced; they are touched once per linear solve / setup);
//   * all scalar bookkeeping sits in registers. One instruction stream advances 32 cells, instead of one cell per warp in
//     cellpop_warp.cuh where every lane repeats the scalar work -- for N ~ 10-20 that is the better trade;
//   * control flow is kept converged the way poppk_kernel does it: one step attempt per loop trip, Newton loops made
//     warp-uniform with votes, the warps of a block in lock-step (__syncthreads_or) to share instruction fetches.
#pragma once

#include <cstdint>

#include "cellpop_args.h"

#ifndef CP_THREADS_PER_BLOCK
#define CP_THREADS_PER_BLOCK 64
#endif

namespace cellpop_thread {

constexpr int N = CP_N;
constexpr int BS = CP_THREADS_PER_BLOCK;
constexpr unsigned FULL = 0xffffffffu;
constexpr double UROUND = DBL_EPSILON;

enum { V_ZN0 = 0, V_EWT = 6, V_ACOR = 7, V_Y = 8, V_FTEMP = 9, V_TEMPV = 10, V_DELTA = 11, NUM_VECTORS = 12 };
enum { T_FIRST_CALL = 101, T_PREV_CONV_FAIL = 102, T_PREV_ERR_FAIL = 103 };
enum { T_NO_FAILURES = 0, T_FAIL_BAD_J = 1, T_FAIL_OTHER = 2 };
enum { T_RETRY = 0, T_DONE = 1, T_FAILED = -1 };

// strided views: element i of a per-thread vector / matrix
struct Vec {
	double* p;
	__device__ __forceinline__ double& operator[](int i) const { return p[i * BS]; }
};
struct Mat { // column-major N x N, element (i, j)
	double* p;
	long long stride;
	__device__ __forceinline__ double& at(int i, int j) const { return p[(long long)(i + j * N) * stride]; }
	__device__ __forceinline__ double& lin(int e) const { return p[(long long)e * stride]; }
};
struct MatColumn { // `out[i]` of generated_derivative writing straight into column j
	Mat m;
	int j;
	__device__ __forceinline__ double& operator[](int i) const { return m.at(i, j); }
};
struct SpeciesPerturbed {
	Vec y;
	int j;
	double yj;
	__device__ __forceinline__ double operator[](int i) const { return (i == j) ? yj : y[i]; }
};
struct CellParameters {
	const double* base;
	double ov[CP_NUM_OVERRIDES > 0 ? CP_NUM_OVERRIDES : 1];
	__device__ __forceinline__ double operator[](int k) const
	{
		CP_PARAM_OVERRIDE_BODY
		return base[k];
	}
};

struct ThreadBdf {
	double* vbase; // shared memory, already offset by the thread index
	Mat A, SJ;
	const double* constant_species;
	const double* non_sampled;
	CellParameters params;
	double reltol, abstol, hmin;

	double tn, h, hprime, hscale, eta, etamax, hu;
	double gamma, gammap, gamrat, rl1, crate, delp, acnrm, saved_tq5, saved_t;
	double tau[6];
	int q, qprime, L, qwait, nst, nstlp, nstlj, nflag, ncf, nef;
	bool nls_jcur;
	int nfe, nsetups, nje;

	__device__ __forceinline__ Vec vec(int k) const { return Vec{ vbase + (long long)k * N * BS }; }
	__device__ __forceinline__ Vec zn(int j) const { return vec(V_ZN0 + j); }

	__device__ __forceinline__ double wrms(const Vec x) const
	{
		const Vec w = vec(V_EWT);
		double s = 0.0;
		for (int i = 0; i < N; i++) {
			const double p = x[i] * w[i];
			s += p * p;
		}
		return sqrt(s / N);
	}
	template <class OUT>
	__device__ __forceinline__ void rhs(const Vec ysrc, OUT out)
	{
		generated_derivative(out, ysrc, constant_species, params, non_sampled); // Cell::solver_rhs_fn, Cell.cpp:423-433
		nfe++;
	}
	__device__ __forceinline__ void set_ewt()
	{
		const Vec z0 = zn(0), w = vec(V_EWT);
		for (int i = 0; i < N; i++) w[i] = 1.0 / (reltol * fabs(z0[i]) + abstol);
	}

	// CVodeCreate zero state + CVodeReInit(0, y0) + the first-call block of CVode (cvode.c:586-665, 1068-1155); y0 is vec(V_Y)
	__device__ __forceinline__ bool start(double tout)
	{
		for (int j = 0; j < 6; j++) {
			tau[j] = 0.0;
			const Vec z = zn(j);
			for (int i = 0; i < N; i++) z[i] = 0.0;
		}
		gammap = 0.0; crate = 1.0; delp = 0.0; acnrm = 0.0; saved_tq5 = 0.0;
		eta = hu = gamma = gamrat = rl1 = 0.0;
		nls_jcur = false; nstlj = 0; nfe = 0; nsetups = 0; nje = 0;
		tn = 0.0; q = 1; L = 2; qwait = 2; etamax = 10000.0; nst = 0; nstlp = 0; qprime = 1;
		const Vec y = vec(V_Y), z0 = zn(0), z1 = zn(1), tempv = vec(V_TEMPV), w = vec(V_EWT);
		for (int i = 0; i < N; i++) z0[i] = y[i];
		set_ewt();
		rhs(z0, z1);
		// cvHin (cvode.c:1884-1984), no tstop
		const double tdiff = tout - tn;
		if (tdiff == 0.0) return false;
		const double sign = (tdiff > 0.0) ? 1.0 : -1.0;
		const double tdist = fabs(tdiff);
		const double tround = UROUND * fmax(fabs(tn), fabs(tout));
		if (tdist < 2.0 * tround) return false;
		const double hlb = 100.0 * tround;
		double hub_inv = -INFINITY;
		for (int i = 0; i < N; i++) {
			double t2 = fabs(z0[i]);
			double t1 = 1.0 / w[i];
			t1 = 0.1 * t2 + t1;
			t2 = fabs(z1[i]);
			t1 = t2 / t1;
			hub_inv = (t1 > hub_inv) ? t1 : hub_inv;
		}
		double hub = 0.1 * tdist;
		if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;
		double hg = sqrt(hlb * hub);
		if (hub < hlb) {
			h = (sign < 0.0) ? -hg : hg;
		} else {
			double hnew = hg;
#pragma unroll 1
			for (int count1 = 1; count1 <= 4; count1++) {
				const double hgs = hg * sign;
				for (int i = 0; i < N; i++) y[i] = hgs * z1[i] + z0[i];
				rhs(y, tempv);
				const double c = 1.0 / hgs;
				for (int i = 0; i < N; i++) tempv[i] = c * (tempv[i] - z1[i]);
				const double yddnrm = wrms(tempv);
				hnew = (yddnrm * hub * hub > 2.0) ? sqrt(2.0 / yddnrm) : sqrt(hg * hub);
				if (count1 == 4) break;
				const double hrat = hnew / hg;
				if ((hrat > 0.5) && (hrat < 2.0)) break;
				if ((count1 > 1) && (hrat > 2.0)) {
					hnew = hg;
					break;
				}
				hg = hnew;
			}
			double h0 = 0.5 * hnew;
			if (h0 < hlb) h0 = hlb;
			if (h0 > hub) h0 = hub;
			if (sign < 0.0) h0 = -h0;
			h = h0;
		}
		if (fabs(h) < hmin) h *= hmin / fabs(h);
		hscale = h;
		hprime = h;
		for (int i = 0; i < N; i++) z1[i] *= h;
		return true;
	}

	__device__ __forceinline__ void rescale()
	{
		double c = eta;
		for (int j = 1; j <= q; j++) {
			const Vec z = zn(j);
			for (int i = 0; i < N; i++) z[i] *= c;
			c = eta * c;
		}
		h = hscale * eta;
		hscale = h;
	}

	// cvAdjustOrder / cvIncreaseBDF / cvDecreaseBDF (cvode.c:2213-2374)
	__device__ __forceinline__ void adjust_order(int deltaq)
	{
		if ((q == 2) && (deltaq != 1)) return;
		double ll[6];
		for (int i = 0; i < 6; i++) ll[i] = 0.0;
		ll[2] = 1.0;
		if (deltaq == 1) {
			double alpha1 = 1.0, prod = 1.0, xiold = 1.0, alpha0 = -1.0, hsum = hscale;
			for (int j = 1; j < q; j++) {
				hsum += tau[j + 1];
				const double xi = hsum / hscale;
				prod *= xi;
				alpha0 -= 1.0 / (j + 1);
				alpha1 += 1.0 / xi;
				for (int i = j + 2; i >= 2; i--) ll[i] = ll[i] * xiold + ll[i - 1];
				xiold = xi;
			}
			const double A1 = (-alpha0 - alpha1) / prod;
			const Vec z5 = zn(5), zL = zn(L);
			for (int i = 0; i < N; i++) {
				const double znL = A1 * z5[i];
				zL[i] = znL;
				for (int j = 2; j <= q; j++) zn(j)[i] += ll[j] * znL;
			}
		} else if (deltaq == -1) {
			double hsum = 0.0;
			for (int j = 1; j <= q - 2; j++) {
				hsum += tau[j];
				const double xi = hsum / hscale;
				for (int i = j + 2; i >= 2; i--) ll[i] = ll[i] * xi + ll[i - 1];
			}
			if (q > 2) {
				const Vec zq = zn(q);
				for (int i = 0; i < N; i++) {
					const double znq = zq[i];
					for (int j = 2; j < q; j++) zn(j)[i] += (-ll[j]) * znq;
				}
			}
		}
	}

	// CVode loop head + cvStep head (cvode.c:1294-1337, 2094-2102). False on CV_TOO_MUCH_ACC.
	__device__ __forceinline__ bool begin_step()
	{
		if (nst > 0) set_ewt();
		if (UROUND * wrms(zn(0)) > 1.0) return false;
		saved_t = tn;
		ncf = 0;
		nef = 0;
		nflag = T_FIRST_CALL;
		if ((nst > 0) && (hprime != h)) {
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

	__device__ __forceinline__ void restore()
	{
		tn = saved_t;
		for (int i = 0; i < N; i++) {
			double z[6];
			for (int j = 0; j <= q; j++) z[j] = zn(j)[i];
			for (int k = 1; k <= q; k++)
				for (int j = q; j >= k; j--) z[j - 1] = z[j - 1] - z[j];
			for (int j = 0; j < q; j++) zn(j)[i] = z[j];
		}
	}

	// ODESolverCVODE::DifferenceQuotientJacobian (ODESolverCVODE.cpp:496-537) into A
	__device__ __forceinline__ void dq_jacobian()
	{
		const Vec y = vec(V_Y), fy = vec(V_FTEMP), w = vec(V_EWT);
		const double srur = sqrt(UROUND);
		const double fnorm = wrms(fy);
		const double minInc = (fnorm != 0.0) ? (1000.0 * fabs(h) * UROUND * N * fnorm) : 1.0;
#pragma unroll 1
		for (int j = 0; j < N; j++) {
			const double inc = fmax(srur * fabs(y[j]), minInc / w[j]);
			SpeciesPerturbed sp{ y, j, y[j] + inc };
			MatColumn col{ A, j };
			generated_derivative(col, sp, constant_species, params, non_sampled);
			const double inc_inv = 1.0 / inc;
			for (int i = 0; i < N; i++) col[i] = inc_inv * (col[i] - fy[i]);
		}
	}

	// PartialPivLUExtended::compute_optimized (EigenPartialPivLUSomewhatSparse.h:38-105); pivots packed 6 bits each
	__device__ __forceinline__ void lu_factor(int (&piv)[N])
	{
#pragma unroll 1
		for (int k = 0; k < N; k++) {
			double best = -1.0;
			int bi = k;
			for (int i = k; i < N; i++) {
				const double v = fabs(A.at(i, k));
				if (v > best) {
					best = v;
					bi = i;
				}
			}
			piv[k] = bi;
			if (best != 0.0) {
				if (bi != k) {
					for (int j = 0; j < N; j++) {
						const double tmp = A.at(k, j);
						A.at(k, j) = A.at(bi, j);
						A.at(bi, j) = tmp;
					}
				}
				const double inv_coeff = 1.0 / A.at(k, k);
				for (int i = k + 1; i < N; i++) A.at(i, k) *= inv_coeff;
			}
			for (int j = k + 1; j < N; j++) {
				const double a_kj = A.at(k, j);
				if (a_kj != 0.0) {
					for (int i = k + 1; i < N; i++) A.at(i, j) -= a_kj * A.at(i, k);
				}
			}
		}
	}

	__device__ __forceinline__ void lu_solve(const Vec b, const int (&piv)[N])
	{
		for (int k = 0; k < N; k++) {
			const int p = piv[k];
			if (p != k) {
				const double tmp = b[k];
				b[k] = b[p];
				b[p] = tmp;
			}
		}
		for (int k = 0; k < N; k++) {
			const double xk = b[k];
			for (int i = k + 1; i < N; i++) b[i] -= xk * A.at(i, k);
		}
		for (int k = N - 1; k >= 0; k--) {
			b[k] /= A.at(k, k);
			const double xk = b[k];
			for (int i = 0; i < k; i++) b[i] -= xk * A.at(i, k);
		}
	}

	__device__ __forceinline__ void residual()
	{
		const Vec y = vec(V_Y), z0 = zn(0), z1 = zn(1), acor = vec(V_ACOR), f = vec(V_FTEMP), delta = vec(V_DELTA);
		for (int i = 0; i < N; i++) y[i] = z0[i] + acor[i];
		rhs(y, f);
		for (int i = 0; i < N; i++) {
			double r = rl1 * z1[i] + acor[i];
			r += -gamma * f[i];
			delta[i] = r;
		}
	}

	__device__ __forceinline__ double root(double base, double inv_k) const
	{
		if (base <= 0.0) return 0.0;
		return pow(base, inv_k);
	}

	// One pass of cvStep's attempt loop (structure of BdfThread::attempt in bdf_thread.cuh); `mask` = lanes in this call.
	__device__ __forceinline__ int attempt(unsigned mask, int (&piv)[N])
	{
		double l[6], tq[6];
		for (int i = 0; i < 6; i++) {
			l[i] = 0.0;
			tq[i] = 0.0;
		}
		// ---- cvPredict ----
		tn += h;
		for (int i = 0; i < N; i++) {
			double z[6];
			for (int j = 0; j <= q; j++) z[j] = zn(j)[i];
			for (int k = 1; k <= q; k++)
				for (int j = q; j >= k; j--) z[j - 1] += z[j];
			for (int j = 0; j < q; j++) zn(j)[i] = z[j];
		}
		// ---- cvSetBDF + cvSetTqBDF (cvode.c:2611-2686) ----
		{
			double alpha0, alpha0_hat, xi_inv, xistar_inv, hsum;
			l[0] = l[1] = xi_inv = xistar_inv = 1.0;
			for (int i = 2; i <= q; i++) l[i] = 0.0;
			alpha0 = alpha0_hat = -1.0;
			hsum = h;
			if (q > 1) {
				for (int j = 2; j < q; j++) {
					hsum += tau[j - 1];
					xi_inv = h / hsum;
					alpha0 -= 1.0 / j;
					for (int i = j; i >= 1; i--) l[i] += l[i - 1] * xi_inv;
				}
				alpha0 -= 1.0 / q;
				xistar_inv = -l[1] - alpha0;
				hsum += tau[q - 1];
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
				hsum += tau[q];
				xi_inv = h / hsum;
				const double A5 = alpha0 - (1.0 / (q + 1));
				const double A6 = alpha0_hat - xi_inv;
				const double Cppinv = (1.0 - A6 + A5) / A2;
				tq[3] = fabs(Cppinv / (xi_inv * (q + 2) * A5));
			}
			tq[4] = 0.1 / tq[2];
			rl1 = 1.0 / l[1];
			gamma = h * rl1;
			if (nst == 0) gammap = gamma;
			gamrat = (nst > 0) ? gamma / gammap : 1.0;
		}

		// ---- cvNls + Newton: both loops warp-uniform via votes ----
		const Vec acor = vec(V_ACOR), delta = vec(V_DELTA);
		int nls_ret = 1;
		{
			int convfail = ((nflag == T_FIRST_CALL) || (nflag == T_PREV_ERR_FAIL)) ? T_NO_FAILURES : T_FAIL_OTHER;
			bool callSetup = (nflag == T_PREV_CONV_FAIL) || (nflag == T_PREV_ERR_FAIL) || (nst == 0) || (nst >= nstlp + 20) ||
			                 (fabs(gamrat - 1.0) > 0.3);
			for (int i = 0; i < N; i++) acor[i] = 0.0;
			const double tol = tq[4];
			bool jbad = false, need_pass = true;
#pragma unroll 1
			for (;;) {
				if (!__any_sync(mask, need_pass)) break;
				if (need_pass) residual();
				// linear setup, converged over the lanes that need one
				const bool do_setup = need_pass && callSetup;
				if (__any_sync(mask, do_setup)) {
					if (do_setup) {
						if (jbad) convfail = T_FAIL_BAD_J;
						const double dgamma = fabs((gamma / gammap) - 1.0);
						const bool jb = (nst == 0) || (nst > nstlj + 50) || ((convfail == T_FAIL_BAD_J) && (dgamma < 0.2)) || (convfail == T_FAIL_OTHER);
						if (jb) {
							dq_jacobian();
							for (int e = 0; e < N * N; e++) SJ.lin(e) = A.lin(e);
							nstlj = nst;
							nje++;
						} else {
							for (int e = 0; e < N * N; e++) A.lin(e) = SJ.lin(e);
						}
						for (int e = 0; e < N * N; e++) A.lin(e) *= -gamma;
						for (int i = 0; i < N; i++) A.at(i, i) += 1.0;
						lu_factor(piv);
						nsetups++;
						nls_jcur = jb;
						gamrat = 1.0;
						gammap = gamma;
						crate = 1.0;
						nstlp = nst;
					}
				}
				bool iter = need_pass;
#pragma unroll 1
				for (int m = 0; m < 3; m++) {
					if (!__any_sync(mask, iter)) break;
					if (iter) {
						for (int i = 0; i < N; i++) delta[i] = -delta[i];
						lu_solve(delta, piv);
						if (gamrat != 1.0) {
							const double sc = 2.0 / (1.0 + gamrat);
							for (int i = 0; i < N; i++) delta[i] *= sc;
						}
						for (int i = 0; i < N; i++) acor[i] += delta[i];
						const double del = wrms(delta);
						if (m > 0) crate = fmax(0.3 * crate, del / delp);
						const double dcon = del * fmin(1.0, crate) / tol;
						if (dcon <= 1.0) {
							acnrm = (m == 0) ? del : wrms(acor);
							nls_jcur = false;
							nls_ret = 0;
							iter = false;
						} else if ((m >= 1) && (del > 2.0 * delp)) {
							iter = false;
						} else {
							delp = del;
							if (m + 1 >= 3) iter = false;
							else residual();
						}
					}
				}
				if (need_pass) {
					if (nls_ret == 0 || nls_jcur) {
						need_pass = false;
					} else {
						callSetup = true;
						jbad = true;
						for (int i = 0; i < N; i++) acor[i] = 0.0;
					}
				}
			}
		}

		int result;
		if (nls_ret != 0) {
			// ---- cvHandleNFlag ----
			restore();
			ncf++;
			etamax = 1.0;
			if ((fabs(h) <= hmin * 1.000001) || (ncf == 10)) {
				result = T_FAILED;
			} else {
				eta = fmax(0.25, hmin / fabs(h));
				nflag = T_PREV_CONV_FAIL;
				rescale();
				result = T_RETRY;
			}
		} else {
			const double dsm = acnrm * tq[2];
			if (!(dsm <= 1.0)) {
				// ---- cvDoErrorTest, failure ----
				nef++;
				nflag = T_PREV_ERR_FAIL;
				restore();
				if ((fabs(h) <= hmin * 1.000001) || (nef == 7)) {
					result = T_FAILED;
				} else {
					result = T_RETRY;
					etamax = 1.0;
					if (nef <= 3) {
						eta = 1.0 / (root(6.0 * dsm, 1.0 / L) + 0.000001);
						eta = fmax(0.1, fmax(eta, hmin / fabs(h)));
						if (nef >= 2) eta = fmin(eta, 0.2);
						rescale();
					} else if (q > 1) {
						eta = fmax(0.1, hmin / fabs(h));
						adjust_order(-1);
						L = q;
						q--;
						qwait = L;
						rescale();
					} else {
						eta = fmax(0.1, hmin / fabs(h));
						h *= eta;
						hscale = h;
						qwait = 10;
						const Vec tempv = vec(V_TEMPV), z1 = zn(1);
						rhs(zn(0), tempv);
						for (int i = 0; i < N; i++) z1[i] = h * tempv[i];
					}
				}
			} else {
				result = T_DONE;
				// ---- cvCompleteStep ----
				nst++;
				hu = h;
				for (int i = q; i >= 2; i--) tau[i] = tau[i - 1];
				if ((q == 1) && (nst > 1)) tau[2] = tau[1];
				tau[1] = h;
				for (int i = 0; i < N; i++) {
					const double ac = acor[i];
					for (int j = 0; j <= q; j++) zn(j)[i] += l[j] * ac;
				}
				qwait--;
				if ((qwait == 1) && (q != 5)) {
					const Vec z5 = zn(5);
					for (int i = 0; i < N; i++) z5[i] = acor[i];
					saved_tq5 = tq[5];
				}
				// ---- cvPrepareNextStep ----
				if (etamax == 1.0) {
					qwait = (qwait > 2) ? qwait : 2;
					qprime = q;
					hprime = h;
					eta = 1.0;
				} else {
					const double etaq = 1.0 / (root(6.0 * dsm, 1.0 / L) + 0.000001);
					eta = etaq;
					qprime = q;
					if (qwait == 0) {
						qwait = 2;
						double etaqm1 = 0.0;
						if (q > 1) {
							const double ddn = wrms(zn(q)) * tq[1];
							etaqm1 = 1.0 / (root(6.0 * ddn, 1.0 / q) + 0.000001);
						}
						double etaqp1 = 0.0;
						if (q != 5) {
							if (saved_tq5 != 0.0) {
								const double base = h / tau[2];
								double pw = 1.0;
								for (int i = 1; i <= L; i++) pw *= base;
								const double cquot = (tq[5] / saved_tq5) * pw;
								const Vec tempv = vec(V_TEMPV), z5 = zn(5);
								for (int i = 0; i < N; i++) tempv[i] = -cquot * z5[i] + acor[i];
								const double dup = wrms(tempv) * tq[3];
								etaqp1 = 1.0 / (root(10.0 * dup, 1.0 / (L + 1)) + 0.000001);
							}
						}
						const double etam = fmax(etaqm1, fmax(etaq, etaqp1));
						if (etam < 1.5) {
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
							const Vec z5 = zn(5);
							for (int i = 0; i < N; i++) z5[i] = acor[i];
						}
					}
					if (eta < 1.5) {
						eta = 1.0;
						hprime = h;
					} else {
						eta = fmin(eta, etamax);
						hprime = h * eta;
					}
				}
				etamax = 10.0;
				for (int i = 0; i < N; i++) acor[i] *= tq[2];
			}
		}
		return result;
	}

	// CVodeGetDky(t, 0) for one component, cvode.c:1467-1524
	__device__ __forceinline__ bool dky_ok(double t) const
	{
		double tfuzz = 100.0 * UROUND * (fabs(tn) + fabs(hu));
		if (hu < 0.0) tfuzz = -tfuzz;
		const double tp = tn - hu - tfuzz, tn1 = tn + tfuzz;
		return !((t - tp) * (t - tn1) > 0.0);
	}
	__device__ __forceinline__ double dky_component(double t, int i) const
	{
		const double s = (t - tn) / h;
		double acc = 0.0;
		for (int j = q; j >= 0; j--) {
			double c = 1.0;
			for (int k = 0; k < j; k++) c *= s;
			acc = (j == q) ? c * zn(j)[i] : acc + c * zn(j)[i];
		}
		return acc;
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

// grid = (ceil(num_cells / BS), C); thread = one cell of chain blockIdx.y
__global__ void __launch_bounds__(CP_THREADS_PER_BLOCK) cellpop_thread_kernel(const CpArgs a, double* __restrict__ scratch)
{
	extern __shared__ double smem_d[];
	const int tid = threadIdx.x, c = blockIdx.y;
	const int cell = blockIdx.x * BS + tid;
	const bool valid = cell < a.num_cells;
	const long long nthreads = (long long)gridDim.x * gridDim.y * BS;
	const long long gtid = ((long long)blockIdx.y * gridDim.x + blockIdx.x) * BS + tid;

	ThreadBdf S;
	S.vbase = smem_d + tid;
	S.A = Mat{ scratch + gtid, nthreads };
	S.SJ = Mat{ scratch + (long long)N * N * nthreads + gtid, nthreads };
	S.constant_species = a.constant_species;
	S.non_sampled = a.non_sampled;
	S.reltol = a.rel_tol;
	S.abstol = a.abs_tol;
	S.hmin = a.min_dt;
	const double* tv = a.transformed + (long long)c * a.nvar;
	S.params.base = tv;
	CP_OVERRIDE_INIT

	// ---- K0: Cell::Initialize (Cell.cpp:150-191) ----
	const Vec y = S.vec(V_Y);
	for (int i = 0; i < N; i++) y[i] = a.initial_conditions[i];
	const long long gcell = (long long)a.cell_offset + (valid ? cell : 0);
	for (int d = 0; d < a.D; d++) {
		double v = cellpop_variability_value(a, tv, c, gcell, d);
		if (a.var_negate[d]) v = -v;
		if (a.var_is_ic[d]) {
			double x = y[a.var_slot[d]];
			apply_variability(x, v, a.var_apply[d]);
			y[a.var_slot[d]] = x;
		} else {
#pragma unroll
			for (int s = 0; s < CP_NUM_OVERRIDES; s++)
				if (a.var_slot[d] == s) apply_variability(S.params.ov[s], v, a.var_apply[d]);
		}
	}

	// ---- Cell::Simulate + ODESolver::SolveReturnSolution + ODESolverCVODE::Solve ----
	const double creation_time = (a.entry_time_ix >= 0) ? tv[a.entry_time_ix] : a.entry_time_fixed;
	const int T = a.T;
	double* out = a.cell_values + ((long long)c * T) * a.num_cells + cell;
	const double nan = __longlong_as_double(0x7ff8000000000000ll);
	bool done = !valid, ok = true;
	int steps = 0, tpi = 0;
	if (!done) {
		while (tpi < T && (a.timepoints[tpi] - creation_time) < DBL_EPSILON) {
			const double cell_time = a.timepoints[tpi] - creation_time;
			double sv = 0.0;
			for (int k = 0; k < a.num_obs_species; k++) sv += y[a.obs_species[k]];
			out[(long long)tpi * a.num_cells] = (cell_time < 0.0) ? nan : sv;
			tpi++;
		}
	}
	const double end_time = a.sim_end_time - creation_time;
	if (end_time < DBL_EPSILON) done = true;
	if (!done) {
		if (!S.start(end_time)) {
			ok = false;
			done = true;
		}
	}
	int piv[N];
	for (int i = 0; i < N; i++) piv[i] = i;
	bool newstep = true;
#pragma unroll 1
	for (;;) {
		if (__syncthreads_or(done ? 0 : 1) == 0) break;
		if (!done && newstep) {
			newstep = false;
			if (!S.begin_step()) {
				ok = false;
				done = true;
			}
		}
		const bool go = !done;
		const unsigned mask = __ballot_sync(FULL, go);
		if (go) {
			const int r = S.attempt(mask, piv);
			if (r == T_FAILED) {
				ok = false;
				done = true;
			} else if (r == T_DONE) {
				steps++;
				const double tret = S.tn;
				while (tpi < T && tret >= (a.timepoints[tpi] - creation_time)) {
					const double tq = a.timepoints[tpi] - creation_time;
					if (!S.dky_ok(tq)) {
						ok = false;
						done = true;
						break;
					}
					double sv = 0.0;
					for (int k = 0; k < a.num_obs_species; k++) sv += S.dky_component(tq, a.obs_species[k]);
					out[(long long)tpi * a.num_cells] = sv;
					tpi++;
				}
				if (ok) {
					if (tret >= end_time) done = true;
					else if (steps == a.max_steps) {
						ok = false;
						done = true;
					}
				}
				newstep = true;
			}
		}
	}
	if (valid) {
		if (!ok)
			for (int k = tpi; k < T; k++) out[(long long)k * a.num_cells] = nan;
		a.cell_status[(long long)c * a.num_cells + cell] = ok ? 1 : 0;
		if (a.cell_steps)
			a.cell_steps[(long long)c * a.num_cells + cell] = (a.debug_report == 1) ? S.nfe : (a.debug_report == 2) ? S.nsetups : (a.debug_report == 3) ? S.nje : steps;
	}
}

} // namespace cellpop_thread

// scratch doubles needed for a launch of num_chains x num_cells
extern "C" long long cellpop_thread_scratch_doubles(int num_chains, int num_cells)
{
	const long long blocks = ((long long)num_cells + CP_THREADS_PER_BLOCK - 1) / CP_THREADS_PER_BLOCK;
	return 2ll * CP_N * CP_N * blocks * num_chains * CP_THREADS_PER_BLOCK;
}

extern "C" int cellpop_thread_launch(const CpArgs* args, double* scratch, void* stream)
{
	const size_t smem = sizeof(double) * cellpop_thread::NUM_VECTORS * CP_N * CP_THREADS_PER_BLOCK;
	static bool attr_set = false;
	if (!attr_set && smem > 48 * 1024) {
		cudaError_t e = cudaFuncSetAttribute(cellpop_thread::cellpop_thread_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) return (int)e;
		attr_set = true;
	}
	dim3 grid((args->num_cells + CP_THREADS_PER_BLOCK - 1) / CP_THREADS_PER_BLOCK, args->num_chains);
	cellpop_thread::cellpop_thread_kernel<<<grid, CP_THREADS_PER_BLOCK, smem, (cudaStream_t)stream>>>(*args, scratch);
	return (int)cudaGetLastError();
}
