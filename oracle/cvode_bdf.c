/* TEST INFRASTRUCTURE ONLY -- see cvode_bdf.h. Each function cites the reference lines it restates.
 * "cvode.c" = /root/reference/dependencies/cvode-5.3.0/src/cvode/cvode.c, etc. */
#include "cvode_bdf.h"

#include <float.h>
#include <math.h>
#include <string.h>

/* cvode.c:142-172 */
#define FUZZ_FACTOR 100.0
#define HLB_FACTOR 100.0
#define HUB_FACTOR 0.1
#define H_BIAS 0.5
#define MAX_ITERS 4
#define CORTES 0.1
#define THRESH 1.5
#define ETAMX1 10000.0
#define ETAMX2 10.0
#define ETAMX3 10.0
#define ETAMXF 0.2
#define ETAMIN 0.1
#define ETACF 0.25
#define ADDON 0.000001
#define BIAS1 6.0
#define BIAS2 6.0
#define BIAS3 10.0
#define ONEPSM 1.000001
#define SMALL_NST 10
#define MXNCF 10
#define MXNEF 7
#define MXNEF1 3
#define SMALL_NEF 2
#define LONG_WAIT 10
#define DGMAX 0.3
#define MSBP 20
/* cvode_nls.c:29-31 */
#define NLS_MAXCOR 3
#define CRDOWN 0.3
#define RDIV 2.0
/* cvode_ls_impl.h:40-42 */
#define CVLS_MSBJ 50
#define CVLS_DGMAX 0.2

#define UROUND DBL_EPSILON

/* control flags, cvode.c */
#define FIRST_CALL 101
#define PREV_CONV_FAIL 102
#define PREV_ERR_FAIL 103
#define DO_ERROR_TEST 2
#define PREDICT_AGAIN 3
#define TRY_AGAIN 5
#define CONV_RECVR 902 /* SUN_NLS_CONV_RECVR */
#define CV_NO_FAILURES 0
#define CV_FAIL_BAD_J 1
#define CV_FAIL_OTHER 2

static double dmax(double a, double b) { return a > b ? a : b; }
static double dmin(double a, double b) { return a < b ? a : b; }

/* sundials_math.c SUNRpowerR / SUNRpowerI */
static double rpower_r(double base, double exponent)
{
	if (base <= 0.0) return 0.0;
	return pow(base, exponent);
}
static double rpower_i(double base, int exponent)
{
	double prod = 1.0;
	int i, expt = exponent < 0 ? -exponent : exponent;
	for (i = 1; i <= expt; i++) prod *= base;
	if (exponent < 0) prod = 1.0 / prod;
	return prod;
}

/* N_VWrmsNorm_Eigen, nvector_serial_eigen.cpp:386-396 */
static double wrms(int N, const double* x, const double* w)
{
	double sum = 0.0;
	for (int i = 0; i < N; i++) {
		double p = x[i] * w[i];
		sum += p * p;
	}
	return sqrt(sum / N);
}

/* cvEwtSetSV, cvode.c:4268-4295 (atol > 0 always here, so no N_VMin test) */
static int ewt_set(const bdf_mem* m, const double* ycur, double* weight)
{
	for (int i = 0; i < m->N; i++) {
		double v = m->reltol * fabs(ycur[i]) + m->abstol[i];
		if (v <= 0.0) return -1;
		weight[i] = 1.0 / v;
	}
	return 0;
}

void bdf_create(bdf_mem* m, int N, bdf_rhs_fn f, bdf_jac_fn jac, void* user)
{
	/* CVodeCreate (cvode.c:320-430) zero-fills and sets defaults; CVodeInit */
	memset(m, 0, sizeof(*m));
	m->N = N;
	m->f = f;
	m->jac = jac;
	m->user = user;
	m->hmin = 0.0;
	m->hmax_inv = 0.0;
	m->q = 1;
	m->L = 2;
	m->qwait = 2;
	m->etamax = ETAMX1;
	m->tolsf = 1.0;
}

void bdf_set_tolerances(bdf_mem* m, double reltol, const double* abstol)
{
	m->reltol = reltol;
	for (int i = 0; i < m->N; i++) m->abstol[i] = abstol[i];
}

/* CVodeReInit, cvode.c:586-665. Note what is NOT reset: tstopset/tstop, tau[], saved_tq5,
 * gammap, crate, savedJ, the linear-solver counters (those are reset by cvLsInitialize at the
 * next first step, cvode_ls.c:1379) */
void bdf_reinit(bdf_mem* m, double t0, const double* y0)
{
	m->tn = t0;
	m->q = 1;
	m->L = 2;
	m->qwait = m->L;
	m->etamax = ETAMX1;
	m->qu = 0;
	m->hu = 0.0;
	m->tolsf = 1.0;
	for (int i = 0; i < m->N; i++) m->zn[0][i] = y0[i];
	m->nst = 0;
	m->nfe = 0;
	m->ncfn = 0;
	m->netf = 0;
	m->nni = 0;
	m->nsetups = 0;
	m->nstlp = 0;
}

/* CVodeSetStopTime, cvode_io.c:384-411 */
void bdf_set_stop_time(bdf_mem* m, double tstop)
{
	if (m->nst > 0) {
		if ((tstop - m->tn) * m->h < 0.0) return;
	}
	m->tstop = tstop;
	m->tstopset = 1;
}

/* CVodeGetDky with k = 0, cvode.c:1467-1524; sum as N_VLinearCombination_Eigen (nvector_serial_eigen.cpp:496-543) */
int bdf_get_dky(const bdf_mem* m, double t, double* dky)
{
	double tfuzz = FUZZ_FACTOR * UROUND * (fabs(m->tn) + fabs(m->hu));
	if (m->hu < 0.0) tfuzz = -tfuzz;
	double tp = m->tn - m->hu - tfuzz;
	double tn1 = m->tn + tfuzz;
	if ((t - tp) * (t - tn1) > 0.0) return BDF_BAD_T;

	double s = (t - m->tn) / m->h;
	double cvals[BDF_LMAX];
	int nvec = 0;
	for (int j = m->q; j >= 0; j--) {
		double c = 1.0;
		for (int i = 0; i < j; i++) c *= s;
		cvals[nvec++] = c;
	}
	/* X[nvec] = zn[q], zn[q-1], ..., zn[0] */
	if (nvec == 1) { /* cannot happen (q >= 1) */
		for (int i = 0; i < m->N; i++) dky[i] = cvals[0] * m->zn[m->q][i];
		return BDF_SUCCESS;
	}
	if (nvec == 2) {
		/* N_VLinearSum(c0, zn[1], c1 = 1, zn[0], z): b == ONE, z != y -> VLin1: z = a*x + y */
		for (int i = 0; i < m->N; i++) dky[i] = cvals[0] * m->zn[1][i] + m->zn[0][i];
		return BDF_SUCCESS;
	}
	for (int i = 0; i < m->N; i++) dky[i] = cvals[0] * m->zn[m->q][i];
	for (int k = 1; k < nvec; k++) {
		const double* x = m->zn[m->q - k];
		for (int i = 0; i < m->N; i++) dky[i] += cvals[k] * x[i];
	}
	return BDF_SUCCESS;
}

/* ---- linear algebra ---- */

/* ODESolverCVODE::DifferenceQuotientJacobian, ODESolverCVODE.cpp:496-537 */
static int dq_jacobian(bdf_mem* m, double t, const double* y, const double* ydot, double* J)
{
	const int N = m->N;
	double ycopy[BDF_NMAX], work[BDF_NMAX];
	memcpy(ycopy, y, sizeof(double) * N);
	double srur = sqrt(UROUND);
	double fnorm = wrms(N, ydot, m->ewt);
	double minInc = (fnorm != 0.0) ? (1000.0 * fabs(m->h) * UROUND * N * fnorm) : 1.0;
	for (int j = 0; j < N; j++) {
		double inc = dmax(srur * fabs(y[j]), minInc / m->ewt[j]);
		ycopy[j] += inc;
		int rv = m->f(t, ycopy, work, m->user);
		m->nfeDQ++;
		if (rv != 0) return -1;
		ycopy[j] = y[j];
		double inc_inv = 1.0 / inc;
		for (int i = 0; i < N; i++) J[i + j * N] = inc_inv * (work[i] - ydot[i]);
	}
	return 0;
}

/* SUNLinSolSetup_Dense_Eigen{2x2,3x3,}: sunlinsol_dense_eigen.cpp:95-145,
 * Eigen compute_inverse<.,.,3> (InverseImpl.h:124-170), PartialPivLUExtended::compute_optimized
 * (EigenPartialPivLUSomewhatSparse.h:38-105). A is column-major and overwritten. */
#define AE(i, j) A[(i) + (j) * N]
static void lin_setup(bdf_mem* m)
{
	const int N = m->N;
	double* A = m->A;
	if (N == 2) {
		double invdet = 1.0 / (AE(0, 0) * AE(1, 1) - AE(0, 1) * AE(1, 0));
		double i00 = AE(1, 1) * invdet, i01 = -AE(0, 1) * invdet, i10 = -AE(1, 0) * invdet, i11 = AE(0, 0) * invdet;
		AE(0, 0) = i00; AE(0, 1) = i01; AE(1, 0) = i10; AE(1, 1) = i11;
	} else if (N == 3) {
		double M[9];
		memcpy(M, A, sizeof(M));
#define ME(i, j) M[(i) + (j) * 3]
#define COF(i, j) (ME(((i) + 1) % 3, ((j) + 1) % 3) * ME(((i) + 2) % 3, ((j) + 2) % 3) - ME(((i) + 1) % 3, ((j) + 2) % 3) * ME(((i) + 2) % 3, ((j) + 1) % 3))
		double c0 = COF(0, 0), c1 = COF(1, 0), c2 = COF(2, 0);
		double det = c0 * ME(0, 0) + c1 * ME(1, 0) + c2 * ME(2, 0);
		double invdet = 1.0 / det;
		AE(0, 0) = c0 * invdet; AE(0, 1) = c1 * invdet; AE(0, 2) = c2 * invdet;
		AE(1, 0) = COF(0, 1) * invdet; AE(1, 1) = COF(1, 1) * invdet; AE(1, 2) = COF(2, 1) * invdet;
		AE(2, 0) = COF(0, 2) * invdet; AE(2, 1) = COF(1, 2) * invdet; AE(2, 2) = COF(2, 2) * invdet;
#undef COF
#undef ME
	} else {
		for (int k = 0; k < N; k++) {
			int piv = k;
			double big = fabs(AE(k, k));
			for (int i = k + 1; i < N; i++) {
				double v = fabs(AE(i, k));
				if (v > big) { big = v; piv = i; }
			}
			m->piv[k] = piv;
			if (big != 0.0) {
				if (piv != k) {
					for (int j = 0; j < N; j++) { double tmp = AE(k, j); AE(k, j) = AE(piv, j); AE(piv, j) = tmp; }
				}
				double inv_coeff = 1.0 / AE(k, k);
				for (int i = k + 1; i < N; i++) AE(i, k) *= inv_coeff;
			}
			for (int j = k + 1; j < N; j++) {
				double a_kj = AE(k, j);
				if (a_kj != 0.0) {
					for (int i = k + 1; i < N; i++) AE(i, j) -= a_kj * AE(i, k);
				}
			}
		}
	}
}

/* SUNLinSolSolve_Dense_Eigen{2x2,3x3,}: sunlinsol_dense_eigen.cpp:147-178; b overwritten with x */
static void lin_solve(const bdf_mem* m, double* b)
{
	const int N = m->N;
	const double* A = m->A;
	if (N == 2) {
		double x0 = AE(0, 0) * b[0] + AE(0, 1) * b[1];
		double x1 = AE(1, 0) * b[0] + AE(1, 1) * b[1];
		b[0] = x0; b[1] = x1;
	} else if (N == 3) {
		double x[3];
		for (int i = 0; i < 3; i++) x[i] = AE(i, 0) * b[0] + AE(i, 1) * b[1] + AE(i, 2) * b[2];
		b[0] = x[0]; b[1] = x[1]; b[2] = x[2];
	} else {
		for (int k = 0; k < N; k++) {
			int p = m->piv[k];
			if (p != k) { double tmp = b[k]; b[k] = b[p]; b[p] = tmp; }
		}
		for (int k = 0; k < N; k++) {
			double xk = b[k];
			for (int i = k + 1; i < N; i++) b[i] -= xk * AE(i, k);
		}
		for (int k = N - 1; k >= 0; k--) {
			b[k] /= AE(k, k);
			double xk = b[k];
			for (int i = 0; i < k; i++) b[i] -= xk * AE(i, k);
		}
	}
}
#undef AE

/* cvLsSetup + cvLsLinSys, cvode_ls.c:1415-1507, 1201-1286 */
static int ls_setup(bdf_mem* m, int convfail, const double* ypred, const double* fpred)
{
	const int N = m->N;
	double dgamma = fabs((m->gamma / m->gammap) - 1.0);
	int jbad = (m->nst == 0) || (m->nst > m->nstlj + CVLS_MSBJ) || ((convfail == CV_FAIL_BAD_J) && (dgamma < CVLS_DGMAX)) ||
	           (convfail == CV_FAIL_OTHER);
	if (!jbad) {
		m->jcur = 0;
		memcpy(m->A, m->savedJ, sizeof(double) * N * N);
	} else {
		m->jcur = 1;
		memset(m->A, 0, sizeof(double) * N * N);
		int rv = m->jac ? m->jac(m->tn, ypred, fpred, m->A, m->user) : dq_jacobian(m, m->tn, ypred, fpred, m->A);
		if (rv != 0) {
			m->nje++;
			m->nstlj = m->nst;
			return -1;
		}
		memcpy(m->savedJ, m->A, sizeof(double) * N * N);
	}
	/* SUNMatScaleAddI(-gamma, A), sunmatrix_dense_eigen.cpp:128-133 */
	for (int i = 0; i < N * N; i++) m->A[i] *= -m->gamma;
	for (int i = 0; i < N; i++) m->A[i + i * N] += 1.0;
	if (m->jcur) {
		m->nje++;
		m->nstlj = m->nst;
	}
	lin_setup(m);
	return 0;
}

/* cvNlsResidual, cvode_nls.c:281-315 */
static int nls_residual(bdf_mem* m, double* res)
{
	const int N = m->N;
	for (int i = 0; i < N; i++) m->y[i] = m->zn[0][i] + m->acor[i];
	int rv = m->f(m->tn, m->y, m->ftemp, m->user);
	m->nfe++;
	if (rv != 0) return BDF_RHSFUNC_FAIL;
	for (int i = 0; i < N; i++) res[i] = m->rl1 * m->zn[1][i] + m->acor[i];
	for (int i = 0; i < N; i++) res[i] += -m->gamma * m->ftemp[i];
	return 0;
}

/* cvNls (cvode.c:2701-2755) + SUNNonlinSolSolve_Newton (sunnonlinsol_newton.c:183-318) +
 * cvNlsLSetup / cvNlsLSolve / cvNlsConvTest (cvode_nls.c:180-279) + cvLsSolve scaling (cvode_ls.c:1603-1604) */
static int nls_solve(bdf_mem* m, int nflag)
{
	const int N = m->N;
	double delta[BDF_NMAX];
	int convfail = ((nflag == FIRST_CALL) || (nflag == PREV_ERR_FAIL)) ? CV_NO_FAILURES : CV_FAIL_OTHER;
	int callSetup = (nflag == PREV_CONV_FAIL) || (nflag == PREV_ERR_FAIL) || (m->nst == 0) || (m->nst >= m->nstlp + MSBP) ||
	                (fabs(m->gamrat - 1.0) > DGMAX);
	m->convfail = convfail;
	for (int i = 0; i < N; i++) m->acor[i] = 0.0;
	const double tol = m->tq[4];

	int jbad = 0;
	int retval;
	for (;;) {
		retval = nls_residual(m, delta);
		if (retval != 0) break;

		if (callSetup) {
			if (jbad) m->convfail = CV_FAIL_BAD_J;
			int rv = ls_setup(m, m->convfail, m->y, m->ftemp);
			m->nsetups++;
			m->nls_jcur = m->jcur;
			m->gamrat = 1.0;
			m->gammap = m->gamma;
			m->crate = 1.0;
			m->nstlp = m->nst;
			if (rv < 0) { retval = BDF_LSETUP_FAIL; break; }
		}

		m->nls_curiter = 0;
		for (;;) {
			m->nni++;
			for (int i = 0; i < N; i++) delta[i] = -delta[i];
			lin_solve(m, delta);
			if (m->gamrat != 1.0) {
				double sc = 2.0 / (1.0 + m->gamrat);
				for (int i = 0; i < N; i++) delta[i] *= sc;
			}
			for (int i = 0; i < N; i++) m->acor[i] += delta[i];

			/* cvNlsConvTest */
			double del = wrms(N, delta, m->ewt);
			int mi = m->nls_curiter;
			if (mi > 0) m->crate = dmax(CRDOWN * m->crate, del / m->delp);
			double dcon = del * dmin(1.0, m->crate) / tol;
			if (dcon <= 1.0) {
				m->acnrm = (mi == 0) ? del : wrms(N, m->acor, m->ewt);
				m->acnrmcur = 1;
				m->nls_jcur = 0;
				return 0;
			}
			if ((mi >= 1) && (del > RDIV * m->delp)) { retval = CONV_RECVR; break; }
			m->delp = del;

			m->nls_curiter++;
			if (m->nls_curiter >= NLS_MAXCOR) { retval = CONV_RECVR; break; }

			retval = nls_residual(m, delta);
			if (retval != 0) break;
		}

		if ((retval > 0) && !m->nls_jcur) {
			callSetup = 1;
			jbad = 1;
			for (int i = 0; i < N; i++) m->acor[i] = 0.0;
			continue;
		}
		break;
	}
	return retval;
}

/* ---- step-size / order machinery ---- */

/* cvRescale, cvode.c:2384-2400 */
static void rescale(bdf_mem* m)
{
	double cv[BDF_LMAX];
	cv[0] = m->eta;
	for (int j = 1; j <= m->q; j++) cv[j] = m->eta * cv[j - 1];
	for (int j = 1; j <= m->q; j++)
		for (int i = 0; i < m->N; i++) m->zn[j][i] *= cv[j - 1];
	m->h = m->hscale * m->eta;
	m->hscale = m->h;
}

/* cvIncreaseBDF, cvode.c:2310-2340 */
static void increase_bdf(bdf_mem* m)
{
	double alpha0, alpha1, prod, xi, xiold, hsum, A1;
	for (int i = 0; i <= BDF_QMAX; i++) m->l[i] = 0.0;
	m->l[2] = alpha1 = prod = xiold = 1.0;
	alpha0 = -1.0;
	hsum = m->hscale;
	if (m->q > 1) {
		for (int j = 1; j < m->q; j++) {
			hsum += m->tau[j + 1];
			xi = hsum / m->hscale;
			prod *= xi;
			alpha0 -= 1.0 / (j + 1);
			alpha1 += 1.0 / xi;
			for (int i = j + 2; i >= 2; i--) m->l[i] = m->l[i] * xiold + m->l[i - 1];
			xiold = xi;
		}
	}
	A1 = (-alpha0 - alpha1) / prod;
	for (int i = 0; i < m->N; i++) m->zn[m->L][i] = A1 * m->zn[BDF_QMAX][i];
	if (m->q > 1) {
		for (int j = 2; j <= m->q; j++)
			for (int i = 0; i < m->N; i++) m->zn[j][i] += m->l[j] * m->zn[m->L][i];
	}
}

/* cvDecreaseBDF, cvode.c:2352-2374 */
static void decrease_bdf(bdf_mem* m)
{
	double hsum, xi;
	for (int i = 0; i <= BDF_QMAX; i++) m->l[i] = 0.0;
	m->l[2] = 1.0;
	hsum = 0.0;
	for (int j = 1; j <= m->q - 2; j++) {
		hsum += m->tau[j];
		xi = hsum / m->hscale;
		for (int i = j + 2; i >= 2; i--) m->l[i] = m->l[i] * xi + m->l[i - 1];
	}
	if (m->q > 2) {
		for (int j = 2; j < m->q; j++)
			for (int i = 0; i < m->N; i++) m->zn[j][i] += (-m->l[j]) * m->zn[m->q][i];
	}
}

/* cvAdjustOrder, cvode.c:2213-2225 */
static void adjust_order(bdf_mem* m, int deltaq)
{
	if ((m->q == 2) && (deltaq != 1)) return;
	if (deltaq == 1) increase_bdf(m);
	else if (deltaq == -1) decrease_bdf(m);
}

/* cvAdjustParams, cvode.c:2192-2201 */
static void adjust_params(bdf_mem* m)
{
	if (m->qprime != m->q) {
		adjust_order(m, m->qprime - m->q);
		m->q = m->qprime;
		m->L = m->q + 1;
		m->qwait = m->L;
	}
	rescale(m);
}

/* cvPredict, cvode.c:2412-2425 */
static void predict(bdf_mem* m)
{
	m->tn += m->h;
	if (m->tstopset) {
		if ((m->tn - m->tstop) * m->h > 0.0) m->tn = m->tstop;
	}
	for (int k = 1; k <= m->q; k++)
		for (int j = m->q; j >= k; j--)
			for (int i = 0; i < m->N; i++) m->zn[j - 1][i] += m->zn[j][i];
}

/* cvRestore, cvode.c:2918-2927 */
static void restore(bdf_mem* m, double saved_t)
{
	m->tn = saved_t;
	for (int k = 1; k <= m->q; k++)
		for (int j = m->q; j >= k; j--)
			for (int i = 0; i < m->N; i++) m->zn[j - 1][i] = m->zn[j - 1][i] - m->zn[j][i];
}

/* cvSet + cvSetBDF + cvSetTqBDF, cvode.c:2445-2460, 2611-2686 */
static void set_bdf(bdf_mem* m)
{
	double alpha0, alpha0_hat, xi_inv, xistar_inv, hsum;
	const int q = m->q;
	m->l[0] = m->l[1] = xi_inv = xistar_inv = 1.0;
	for (int i = 2; i <= q; i++) m->l[i] = 0.0;
	alpha0 = alpha0_hat = -1.0;
	hsum = m->h;
	if (q > 1) {
		for (int j = 2; j < q; j++) {
			hsum += m->tau[j - 1];
			xi_inv = m->h / hsum;
			alpha0 -= 1.0 / j;
			for (int i = j; i >= 1; i--) m->l[i] += m->l[i - 1] * xi_inv;
		}
		alpha0 -= 1.0 / q;
		xistar_inv = -m->l[1] - alpha0;
		hsum += m->tau[q - 1];
		xi_inv = m->h / hsum;
		alpha0_hat = -m->l[1] - xi_inv;
		for (int i = q; i >= 1; i--) m->l[i] += m->l[i - 1] * xistar_inv;
	}
	/* cvSetTqBDF */
	{
		double A1, A2, A3, A4, A5, A6, C, Cpinv, Cppinv;
		A1 = 1.0 - alpha0_hat + alpha0;
		A2 = 1.0 + q * A1;
		m->tq[2] = fabs(A1 / (alpha0 * A2));
		m->tq[5] = fabs(A2 * xistar_inv / (m->l[q] * xi_inv));
		if (m->qwait == 1) {
			if (q > 1) {
				C = xistar_inv / m->l[q];
				A3 = alpha0 + 1.0 / q;
				A4 = alpha0_hat + xi_inv;
				Cpinv = (1.0 - A4 + A3) / A3;
				m->tq[1] = fabs(C * Cpinv);
			} else
				m->tq[1] = 1.0;
			hsum += m->tau[q];
			xi_inv = m->h / hsum;
			A5 = alpha0 - (1.0 / (q + 1));
			A6 = alpha0_hat - xi_inv;
			Cppinv = (1.0 - A6 + A5) / A2;
			m->tq[3] = fabs(Cppinv / (xi_inv * (q + 2) * A5));
		}
		m->tq[4] = CORTES / m->tq[2];
	}
	m->rl1 = 1.0 / m->l[1];
	m->gamma = m->h * m->rl1;
	if (m->nst == 0) m->gammap = m->gamma;
	m->gamrat = (m->nst > 0) ? m->gamma / m->gammap : 1.0;
}

/* cvHandleNFlag, cvode.c:2865-2908 */
static int handle_nflag(bdf_mem* m, int* nflag, double saved_t, int* ncf)
{
	int nf = *nflag;
	if (nf == 0) return DO_ERROR_TEST;
	m->ncfn++;
	restore(m, saved_t);
	if (nf < 0) return nf; /* LSETUP_FAIL / RHSFUNC_FAIL */
	(*ncf)++;
	m->etamax = 1.0;
	if ((fabs(m->h) <= m->hmin * ONEPSM) || (*ncf == MXNCF)) return BDF_CONV_FAILURE;
	m->eta = dmax(ETACF, m->hmin / fabs(m->h));
	*nflag = PREV_CONV_FAIL;
	rescale(m);
	return PREDICT_AGAIN;
}

/* cvDoErrorTest, cvode.c:2958-3023 */
static int do_error_test(bdf_mem* m, int* nflag, double saved_t, int* nef, double* dsm_out)
{
	double dsm = m->acnrm * m->tq[2];
	*dsm_out = dsm;
	if (dsm <= 1.0) return 0;

	(*nef)++;
	m->netf++;
	*nflag = PREV_ERR_FAIL;
	restore(m, saved_t);

	if ((fabs(m->h) <= m->hmin * ONEPSM) || (*nef == MXNEF)) return BDF_ERR_FAILURE;

	m->etamax = 1.0;

	if (*nef <= MXNEF1) {
		m->eta = 1.0 / (rpower_r(BIAS2 * dsm, 1.0 / m->L) + ADDON);
		m->eta = dmax(ETAMIN, dmax(m->eta, m->hmin / fabs(m->h)));
		if (*nef >= SMALL_NEF) m->eta = dmin(m->eta, ETAMXF);
		rescale(m);
		return TRY_AGAIN;
	}

	if (m->q > 1) {
		m->eta = dmax(ETAMIN, m->hmin / fabs(m->h));
		adjust_order(m, -1);
		m->L = m->q;
		m->q--;
		m->qwait = m->L;
		rescale(m);
		return TRY_AGAIN;
	}

	m->eta = dmax(ETAMIN, m->hmin / fabs(m->h));
	m->h *= m->eta;
	m->hscale = m->h;
	m->qwait = LONG_WAIT;

	int rv = m->f(m->tn, m->zn[0], m->tempv, m->user);
	m->nfe++;
	if (rv != 0) return BDF_RHSFUNC_FAIL;
	for (int i = 0; i < m->N; i++) m->zn[1][i] = m->h * m->tempv[i];
	return TRY_AGAIN;
}

/* cvCompleteStep, cvode.c:3043-3073 */
static void complete_step(bdf_mem* m)
{
	m->nst++;
	m->hu = m->h;
	m->qu = m->q;
	for (int i = m->q; i >= 2; i--) m->tau[i] = m->tau[i - 1];
	if ((m->q == 1) && (m->nst > 1)) m->tau[2] = m->tau[1];
	m->tau[1] = m->h;
	for (int j = 0; j <= m->q; j++)
		for (int i = 0; i < m->N; i++) m->zn[j][i] += m->l[j] * m->acor[i];
	m->qwait--;
	if ((m->qwait == 1) && (m->q != BDF_QMAX)) {
		for (int i = 0; i < m->N; i++) m->zn[BDF_QMAX][i] = m->acor[i];
		m->saved_tq5 = m->tq[5];
	}
}

/* cvSetEta, cvode.c:3132-3147 */
static void set_eta(bdf_mem* m)
{
	if (m->eta < THRESH) {
		m->eta = 1.0;
		m->hprime = m->h;
	} else {
		m->eta = dmin(m->eta, m->etamax);
		m->eta /= dmax(1.0, fabs(m->h) * m->hmax_inv * m->eta);
		m->hprime = m->h * m->eta;
	}
}

/* cvPrepareNextStep + cvComputeEtaqm1/qp1 + cvChooseEta, cvode.c:3093-3243 */
static void prepare_next_step(bdf_mem* m, double dsm)
{
	const int N = m->N;
	if (m->etamax == 1.0) {
		m->qwait = m->qwait > 2 ? m->qwait : 2;
		m->qprime = m->q;
		m->hprime = m->h;
		m->eta = 1.0;
		return;
	}
	m->etaq = 1.0 / (rpower_r(BIAS2 * dsm, 1.0 / m->L) + ADDON);
	if (m->qwait != 0) {
		m->eta = m->etaq;
		m->qprime = m->q;
		set_eta(m);
		return;
	}
	m->qwait = 2;
	/* cvComputeEtaqm1 */
	m->etaqm1 = 0.0;
	if (m->q > 1) {
		double ddn = wrms(N, m->zn[m->q], m->ewt) * m->tq[1];
		m->etaqm1 = 1.0 / (rpower_r(BIAS1 * ddn, 1.0 / m->q) + ADDON);
	}
	/* cvComputeEtaqp1 */
	m->etaqp1 = 0.0;
	if (m->q != BDF_QMAX) {
		if (m->saved_tq5 != 0.0) {
			double cquot = (m->tq[5] / m->saved_tq5) * rpower_i(m->h / m->tau[2], m->L);
			/* N_VLinearSum(-cquot, zn[qmax], ONE, acor, tempv): b == ONE, z != y -> VLin1: a*x + y */
			for (int i = 0; i < N; i++) m->tempv[i] = -cquot * m->zn[BDF_QMAX][i] + m->acor[i];
			double dup = wrms(N, m->tempv, m->ewt) * m->tq[3];
			m->etaqp1 = 1.0 / (rpower_r(BIAS3 * dup, 1.0 / (m->L + 1)) + ADDON);
		}
	}
	/* cvChooseEta */
	double etam = dmax(m->etaqm1, dmax(m->etaq, m->etaqp1));
	if (etam < THRESH) {
		m->eta = 1.0;
		m->qprime = m->q;
	} else if (etam == m->etaq) {
		m->eta = m->etaq;
		m->qprime = m->q;
	} else if (etam == m->etaqm1) {
		m->eta = m->etaqm1;
		m->qprime = m->q - 1;
	} else {
		m->eta = m->etaqp1;
		m->qprime = m->q + 1;
		for (int i = 0; i < N; i++) m->zn[BDF_QMAX][i] = m->acor[i];
	}
	set_eta(m);
}

/* cvStep, cvode.c:2082-2174 */
static int take_step(bdf_mem* m)
{
	double saved_t = m->tn;
	double dsm = 0.0;
	int ncf = 0, nef = 0;
	int nflag = FIRST_CALL;

	if ((m->nst > 0) && (m->hprime != m->h)) adjust_params(m);

	for (;;) {
		predict(m);
		set_bdf(m);
		nflag = nls_solve(m, nflag);
		int kflag = handle_nflag(m, &nflag, saved_t, &ncf);
		if (kflag == PREDICT_AGAIN) continue;
		if (kflag != DO_ERROR_TEST) return kflag;
		int eflag = do_error_test(m, &nflag, saved_t, &nef, &dsm);
		if (eflag == TRY_AGAIN) continue;
		if (eflag != 0) return eflag;
		break;
	}
	complete_step(m);
	prepare_next_step(m, dsm);
	m->etamax = (m->nst <= SMALL_NST) ? ETAMX2 : ETAMX3;
	for (int i = 0; i < m->N; i++) m->acor[i] *= m->tq[2];
	return 0;
}

/* cvUpperBoundH0, cvode.c:1993-2029 (N_VMaxNorm_Eigen is maxCoeff(), nvector_serial_eigen.cpp:381-384) */
static double upper_bound_h0(bdf_mem* m, double tdist)
{
	double hub_inv = -INFINITY;
	for (int i = 0; i < m->N; i++) {
		double t2 = fabs(m->zn[0][i]);
		double t1 = 1.0 / (m->reltol * fabs(m->zn[0][i]) + m->abstol[i]); /* efun */
		t1 = 1.0 / t1;
		t1 = HUB_FACTOR * t2 + t1;
		t2 = fabs(m->zn[1][i]);
		t1 = t2 / t1;
		if (t1 > hub_inv) hub_inv = t1;
	}
	double hub = HUB_FACTOR * tdist;
	if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;
	return hub;
}

/* cvYddNorm, cvode.c:2038-2054 */
static int ydd_norm(bdf_mem* m, double hg, double* yddnrm)
{
	const int N = m->N;
	for (int i = 0; i < N; i++) m->y[i] = hg * m->zn[1][i] + m->zn[0][i];
	int rv = m->f(m->tn + hg, m->y, m->tempv, m->user);
	m->nfe++;
	if (rv != 0) return BDF_RHSFUNC_FAIL;
	/* N_VLinearSum(1/hg, tempv, -1/hg, zn[1], tempv): a == -b -> VScaleDiff: c*(x - y) */
	double c = 1.0 / hg;
	for (int i = 0; i < N; i++) m->tempv[i] = c * (m->tempv[i] - m->zn[1][i]);
	*yddnrm = wrms(N, m->tempv, m->ewt);
	return 0;
}

/* cvHin, cvode.c:1884-1984 */
static int hin(bdf_mem* m, double tout)
{
	double tdiff = tout - m->tn;
	if (tdiff == 0.0) return BDF_TOO_CLOSE;
	int sign = (tdiff > 0.0) ? 1 : -1;
	double tdist = fabs(tdiff);
	double tround = UROUND * dmax(fabs(m->tn), fabs(tout));
	if (tdist < 2.0 * tround) return BDF_TOO_CLOSE;

	double hlb = HLB_FACTOR * tround;
	double hub = upper_bound_h0(m, tdist);
	double hg = sqrt(hlb * hub);
	if (hub < hlb) {
		m->h = (sign == -1) ? -hg : hg;
		return 0;
	}
	double hnew = hg, hrat, yddnrm;
	for (int count1 = 1; count1 <= MAX_ITERS; count1++) {
		double hgs = hg * sign;
		int rv = ydd_norm(m, hgs, &yddnrm);
		if (rv < 0) return BDF_RHSFUNC_FAIL;
		hnew = (yddnrm * hub * hub > 2.0) ? sqrt(2.0 / yddnrm) : sqrt(hg * hub);
		if (count1 == MAX_ITERS) break;
		hrat = hnew / hg;
		if ((hrat > 0.5) && (hrat < 2.0)) break;
		if ((count1 > 1) && (hrat > 2.0)) {
			hnew = hg;
			break;
		}
		hg = hnew;
	}
	double h0 = H_BIAS * hnew;
	if (h0 < hlb) h0 = hlb;
	if (h0 > hub) h0 = hub;
	if (sign == -1) h0 = -h0;
	m->h = h0;
	return 0;
}

/* CVode(..., CV_ONE_STEP), cvode.c:1006-1443 */
int bdf_step(bdf_mem* m, double tout, double* yout, double* tret)
{
	const int N = m->N;
	double troundoff;

	if (m->nst == 0) {
		m->tretlast = *tret = m->tn;
		/* cvInitialSetup: initial error weights; linit resets the LS counters; NLS initialise */
		if (ewt_set(m, m->zn[0], m->ewt) != 0) return BDF_ILL_INPUT;
		m->nje = 0;
		m->nfeDQ = 0;
		m->nstlj = 0;
		m->nls_jcur = 0;

		int rv = m->f(m->tn, m->zn[0], m->zn[1], m->user);
		m->nfe++;
		if (rv != 0) return BDF_RHSFUNC_FAIL;

		if (m->tstopset) {
			if ((m->tstop - m->tn) * (tout - m->tn) <= 0.0) return BDF_ILL_INPUT;
		}
		double tout_hin = tout;
		if (m->tstopset && (tout - m->tn) * (tout - m->tstop) > 0.0) tout_hin = m->tstop;
		int hflag = hin(m, tout_hin);
		if (hflag != 0) return hflag;
		double rh = fabs(m->h) * m->hmax_inv;
		if (rh > 1.0) m->h /= rh;
		if (fabs(m->h) < m->hmin) m->h *= m->hmin / fabs(m->h);
		if (m->tstopset) {
			if ((m->tn + m->h - m->tstop) * m->h > 0.0) m->h = (m->tstop - m->tn) * (1.0 - 4.0 * UROUND);
		}
		m->hscale = m->h;
		m->hprime = m->h;
		for (int i = 0; i < N; i++) m->zn[1][i] *= m->h;
	}

	if (m->nst > 0) {
		troundoff = FUZZ_FACTOR * UROUND * (fabs(m->tn) + fabs(m->h));
		if (fabs(m->tn - m->tretlast) > troundoff) {
			m->tretlast = *tret = m->tn;
			for (int i = 0; i < N; i++) yout[i] = m->zn[0][i];
			return BDF_SUCCESS;
		}
		if (m->tstopset) {
			if (fabs(m->tn - m->tstop) <= troundoff) {
				if (bdf_get_dky(m, m->tstop, yout) != 0) return BDF_ILL_INPUT;
				m->tretlast = *tret = m->tstop;
				m->tstopset = 0;
				return BDF_TSTOP_RETURN;
			}
			if ((m->tn + m->hprime - m->tstop) * m->h > 0.0) {
				m->hprime = (m->tstop - m->tn) * (1.0 - 4.0 * UROUND);
				m->eta = m->hprime / m->h;
			}
		}
	}

	/* one internal step */
	if (m->nst > 0) {
		if (ewt_set(m, m->zn[0], m->ewt) != 0) {
			m->tretlast = *tret = m->tn;
			for (int i = 0; i < N; i++) yout[i] = m->zn[0][i];
			return BDF_ILL_INPUT;
		}
	}
	double nrm = wrms(N, m->zn[0], m->ewt);
	m->tolsf = UROUND * nrm;
	if (m->tolsf > 1.0) {
		m->tretlast = *tret = m->tn;
		for (int i = 0; i < N; i++) yout[i] = m->zn[0][i];
		m->tolsf *= 2.0;
		return BDF_TOO_MUCH_ACC;
	} else {
		m->tolsf = 1.0;
	}

	int kflag = take_step(m);
	if (kflag != 0) {
		m->tretlast = *tret = m->tn;
		for (int i = 0; i < N; i++) yout[i] = m->zn[0][i];
		return kflag;
	}

	if (m->tstopset) {
		troundoff = FUZZ_FACTOR * UROUND * (fabs(m->tn) + fabs(m->h));
		if (fabs(m->tn - m->tstop) <= troundoff) {
			(void)bdf_get_dky(m, m->tstop, yout);
			m->tretlast = *tret = m->tstop;
			m->tstopset = 0;
			return BDF_TSTOP_RETURN;
		}
		if ((m->tn + m->hprime - m->tstop) * m->h > 0.0) {
			m->hprime = (m->tstop - m->tn) * (1.0 - 4.0 * UROUND);
			m->eta = m->hprime / m->h;
		}
	}
	m->tretlast = *tret = m->tn;
	for (int i = 0; i < N; i++) yout[i] = m->zn[0][i];
	return BDF_SUCCESS;
}
