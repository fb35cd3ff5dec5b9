/* TEST INFRASTRUCTURE ONLY -- not part of the product; nothing under bcm3_b200/ may use it.
 *
 * Plain-C restatement of the slice of SUNDIALS CVODE 5.3.0 that BCM3 drives
 * (reference: dependencies/cvode-5.3.0/src/cvode/{cvode.c,cvode_nls.c,cvode_ls.c},
 * src/sunnonlinsol/newton/sunnonlinsol_newton.c) with BCM3's Eigen-backed linear
 * algebra (src/odecommon/{nvector_serial_eigen,sunmatrix_dense_eigen,sunlinsol_dense_eigen}.cpp,
 * src/utils/EigenPartialPivLUSomewhatSparse.h): variable-order (1..5) variable-step BDF
 * in Nordsieck form, modified Newton with CVODE's Jacobian / gamma reuse heuristics,
 * direct dense solve (explicit inverse for N=2,3; zero-skipping partial-pivot LU above),
 * CV_ONE_STEP stepping with tstop, CVodeReInit, CVodeGetDky.
 * Fixed settings as BCM3 uses them: CV_BDF, Newton, qmax=5, hin=0, no root finding,
 * no constraints, no stability-limit detection, SV tolerances.
 * Parity of this restatement is pinned by tests/golden/ (generated with oracle/_ref,
 * i.e. the reference's own compiled code) -- see tests/test_oracle.py.
 */
#ifndef BCM3_ORACLE_CVODE_BDF_H
#define BCM3_ORACLE_CVODE_BDF_H

#ifdef __cplusplus
extern "C" {
#endif

#define BDF_NMAX 64
#define BDF_QMAX 5
#define BDF_LMAX 6

/* return values of bdf_step (subset of CVODE's) */
#define BDF_SUCCESS 0
#define BDF_TSTOP_RETURN 1
#define BDF_TOO_MUCH_ACC (-2)
#define BDF_ERR_FAILURE (-3)
#define BDF_CONV_FAILURE (-4)
#define BDF_LSETUP_FAIL (-6)
#define BDF_RHSFUNC_FAIL (-8)
#define BDF_ILL_INPUT (-22)
#define BDF_BAD_T (-25)
#define BDF_TOO_CLOSE (-27)

/* return 0 = ok, <0 unrecoverable (BCM3's callbacks never report recoverable errors, ODESolverCVODE.cpp:465-477) */
typedef int (*bdf_rhs_fn)(double t, const double* y, double* ydot, void* user);
/* J is N x N column-major, zeroed before the call (cvLsLinSys, cvode_ls.c:1236-1244) */
typedef int (*bdf_jac_fn)(double t, const double* y, const double* fy, double* J, void* user);

typedef struct bdf_mem {
	int N;
	bdf_rhs_fn f;
	bdf_jac_fn jac; /* NULL => difference quotients, ODESolverCVODE.cpp:496-537 */
	void* user;
	double reltol;
	double abstol[BDF_NMAX];
	double hmin, hmax_inv;

	double zn[BDF_LMAX][BDF_NMAX];
	double ewt[BDF_NMAX], y[BDF_NMAX], acor[BDF_NMAX], tempv[BDF_NMAX], ftemp[BDF_NMAX];
	double tn, h, hprime, hscale, eta, etamax, hu, tolsf;
	double etaq, etaqm1, etaqp1;
	int q, qprime, L, qwait, qu;
	double tau[BDF_LMAX + 1], tq[BDF_LMAX], l[BDF_LMAX];
	double rl1, gamma, gammap, gamrat, crate, delp, acnrm;
	int acnrmcur;
	double saved_tq5;
	int tstopset;
	double tstop, tretlast;
	long nst, nfe, ncfn, netf, nni, nsetups, nstlp, nje, nstlj, nfeDQ;
	int convfail, jcur, nls_jcur, nls_curiter;
	double savedJ[BDF_NMAX * BDF_NMAX];
	double A[BDF_NMAX * BDF_NMAX]; /* I - gamma J, then its inverse (N<=3) or LU factors */
	int piv[BDF_NMAX];
} bdf_mem;

void bdf_create(bdf_mem* m, int N, bdf_rhs_fn f, bdf_jac_fn jac, void* user);
void bdf_set_tolerances(bdf_mem* m, double reltol, const double* abstol);
void bdf_reinit(bdf_mem* m, double t0, const double* y0);
void bdf_set_stop_time(bdf_mem* m, double tstop);
/* CVode(mem, tout, yout, &tret, CV_ONE_STEP) */
int bdf_step(bdf_mem* m, double tout, double* yout, double* tret);
/* CVodeGetDky(mem, t, 0, dky) */
int bdf_get_dky(const bdf_mem* m, double t, double* dky);

#ifdef __cplusplus
}
#endif
#endif
