// bdf_thread.cuh -- one ODE system per CUDA thread: variable-order (1..5) variable-step BDF in
// Nordsieck form with modified Newton, reproducing the control flow of SUNDIALS CVODE 5.3.0 as
// BCM3 drives it (CV_BDF, Newton, direct dense solve, CV_ONE_STEP + tstop, CVodeReInit at every
// discontinuity). Written for small systems (N = 2, 3): the whole integrator state lives in
// registers -- every loop over the order q or the dimension N is fully unrolled with run-time
// predicates, so no array is dynamically indexed and nothing spills to local memory by design.
//
// What each block follows in the reference (dependencies/cvode-5.3.0/src/cvode/):
//   restart()        CVodeReInit cvode.c:586-665 + first-call block of CVode cvode.c:1068-1155
//                    (cvInitialSetup :1757, cvHin :1884-1984, cvUpperBoundH0 :1993-2029, cvYddNorm :2038-2054)
//   begin_step()     CVode loop head cvode.c:1294-1337 + cvStep head :2094-2102 (cvAdjustParams :2192,
//                    cvIncreaseBDF :2310, cvDecreaseBDF :2352, cvRescale :2384)
//   attempt()        cvPredict :2412, cvSet/cvSetBDF/cvSetTqBDF :2445-2686, cvNls :2701 +
//                    SUNNonlinSolSolve_Newton (sunnonlinsol_newton.c:183-318) + cvNlsLSetup/cvNlsConvTest/
//                    cvNlsResidual (cvode_nls.c:180-315) + cvLsSetup/cvLsLinSys/cvLsSolve (cvode_ls.c:1201-1286,
//                    1415-1507, 1603-1604), cvHandleNFlag :2865, cvDoErrorTest :2958, cvCompleteStep :3043,
//                    cvPrepareNextStep..cvChooseEta :3093-3243, etamax/acor scaling :2166-2171
//   after_step()     tstop handling of CVode cvode.c:1410-1438
//   dky()            CVodeGetDky (k = 0) cvode.c:1467-1524
// Linear algebra follows BCM3's Eigen-backed objects: explicit 2x2 inverse / 3x3 cofactor inverse without
// pivoting (src/odecommon/sunlinsol_dense_eigen.cpp:111-145, Eigen InverseImpl.h:124-170), WRMS norm
// summed left to right (src/odecommon/nvector_serial_eigen.cpp:386-396).
#pragma once

#include <cfloat>
#include <cmath>
#include <type_traits>
#include <utility>

namespace bcm3b200 {

// Compile-time loops over the order index. Written as fold expressions rather than `#pragma unroll`
// loops on purpose: with a run-time guard such as `if (j == q)` inside a real loop the optimiser
// rewrites the loop into one dynamically indexed access (zn[q]) before unrolling, which forces the
// whole integrator state into local memory.
template <class F, int... Js>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, Js...>)
{
	(f(std::integral_constant<int, Js>{}), ...);
}
// f(J) for J = B, B+1, ..., E-1
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f)
{
	if constexpr (E > B) static_for_impl([&](auto K) { f(std::integral_constant<int, B + decltype(K)::value>{}); }, std::make_integer_sequence<int, E - B>{});
}
// f(J) for J = E-1, E-2, ..., B
template <int B, int E, class F>
__device__ __forceinline__ void static_rfor(F&& f)
{
	if constexpr (E > B) static_for_impl([&](auto K) { f(std::integral_constant<int, E - 1 - decltype(K)::value>{}); }, std::make_integer_sequence<int, E - B>{});
}

// cvode.c:142-172, cvode_nls.c:29-31, cvode_ls_impl.h:40-42
#define BDF_FUZZ_FACTOR 100.0
#define BDF_HLB_FACTOR 100.0
#define BDF_HUB_FACTOR 0.1
#define BDF_H_BIAS 0.5
#define BDF_MAX_ITERS 4
#define BDF_CORTES 0.1
#define BDF_THRESH 1.5
#define BDF_ETAMX1 10000.0
#define BDF_ETAMX2 10.0
#define BDF_ETAMX3 10.0
#define BDF_ETAMXF 0.2
#define BDF_ETAMIN 0.1
#define BDF_ETACF 0.25
#define BDF_ADDON 0.000001
#define BDF_BIAS1 6.0
#define BDF_BIAS2 6.0
#define BDF_BIAS3 10.0
#define BDF_ONEPSM 1.000001
#define BDF_SMALL_NST 10
#define BDF_MXNCF 10
#define BDF_MXNEF 7
#define BDF_MXNEF1 3
#define BDF_SMALL_NEF 2
#define BDF_LONG_WAIT 10
#define BDF_DGMAX 0.3
#define BDF_MSBP 20
#define BDF_NLS_MAXCOR 3
#define BDF_CRDOWN 0.3
#define BDF_RDIV 2.0
#define BDF_MSBJ 50
#define BDF_LS_DGMAX 0.2
#define BDF_UROUND DBL_EPSILON

enum : int {
	BDF_FIRST_CALL = 101,
	BDF_PREV_CONV_FAIL = 102,
	BDF_PREV_ERR_FAIL = 103,
};
enum : int { BDF_NO_FAILURES = 0, BDF_FAIL_BAD_J = 1, BDF_FAIL_OTHER = 2 };

// outcome of one attempt()
enum : int {
	BDF_ATTEMPT_RETRY = 0,   // step rejected, state rescaled, call attempt() again
	BDF_ATTEMPT_DONE = 1,    // step accepted
	BDF_ATTEMPT_FAILED = -1, // CVode would return a negative flag
};

struct BdfCounters {
	int nfe, nsetups, nje, netf, ncfn, nni;
};

#ifndef BCM3_ROOT_INLINE
#define BCM3_ROOT_INLINE __noinline__
#endif
// SUNRpowerR(base, 1/k) for k = 1..6 (sundials_math.c): the k-th root in CVODE's step-size formulas.
// Evaluated as exp(log(base) / k) with the correctly rounded constant 1/k instead of pow(): a few ulp from pow's
// result, i.e. the same size as the FMA-contraction differences between two builds of the reference itself, and
// several times cheaper than the generic double-precision pow on the GPU.
static __device__ BCM3_ROOT_INLINE double bdf_root(double base, int k)
{
	if (base <= 0.0) return 0.0;
	double inv = 1.0;
	inv = (k == 2) ? (1.0 / 2.0) : inv;
	inv = (k == 3) ? (1.0 / 3.0) : inv;
	inv = (k == 4) ? (1.0 / 4.0) : inv;
	inv = (k == 5) ? (1.0 / 5.0) : inv;
	inv = (k == 6) ? (1.0 / 6.0) : inv;
	return exp(log(base) * inv);
}

// The same root without the double-precision log/exp pair: a single-precision seed exp2(log2(x) / k) (relative error
// ~3e-7) refined by one Halley step for y^k = x in double precision (cubic convergence: ~(k^2 - 1) / 12 * e^3 < 1e-17),
// so the result is within an ulp or two of pow(x, 1/k), like the exp/log form, at about a quarter of its instructions
// and with no dependence on k in the control flow. Outside the single-precision range it falls back to exp/log.
static __device__ BCM3_ROOT_INLINE double bdf_root_halley(double base, int k)
{
	if (base <= 0.0) return 0.0;
	if (!(base > 1e-30 && base < 1e30)) return bdf_root(base, k);
	const float kf = (float)k;
	const double y = (double)exp2f(__fdividef(log2f((float)base), kf));
	const double y2 = y * y, y3 = y2 * y, y4 = y2 * y2;
	double yk = y2;
	yk = (k == 3) ? y3 : yk;
	yk = (k == 4) ? y4 : yk;
	yk = (k == 5) ? y4 * y : yk;
	yk = (k == 6) ? y3 * y3 : yk;
	const double km = (double)(k - 1), kp = (double)(k + 1);
	const double num = km * yk + kp * base;
	const double den = kp * yk + km * base;
	return y * (num / den);
}

#ifndef BCM3_ROOT_EXPLOG
#define bdf_step_root bdf_root_halley
#else
#define bdf_step_root bdf_root
#endif

// Where the state lives. Hot vectors and scalars (Nordsieck columns 0 and 1, weights, correction, t, h, gamma, ...) are
// registers. Everything that is touched a few times per step at most -- Nordsieck columns 2..5, the step history tau,
// the inverse Newton matrix, the step-size bookkeeping, and l[] / tq[] while the Newton loop runs -- sits in a
// thread-private column of shared memory (word `slot` of thread `tid` at shared[slot * STRIDE + tid]: conflict-free),
// which brings the kernel from 255 registers (8 warps per SM) to 168 (12 warps per SM) without spills.
template <int N>
struct BdfSlots {
	enum : int {
		ZNH = 0,                  // zn[2..5][N]
		TAU = ZNH + 4 * N,        // tau[0..5] (tau[0] unused)
		MINV = TAU + 6,           // N * N
		HPRIME = MINV + N * N, HSCALE, ETA, ETAMAX, HU, GAMMAP, CRATE, DELP, ACNRM, SAVED_TQ5, SAVED_T, TSTOP,
		LSTASH,                   // l[0..5]
		TQSTASH = LSTASH + 6,     // tq[0..5]
		COUNT = TQSTASH + 6
	};
};

template <int N, class Model, bool STATS, int STRIDE>
struct BdfThread {
	static constexpr int QMAX = 5;
	using SL = BdfSlots<N>;

	double zn01[2][N];
	double ewt[N], acor[N];
	double tn, h;
	double gamma, gamrat, rl1;
	double* sh; // shared memory, already offset by the thread index
	int q, qprime, L, qwait, nst, nstlp, nstlj;
	int nflag, ncf, nef;
	// cvAdjustOrder + cvRescale requested by begin_step (cvAdjustParams) or by a failed attempt: carried out at the top of the
	// next attempt(), so that the kernel holds ONE copy of that code
	int pending_adjust;
	bool pending_rescale;
	bool tstopset, nls_jcur;
	BdfCounters cnt;

	__device__ __forceinline__ double& slot(int k) const { return sh[k * STRIDE]; }
	template <int J>
	__device__ __forceinline__ double& Z(int i)
	{
		if constexpr (J < 2) return zn01[J][i];
		else return sh[(SL::ZNH + (J - 2) * N + i) * STRIDE];
	}
	__device__ __forceinline__ double& tau(int j) const { return slot(SL::TAU + j); }
	__device__ __forceinline__ double& Minv(int k) const { return slot(SL::MINV + k); }
	__device__ __forceinline__ double& hprime() const { return slot(SL::HPRIME); }
	__device__ __forceinline__ double& hscale() const { return slot(SL::HSCALE); }
	__device__ __forceinline__ double& eta() const { return slot(SL::ETA); }
	__device__ __forceinline__ double& etamax() const { return slot(SL::ETAMAX); }
	__device__ __forceinline__ double& hu() const { return slot(SL::HU); }
	__device__ __forceinline__ double& gammap() const { return slot(SL::GAMMAP); }
	__device__ __forceinline__ double& crate() const { return slot(SL::CRATE); }
	__device__ __forceinline__ double& delp() const { return slot(SL::DELP); }
	__device__ __forceinline__ double& acnrm() const { return slot(SL::ACNRM); }
	__device__ __forceinline__ double& saved_tq5() const { return slot(SL::SAVED_TQ5); }
	__device__ __forceinline__ double& saved_t() const { return slot(SL::SAVED_T); }
	__device__ __forceinline__ double& tstop() const { return slot(SL::TSTOP); }

	__device__ __forceinline__ void count_reset()
	{
		if (STATS) { cnt.nfe = cnt.nsetups = cnt.nje = cnt.netf = cnt.ncfn = cnt.nni = 0; }
	}

	__device__ __forceinline__ double wrms(const double (&x)[N]) const
	{
		double sum = 0.0;
#pragma unroll
		for (int i = 0; i < N; i++) {
			double p = x[i] * ewt[i];
			sum += p * p;
		}
		return sqrt(sum * (1.0 / N));
	}

	// cvEwtSetSV, cvode.c:4268-4295 (atol > 0 so no N_VMin test)
	__device__ __forceinline__ void set_ewt(double rtol, double atol)
	{
#pragma unroll
		for (int i = 0; i < N; i++) ewt[i] = 1.0 / (rtol * fabs(Z<0>(i)) + atol);
	}

	// persistent members that CVodeCreate zero-fills once and CVodeReInit never touches
	__device__ __forceinline__ void create()
	{
		static_for<0, 6>([&](auto J) {
			constexpr int j = decltype(J)::value;
			tau(j) = 0.0;
#pragma unroll
			for (int i = 0; i < N; i++) Z<j>(i) = 0.0;
		});
#pragma unroll
		for (int i = 0; i < N; i++) acor[i] = 0.0;
#pragma unroll
		for (int i = 0; i < N * N; i++) Minv(i) = 0.0;
		gammap() = 0.0;
		crate() = 1.0;
		delp() = 0.0;
		acnrm() = 0.0;
		saved_tq5() = 0.0;
		tstopset = false;
		tstop() = 0.0;
		h = hprime() = hscale() = eta() = hu() = gamma = gamrat = rl1 = 0.0;
		nls_jcur = false;
		nstlj = 0;
		count_reset();
	}

	// CVodeReInit(t0, y0) followed by the first-call block of CVode(tout, CV_ONE_STEP).
	// tstop() (if any) must have been set by the caller. Returns false where CVode returns < 0.
	__device__ __forceinline__ bool restart(double t0, const double (&y0)[N], double tout, const Model& model, double rtol,
	                                        double atol)
	{
		tn = t0;
		q = 1;
		L = 2;
		qwait = 2;
		pending_adjust = 0;
		pending_rescale = false;
		etamax() = BDF_ETAMX1;
		hu() = 0.0;
		nst = 0;
		nstlp = 0;
#pragma unroll
		for (int i = 0; i < N; i++) Z<0>(i) = y0[i];

		// cvInitialSetup: error weights; cvLsInitialize resets nstlj; SUNNonlinSolInitialize resets jcur
		set_ewt(rtol, atol);
		nstlj = 0;
		nls_jcur = false;

		// zn[1] = f(t0, y0)
		model.rhs(tn, zn01[0], zn01[1]);
		if (STATS) cnt.nfe++;

		if (tstopset) {
			if ((tstop() - tn) * (tout - tn) <= 0.0) return false; // CV_ILL_INPUT
		}
		double tout_hin = tout;
		if (tstopset && (tout - tn) * (tout - tstop()) > 0.0) tout_hin = tstop();

		// ---- cvHin ----
		{
			double tdiff = tout_hin - tn;
			if (tdiff == 0.0) return false; // CV_TOO_CLOSE
			double sign = (tdiff > 0.0) ? 1.0 : -1.0;
			double tdist = fabs(tdiff);
			double tround = BDF_UROUND * fmax(fabs(tn), fabs(tout_hin));
			if (tdist < 2.0 * tround) return false;
			double hlb = BDF_HLB_FACTOR * tround;
			// cvUpperBoundH0 (N_VMaxNorm_Eigen is a plain maxCoeff, nvector_serial_eigen.cpp:381-384)
			double hub_inv = -INFINITY;
#pragma unroll
			for (int i = 0; i < N; i++) {
				double t2 = fabs(Z<0>(i));
				double t1 = 1.0 / ewt[i]; // N_VInv of the freshly computed weights
				t1 = BDF_HUB_FACTOR * t2 + t1;
				t2 = fabs(Z<1>(i));
				t1 = t2 / t1;
				hub_inv = (t1 > hub_inv) ? t1 : hub_inv;
			}
			double hub = BDF_HUB_FACTOR * tdist;
			if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;

			double hg = sqrt(hlb * hub);
			if (hub < hlb) {
				h = (sign < 0.0) ? -hg : hg;
			} else {
				double hnew = hg;
#pragma unroll 1
				for (int count1 = 1; count1 <= BDF_MAX_ITERS; count1++) {
					double hgs = hg * sign;
					// cvYddNorm
					double ytmp[N], ftmp[N];
#pragma unroll
					for (int i = 0; i < N; i++) ytmp[i] = hgs * Z<1>(i) + Z<0>(i);
					model.rhs(tn + hgs, ytmp, ftmp);
					if (STATS) cnt.nfe++;
					double c = 1.0 / hgs;
#pragma unroll
					for (int i = 0; i < N; i++) ftmp[i] = c * (ftmp[i] - Z<1>(i));
					double yddnrm = wrms(ftmp);

					hnew = (yddnrm * hub * hub > 2.0) ? sqrt(2.0 / yddnrm) : sqrt(hg * hub);
					if (count1 == BDF_MAX_ITERS) break;
					double hrat = hnew / hg;
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
		}
		// hmax_inv = 0, hmin = 0 (PopPK leaves CVODE's defaults)
		if (tstopset) {
			if ((tn + h - tstop()) * h > 0.0) h = (tstop() - tn) * (1.0 - 4.0 * BDF_UROUND);
		}
		hscale() = h;
		hprime() = h;
#pragma unroll
		for (int i = 0; i < N; i++) Z<1>(i) *= h;
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
				for (int i = 0; i < N; i++) Z<j>(i) *= c;
				c = eta() * c;
			}
		});
		h = hscale() * eta();
		hscale() = h;
	}

	// cvIncreaseBDF, cvode.c:2310-2340 (zn[L] <- A1 * saved acor in zn[qmax]; zn[2..q] += l[j] zn[L])
	__device__ __forceinline__ void increase_bdf()
	{
		double ll[6];
#pragma unroll
		for (int i = 0; i < 6; i++) ll[i] = 0.0;
		double alpha1 = 1.0, prod = 1.0, xiold = 1.0, alpha0 = -1.0;
		ll[2] = 1.0;
		double hsum = hscale();
		// an increase only happens for q < qmax, so j <= 3
		static_for<1, QMAX - 1>([&](auto J) {
			constexpr int j = decltype(J)::value;
			if (j < q) {
				hsum += tau(j + 1);
				double xi = hsum / hscale();
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
		double A1 = (-alpha0 - alpha1) / prod;
		double znL[N];
#pragma unroll
		for (int i = 0; i < N; i++) znL[i] = A1 * Z<QMAX>(i);
		// zn[j] += l[j] * zn[L] for j = 2..q, then the new column zn[L] (L = q + 1 <= 5)
		static_for<2, QMAX + 1>([&](auto J) {
			constexpr int j = decltype(J)::value;
			if (j <= q) {
#pragma unroll
				for (int i = 0; i < N; i++) Z<j>(i) += ll[j] * znL[i];
			} else if (j <= L) {
#pragma unroll
				for (int i = 0; i < N; i++) Z<j>(i) = znL[i];
			}
		});
	}

	// cvDecreaseBDF, cvode.c:2352-2374 (zn[2..q-1] -= l[j] zn[q])
	__device__ __forceinline__ void decrease_bdf()
	{
		double ll[6];
#pragma unroll
		for (int i = 0; i < 6; i++) ll[i] = 0.0;
		ll[2] = 1.0;
		double hsum = 0.0;
		static_for<1, QMAX - 1>([&](auto J) {
			constexpr int j = decltype(J)::value;
			if (j <= q - 2) {
				hsum += tau(j);
				double xi = hsum / hscale();
				static_rfor<2, j + 3>([&](auto I) {
					constexpr int i = decltype(I)::value;
					ll[i] = ll[i] * xi + ll[i - 1];
				});
			}
		});
		// znq = zn[q] (q >= 3 here): the last column with j <= q wins
		double znq[N];
#pragma unroll
		for (int i = 0; i < N; i++) znq[i] = Z<2>(i);
		static_for<3, QMAX + 1>([&](auto J) {
			constexpr int j = decltype(J)::value;
			if (j <= q) {
#pragma unroll
				for (int i = 0; i < N; i++) znq[i] = Z<j>(i);
			}
		});
		if (q > 2) {
			static_for<2, QMAX>([&](auto J) {
				constexpr int j = decltype(J)::value;
				if (j < q) {
#pragma unroll
					for (int i = 0; i < N; i++) Z<j>(i) += (-ll[j]) * znq[i];
				}
			});
		}
	}

	// cvAdjustOrder, cvode.c:2213-2225
	__device__ __forceinline__ void adjust_order(int deltaq)
	{
		if ((q == 2) && (deltaq != 1)) return;
		if (deltaq == 1) increase_bdf();
		else if (deltaq == -1) decrease_bdf();
	}

	// head of one step: CVode loop head (ewt reset, too-much-accuracy test) + cvStep head.
	// Returns false on CV_TOO_MUCH_ACC.
	__device__ __forceinline__ bool begin_step(double rtol, double atol)
	{
		if (nst > 0) set_ewt(rtol, atol);
		double tolsf = BDF_UROUND * wrms(zn01[0]);
		if (tolsf > 1.0) return false;

		saved_t() = tn;
		ncf = 0;
		nef = 0;
		nflag = BDF_FIRST_CALL;
		if ((nst > 0) && (hprime() != h)) {
			// cvAdjustParams: order change, then rescale -- at the top of attempt()
			pending_adjust = qprime - q;
			pending_rescale = true;
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
				if (j <= q) { // implies k <= q
#pragma unroll
					for (int i = 0; i < N; i++) Z<j - 1>(i) = Z<j - 1>(i) - Z<j>(i);
				}
			});
		});
	}

	// SUNLinSolSetup_Dense_Eigen2x2/3x3 on A = I - gamma*J (cvLsLinSys + SUNMatScaleAddI)
	__device__ __forceinline__ void linear_setup(const Model& model)
	{
		double A[N * N];
#pragma unroll
		for (int i = 0; i < N * N; i++) A[i] = 0.0;
		model.jac(A); // row-major A[i * N + j]
#pragma unroll
		for (int i = 0; i < N * N; i++) A[i] *= -gamma;
#pragma unroll
		for (int i = 0; i < N; i++) A[i * N + i] += 1.0;
		if (N == 2) {
			double invdet = 1.0 / (A[0] * A[3] - A[1] * A[2]);
			Minv(0) = A[3] * invdet;
			Minv(1) = -A[1] * invdet;
			Minv(2) = -A[2] * invdet;
			Minv(3) = A[0] * invdet;
		} else {
#define BDF_A(i, j) A[(i) * N + (j)]
#define BDF_COF(i, j) \
	(BDF_A(((i) + 1) % 3, ((j) + 1) % 3) * BDF_A(((i) + 2) % 3, ((j) + 2) % 3) - BDF_A(((i) + 1) % 3, ((j) + 2) % 3) * BDF_A(((i) + 2) % 3, ((j) + 1) % 3))
			double c0 = BDF_COF(0, 0), c1 = BDF_COF(1, 0), c2 = BDF_COF(2, 0);
			double det = c0 * BDF_A(0, 0) + c1 * BDF_A(1, 0) + c2 * BDF_A(2, 0);
			double invdet = 1.0 / det;
			Minv(0 * N + 0) = c0 * invdet;
			Minv(0 * N + 1) = c1 * invdet;
			Minv(0 * N + 2) = c2 * invdet;
			Minv(1 * N + 0) = BDF_COF(0, 1) * invdet;
			Minv(1 * N + 1) = BDF_COF(1, 1) * invdet;
			Minv(1 * N + 2) = BDF_COF(2, 1) * invdet;
			Minv(2 * N + 0) = BDF_COF(0, 2) * invdet;
			Minv(2 * N + 1) = BDF_COF(1, 2) * invdet;
			Minv(2 * N + 2) = BDF_COF(2, 2) * invdet;
#undef BDF_COF
#undef BDF_A
		}
	}

	// One pass of cvStep's attempt loop.
	__device__ __forceinline__ int attempt(const Model& model, unsigned mask)
	{
		// l[] and tq[] never outlive one attempt: cvSetBDF rewrites l[0..q], tq[2], tq[4], tq[5] every time and
		// tq[1], tq[3] are written (qwait == 1) in the same step that reads them (qwait == 0 after cvCompleteStep)
		double l[6], tq[6];
#pragma unroll
		for (int i = 0; i < 6; i++) {
			l[i] = 0.0;
			tq[i] = 0.0;
		}
		// ---- pending cvAdjustOrder (q, L, qwait follow as in cvAdjustParams :2192-2204 / cvDoErrorTest :3017-3022) + cvRescale ----
		if (pending_adjust != 0) {
			adjust_order(pending_adjust);
			q += pending_adjust;
			L = q + 1;
			qwait = L;
			pending_adjust = 0;
		}
		if (pending_rescale) {
			rescale();
			pending_rescale = false;
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
				if (j <= q) { // implies k <= q
#pragma unroll
					for (int i = 0; i < N; i++) Z<j - 1>(i) += Z<j>(i);
				}
			});
		});

		// ---- cvSetBDF + cvSetTqBDF ----
		{
			double xi_inv = 1.0, xistar_inv = 1.0, alpha0 = -1.0, alpha0_hat = -1.0;
			double hsum = h;
			l[0] = 1.0;
			l[1] = 1.0;
			static_for<2, QMAX + 1>([&](auto I) {
				constexpr int i = decltype(I)::value;
				if (i <= q) l[i] = 0.0;
			});
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
				// tau(q - 1): the last j <= q - 1 wins
				hsum += tau(q - 1);
				xi_inv = h / hsum;
				alpha0_hat = -l[1] - xi_inv;
				static_rfor<1, QMAX + 1>([&](auto I) {
					constexpr int i = decltype(I)::value;
					if (i <= q) l[i] += l[i - 1] * xistar_inv;
				});
			}
			// l[q], tau(q): the last j <= q wins
			double lq = l[1];
			static_for<2, QMAX + 1>([&](auto J) {
				constexpr int j = decltype(J)::value;
				if (j <= q) lq = l[j];
			});
			const double tauq = tau(q);
			double A1 = 1.0 - alpha0_hat + alpha0;
			double A2 = 1.0 + q * A1;
			tq[2] = fabs(A1 / (alpha0 * A2));
			tq[5] = fabs(A2 * xistar_inv / (lq * xi_inv));
			if (qwait == 1) {
				if (q > 1) {
					double C = xistar_inv / lq;
					double A3 = alpha0 + 1.0 / q;
					double A4 = alpha0_hat + xi_inv;
					double Cpinv = (1.0 - A4 + A3) / A3;
					tq[1] = fabs(C * Cpinv);
				} else {
					tq[1] = 1.0;
				}
				hsum += tauq;
				xi_inv = h / hsum;
				double A5 = alpha0 - (1.0 / (q + 1));
				double A6 = alpha0_hat - xi_inv;
				double Cppinv = (1.0 - A6 + A5) / A2;
				tq[3] = fabs(Cppinv / (xi_inv * (q + 2) * A5));
			}
			tq[4] = BDF_CORTES / tq[2];
			// cvSet tail
			rl1 = 1.0 / l[1];
			gamma = h * rl1;
			if (nst == 0) gammap() = gamma;
			gamrat = (nst > 0) ? gamma / gammap() : 1.0;
		}

		// ---- cvNls / Newton ----
		// Both loops below are made warp-uniform with votes over `mask` (the lanes that entered this attempt):
		// every lane stays in a loop until no lane needs another trip, so the warp is converged again when the
		// loop ends and the step-completion code that follows runs once for all lanes, not once per exit path.
		// l[] and tq[1, 2, 3, 5] are not needed until the step is accepted: parked in shared memory over the Newton loop
#pragma unroll
		for (int i = 0; i < 6; i++) {
			slot(SL::LSTASH + i) = l[i];
			slot(SL::TQSTASH + i) = tq[i];
		}
		int nls_ret = 1; // 0 ok, 1 recoverable convergence failure
		{
			int convfail = ((nflag == BDF_FIRST_CALL) || (nflag == BDF_PREV_ERR_FAIL)) ? BDF_NO_FAILURES : BDF_FAIL_OTHER;
			bool callSetup = (nflag == BDF_PREV_CONV_FAIL) || (nflag == BDF_PREV_ERR_FAIL) || (nst == 0) ||
			                 (nst >= nstlp + BDF_MSBP) || (fabs(gamrat - 1.0) > BDF_DGMAX);
#pragma unroll
			for (int i = 0; i < N; i++) acor[i] = 0.0;
			const double tol = tq[4];
			bool jbad = false;
			bool need_pass = true;
#pragma unroll 1
			for (;;) {
				if (!__any_sync(mask, need_pass)) break;
				double delta[N];
#pragma unroll
				for (int i = 0; i < N; i++) delta[i] = 0.0;
				if (need_pass) {
					// cvNlsResidual
					double y[N], f[N];
#pragma unroll
					for (int i = 0; i < N; i++) y[i] = Z<0>(i) + acor[i];
					model.rhs(tn, y, f);
					if (STATS) cnt.nfe++;
#pragma unroll
					for (int i = 0; i < N; i++) delta[i] = rl1 * Z<1>(i) + acor[i];
#pragma unroll
					for (int i = 0; i < N; i++) delta[i] += -gamma * f[i];

					if (callSetup) {
						// cvNlsLSetup + cvLsSetup
						if (jbad) convfail = BDF_FAIL_BAD_J;
						double dgamma = fabs((gamma / gammap()) - 1.0);
						bool jb = (nst == 0) || (nst > nstlj + BDF_MSBJ) || ((convfail == BDF_FAIL_BAD_J) && (dgamma < BDF_LS_DGMAX)) ||
						          (convfail == BDF_FAIL_OTHER);
						if (jb) {
							nstlj = nst;
							if (STATS) cnt.nje++;
						}
						linear_setup(model);
						if (STATS) cnt.nsetups++;
						nls_jcur = jb;
						gamrat = 1.0;
						gammap() = gamma;
						crate() = 1.0;
						nstlp = nst;
					}
				}

				bool iter = need_pass;
#pragma unroll 1
				for (int m = 0; m < BDF_NLS_MAXCOR; m++) {
					if (!__any_sync(mask, iter)) break;
					if (iter) {
						if (STATS) cnt.nni++;
						// delta <- A^-1 (-delta), scaled for a changed gamma (cvLsSolve)
						double b[N];
#pragma unroll
						for (int i = 0; i < N; i++) b[i] = -delta[i];
#pragma unroll
						for (int i = 0; i < N; i++) {
							double sx = Minv(i * N + 0) * b[0];
#pragma unroll
							for (int j = 1; j < N; j++) sx += Minv(i * N + j) * b[j];
							delta[i] = sx;
						}
						if (gamrat != 1.0) {
							double sc = 2.0 / (1.0 + gamrat);
#pragma unroll
							for (int i = 0; i < N; i++) delta[i] *= sc;
						}
#pragma unroll
						for (int i = 0; i < N; i++) acor[i] += delta[i];

						// cvNlsConvTest
						double del = wrms(delta);
						if (m > 0) crate() = fmax(BDF_CRDOWN * crate(), del / delp());
						double dcon = del * fmin(1.0, crate()) / tol;
						if (dcon <= 1.0) {
							acnrm() = (m == 0) ? del : wrms(acor);
							nls_jcur = false;
							nls_ret = 0;
							iter = false;
						} else if ((m >= 1) && (del > BDF_RDIV * delp())) {
							iter = false;
						} else {
							delp() = del;
							if (m + 1 >= BDF_NLS_MAXCOR) {
								iter = false;
							} else {
								// next residual
								double y[N], f[N];
#pragma unroll
								for (int i = 0; i < N; i++) y[i] = Z<0>(i) + acor[i];
								model.rhs(tn, y, f);
								if (STATS) cnt.nfe++;
#pragma unroll
								for (int i = 0; i < N; i++) delta[i] = rl1 * Z<1>(i) + acor[i];
#pragma unroll
								for (int i = 0; i < N; i++) delta[i] += -gamma * f[i];
							}
						}
					}
				}

				if (need_pass) {
					if (nls_ret == 0 || nls_jcur) {
						need_pass = false;
					} else {
						// stale Jacobian data: redo with a forced setup (sunnonlinsol_newton.c:301-312)
						callSetup = true;
						jbad = true;
#pragma unroll
						for (int i = 0; i < N; i++) acor[i] = 0.0;
					}
				}
			}
		}

#pragma unroll
		for (int i = 0; i < 6; i++) {
			l[i] = slot(SL::LSTASH + i);
			tq[i] = slot(SL::TQSTASH + i);
		}
		int result;
		// Both failure paths (cvHandleNFlag :2865, cvDoErrorTest :2958) restore the Nordsieck array and, unless the step is
		// abandoned, rescale it: written with ONE restore() and ONE rescale() site -- per warp a failure happens in half of
		// the trips, so this is hot code, and the kernel is bound by its instruction footprint.
		const double dsm = acnrm() * tq[2];
		const bool conv_fail = (nls_ret != 0);
		const bool err_fail = !conv_fail && !(dsm <= 1.0);
		if (conv_fail || err_fail) {
			if (conv_fail) {
				if (STATS) cnt.ncfn++;
				ncf++;
			} else {
				nef++;
				if (STATS) cnt.netf++;
				nflag = BDF_PREV_ERR_FAIL;
			}
			restore();
			if (conv_fail) {
				// ---- cvHandleNFlag ----
				etamax() = 1.0;
				// hmin = 0: |h| <= hmin * ONEPSM only for h == 0
				if ((fabs(h) <= 0.0) || (ncf == BDF_MXNCF)) {
					result = BDF_ATTEMPT_FAILED;
				} else {
					eta() = BDF_ETACF;
					nflag = BDF_PREV_CONV_FAIL;
					pending_rescale = true;
					result = BDF_ATTEMPT_RETRY;
				}
			} else if ((fabs(h) <= 0.0) || (nef == BDF_MXNEF)) {
				// ---- cvDoErrorTest ----
				result = BDF_ATTEMPT_FAILED;
			} else {
				result = BDF_ATTEMPT_RETRY;
				etamax() = 1.0;
				if (nef <= BDF_MXNEF1) {
					eta() = 1.0 / (bdf_step_root(BDF_BIAS2 * dsm, L) + BDF_ADDON);
					eta() = fmax(BDF_ETAMIN, eta());
					if (nef >= BDF_SMALL_NEF) eta() = fmin(eta(), BDF_ETAMXF);
					pending_rescale = true;
				} else if (q > 1) {
					eta() = BDF_ETAMIN;
					pending_adjust = -1;
					pending_rescale = true;
				} else {
					eta() = BDF_ETAMIN;
					h *= eta();
					hscale() = h;
					qwait = BDF_LONG_WAIT;
					double f[N];
					model.rhs(tn, zn01[0], f);
					if (STATS) cnt.nfe++;
#pragma unroll
					for (int i = 0; i < N; i++) Z<1>(i) = h * f[i];
				}
			}
		} else {
			{
				result = BDF_ATTEMPT_DONE;
				// ---- cvCompleteStep ----
				nst++;
				hu() = h;
				static_rfor<2, QMAX + 1>([&](auto I) {
					constexpr int i = decltype(I)::value;
					if (i <= q) tau(i) = tau(i - 1);
				});
				if ((q == 1) && (nst > 1)) tau(2) = tau(1);
				tau(1) = h;
				static_for<0, QMAX + 1>([&](auto J) {
					constexpr int j = decltype(J)::value;
					if (j <= q) {
#pragma unroll
						for (int i = 0; i < N; i++) Z<j>(i) += l[j] * acor[i];
					}
				});
				qwait--;
				if ((qwait == 1) && (q != QMAX)) {
#pragma unroll
					for (int i = 0; i < N; i++) Z<QMAX>(i) = acor[i];
					saved_tq5() = tq[5];
				}

				// ---- cvPrepareNextStep ----
				if (etamax() == 1.0) {
					qwait = (qwait > 2) ? qwait : 2;
					qprime = q;
					hprime() = h;
					eta() = 1.0;
				} else {
					const double etaq = 1.0 / (bdf_step_root(BDF_BIAS2 * dsm, L) + BDF_ADDON);
					eta() = etaq;
					qprime = q;
					if (qwait == 0) {
						qwait = 2;
						// cvComputeEtaqm1
						double etaqm1 = 0.0;
						if (q > 1) {
							double znq[N];
#pragma unroll
							for (int i = 0; i < N; i++) znq[i] = Z<2>(i);
							static_for<3, QMAX + 1>([&](auto J) {
								constexpr int j = decltype(J)::value;
								if (j <= q) { // the last j <= q wins: znq = zn[q]
#pragma unroll
									for (int i = 0; i < N; i++) znq[i] = Z<j>(i);
								}
							});
							double ddn = wrms(znq) * tq[1];
							etaqm1 = 1.0 / (bdf_step_root(BDF_BIAS1 * ddn, q) + BDF_ADDON);
						}
						// cvComputeEtaqp1
						double etaqp1 = 0.0;
						if (q != QMAX) {
							if (saved_tq5() != 0.0) {
								double base = h / tau(2);
								double pw = 1.0;
								static_for<1, QMAX + 1>([&](auto I) {
									if (decltype(I)::value <= L) pw *= base;
								});
								double cquot = (tq[5] / saved_tq5()) * pw;
								double tmp[N];
#pragma unroll
								for (int i = 0; i < N; i++) tmp[i] = -cquot * Z<QMAX>(i) + acor[i];
								double dup = wrms(tmp) * tq[3];
								etaqp1 = 1.0 / (bdf_step_root(BDF_BIAS3 * dup, L + 1) + BDF_ADDON);
							}
						}
						// cvChooseEta
						double etam = fmax(etaqm1, fmax(etaq, etaqp1));
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
							for (int i = 0; i < N; i++) Z<QMAX>(i) = acor[i];
						}
					}
					// cvSetEta (hmax_inv = 0)
					if (eta() < BDF_THRESH) {
						eta() = 1.0;
						hprime() = h;
					} else {
						eta() = fmin(eta(), etamax());
						hprime() = h * eta();
					}
				}
				etamax() = (nst <= BDF_SMALL_NST) ? BDF_ETAMX2 : BDF_ETAMX3;
#pragma unroll
				for (int i = 0; i < N; i++) acor[i] *= tq[2];
			}
		}
		return result;
	}

	// CVodeGetDky(t, k = 0) for all components. Returns false on CV_BAD_T.
	__device__ __forceinline__ bool dky(double t, double (&out)[N])
	{
		double tfuzz = BDF_FUZZ_FACTOR * BDF_UROUND * (fabs(tn) + fabs(hu()));
		if (hu() < 0.0) tfuzz = -tfuzz;
		double tp = tn - hu() - tfuzz;
		double tn1 = tn + tfuzz;
		if ((t - tp) * (t - tn1) > 0.0) return false;
		double s = (t - tn) / h;
		// z = c_q zn[q] + c_{q-1} zn[q-1] + ... + zn[0], c_j = s^j built by repeated multiplication,
		// accumulated from the highest order down (N_VLinearCombination_Eigen, nvector_serial_eigen.cpp:496-543)
		double acc[N];
#pragma unroll
		for (int i = 0; i < N; i++) acc[i] = 0.0;
		bool first = true;
		static_rfor<0, QMAX + 1>([&](auto J) {
			constexpr int j = decltype(J)::value;
			if (j <= q) {
				double c = 1.0;
#pragma unroll
				for (int i = 0; i < j; i++) c *= s;
				if (first) {
#pragma unroll
					for (int i = 0; i < N; i++) acc[i] = c * Z<j>(i);
					first = false;
				} else {
#pragma unroll
					for (int i = 0; i < N; i++) acc[i] += c * Z<j>(i);
				}
			}
		});
#pragma unroll
		for (int i = 0; i < N; i++) out[i] = acc[i];
		return true;
	}

	// tstop() handling after an accepted step (cvode.c:1410-1438). Returns true for CV_TSTOP_RETURN, in which
	// case yout = Dky(tstop()) and tret = tstop(); otherwise yout = zn[0], tret = tn.
	__device__ __forceinline__ bool after_step(double (&yout)[N], double& tret)
	{
		if (tstopset) {
			double troundoff = BDF_FUZZ_FACTOR * BDF_UROUND * (fabs(tn) + fabs(h));
			if (fabs(tn - tstop()) <= troundoff) {
				(void)dky(tstop(), yout);
				tret = tstop();
				tstopset = false;
				return true;
			}
			if ((tn + hprime() - tstop()) * h > 0.0) {
				hprime() = (tstop() - tn) * (1.0 - 4.0 * BDF_UROUND);
				eta() = hprime() / h;
			}
		}
		tret = tn;
#pragma unroll
		for (int i = 0; i < N; i++) yout[i] = Z<0>(i);
		return false;
	}
};

} // namespace bcm3b200
