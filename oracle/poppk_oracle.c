/* TEST INFRASTRUCTURE ONLY -- not part of the product; nothing under bcm3_b200/ may use it.
 *
 * Plain-C restatement of LikelihoodPopPKTrajectory::EvaluateLogProbability
 * (reference: src/likelihoods/LikelihoodPopPKTrajectory.cpp:259-444) on top of
 * oracle/cvode_bdf.c, including the ODESolverCVODE driver loop
 * (src/odecommon/ODESolverCVODE.cpp:322-463, src/odecommon/ODESolver.cpp:93-134).
 * Pinned against oracle/_ref (the reference's own compiled solver) via tests/golden/.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "cvode_bdf.h"
#include "oracle_api.h"

/* Phi^-1: stands in for boost::math::quantile(normal) (ProbabilityDistributions.cpp:359-363);
 * rational start (Acklam) + two Halley steps on libm erfc => full double accuracy. */
static double ndtri(double p)
{
	static const double a[6] = { -3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
	                             1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00 };
	static const double b[5] = { -5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
	                             6.680131188771972e+01, -1.328068155288572e+01 };
	static const double c[6] = { -7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
	                             -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00 };
	static const double d[4] = { 7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
	                             3.754408661907416e+00 };
	if (!(p > 0.0 && p < 1.0)) {
		if (p == 0.0) return -INFINITY;
		if (p == 1.0) return INFINITY;
		return NAN;
	}
	double x;
	if (p < 0.02425) {
		double q = sqrt(-2.0 * log(p));
		x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
	} else if (p > 1.0 - 0.02425) {
		double q = sqrt(-2.0 * log(1.0 - p));
		x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
	} else {
		double q = p - 0.5, r = q * q;
		x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q / (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1.0);
	}
	for (int it = 0; it < 2; it++) {
		double e = (x < 0.0) ? 0.5 * erfc(-x * M_SQRT1_2) - p : (1.0 - p) - 0.5 * erfc(x * M_SQRT1_2);
		double u = e * 2.5066282746310002 * exp(0.5 * x * x);
		x = x - u / (1.0 + 0.5 * x * u);
	}
	return x;
}

double oracle_ndtri(double p) { return ndtri(p); }

static double quantile_normal(double p, double mu, double sigma)
{
	double r = ndtri(p);
	r *= sigma;
	r += mu;
	return r;
}

/* bcm3::fastpow10, MathFunctions.h:13 */
static double fastpow10(double x) { return exp(x * 2.3025850929940459); }

/* VariableSet::TransformVariable, VariableSet.cpp:97-124 */
static double transform_variable(int transform, double x)
{
	switch (transform) {
	case ORACLE_TRANSFORM_LOG:
		return exp(x);
	case ORACLE_TRANSFORM_LOG10:
		return fastpow10(x);
	case ORACLE_TRANSFORM_LOGIT:
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

/* bcm3::LogPdfTnu4, ProbabilityDistributions.cpp:216-224 */
static double logpdf_tnu4(double x, double mu, double sigma)
{
	double xn = (x - mu) / sigma;
	return -0.9808292530117262 - 2.5 * log1p(0.25 * xn * xn) - log(sigma);
}

typedef struct {
	double dose, dosing_interval, dose_after_dose_change, dose_change_time;
	unsigned intermittent;
	uint32_t skipped_days;
	double k_absorption, k_excretion, k_elimination, k_vod, k_periphery_fwd, k_periphery_bwd;
	double k_transit, n_transit, k_biphasic_switch_time, k_absorption2, last_treatment;
	int biphasic_switch;
	int two, biphasic, transit; /* model type flags */
	double current_dose_time;
} patient_data;

/* CalculateDerivative_* of the biphasic-uptake and transit types, cpp:496-642 (the plain types keep their own functions) */
static int rhs_variant(double t, const double* y, double* dydt, void* user)
{
	const patient_data* pd = (const patient_data*)user;
	double ka = pd->k_absorption;
	if (pd->biphasic && !pd->biphasic_switch) ka = pd->k_absorption2;
	if (pd->transit) {
		double dose = pd->dose;
		if (t >= pd->dose_change_time) dose = pd->dose_after_dose_change;
		double t_since_treatment = t - pd->last_treatment;
		double log_n_transit_factorial = 0.9189385332046727 + (pd->n_transit + 0.5) * log(pd->n_transit) - pd->n_transit + log(1 + 1 / (12.0 * pd->n_transit));
		double transit = exp((pd->n_transit * log(pd->k_transit * t_since_treatment) - pd->k_transit * t_since_treatment) - log_n_transit_factorial);
		transit = pd->k_transit * transit * dose;
		dydt[0] = transit - (ka + pd->k_excretion) * y[0];
	} else {
		dydt[0] = -(ka + pd->k_excretion) * y[0];
	}
	if (!pd->two) {
		dydt[1] = ka * y[0] - pd->k_elimination * y[1];
	} else {
		dydt[1] = ka * y[0] - pd->k_elimination * y[1] - pd->k_periphery_fwd * y[1] + pd->k_periphery_bwd * y[2];
		dydt[2] = pd->k_periphery_fwd * y[1] - pd->k_periphery_bwd * y[2];
	}
	return 0;
}
static int jac_variant(double t, const double* y, const double* fy, double* J, void* user)
{
	const patient_data* pd = (const patient_data*)user;
	(void)t; (void)y; (void)fy;
	double ka = pd->k_absorption;
	if (pd->biphasic && !pd->biphasic_switch) ka = pd->k_absorption2;
	const int N = pd->two ? 3 : 2;
	J[0 + 0 * N] = -(ka + pd->k_excretion);
	J[1 + 0 * N] = ka;
	if (!pd->two) {
		J[1 + 1 * N] = -pd->k_elimination;
	} else {
		J[1 + 1 * N] = -(pd->k_elimination + pd->k_periphery_fwd);
		J[1 + 2 * N] = pd->k_periphery_bwd;
		J[2 + 1 * N] = pd->k_periphery_fwd;
		J[2 + 2 * N] = -pd->k_periphery_bwd;
	}
	return 0;
}

/* cpp:446-455 */
static int rhs_one(double t, const double* y, double* dydt, void* user)
{
	const patient_data* pd = (const patient_data*)user;
	(void)t;
	dydt[0] = -(pd->k_absorption + pd->k_excretion) * y[0];
	dydt[1] = pd->k_absorption * y[0] - pd->k_elimination * y[1];
	return 0;
}
/* cpp:457-467 */
static int jac_one(double t, const double* y, const double* fy, double* J, void* user)
{
	const patient_data* pd = (const patient_data*)user;
	(void)t; (void)y; (void)fy;
	J[0 + 0 * 2] = -(pd->k_absorption + pd->k_excretion);
	J[1 + 0 * 2] = pd->k_absorption;
	J[1 + 1 * 2] = -pd->k_elimination;
	return 0;
}
/* cpp:469-479 */
static int rhs_two(double t, const double* y, double* dydt, void* user)
{
	const patient_data* pd = (const patient_data*)user;
	(void)t;
	dydt[0] = -(pd->k_absorption + pd->k_excretion) * y[0];
	dydt[1] = pd->k_absorption * y[0] - pd->k_elimination * y[1] - pd->k_periphery_fwd * y[1] + pd->k_periphery_bwd * y[2];
	dydt[2] = pd->k_periphery_fwd * y[1] - pd->k_periphery_bwd * y[2];
	return 0;
}
/* cpp:481-494 */
static int jac_two(double t, const double* y, const double* fy, double* J, void* user)
{
	const patient_data* pd = (const patient_data*)user;
	(void)t; (void)y; (void)fy;
	J[0 + 0 * 3] = -(pd->k_absorption + pd->k_excretion);
	J[1 + 0 * 3] = pd->k_absorption;
	J[1 + 1 * 3] = -(pd->k_elimination + pd->k_periphery_fwd);
	J[1 + 2 * 3] = pd->k_periphery_bwd;
	J[2 + 1 * 3] = pd->k_periphery_fwd;
	J[2 + 2 * 3] = -pd->k_periphery_bwd;
	return 0;
}

/* cpp:644-671 */
static int check_give_treatment(double t, const patient_data* pd)
{
	int give = 1;
	int day = (int)floor(t / 24.0);
	if (day >= 0 && day < 29 && ((pd->skipped_days >> day) & 1u)) give = 0;
	if (pd->intermittent == 1) {
		double tw = t - 7.0 * 24.0 * floor(t / (7.0 * 24.0));
		if (tw >= 5.0 * 24.0) give = 0;
	} else if (pd->intermittent == 2) {
		double tc = t - 28.0 * 24.0 * floor(t / (28.0 * 24.0));
		if (tc >= 21.0 * 24.0) give = 0;
	} else if (pd->intermittent == 3) {
		double tw = t - 7.0 * 24.0 * floor(t / (7.0 * 24.0));
		if (tw >= 4.0 * 24.0) give = 0;
	}
	return give;
}

static void add_counters(const bdf_mem* m, int64_t* cnt)
{
	cnt[ORACLE_CNT_NFE] += m->nfe;
	cnt[ORACLE_CNT_NSETUPS] += m->nsetups;
	cnt[ORACLE_CNT_NJE] += m->nje;
	cnt[ORACLE_CNT_NETF] += m->netf;
	cnt[ORACLE_CNT_NCFN] += m->ncfn;
	cnt[ORACLE_CNT_NNI] += m->nni;
}

/* ODESolver::SolveReturnSolution + ODESolverCVODE::Solve with the PopPK treatment callback.
 * out[N][ntp] column-major like OdeMatrixReal (out[i + tpi*N]); returns 1 on success. */
static int solve_patient(bdf_mem* m, patient_data* pd, int N, const double* y0, const double* tp, int ntp, int max_steps,
                         double* out, int64_t* cnt)
{
	/* ODESolver.cpp:109-118: timepoints at t ~ 0 get the initial condition */
	int ti = 0;
	while (tp[ti] < 2.220446049250313e-16) {
		for (int i = 0; i < N; i++) out[i + ti * N] = y0[i];
		ti++;
		if (ti == ntp) return 1;
	}
	double end_time = tp[ntp - 1];

	/* cpp:362-363: SetDiscontinuity(dosing_interval, ...) is ignored for time <= 0 (ODESolver.cpp:62-71);
	 * the stale value of a previous solve would be used by the reference -- not reproduced, inputs keep interval > 0 */
	double next_disc = pd->dosing_interval > 0.0 ? pd->dosing_interval : NAN;
	if (pd->biphasic) next_disc = pd->k_biphasic_switch_time > 0.0 ? pd->k_biphasic_switch_time : NAN; /* cpp:357-360 */

	double y[BDF_NMAX];
	for (int i = 0; i < N; i++) y[i] = y0[i];
	bdf_reinit(m, 0.0, y);
	if (!isnan(next_disc)) bdf_set_stop_time(m, next_disc);

	int current_step = 0;
	double t = 0.0;
	int tpi = ti;
	double tmp[BDF_NMAX];
	for (;;) {
		double tret;
		int result = bdf_step(m, end_time, y, &tret);
		if (result < 0) {
			if (cnt) { add_counters(m, cnt); cnt[ORACLE_CNT_STEPS] = current_step; }
			return 0;
		}
		t = tret;
		current_step++;

		while (tret >= tp[tpi]) {
			if (bdf_get_dky(m, tp[tpi], tmp) != BDF_SUCCESS) {
				if (cnt) { add_counters(m, cnt); cnt[ORACLE_CNT_STEPS] = current_step; }
				return 0;
			}
			for (int i = 0; i < N; i++) out[i + tpi * N] = tmp[i];
			tpi++;
			if (tpi >= ntp) break;
		}

		if (t >= end_time) break;

		if (current_step == max_steps) {
			if (cnt) { add_counters(m, cnt); cnt[ORACLE_CNT_STEPS] = current_step; }
			return 0;
		}

		if (!isnan(next_disc) && (result == BDF_TSTOP_RETURN || next_disc == t)) {
			if (cnt) add_counters(m, cnt);
			if (!pd->biphasic) {
				/* TreatmentCallback, cpp:673-690 */
				pd->current_dose_time += pd->dosing_interval;
				if (check_give_treatment(t, pd)) {
					double dose = pd->dose;
					if (t >= pd->dose_change_time) dose = pd->dose_after_dose_change;
					if (pd->transit) pd->last_treatment = t;
					else y[0] = y[0] + dose;
				}
				next_disc = pd->current_dose_time;
			} else if (pd->biphasic_switch) {
				/* TreatmentCallbackBiphasic, cpp:692-718 */
				pd->biphasic_switch = 0;
				pd->current_dose_time += pd->dosing_interval;
				next_disc = pd->current_dose_time;
			} else if (check_give_treatment(t, pd)) {
				double dose = pd->dose;
				if (t >= pd->dose_change_time) dose = pd->dose_after_dose_change;
				y[0] = y[0] + dose;
				pd->biphasic_switch = 1;
				next_disc = pd->current_dose_time + pd->k_biphasic_switch_time;
			} else {
				pd->current_dose_time += pd->dosing_interval;
				next_disc = pd->current_dose_time;
			}
			bdf_reinit(m, t, y);
			if (!isnan(next_disc) && next_disc < INFINITY) bdf_set_stop_time(m, next_disc);
		}
	}
	if (cnt) { add_counters(m, cnt); cnt[ORACLE_CNT_STEPS] = current_step; }
	return 1;
}

static void evaluate_chain(const oracle_poppk_problem* pr, const double* values, double* logp_out, double* conc,
                           double* patient_ll, int64_t* counters)
{
	const int P = pr->num_patients, T = pr->num_timepoints;
	const int two = ORACLE_PK_IS_TWO(pr->pk_type);
	const int biphasic = ORACLE_PK_IS_BIPHASIC(pr->pk_type), transit = ORACLE_PK_IS_TRANSIT(pr->pk_type);
	const int N = two ? 3 : 2;
	static const size_t pk_params_of_type[6] = { 4, 6, 7, 7, 6, 8 }; /* cpp:99-120 */
	const size_t npk = pk_params_of_type[pr->pk_type];
	const int report_all = conc || patient_ll || counters;

	patient_data pd;
	memset(&pd, 0, sizeof(pd));
	pd.two = two;
	pd.biphasic = biphasic;
	pd.transit = transit;
	bdf_mem* m = (bdf_mem*)malloc(sizeof(bdf_mem));
	if (biphasic || transit) bdf_create(m, N, rhs_variant, jac_variant, &pd);
	else bdf_create(m, N, two ? rhs_two : rhs_one, two ? jac_two : jac_one, &pd);
	double atolv[3] = { pr->atol, pr->atol, pr->atol };
	bdf_set_tolerances(m, pr->rtol, atolv);

	double* traj = (double*)malloc(sizeof(double) * N * (T > 0 ? T : 1));

	double logp = 0.0;
	int stopped = 0;
	const size_t sdix = (size_t)pr->sd_ix;
	double sd = transform_variable(pr->transforms[sdix], values[sdix]);
	double sd2 = transform_variable(pr->transforms[sdix + 1], values[sdix + 1]);

	for (int j = 0; j < P; j++) {
		if (conc)
			for (int i = 0; i < T; i++) conc[(size_t)j * T + i] = NAN;
		int64_t* cnt = counters ? counters + (size_t)j * ORACLE_NUM_COUNTERS : NULL;
		if (cnt) memset(cnt, 0, sizeof(int64_t) * ORACLE_NUM_COUNTERS);
		if (stopped && !report_all) break;

		pd.dose = pr->dose[j];
		pd.dosing_interval = pr->dosing_interval[j];
		pd.dose_after_dose_change = pr->dose_after_dose_change[j];
		pd.dose_change_time = pr->dose_change_time[j];
		pd.intermittent = (unsigned)pr->intermittent[j];
		if (pr->single) pd.intermittent = pr->intermittent[j] ? 1u : 0u;
		pd.skipped_days = pr->skipped_days[j];

		if (!pr->single) pd.k_absorption = fastpow10(quantile_normal(values[npk + 2 * (j + 1) + 0], values[0], values[npk + 0]));
		pd.k_excretion = transform_variable(pr->transforms[1], values[1]);
		pd.k_vod = isnan(pr->fixed_vod) ? transform_variable(pr->transforms[3], values[3]) : pr->fixed_vod;
		if (pr->single) {
			pd.k_absorption = transform_variable(pr->transforms[0], values[0]);
			pd.k_elimination = transform_variable(pr->transforms[2], values[2]) / pd.k_vod;
		} else
			pd.k_elimination = fastpow10(quantile_normal(values[npk + 2 * (j + 1) + 1], values[2], values[npk + 1])) / pd.k_vod;
		if (two) {
			if (isnan(pr->fixed_periphery_fwd)) {
				pd.k_periphery_fwd = transform_variable(pr->transforms[4], values[4]);
				pd.k_periphery_bwd = transform_variable(pr->transforms[5], values[5]);
			} else {
				pd.k_periphery_fwd = pr->fixed_periphery_fwd;
				pd.k_periphery_bwd = pr->fixed_periphery_bwd;
			}
		}
		if (transit) { /* cpp:296-301 */
			pd.n_transit = transform_variable(pr->transforms[pr->n_transit_ix], values[pr->n_transit_ix]);
			pd.k_transit = (pd.n_transit + 1) / transform_variable(pr->transforms[pr->mean_transit_time_ix], values[pr->mean_transit_time_ix]);
		}
		if (biphasic) { /* cpp:302-310 */
			pd.k_biphasic_switch_time = transform_variable(pr->transforms[pr->biphasic_uptake_time_ix], values[pr->biphasic_uptake_time_ix]);
			if (!pr->single && pd.dosing_interval - 1e-2 < pd.k_biphasic_switch_time) pd.k_biphasic_switch_time = pd.dosing_interval - 1e-2;
			pd.k_absorption2 = transform_variable(pr->transforms[pr->mean_absorption2_ix], values[pr->mean_absorption2_ix]);
		}
		pd.last_treatment = 0.0;
		pd.biphasic_switch = 1;                                   /* cpp:358 */
		pd.current_dose_time = biphasic ? 0.0 : pd.dosing_interval; /* cpp:359-362 */

		double y0[3] = { transit ? 0.0 : pd.dose, 0.0, 0.0 }; /* cpp:366-373 */
		double conversion = (1e6 / pr->mol_weight) / pd.k_vod;
		int ntp = pr->simulate_until[j];

		double patient_logllh = 0.0;
		if (ntp > 0) {
			int ok = solve_patient(m, &pd, N, y0, pr->time, ntp, pr->max_steps, traj, cnt);
			if (cnt) cnt[ORACLE_CNT_OK] = ok;
			if (!ok) {
				patient_logllh = -INFINITY;
			} else {
				for (int i = 0; i < ntp; i++) {
					double x = conversion * traj[1 + i * N];
					double yobs = pr->observed_concentration[(size_t)j * T + i];
					if (conc) conc[(size_t)j * T + i] = x;
					if (!isnan(yobs)) patient_logllh += logpdf_tnu4(x, yobs, sd + sd2 * (x > 0.0 ? x : 0.0));
					if (isnan(x) && !pr->single) {
						patient_logllh = -INFINITY;
						break;
					}
				}
			}
		}
		if (patient_ll) patient_ll[j] = patient_logllh;
		if (!stopped) {
			logp += patient_logllh;
			if (logp == -INFINITY) stopped = 1;
		}
	}
	*logp_out = logp;
	free(traj);
	free(m);
}

typedef struct {
	const oracle_poppk_problem* prob;
	size_t num_chains;
	const double* values;
	double *logp, *conc, *patient_ll;
	int64_t* counters;
	size_t* next;
	pthread_mutex_t* mu;
} work_t;

static void* worker(void* arg)
{
	work_t* w = (work_t*)arg;
	const size_t P = (size_t)w->prob->num_patients, T = (size_t)w->prob->num_timepoints, nvar = (size_t)w->prob->num_variables;
	for (;;) {
		pthread_mutex_lock(w->mu);
		size_t c = (*w->next)++;
		pthread_mutex_unlock(w->mu);
		if (c >= w->num_chains) break;
		evaluate_chain(w->prob, w->values + c * nvar, w->logp + c, w->conc ? w->conc + c * P * T : NULL,
		               w->patient_ll ? w->patient_ll + c * P : NULL,
		               w->counters ? w->counters + c * P * ORACLE_NUM_COUNTERS : NULL);
	}
	return NULL;
}

int oracle_poppk_evaluate(const oracle_poppk_problem* prob, size_t num_chains, const double* values, double* logp,
                          double* conc, double* patient_ll, int64_t* counters, int num_threads)
{
	if (!prob || !values || !logp) return -1;
	if (prob->pk_type < ORACLE_PK_ONE || prob->pk_type > ORACLE_PK_TWO_TRANSIT) return -2;
	if (num_threads < 1) num_threads = 1;
	if ((size_t)num_threads > num_chains) num_threads = (int)num_chains;
	size_t next = 0;
	pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
	work_t w = { prob, num_chains, values, logp, conc, patient_ll, counters, &next, &mu };
	if (num_threads <= 1) {
		worker(&w);
	} else {
		pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * num_threads);
		for (int i = 0; i < num_threads; i++) pthread_create(&th[i], NULL, worker, &w);
		for (int i = 0; i < num_threads; i++) pthread_join(th[i], NULL);
		free(th);
	}
	return 0;
}

const char* oracle_kind(void) { return "port"; }
