// pharmaco_kernel.cuh -- K1 of the pharmaco_population path: one (chain, patient) pair per thread.
//
// Reference: PharmacoLikelihoodPopulation::EvaluateLogProbability / SetupSimulation (src/pharmaco/
// PharmacoLikelihoodPopulation.cpp:202-340) around PharmacokineticModel::Solve / ConstructMatrix
// (src/pharmaco/PharmacokineticModel.cpp:111-247): a LINEAR compartment model -- depot, central, optional peripheral and
// transit compartments -- advanced from dose to dose with the matrix exponential, y(t + dt) = exp(A dt) y(t), each
// observation evaluated with its own exponential from the last dose. No ODE solver is involved; the exponential is
// Eigen's (unsupported/Eigen/src/MatrixFunctions/MatrixExponential.h:64-345, double): Pade approximants of degree 3, 5,
// 7, 9 chosen by the 1-norm, degree 13 with scaling and squaring above 2.0978, (V - U) X = (V + U) solved by LU with
// partial pivoting. The same degrees, coefficients and scaling are used here so that the results agree to round-off.
//
// The matrices are N x N with N = 2 + peripheral + transit compartments, N a template parameter: everything lives in
// registers for the common N = 2, 3. An exponential whose argument has the same bits as the previous one (successive
// dosing intervals of equal length) is not recomputed -- same input, same output.
#pragma once

#include <cfloat>
#include <cmath>
#include <cstdint>

namespace bcm3b200 {

struct PhArgs {
	int P_local, patient_offset, num_chains, nvar;
	const double* values;       // [C][nvar] sampled values as the sampler holds them (untransformed)
	const int32_t* transforms;  // [nvar] VariableSet transform codes
	// variable indices, -1 = not in the prior (PharmacoLikelihoodPopulation::PostInitialize, cpp:102-188)
	int additive_sd_ix, proportional_sd_ix, mean_absorption_ix, mean_excretion_ix, mean_clearance_ix, mean_vod_ix;
	int sigma_absorption_ix, sigma_excretion_ix, sigma_clearance_ix, sigma_vod_ix, sigma_transit_ix;
	int periph_fwd_ix, periph_bwd_ix, mean_transit_time_ix;
	int use_peripheral, num_transit, use_bioavailability;
	// model kind "pharmaco_single" (PharmacoLikelihoodSingle.cpp): ONE patient whose rates are the chain's variables themselves
	// (mean_*_ix hold the indices of "absorption", "excretion", "clearance", "volume_of_distribution"), optionally with the
	// direct-absorption route and the metabolite compartment that only this likelihood switches on
	int single, use_biphasic, use_metabolite, direct_absorption_ix, metabolite_conversion_ix;
	// per-patient marginal variables p<i>_<name> (InitializePatientMarginals, cpp:342-354): [P] variable indices, or null
	const int32_t *p_absorption_ix, *p_excretion_ix, *p_clearance_ix, *p_vod_ix, *p_transit_ix, *p_bioavailability_ix;
	double conv_base; // 1e6 / molecular weight (cpp:339)
	// per patient (global index): dose times / amounts and the observations with a value (Patient::Load, PharmacoPatient.cpp:8-116)
	const int32_t* treat_begin; // [P + 1]
	const double *treat_time, *treat_dose;
	const int32_t* obs_begin; // [P + 1]
	const double *obs_time, *obs_value;
	const int32_t* obs_grid; // index of the observation in the trial's time grid (diagnostics)
	int T;                   // size of that grid
	double* patient_ll;      // [C][P_local]
	double* diag_conc;       // [C][P_local][T] or null: conversion * simulated concentration, NaN where nothing was observed
};

__device__ __forceinline__ double ph_transform(int tr, double x)
{
	// VariableSet::TransformVariable, VariableSet.cpp:97-124
	switch (tr) {
	case 1: return exp(x);
	case 2: return exp(x * 2.3025850929940459);
	case 3:
		if (x > 0) {
			const double z = exp(-x);
			return 1.0 / (1.0 + z);
		} else {
			const double z = exp(x);
			return z / (1.0 + z);
		}
	default: return x;
	}
}
__device__ __forceinline__ double ph_pow10(double x) { return exp(x * 2.3025850929940459); } // bcm3::fastpow10, MathFunctions.h:13

template <int N>
struct PhMat {
	double m[N * N];
	__device__ __forceinline__ double& operator()(int i, int j) { return m[i * N + j]; }
	__device__ __forceinline__ double operator()(int i, int j) const { return m[i * N + j]; }
};

template <int N>
__device__ __forceinline__ void ph_mul(const PhMat<N>& a, const PhMat<N>& b, PhMat<N>& out)
{
#pragma unroll
	for (int i = 0; i < N; i++)
#pragma unroll
		for (int j = 0; j < N; j++) {
			double s = a(i, 0) * b(0, j);
#pragma unroll
			for (int k = 1; k < N; k++) s = fma(a(i, k), b(k, j), s);
			out(i, j) = s;
		}
}

// out = c2 * A2 + c4 * A4 + c6 * A6 + c0 * I (terms with a zero coefficient are skipped at compile time by the callers' constants)
template <int N>
__device__ __forceinline__ void ph_poly(const PhMat<N>& A2, const PhMat<N>& A4, const PhMat<N>& A6, double c6, double c4, double c2, double c0, PhMat<N>& out)
{
#pragma unroll
	for (int i = 0; i < N; i++)
#pragma unroll
		for (int j = 0; j < N; j++) {
			double v = c6 * A6(i, j) + c4 * A4(i, j) + c2 * A2(i, j);
			if (i == j) v += c0;
			out(i, j) = v;
		}
}

// X = (V - U)^-1 (V + U): PartialPivLU of the denominator, then the two triangular solves, column by column
template <int N>
__device__ __forceinline__ void ph_pade_solve(const PhMat<N>& U, const PhMat<N>& V, PhMat<N>& X)
{
	PhMat<N> D;
	int perm[N];
#pragma unroll
	for (int i = 0; i < N; i++) {
		perm[i] = i;
#pragma unroll
		for (int j = 0; j < N; j++) {
			D(i, j) = V(i, j) - U(i, j);
			X(i, j) = V(i, j) + U(i, j);
		}
	}
#pragma unroll
	for (int k = 0; k < N; k++) {
		int piv = k;
		double best = fabs(D(k, k));
#pragma unroll
		for (int i = k + 1; i < N; i++) {
			const double v = fabs(D(i, k));
			if (v > best) {
				best = v;
				piv = i;
			}
		}
		if (piv != k) {
#pragma unroll
			for (int j = 0; j < N; j++) {
				// register arrays: the swap runs over every row pair with a select, no dynamic indexing
#pragma unroll
				for (int i = k + 1; i < N; i++) {
					if (i == piv) {
						const double t = D(k, j);
						D(k, j) = D(i, j);
						D(i, j) = t;
						const double tx = X(k, j);
						X(k, j) = X(i, j);
						X(i, j) = tx;
					}
				}
			}
		}
		const double inv = 1.0 / D(k, k);
#pragma unroll
		for (int i = k + 1; i < N; i++) {
			const double l = D(i, k) * inv;
			D(i, k) = l;
#pragma unroll
			for (int j = k + 1; j < N; j++) D(i, j) = fma(-l, D(k, j), D(i, j));
#pragma unroll
			for (int j = 0; j < N; j++) X(i, j) = fma(-l, X(k, j), X(i, j)); // forward substitution on all right-hand sides
		}
	}
	(void)perm;
#pragma unroll
	for (int k = N - 1; k >= 0; k--) {
		const double inv = 1.0 / D(k, k);
#pragma unroll
		for (int j = 0; j < N; j++) {
			double s = X(k, j);
#pragma unroll
			for (int i = k + 1; i < N; i++) s = fma(-D(k, i), X(i, j), s);
			X(k, j) = s * inv;
		}
	}
}

// exp(M), MatrixExponential.h:230-345 (matrix_exp_computeUV<double>, matrix_exp_compute)
template <int N>
__device__ __noinline__ void ph_expm(const PhMat<N>& M, PhMat<N>& R)
{
	double l1 = 0.0; // arg.cwiseAbs().colwise().sum().maxCoeff()
#pragma unroll
	for (int j = 0; j < N; j++) {
		double s = 0.0;
#pragma unroll
		for (int i = 0; i < N; i++) s += fabs(M(i, j));
		l1 = fmax(l1, s);
	}
	PhMat<N> A = M, A2, A4, A6, U, V, tmp;
	int squarings = 0;
	if (!(l1 < 2.097847961257068e+000)) {
		const double maxnorm = 5.371920351148152;
		(void)frexp(l1 / maxnorm, &squarings);
		if (squarings < 0) squarings = 0;
		const double sc = ldexp(1.0, -squarings); // MatrixExponentialScalingOp: ldexp(x, -squarings), exact
#pragma unroll
		for (int i = 0; i < N * N; i++) A.m[i] = M.m[i] * sc;
	}
	ph_mul(A, A, A2);
	if (l1 < 1.495585217958292e-002) { // pade3
		ph_poly(A2, A2, A2, 0.0, 0.0, 1.0, 60.0, tmp);
		ph_mul(A, tmp, U);
		ph_poly(A2, A2, A2, 0.0, 0.0, 12.0, 120.0, V);
	} else if (l1 < 2.539398330063230e-001) { // pade5
		ph_mul(A2, A2, A4);
		ph_poly(A2, A4, A4, 0.0, 1.0, 420.0, 15120.0, tmp);
		ph_mul(A, tmp, U);
		ph_poly(A2, A4, A4, 0.0, 30.0, 3360.0, 30240.0, V);
	} else if (l1 < 9.504178996162932e-001) { // pade7
		ph_mul(A2, A2, A4);
		ph_mul(A4, A2, A6);
		ph_poly(A2, A4, A6, 1.0, 1512.0, 277200.0, 8648640.0, tmp);
		ph_mul(A, tmp, U);
		ph_poly(A2, A4, A6, 56.0, 25200.0, 1995840.0, 17297280.0, V);
	} else if (l1 < 2.097847961257068e+000) { // pade9
		PhMat<N> A8;
		ph_mul(A2, A2, A4);
		ph_mul(A4, A2, A6);
		ph_mul(A6, A2, A8);
		ph_poly(A2, A4, A6, 3960.0, 2162160.0, 302702400.0, 8821612800.0, tmp);
#pragma unroll
		for (int i = 0; i < N * N; i++) tmp.m[i] += 1.0 * A8.m[i];
		ph_mul(A, tmp, U);
		ph_poly(A2, A4, A6, 110880.0, 30270240.0, 2075673600.0, 17643225600.0, V);
#pragma unroll
		for (int i = 0; i < N * N; i++) V.m[i] += 90.0 * A8.m[i];
	} else { // pade13 on the scaled matrix
		ph_mul(A2, A2, A4);
		ph_mul(A4, A2, A6);
		ph_poly(A2, A4, A6, 1.0, 16380.0, 40840800.0, 0.0, V); // b13 A6 + b11 A4 + b9 A2
		ph_mul(A6, V, tmp);
		PhMat<N> w;
		ph_poly(A2, A4, A6, 33522128640.0, 10559470521600.0, 1187353796428800.0, 32382376266240000.0, w);
#pragma unroll
		for (int i = 0; i < N * N; i++) tmp.m[i] += w.m[i];
		ph_mul(A, tmp, U);
		ph_poly(A2, A4, A6, 182.0, 960960.0, 1323241920.0, 0.0, tmp); // b12 A6 + b10 A4 + b8 A2
		ph_mul(A6, tmp, V);
		ph_poly(A2, A4, A6, 670442572800.0, 129060195264000.0, 7771770303897600.0, 64764752532480000.0, w);
#pragma unroll
		for (int i = 0; i < N * N; i++) V.m[i] += w.m[i];
	}
	ph_pade_solve(U, V, R);
#pragma unroll 1
	for (int s = 0; s < squarings; s++) {
		ph_mul(R, R, tmp);
		R = tmp;
	}
}

// PharmacokineticModel::ConstructMatrix, PharmacokineticModel.cpp:188-247 (the metabolite compartment and the direct
// absorption route are switched on by the single-patient likelihood only)
template <int N>
__device__ __forceinline__ void ph_construct(PhMat<N>& A, double absorption, double excretion, double elimination, bool periph, double kf, double kb,
                                             int ntransit, double transit_rate, bool biphasic = false, double direct_absorption = 0.0,
                                             bool metabolite = false, double metabolite_conversion = 0.0, double metabolite_elimination = 0.0)
{
#pragma unroll
	for (int i = 0; i < N * N; i++) A.m[i] = 0.0;
	const int metabolite_ix = periph ? 3 : 2;
	const int first_transit = metabolite_ix + (metabolite ? 1 : 0);
	A(0, 0) -= excretion;
	A(0, 0) -= absorption;
	if (ntransit > 0) {
		// indices are run-time values bounded by N: the selects keep the matrix in registers
		auto add = [&](int r, int c, double v, bool assign) {
#pragma unroll
			for (int i = 0; i < N; i++)
#pragma unroll
				for (int j = 0; j < N; j++)
					if (i == r && j == c) A(i, j) = assign ? v : A(i, j) + v;
		};
		add(first_transit, 0, absorption, false);
		if (ntransit > 2) { // the reference's condition (:215): with one or two transit compartments the chain is not linked
			for (int i = 0; i < ntransit - 1; i++) {
				add(first_transit + i, first_transit + i, -transit_rate, false);
				add(first_transit + i + 1, first_transit + i, transit_rate, false);
			}
		}
		add(first_transit + ntransit - 1, first_transit + ntransit - 1, -transit_rate, true);
		add(1, first_transit + ntransit - 1, transit_rate, false);
	} else {
		A(1, 0) += absorption;
	}
	if (periph) {
		if constexpr (N >= 3) {
			A(1, 1) -= kf;
			A(2, 1) += kf;
			A(1, 2) += kb;
			A(2, 2) -= kb;
		}
	}
	if (biphasic) { // :233-236
		A(0, 0) -= direct_absorption;
		A(1, 0) += direct_absorption;
	}
	if (metabolite) { // :238-242
		A(1, 1) -= metabolite_conversion;
#pragma unroll
		for (int i = 2; i < N; i++)
			if (i == metabolite_ix) {
				A(i, 1) += metabolite_conversion;
				A(i, i) -= metabolite_elimination;
			}
	}
	A(1, 1) -= elimination;
}

template <int N>
__global__ void __launch_bounds__(128) pharmaco_kernel(const PhArgs a)
{
	const int jl = blockIdx.x * blockDim.x + threadIdx.x;
	const int c = blockIdx.y;
	if (jl >= a.P_local) return;
	const int j = a.patient_offset + jl;
	const double* v = a.values + (long long)c * a.nvar;
	auto tv = [&](int ix) { return ph_transform(a.transforms[ix], v[ix]); };
	// QuantileNormal(p, mu, sigma) = mu + sigma * Phi^-1(p) (ProbabilityDistributions.cpp, boost::math::quantile(normal))
	auto marginal = [&](int mean_ix, int sigma_ix, const int32_t* pix) {
		if (sigma_ix < 0) return ph_pow10(v[mean_ix]);
		return ph_pow10(fma(v[sigma_ix], normcdfinv(v[pix[j]]), v[mean_ix]));
	};
	// ---- SetupSimulation, cpp:259-340 ----
	const double additive_sd = a.additive_sd_ix >= 0 ? tv(a.additive_sd_ix) : 0.0;
	const double proportional_sd = a.proportional_sd_ix >= 0 ? tv(a.proportional_sd_ix) : 0.0;
	// pharmaco_single (PharmacoLikelihoodSingle.cpp:163-178): the variables, transformed, are the rates
	const double absorption = a.single ? tv(a.mean_absorption_ix) : marginal(a.mean_absorption_ix, a.sigma_absorption_ix, a.p_absorption_ix);
	const double excretion = a.mean_excretion_ix >= 0 ? (a.single ? tv(a.mean_excretion_ix) : marginal(a.mean_excretion_ix, a.sigma_excretion_ix, a.p_excretion_ix)) : 0.0;
	const double clearance = a.single ? tv(a.mean_clearance_ix) : marginal(a.mean_clearance_ix, a.sigma_clearance_ix, a.p_clearance_ix);
	const double vod = a.single ? tv(a.mean_vod_ix) : marginal(a.mean_vod_ix, a.sigma_vod_ix, a.p_vod_ix);
	double kf = 0.0, kb = 0.0, transit_rate = 0.0, bioavailability = 1.0;
	if (a.use_peripheral) {
		kf = tv(a.periph_fwd_ix);
		kb = tv(a.periph_bwd_ix);
	}
	if (a.num_transit > 0) {
		const double transit_time = (a.sigma_transit_ix < 0) ? tv(a.mean_transit_time_ix)
		                                                      : ph_pow10(fma(v[a.sigma_transit_ix], normcdfinv(v[a.p_transit_ix[j]]), v[a.mean_transit_time_ix]));
		transit_rate = (a.num_transit + 1.0) / transit_time;
	}
	if (a.use_bioavailability) bioavailability = v[a.p_bioavailability_ix[j]];
	const double conversion = a.conv_base / vod;
	PhMat<N> A;
	const double direct_absorption = a.use_biphasic ? tv(a.direct_absorption_ix) : 0.0;             // Single.cpp:192-195
	const double metabolite_conversion = a.use_metabolite ? tv(a.metabolite_conversion_ix) : 0.0;   // :196-199; elimination fixed at 1 (:143)
	ph_construct(A, absorption, excretion, clearance / vod, a.use_peripheral != 0, kf, kb, a.num_transit, transit_rate, a.use_biphasic != 0, direct_absorption,
	             a.use_metabolite != 0, metabolite_conversion, 1.0);

	// ---- PharmacokineticModel::Solve, PharmacokineticModel.cpp:111-177 ----
	const int t0 = a.treat_begin[j], t1 = a.treat_begin[j + 1], o0 = a.obs_begin[j], o1 = a.obs_begin[j + 1];
	double y[N];
#pragma unroll
	for (int i = 0; i < N; i++) y[i] = 0.0;
	double ll = 0.0;
	bool ok = true, broken = false;
	double* conc = a.diag_conc ? a.diag_conc + ((long long)c * a.P_local + jl) * a.T : nullptr;
	if (conc)
		for (int i = 0; i < a.T; i++) conc[i] = __longlong_as_double(0x7ff8000000000000ll);
	if (o1 > o0) {
		const double simulate_until = a.obs_time[o1 - 1];
		int oti = o0;
		double current_t = 0.0, last_dt = __longlong_as_double(0x7ff8000000000000ll);
		PhMat<N> E, Edt;
		for (int tti = t0; tti < t1 && current_t < simulate_until; tti++) {
			const double target_t = (tti < t1 - 1) ? a.treat_time[tti + 1] : simulate_until;
			y[0] += a.treat_dose[tti] * bioavailability;
			while (oti < o1 && a.obs_time[oti] <= target_t) {
				const double offset_t = a.obs_time[oti] - current_t;
				PhMat<N> Mt;
#pragma unroll
				for (int i = 0; i < N * N; i++) Mt.m[i] = A.m[i] * offset_t;
				ph_expm(Mt, E);
				double central = E(1, 0) * y[0];
#pragma unroll
				for (int k = 1; k < N; k++) central = fma(E(1, k), y[k], central);
				// cpp:222-236
				const double x = conversion * central;
				if (conc) conc[a.obs_grid[oti]] = x;
				if (!broken) {
					if (isnan(x) || isinf(x)) {
						ll = -INFINITY;
						broken = true;
					} else {
						const double yobs = a.obs_value[oti];
						const double sigma = additive_sd + proportional_sd * fmax(x, 0.0);
						const double xn = (x - yobs) / sigma;
						ll += -0.9808292530117262 - 2.5 * log1p(0.25 * xn * xn) - log(sigma); // bcm3::LogPdfTnu4
					}
				}
				oti++;
			}
			const double dt = target_t - current_t;
			if (!(dt == last_dt)) { // the same bits give the same exponential
				PhMat<N> Mt;
#pragma unroll
				for (int i = 0; i < N * N; i++) Mt.m[i] = A.m[i] * dt;
				ph_expm(Mt, Edt);
				last_dt = dt;
			}
			double yn[N];
#pragma unroll
			for (int i = 0; i < N; i++) {
				double s = Edt(i, 0) * y[0];
#pragma unroll
				for (int k = 1; k < N; k++) s = fma(Edt(i, k), y[k], s);
				yn[i] = s;
			}
			bool nan = false;
#pragma unroll
			for (int i = 0; i < N; i++) {
				nan = nan || isnan(yn[i]);
				y[i] = yn[i];
			}
			if (nan) { // Solve returns false: the patient's term is -inf (cpp:238-240)
				ok = false;
				break;
			}
			current_t = target_t;
		}
	}
	if (!ok) ll = -INFINITY;
	a.patient_ll[(long long)c * a.P_local + jl] = ll;
}

} // namespace bcm3b200
