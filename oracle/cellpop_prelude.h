// TEST INFRASTRUCTURE ONLY -- cellpop_prelude.h -- host restatement of the helper functions the reference prepends to its generated
// derivative code (src/cellpop/SolverCodeGenerator.cpp:122-295), including their quirks (SURVEY.md App. D #6:
// hill_function_fixedn2 returns 10.0 on overflow, hill_function_fixedn16 has a no-op overflow test). The
// *_derivative helpers are only called by generated_jacobian, which the reference compiles into the same file and never
// installs (src/cellpop/Cell.cpp:57-76); they are here so that real generator output compiles as it does on the reference side.
#pragma once

#include <cfloat>
#include <cmath>

typedef double OdeReal;

#define CP_REAL_MIN DBL_MIN /* std::numeric_limits<OdeReal>::min() */

static inline OdeReal square(OdeReal x) { return x * x; }

static inline OdeReal hill_function(OdeReal x, OdeReal k, OdeReal n)
{
	if (x <= 0.0) return 0.0;
	OdeReal xn = pow(x, n);
	OdeReal kn = pow(k, n);
	OdeReal xnpkn = xn + kn;
	if (xnpkn < CP_REAL_MIN) return 0.0;
	if (xnpkn > 3e38f) return 1.0;
	return xn / xnpkn;
}
static inline OdeReal hill_function_fixedn2(OdeReal x, OdeReal k)
{
	if (x <= 0.0) return 0.0;
	OdeReal x2 = x * x;
	OdeReal k2 = k * k;
	OdeReal xnpkn = x2 + k2;
	if (xnpkn < CP_REAL_MIN) return 0.0;
	if (xnpkn > 3e38f) return 10.0;
	return x2 / xnpkn;
}
static inline OdeReal hill_function_fixedn4(OdeReal x, OdeReal k)
{
	if (x <= 0.0) return 0.0;
	OdeReal x2 = x * x;
	OdeReal x4 = x2 * x2;
	OdeReal k2 = k * k;
	OdeReal k4 = k2 * k2;
	OdeReal xnpkn = x4 + k4;
	if (xnpkn < CP_REAL_MIN) return 0.0;
	if (xnpkn > 3e38f) return 1.0;
	return x4 / xnpkn;
}
static inline OdeReal hill_function_fixedn10(OdeReal x, OdeReal k)
{
	if (x <= 0.0) return 0.0;
	OdeReal x2 = x * x;
	OdeReal x4 = x2 * x2;
	OdeReal x8 = x4 * x4;
	OdeReal x10 = x2 * x8;
	OdeReal k2 = k * k;
	OdeReal k4 = k2 * k2;
	OdeReal k8 = k4 * k4;
	OdeReal k10 = k2 * k8;
	OdeReal xnpkn = x10 + k10;
	if (xnpkn < CP_REAL_MIN) return 0.0;
	if (xnpkn > 3e38f) return 1.0;
	return x10 / xnpkn;
}
static inline OdeReal hill_function_fixedn16(OdeReal x, OdeReal k)
{
	if (x <= 0.0) return 0.0;
	OdeReal x2 = x * x;
	OdeReal x4 = x2 * x2;
	OdeReal x8 = x4 * x4;
	OdeReal x16 = x8 * x8;
	OdeReal k2 = k * k;
	OdeReal k4 = k2 * k2;
	OdeReal k8 = k4 * k4;
	OdeReal k16 = k8 * k8;
	OdeReal xnpkn = x16 + k16;
	if (xnpkn < CP_REAL_MIN) return 0.0;
	// the reference's `if (xnpkn > 3e38f) 1.0;` has no `return`: no effect
	return x16 / xnpkn;
}
static inline OdeReal hill_function_fixedn100(OdeReal x, OdeReal k)
{
	if (x <= 0.0) return 0.0;
	OdeReal x2 = x * x;
	OdeReal x4 = x2 * x2;
	OdeReal x8 = x4 * x4;
	OdeReal x16 = x8 * x8;
	OdeReal x32 = x16 * x16;
	OdeReal x64 = x32 * x32;
	OdeReal x100 = x64 * x32 * x4;
	OdeReal k2 = k * k;
	OdeReal k4 = k2 * k2;
	OdeReal k8 = k4 * k4;
	OdeReal k16 = k8 * k8;
	OdeReal k32 = k16 * k16;
	OdeReal k64 = k32 * k32;
	OdeReal k100 = k64 * k32 * k4;
	OdeReal xnpkn = x100 + k100;
	if (xnpkn < CP_REAL_MIN) return 0.0;
	if (xnpkn > 3e38f) return 1.0;
	return x100 / xnpkn;
}
static inline OdeReal michaelis_menten_function(OdeReal kcat, OdeReal KM, OdeReal e, OdeReal s)
{
	if (e <= 0) return 0.0;
	if (s + KM < 0.1 * KM) {
		OdeReal bound = -KM + 0.1 * KM;
		OdeReal offset = (e * kcat * bound / (0.01 * KM) - e * kcat * bound / (KM + bound));
		return e * kcat * s / (0.01 * KM) - offset;
	}
	return kcat * e * s / (KM + s);
}
static inline OdeReal safepow(OdeReal x, OdeReal n)
{
	if (x <= 0) {
		return 0.0;
	} else {
		return pow(x, n);
	}
}
static inline OdeReal synthcap(OdeReal x)
{
	if (x <= 0) {
		return 1.0;
	} else {
		OdeReal x2 = x * x;
		OdeReal x4 = x2 * x2;
		OdeReal x8 = x4 * x4;
		return 1.0 - x8 * x2;
	}
}
static inline OdeReal tQSSA(OdeReal k, OdeReal km, OdeReal e, OdeReal s)
{
	OdeReal ekms = e + km + s;
	return 0.5 * k * (ekms - sqrt(ekms * ekms - 4 * e * s));
}

// ---- helpers of generated_jacobian (SolverCodeGenerator.cpp:218-230, 241-258, 268-279, 285-294); compiled, never called ----
static inline OdeReal hill_function_derivative(OdeReal x, OdeReal k, OdeReal n)
{
	if (x <= 0.0) return 0.0;
	OdeReal xn = pow(x, n);
	OdeReal kn = pow(k, n);
	OdeReal denom = (square(xn + kn));
	if (denom < CP_REAL_MIN) return 0.0;
	if (denom > 3e38f) return 0.0;
	OdeReal xnm1 = pow(x, n - 1);
	return kn * n * xnm1 / denom;
}
static inline OdeReal michaelis_menten_derivative_enzyme(OdeReal kcat, OdeReal KM, OdeReal e, OdeReal s)
{
	if (e <= 0) return 0.0;
	if (s + KM < 0.1 * KM) {
		OdeReal bound = -KM + 0.1 * KM;
		OdeReal offset = (kcat * bound / (0.01 * KM) - kcat * bound / (KM + bound));
		return kcat * s / (0.01 * KM) - offset;
	}
	return kcat * s / (KM + s);
}
static inline OdeReal michaelis_menten_derivative_substrate(OdeReal kcat, OdeReal KM, OdeReal e, OdeReal s)
{
	if (e <= 0) return 0.0;
	if (s + KM <= 0.1 * KM) return e * kcat / (0.01 * KM);
	return e * kcat * KM / (square(KM + s));
}
static inline OdeReal synthcap_derivative(OdeReal x, OdeReal dx)
{
	if (x <= 0) return 0.0;
	OdeReal x2 = x * x;
	OdeReal x4 = x2 * x2;
	OdeReal x8 = x4 * x4;
	return -10.0 * x8 * x * dx;
}
static inline OdeReal tQSSA_derivative_enzyme(OdeReal k, OdeReal km, OdeReal e, OdeReal s)
{
	OdeReal ekms = e + km + s;
	return k * (0.5 - 0.5 * (km - s + e) / (sqrt(ekms * ekms - 4 * e * s)));
}
static inline OdeReal tQSSA_derivative_substrate(OdeReal k, OdeReal km, OdeReal e, OdeReal s)
{
	OdeReal ekms = e + km + s;
	return k * (0.5 - 0.5 * (km + s - e) / (sqrt(ekms * ekms - 4 * e * s)));
}
