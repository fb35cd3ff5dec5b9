// cellpop_prelude.cuh -- device restatement of the helper functions the reference prepends to its generated
// derivative code (src/cellpop/SolverCodeGenerator.cpp:122-295), including their quirks (SURVEY.md App. D #6:
// hill_function_fixedn2 returns 10.0 on overflow, hill_function_fixedn16 has a no-op overflow test). Only the
// helpers used by generated_derivative are needed: the generated Jacobian is never installed by the reference
// (src/cellpop/Cell.cpp:57-76), so its *_derivative helpers are omitted.
#pragma once

#include <cfloat>
#include <cmath>

typedef double OdeReal;

// The rate-law helpers are real functions by default: a 50-species model calls them ~100 times per right-hand side, and
// inlined (divisions and square roots expand to dozens of instructions each) the right-hand side alone outgrows the
// 32 KB instruction cache. CP_HELPER_INLINE=1 restores inlining.
#ifndef CP_HELPER_INLINE
#define CP_HELPER_INLINE 0
#endif
#if CP_HELPER_INLINE
#define CP_HELPER __device__ __forceinline__
#else
#define CP_HELPER __device__ __noinline__
#endif

#define CP_REAL_MIN DBL_MIN /* std::numeric_limits<OdeReal>::min() */

__device__ __forceinline__ OdeReal square(OdeReal x) { return x * x; }

CP_HELPER OdeReal hill_function(OdeReal x, OdeReal k, OdeReal n)
{
	if (x <= 0.0) return 0.0;
	OdeReal xn = pow(x, n);
	OdeReal kn = pow(k, n);
	OdeReal xnpkn = xn + kn;
	if (xnpkn < CP_REAL_MIN) return 0.0;
	if (xnpkn > 3e38f) return 1.0;
	return xn / xnpkn;
}
CP_HELPER OdeReal hill_function_fixedn2(OdeReal x, OdeReal k)
{
	if (x <= 0.0) return 0.0;
	OdeReal x2 = x * x;
	OdeReal k2 = k * k;
	OdeReal xnpkn = x2 + k2;
	if (xnpkn < CP_REAL_MIN) return 0.0;
	if (xnpkn > 3e38f) return 10.0;
	return x2 / xnpkn;
}
CP_HELPER OdeReal hill_function_fixedn4(OdeReal x, OdeReal k)
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
CP_HELPER OdeReal hill_function_fixedn10(OdeReal x, OdeReal k)
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
CP_HELPER OdeReal hill_function_fixedn16(OdeReal x, OdeReal k)
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
CP_HELPER OdeReal hill_function_fixedn100(OdeReal x, OdeReal k)
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
CP_HELPER OdeReal michaelis_menten_function(OdeReal kcat, OdeReal KM, OdeReal e, OdeReal s)
{
	if (e <= 0) return 0.0;
	if (s + KM < 0.1 * KM) {
		OdeReal bound = -KM + 0.1 * KM;
		OdeReal offset = (e * kcat * bound / (0.01 * KM) - e * kcat * bound / (KM + bound));
		return e * kcat * s / (0.01 * KM) - offset;
	}
	return kcat * e * s / (KM + s);
}
CP_HELPER OdeReal safepow(OdeReal x, OdeReal n)
{
	if (x <= 0) {
		return 0.0;
	} else {
		return pow(x, n);
	}
}
CP_HELPER OdeReal synthcap(OdeReal x)
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
CP_HELPER OdeReal tQSSA(OdeReal k, OdeReal km, OdeReal e, OdeReal s)
{
	OdeReal ekms = e + km + s;
	return 0.5 * k * (ekms - sqrt(ekms * ekms - 4 * e * s));
}
