// Deterministic per-chain random streams. The reference seeds nothing reproducibly (its worker RNGs are per THREAD
// and copies are seeded, SURVEY.md App. D #13: Sampler.cpp:95-98, SamplerPTChain.cpp:219,223), so which random
// numbers a chain sees depends on task scheduling. Here every chain (and the swap logic) owns a counter-based stream
// keyed by (seed, stream id): results do not depend on thread counts or on whether the evaluation is batched.
#pragma once

#include <cmath>
#include <cstdint>

namespace bcm3 {

class RNG {
public:
	RNG(uint64_t seed = 0, uint64_t stream = 0) { Seed(seed, stream); }
	void Seed(uint64_t seed, uint64_t stream)
	{
		key = mix(seed ^ 0x9E3779B97F4A7C15ull) ^ mix(stream * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull);
		counter = 0;
		have_spare = false;
	}
	uint64_t GetUnsigned64() { return mix(key + 0x9E3779B97F4A7C15ull * ++counter); }
	// uniform in [0, 1) with 53 random bits
	double GetReal() { return (GetUnsigned64() >> 11) * (1.0 / 9007199254740992.0); }
	unsigned int GetUnsignedInt(unsigned int max_inclusive) { return (unsigned int)(GetUnsigned64() % ((uint64_t)max_inclusive + 1)); }
	double GetUniform(double lower, double upper) { return lower + (upper - lower) * GetReal(); }
	double GetNormal(double mu = 0.0, double sigma = 1.0)
	{
		if (have_spare) {
			have_spare = false;
			return mu + sigma * spare;
		}
		double u1, u2;
		do { u1 = GetReal(); } while (u1 <= 0.0);
		u2 = GetReal();
		const double r = std::sqrt(-2.0 * std::log(u1));
		spare = r * std::sin(6.283185307179586 * u2);
		have_spare = true;
		return mu + sigma * r * std::cos(6.283185307179586 * u2);
	}

	// RNG::GetGamma (src/utils/RNG.cpp:84-111): Marsaglia-Tsang, shape k, SCALE theta
	double GetGamma(double k, double theta)
	{
		if (k < 1) {
			const double u = GetReal();
			return GetGamma(1.0 + k, theta) * std::pow(u, 1.0 / k);
		}
		const double d = k - 0.33333333333333333333333333333333;
		const double c = 0.33333333333333333333333333333333 / std::sqrt(d);
		double x, v, u;
		for (;;) {
			do {
				x = GetNormal(0.0, 1.0);
				v = 1.0 + c * x;
			} while (v <= 0.0);
			v = v * v * v;
			u = GetReal();
			if (u < 1 - 0.0331 * x * x * x * x) break;
			if (std::log(u) < 0.5 * x * x + d * (1 - v + std::log(v))) break;
		}
		return theta * d * v;
	}

private:
	static uint64_t mix(uint64_t z)
	{
		// splitmix64 finaliser
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
		return z ^ (z >> 31);
	}
	uint64_t key = 0, counter = 0;
	bool have_spare = false;
	double spare = 0.0;
};

} // namespace bcm3
