#include "GaussianMixture.h"

#include <algorithm>
#include <set>

namespace bcm3 {

namespace {
const Real kLog2Pi = 1.8378770664093454835606594728112;
}

size_t SampleIndex(RNG& rng, const VectorReal& probabilities)
{
	const Real t = rng.GetReal();
	Real p = 0.0;
	for (size_t i = 0; i < probabilities.size(); i++) {
		p += probabilities[i];
		if (t < p) return i;
	}
	return probabilities.empty() ? 0 : probabilities.size() - 1;
}

// ------------------------------------------------------------------------------------------------ dense helpers
bool GaussianMixture::Cholesky(const std::vector<Real>& a, size_t n, std::vector<Real>& l)
{
	l.assign(n * n, 0.0);
	for (size_t j = 0; j < n; j++) {
		Real d = a[j + j * n];
		for (size_t k = 0; k < j; k++) d -= l[j + k * n] * l[j + k * n];
		if (!(d > 0.0)) return false;
		const Real ljj = sqrt(d);
		l[j + j * n] = ljj;
		for (size_t i = j + 1; i < n; i++) {
			Real v = a[i + j * n];
			for (size_t k = 0; k < j; k++) v -= l[i + k * n] * l[j + k * n];
			l[i + j * n] = v / ljj;
		}
	}
	return true;
}

void GaussianMixture::SolveLower(const std::vector<Real>& l, size_t n, VectorReal& v)
{
	for (size_t i = 0; i < n; i++) {
		Real s = v[i];
		for (size_t k = 0; k < i; k++) s -= l[i + k * n] * v[k];
		v[i] = s / l[i + i * n];
	}
}

// cyclic Jacobi rotations; fine for the few tens of dimensions of a sampler block
void GaussianMixture::SymmetricEigen(std::vector<Real> a, size_t n, VectorReal& values, std::vector<Real>& vectors)
{
	vectors.assign(n * n, 0.0);
	for (size_t i = 0; i < n; i++) vectors[i + i * n] = 1.0;
	for (int sweep = 0; sweep < 64; sweep++) {
		Real off = 0.0, diag = 0.0;
		for (size_t i = 0; i < n; i++) {
			diag += a[i + i * n] * a[i + i * n];
			for (size_t j = i + 1; j < n; j++) off += a[i + j * n] * a[i + j * n];
		}
		if (off <= 1e-30 * (diag + 1e-300)) break;
		for (size_t p = 0; p + 1 < n; p++)
			for (size_t q = p + 1; q < n; q++) {
				const Real apq = a[p + q * n];
				if (apq == 0.0) continue;
				const Real theta = (a[q + q * n] - a[p + p * n]) / (2.0 * apq);
				const Real t = ((theta >= 0.0) ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
				const Real c = 1.0 / sqrt(t * t + 1.0), s = t * c;
				for (size_t k = 0; k < n; k++) { // columns p, q
					const Real akp = a[k + p * n], akq = a[k + q * n];
					a[k + p * n] = c * akp - s * akq;
					a[k + q * n] = s * akp + c * akq;
				}
				for (size_t k = 0; k < n; k++) { // rows p, q
					const Real apk = a[p + k * n], aqk = a[q + k * n];
					a[p + k * n] = c * apk - s * aqk;
					a[q + k * n] = s * apk + c * aqk;
				}
				for (size_t k = 0; k < n; k++) {
					const Real vkp = vectors[k + p * n], vkq = vectors[k + q * n];
					vectors[k + p * n] = c * vkp - s * vkq;
					vectors[k + q * n] = s * vkp + c * vkq;
				}
			}
	}
	std::vector<size_t> order(n);
	for (size_t i = 0; i < n; i++) order[i] = i;
	std::sort(order.begin(), order.end(), [&](size_t x, size_t y) { return a[x + x * n] < a[y + y * n]; });
	values.resize(n);
	std::vector<Real> sorted(n * n);
	for (size_t j = 0; j < n; j++) {
		values[j] = a[order[j] + order[j] * n];
		for (size_t k = 0; k < n; k++) sorted[k + j * n] = vectors[k + order[j] * n];
	}
	vectors.swap(sorted);
}

Real GaussianMixture::LogSum(Real loga, Real logb)
{
	// MathFunctions.h:67-82
	if (logb > loga) std::swap(loga, logb);
	if (loga == -kInf) return loga;
	const Real diff = logb - loga;
	if (diff < -500) return loga;
	return loga + log1p(exp(diff));
}

// ------------------------------------------------------------------------------------------------ mixture
bool GaussianMixture::Finish(Component& c) const
{
	if (!Cholesky(c.covariance, D, c.chol)) return false;
	Real det = 0.0;
	for (size_t j = 0; j < D; j++) det += log(c.chol[j + j * D]);
	c.logC = -det - 0.5 * D * kLog2Pi;
	return true;
}

bool GaussianMixture::Set(const std::vector<VectorReal>& means, const std::vector<std::vector<Real>>& covariances, const VectorReal& w)
{
	if (means.empty() || means.size() != covariances.size() || means.size() != w.size()) return false;
	D = means[0].size();
	components.assign(means.size(), Component());
	for (size_t i = 0; i < means.size(); i++) {
		if (means[i].size() != D || covariances[i].size() != D * D) return false;
		components[i].mean = means[i];
		components[i].covariance = covariances[i];
		if (!Finish(components[i])) return false;
	}
	weights = w;
	return true;
}

Real GaussianMixture::ComponentLogPdf(const Component& c, const VectorReal& x, VectorReal& v) const
{
	v.resize(D);
	for (size_t i = 0; i < D; i++) v[i] = x[i] - c.mean[i];
	SolveLower(c.chol, D, v);
	Real q = 0.0;
	for (size_t i = 0; i < D; i++) q += v[i] * v[i];
	return c.logC - 0.5 * q;
}

Real GaussianMixture::LogPdf(const VectorReal& x) const
{
	Real logp = -kInf;
	VectorReal v;
	for (size_t i = 0; i < components.size(); i++) logp = LogSum(logp, ComponentLogPdf(components[i], x, v) + log(weights[i]));
	return logp;
}

void GaussianMixture::CalculateResponsibilities(const VectorReal& x, VectorReal& out) const
{
	const size_t K = components.size();
	out.resize(K);
	VectorReal v;
	Real m = -kInf;
	for (size_t i = 0; i < K; i++) {
		out[i] = ComponentLogPdf(components[i], x, v) + log(weights[i]);
		m = std::max(m, out[i]);
	}
	if (m == -kInf) { // x is out of every component's double-precision reach
		std::fill(out.begin(), out.end(), 1.0 / K);
		return;
	}
	Real sum = 0.0;
	for (size_t i = 0; i < K; i++) {
		out[i] = exp(out[i] - m);
		sum += out[i];
	}
	for (size_t i = 0; i < K; i++) out[i] /= sum;
}

// k-means++ seeding and a hard assignment to the nearest seed (GMM.cpp:188-243)
bool GaussianMixture::KMeanspp(const std::vector<VectorReal>& samples, size_t K, RNG& rng, std::vector<VectorReal>& resp)
{
	if (K < 2) return false;
	const size_t n = samples.size();
	components.assign(K, Component());
	auto distsq = [&](const VectorReal& a, const VectorReal& b) {
		Real s = 0.0;
		for (size_t i = 0; i < D; i++) s += (a[i] - b[i]) * (a[i] - b[i]);
		return s;
	};
	std::set<size_t> used;
	size_t ix = rng.GetUnsignedInt((unsigned)n - 1);
	components[0].mean = samples[ix];
	used.insert(ix);
	VectorReal mind(n);
	for (size_t i = 1; i < K; i++) {
		Real total = 0.0;
		for (size_t j = 0; j < n; j++) {
			mind[j] = 0.0;
			if (used.count(j)) continue;
			Real best = std::numeric_limits<Real>::max();
			for (size_t l = 0; l < i; l++) best = std::min(best, distsq(samples[j], components[l].mean));
			mind[j] = best;
			total += best;
		}
		if (!(total > 0.0)) return false; // fewer distinct points than components
		for (Real& v : mind) v /= total;
		ix = SampleIndex(rng, mind);
		components[i].mean = samples[ix];
		used.insert(ix);
	}
	resp.assign(K, VectorReal(n, 0.0));
	for (size_t j = 0; j < n; j++) {
		size_t which = 0;
		Real best = std::numeric_limits<Real>::max();
		for (size_t l = 0; l < K; l++) {
			const Real d = distsq(samples[j], components[l].mean);
			if (d < best) {
				best = d;
				which = l;
			}
		}
		resp[which][j] = 1.0;
	}
	return true;
}

// weighted mean and covariance in one pass, then shrinkage of the correlation eigenvalues for the effective number of
// samples behind the estimate (GMM.cpp:245-336)
void GaussianMixture::MeanCovariance(const std::vector<VectorReal>& samples, const VectorReal& resp, Component& c, Real ess_factor) const
{
	c.mean.assign(D, 0.0);
	std::vector<Real>& cov = c.covariance;
	cov.assign(D * D, 0.0);
	VectorReal d(D), d2(D);
	Real wsum = 0.0;
	for (size_t r = 0; r < samples.size(); r++) {
		const Real w = resp[r];
		if (w < std::numeric_limits<Real>::epsilon()) continue;
		wsum += w;
		for (size_t i = 0; i < D; i++) {
			d[i] = samples[r][i] - c.mean[i];
			c.mean[i] += (w / wsum) * d[i];
			d2[i] = samples[r][i] - c.mean[i];
		}
		for (size_t j = 0; j < D; j++)
			for (size_t i = 0; i < D; i++) cov[i + j * D] += w * d[i] * d2[j];
	}
	if (wsum < 2.0) {
		cov.assign(D * D, 0.0);
		for (size_t i = 0; i < D; i++) cov[i + i * D] = 1.0;
		return;
	}
	for (Real& v : cov) v /= (wsum - 1.0);

	Real n_eff = wsum / ess_factor;
	if (n_eff < 2.0) {
		for (size_t j = 0; j < D; j++)
			for (size_t i = 0; i < D; i++)
				if (i != j) cov[i + j * D] = 0.0;
		return;
	}
	n_eff = std::max(n_eff, (Real)D);
	VectorReal sd(D);
	bool degenerate = false;
	for (size_t i = 0; i < D; i++) {
		sd[i] = sqrt(cov[i + i * D]);
		if (!(sd[i] > 0.0) || !std::isfinite(sd[i])) degenerate = true;
	}
	if (degenerate) {
		// a variable that did not move in the history has no correlation to shrink (the reference ends up with NaNs and a
		// failed factorisation here): keep the variances, floor them like the regular path does
		for (size_t j = 0; j < D; j++)
			for (size_t i = 0; i < D; i++)
				if (i != j) cov[i + j * D] = 0.0;
		for (size_t i = 0; i < D; i++) cov[i + i * D] = (std::isfinite(cov[i + i * D]) && cov[i + i * D] > 0.0 ? cov[i + i * D] : 0.0) + 1e-8;
		return;
	}
	std::vector<Real> corr(D * D);
	for (size_t i = 0; i < D; i++) {
		corr[i + i * D] = 1.0;
		for (size_t j = i + 1; j < D; j++) {
			// the reference reads the upper triangle of an estimate that is symmetric only up to rounding (d d2^T)
			const Real v = cov[i + j * D] / (sd[i] * sd[j]);
			corr[i + j * D] = v;
			corr[j + i * D] = v;
		}
	}
	// Stein-type shrinkage of the eigenvalues (Dey & Srinivasan 1985, Thm 3.1), largest eigenvalue first, with the
	// effective sample size in place of the sample size
	VectorReal eigval;
	std::vector<Real> eigvec;
	SymmetricEigen(corr, D, eigval, eigvec);
	const size_t n_eff_int = (size_t)floor(n_eff);
	for (size_t i = 0; i < D; i++) {
		Real& ev = eigval[(D - 1) - i];
		if (i < n_eff_int) ev *= n_eff / (n_eff + (Real)D + 1.0 - 2.0 * (Real)i);
		else ev = 0.0;
	}
	for (size_t j = 0; j < D; j++)
		for (size_t i = 0; i < D; i++) {
			Real s = 0.0;
			for (size_t k = 0; k < D; k++) s += eigvec[i + k * D] * eigval[k] * eigvec[j + k * D];
			cov[i + j * D] = sd[i] * s * sd[j];
		}
	for (size_t i = 0; i < D; i++) cov[i + i * D] += 1e-8;
}

bool GaussianMixture::Expectation(const std::vector<VectorReal>& samples, std::vector<VectorReal>& resp, Real& logl)
{
	const size_t n = samples.size(), K = components.size();
	VectorReal sample_logl(n, -kInf), v;
	for (size_t i = 0; i < K; i++) {
		if (!Finish(components[i])) return false;
		const Real log_weight = log(weights[i]);
		for (size_t j = 0; j < n; j++) {
			const Real p = ComponentLogPdf(components[i], samples[j], v) + log_weight;
			resp[i][j] = exp(p);
			sample_logl[j] = LogSum(sample_logl[j], p);
		}
	}
	logl = 0.0;
	for (size_t j = 0; j < n; j++) logl += sample_logl[j];
	for (size_t j = 0; j < n; j++) {
		Real total = 0.0;
		for (size_t i = 0; i < K; i++) total += resp[i][j];
		for (size_t i = 0; i < K; i++) resp[i][j] = (total == 0.0) ? 1.0 / K : resp[i][j] / total;
	}
	return true;
}

bool GaussianMixture::Fit(const std::vector<VectorReal>& samples, size_t K, RNG& rng, Real ess_factor)
{
	const size_t maxsteps = 100, retries = 4;
	const Real logl_epsilon = 1e-5;
	const size_t n = samples.size();
	if (n == 0 || K == 0) return false;
	D = samples[0].size();
	Real logl = 0.0;
	bool singular = false;
	if (K == 1) {
		components.assign(1, Component());
		MeanCovariance(samples, VectorReal(n, 1.0), components[0], ess_factor);
		if (!Finish(components[0])) return false;
		VectorReal v;
		for (size_t j = 0; j < n; j++) logl += ComponentLogPdf(components[0], samples[j], v);
		weights.assign(1, 1.0);
	} else {
		if ((Real)n < 2.0 * D * K) return false; // every component needs more than D samples for the regularisation to work
		std::vector<VectorReal> resp;
		for (size_t attempt = 0; attempt < retries; attempt++) {
			singular = false;
			bool converged = false;
			if (!KMeanspp(samples, K, rng, resp)) return false;
			for (size_t i = 0; i < K; i++) MeanCovariance(samples, resp[i], components[i], ess_factor);
			weights.assign(K, 1.0 / K);
			Real prev = -kInf;
			for (size_t step = 0; step < maxsteps; step++) {
				if (!Expectation(samples, resp, logl)) {
					singular = true;
					break;
				}
				if (logl < prev) {
					// a small decrease is rounding: converged; a large one means trouble: start over
					converged = (prev - logl < fabs(logl * logl_epsilon * 10));
					break;
				}
				if (logl - prev < fabs(logl * logl_epsilon)) {
					converged = true;
					break;
				}
				prev = logl;
				for (size_t i = 0; i < K; i++) { // maximisation
					Real s = 0.0;
					for (size_t j = 0; j < n; j++) s += resp[i][j];
					weights[i] = s / (Real)n;
					MeanCovariance(samples, resp[i], components[i], ess_factor);
				}
			}
			if (converged) break;
		}
		if (!singular)
			for (auto& c : components)
				if (!Finish(c)) return false; // the factors of the covariances the last maximisation left behind
	}
	const size_t nparam = K * (D + D * (D + 1) / 2) + K - 1;
	full_logl = logl;
	aic = 2.0 * (Real)nparam - 2.0 * logl;
	return !singular;
}

} // namespace bcm3
