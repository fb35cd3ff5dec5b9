// Gaussian mixture model for the adaptive proposal of the PT sampler: what src/stats/GMM.{h,cpp} provides to
// ProposalGaussianMixture (src/sampler/ProposalGaussianMixture.cpp) -- k-means++ start, EM with a covariance estimate whose
// correlation eigenvalues are shrunk for the effective sample size (GMM.cpp:245-336), AIC. Dense algebra on small flat
// column-major arrays (the variable blocks of a sampler have tens of dimensions): Cholesky and a cyclic Jacobi
// eigen-decomposition written here, no Eigen.
#pragma once

#include <vector>

#include "RNG.h"
#include "Types.h"

namespace bcm3 {

class GaussianMixture {
public:
	struct Component {
		VectorReal mean;
		std::vector<Real> covariance, chol; // D x D column-major; chol = lower Cholesky factor
		Real logC = 0.0;                    // -sum log L_ii - D/2 log(2 pi)
	};

	// explicit mixture (GMM::Set, GMM.cpp:14-46); false when a covariance is not positive definite
	bool Set(const std::vector<VectorReal>& means, const std::vector<std::vector<Real>>& covariances, const VectorReal& weights);
	// GMM::Fit (GMM.cpp:48-158): samples[r] = one D-vector; ess_factor = samples per effective sample
	bool Fit(const std::vector<VectorReal>& samples, size_t num_components, RNG& rng, Real ess_factor);

	size_t GetNumComponents() const { return components.size(); }
	const Component& GetComponent(size_t k) const { return components[k]; }
	const VectorReal& GetWeights() const { return weights; }
	Real GetAIC() const { return aic; }
	Real GetLogLikelihood() const { return full_logl; }
	Real LogPdf(const VectorReal& x) const;
	// posterior component probabilities of x (GMM.cpp:172-186)
	void CalculateResponsibilities(const VectorReal& x, VectorReal& out) const;

	// ---- dense helpers (also used by the proposal) ----
	static bool Cholesky(const std::vector<Real>& a, size_t n, std::vector<Real>& l);
	static void SolveLower(const std::vector<Real>& l, size_t n, VectorReal& v); // v <- L^-1 v
	// symmetric eigen-decomposition, eigenvalues ascending, eigenvectors in the columns of `vectors`
	static void SymmetricEigen(std::vector<Real> a, size_t n, VectorReal& values, std::vector<Real>& vectors);
	static Real LogSum(Real loga, Real logb);

private:
	bool Finish(Component& c) const;
	bool KMeanspp(const std::vector<VectorReal>& samples, size_t K, RNG& rng, std::vector<VectorReal>& resp);
	void MeanCovariance(const std::vector<VectorReal>& samples, const VectorReal& resp, Component& c, Real ess_factor) const;
	bool Expectation(const std::vector<VectorReal>& samples, std::vector<VectorReal>& resp, Real& logl);
	Real ComponentLogPdf(const Component& c, const VectorReal& x, VectorReal& scratch) const;

	size_t D = 0;
	std::vector<Component> components;
	VectorReal weights;
	Real aic = 0.0, full_logl = 0.0;
};

// RNG::Sample (RNG.cpp:41-56): index drawn with the given probabilities
size_t SampleIndex(RNG& rng, const VectorReal& probabilities);

} // namespace bcm3
