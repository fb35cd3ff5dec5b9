#include "TestLikelihoodBanana.h"

using bcm3::Real;

static Real PdfNormal(Real x, Real mu, Real sigma)
{
	// bcm3::PdfNormal (ProbabilityDistributions.cpp:51-56) uses a hand-rolled rsqrt accurate to ~1e-14 (App. D #15)
	const Real d = (x - mu) / sigma;
	return 0.3989422804014327 / sigma * exp(-0.5 * d * d);
}

bool TestLikelihoodBanana::Initialize(std::shared_ptr<const bcm3::VariableSet> varset, const bcm3::XmlNode& node)
{
	if (!node.has("dimension") || !node.has("sd1") || !node.has("sd2")) return false;
	dim = (size_t)node.get_int("dimension", 0);
	if (dim != varset->GetNumVariables() || dim < 2) return false;
	sd1 = node.get_real("sd1", 0.0);
	sd2 = node.get_real("sd2", 0.0);
	return sd1 > 0.0 && sd2 > 0.0;
}

bool TestLikelihoodBanana::EvaluateLogProbability(size_t, const bcm3::VectorReal& values, Real& logp)
{
	Real p = 1.0;
	for (size_t i = 0; i < dim - 1; i++) p = p * PdfNormal(values[i], 0, sd1);
	Real y = values[0];
	for (size_t i = 1; i < dim - 1; i++) y += values[i];
	p *= PdfNormal(values[dim - 1], y + 3 * y + (1 - y) * (1 - y), sd2);
	logp = log(p);
	return true;
}
