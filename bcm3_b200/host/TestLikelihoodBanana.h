// Mirror of TestLikelihoodBanana (src/likelihoods/TestLikelihoodBanana.cpp:14-55): analytic, CPU, config 1.
#pragma once

#include "Likelihood.h"

class TestLikelihoodBanana : public bcm3::Likelihood {
public:
	TestLikelihoodBanana(size_t sampling_threads, size_t evaluation_threads) { (void)sampling_threads; (void)evaluation_threads; }
	bool Initialize(std::shared_ptr<const bcm3::VariableSet> varset, const bcm3::XmlNode& likelihood_node) override;
	bool IsReentrant() override { return true; }
	bool EvaluateLogProbability(size_t threadix, const bcm3::VectorReal& values, bcm3::Real& logp) override;

private:
	size_t dim = 0;
	bcm3::Real sd1 = 1.0, sd2 = 1.0;
};
