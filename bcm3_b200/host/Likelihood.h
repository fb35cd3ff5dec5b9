// Mirror of bcm3::Likelihood (src/sampler/Likelihood.h:9-35) with the ONE added entry of the batched design.
#pragma once

#include "Types.h"
#include "VariableSet.h"
#include "Xml.h"

namespace bcm3 {

class Likelihood {
public:
	virtual ~Likelihood() {}

	bool SetLearningRate(Real lr)
	{
		if (lr < 0.0 || lr > 1.0) return false; // Likelihood.cpp:17-20
		learning_rate = lr;
		return true;
	}
	Real GetLearningRate() const { return learning_rate; }

	// `likelihood_node` = the <bcm_likelihood> element (boost ptree in the reference)
	virtual bool Initialize(std::shared_ptr<const VariableSet> varset, const XmlNode& likelihood_node) { (void)varset; (void)likelihood_node; return true; }
	// The optimiser's hooks (src/sampler/Likelihood.h:18-19, used by bcmopt/main.cpp:156-234): parameters that are not
	// sampled but set from outside between evaluations. Defaults as in Likelihood.cpp:31-38.
	virtual bool AddNonSampledParameters(const std::vector<std::string>& variable_names) { (void)variable_names; return true; }
	virtual void SetNonSampledParameters(const VectorReal& values) { (void)values; }
	virtual bool PostInitialize() { return true; }
	virtual bool IsReentrant() = 0;
	// Likelihood.h:22, called once after sampling (bcminf/main.cpp:133)
	virtual void OutputEvaluationStatistics(const std::string& path) const { (void)path; }

	//! As in the reference: evaluate one variable vector. False = unrecoverable, the sampler stops.
	virtual bool EvaluateLogProbability(size_t threadix, const VectorReal& values, Real& logp) = 0;

	//! NEW: evaluate all chains' proposals at once. values is nvar x C (column = chain), logp gets C entries.
	//! The default loops over EvaluateLogProbability so every existing likelihood keeps working.
	virtual bool EvaluateLogProbabilityBatch(const MatrixReal& values, VectorReal& logp)
	{
		logp.assign(values.cols(), -kInf);
		VectorReal v(values.rows());
		for (size_t c = 0; c < values.cols(); c++) {
			v.assign(values.col(c), values.col(c) + values.rows());
			if (!EvaluateLogProbability(0, v, logp[c])) return false;
		}
		return true;
	}

protected:
	Likelihood() : learning_rate(1.0) {}
	Real learning_rate;
};

} // namespace bcm3
