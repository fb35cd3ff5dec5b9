// Mirror of the independence prior of the reference (src/sampler/PriorIndependence.cpp:20-157,
// UnivariateMarginal.cpp:28-94) for the marginals the synthetic configs use: uniform and normal.
#pragma once

#include "RNG.h"
#include "Types.h"
#include "VariableSet.h"
#include "Xml.h"

namespace bcm3 {

class Prior {
public:
	struct Marginal {
		enum Kind { Uniform, Normal } kind = Uniform;
		Real a = 0.0, b = 1.0; // uniform: lower, upper; normal: mu, sigma
	};
	static std::shared_ptr<Prior> Create(const std::string& prior_xml_fn, std::shared_ptr<const VariableSet> varset);
	static std::shared_ptr<Prior> CreateFromNode(const XmlNode& prior_node, std::shared_ptr<const VariableSet> varset);

	bool EvaluateLogPDF(size_t threadix, const Real* values, Real& logp) const; // PriorIndependence.cpp:129-157
	bool Sample(Real* values, RNG* rng) const;                                  // :159-179
	Real GetLowerBound(size_t i) const;
	Real GetUpperBound(size_t i) const;
	bool EvaluateMarginalMean(size_t i, Real& mean) const; // PriorIndependence.cpp:181-193
	bool EvaluateMarginalVariance(size_t i, Real& var) const;
	size_t GetNumVariables() const { return marginals.size(); }

private:
	std::vector<Marginal> marginals;
};

} // namespace bcm3
