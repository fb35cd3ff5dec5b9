// Mirror of bcm3::LikelihoodFactory::CreateLikelihood (src/likelihoods/LikelihoodFactory.cpp:31-101): same type strings.
#pragma once

#include "Likelihood.h"

namespace bcm3 {

class LikelihoodFactory {
public:
	static std::shared_ptr<Likelihood> CreateLikelihood(const std::string& likelihood_xml_fn, std::shared_ptr<const VariableSet> varset,
	                                                    size_t sampling_threads, size_t evaluation_threads, std::string* error = nullptr);
	static std::shared_ptr<Likelihood> CreateLikelihoodFromText(const std::string& xml_text, std::shared_ptr<const VariableSet> varset,
	                                                            size_t sampling_threads, size_t evaluation_threads, std::string* error = nullptr);
};

} // namespace bcm3
