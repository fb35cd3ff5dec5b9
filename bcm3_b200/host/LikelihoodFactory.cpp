#include "LikelihoodFactory.h"

#include <fstream>
#include <sstream>

#include "CellPopulationLikelihoodB200.h"
#include "LikelihoodPopPKTrajectoryB200.h"
#include "PharmacoLikelihoodPopulationB200.h"
#include "TestLikelihoodBanana.h"

namespace bcm3 {

std::shared_ptr<Likelihood> LikelihoodFactory::CreateLikelihood(const std::string& fn, std::shared_ptr<const VariableSet> varset,
                                                                size_t sampling_threads, size_t evaluation_threads, std::string* error)
{
	std::ifstream f(fn);
	if (!f) {
		if (error) *error = "Error loading likelihood file: cannot open " + fn;
		return nullptr;
	}
	std::stringstream ss;
	ss << f.rdbuf();
	return CreateLikelihoodFromText(ss.str(), varset, sampling_threads, evaluation_threads, error);
}

std::shared_ptr<Likelihood> LikelihoodFactory::CreateLikelihoodFromText(const std::string& text, std::shared_ptr<const VariableSet> varset,
                                                                        size_t sampling_threads, size_t evaluation_threads, std::string* error)
{
	std::shared_ptr<Likelihood> ll;
	XmlNode root;
	std::string err;
	if (!ParseXml(text, root, err)) {
		if (error) *error = "Error loading likelihood file: " + err;
		return ll;
	}
	const XmlNode* node = root.child("bcm_likelihood");
	if (!node || !node->has("type")) {
		if (error) *error = "Error parsing likelihood file: no bcm_likelihood type";
		return ll;
	}
	const std::string type = node->get("type");
	if (type == "banana") {
		ll = std::make_shared<TestLikelihoodBanana>(sampling_threads, evaluation_threads);
	} else if (type == "pop_pk_trajectory") {
		ll = std::make_shared<LikelihoodPopPKTrajectoryB200>(sampling_threads, evaluation_threads);
	} else if (type == "pharmacokinetic_trajectory") { // LikelihoodFactory.cpp:60
		ll = std::make_shared<LikelihoodPopPKTrajectoryB200>(sampling_threads, evaluation_threads, true);
	} else if (type == "pharmaco_single") { // LikelihoodFactory.cpp:64
		ll = std::make_shared<PharmacoLikelihoodPopulationB200>(sampling_threads, evaluation_threads, true);
	} else if (type == "pharmaco_population") { // LikelihoodFactory.cpp:66
		ll = std::make_shared<PharmacoLikelihoodPopulationB200>(sampling_threads, evaluation_threads);
	} else if (type == "cell_population") { // LikelihoodFactory.cpp:81
		ll = std::make_shared<CellPopulationLikelihoodB200>(sampling_threads, evaluation_threads);
	} else {
		if (error) *error = "Unknown likelihood type \"" + type + "\"";
		return ll;
	}
	if (!ll->Initialize(varset, *node)) {
		if (error) {
			*error = "Failed to initialize likelihood";
			if (auto* cp = dynamic_cast<CellPopulationLikelihoodB200*>(ll.get())) *error += ": " + cp->LastError();
			if (auto* ph = dynamic_cast<PharmacoLikelihoodPopulationB200*>(ll.get())) *error += ": " + ph->LastError();
		}
		ll.reset();
	}
	return ll;
}

} // namespace bcm3
