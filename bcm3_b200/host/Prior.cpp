#include "Prior.h"

namespace bcm3 {

std::shared_ptr<Prior> Prior::Create(const std::string& prior_xml_fn, std::shared_ptr<const VariableSet> varset)
{
	XmlNode root;
	std::string err;
	if (!LoadXmlFile(prior_xml_fn, root, err)) return nullptr;
	const XmlNode* node = root.child("prior");
	if (!node) node = root.child("variableset");
	if (!node) return nullptr;
	return CreateFromNode(*node, varset);
}

std::shared_ptr<Prior> Prior::CreateFromNode(const XmlNode& prior_node, std::shared_ptr<const VariableSet> varset)
{
	auto p = std::make_shared<Prior>();
	for (const XmlNode& var : prior_node.children) {
		if (var.name != "variable") continue;
		Marginal m;
		const std::string dist = var.get("distribution");
		if (dist == "uniform") {
			m.kind = Marginal::Uniform;
			m.a = var.get_real("lower", 0.0);
			m.b = var.get_real("upper", 1.0);
			if (!(m.b > m.a)) return nullptr;
		} else if (dist == "normal") {
			m.kind = Marginal::Normal;
			m.a = var.get_real("mu", 0.0);
			m.b = var.get_real("sigma", 1.0);
			if (!(m.b > 0.0)) return nullptr;
		} else {
			return nullptr; // other marginals of UnivariateMarginal.cpp are not on the path
		}
		const long repeat = var.get_int("repeat", 1);
		for (long i = 0; i < repeat; i++) p->marginals.push_back(m);
	}
	if (varset && p->marginals.size() != varset->GetNumVariables()) return nullptr;
	return p;
}

bool Prior::EvaluateLogPDF(size_t, const Real* values, Real& logp) const
{
	logp = 0.0;
	for (size_t i = 0; i < marginals.size(); i++) {
		const Marginal& m = marginals[i];
		const Real x = values[i];
		if (m.kind == Marginal::Uniform) {
			if (x < m.a || x > m.b) logp += -kInf;
			else logp += -log(m.b - m.a);
		} else {
			const Real d = x - m.a;
			logp += -log(m.b) - 0.91893853320467274178 - d * d / (2.0 * m.b * m.b); // LogPdfNormal, ProbabilityDistributions.cpp:129-138
		}
	}
	return true;
}

bool Prior::Sample(Real* values, RNG* rng) const
{
	for (size_t i = 0; i < marginals.size(); i++) {
		const Marginal& m = marginals[i];
		values[i] = (m.kind == Marginal::Uniform) ? rng->GetUniform(m.a, m.b) : rng->GetNormal(m.a, m.b);
	}
	return true;
}

Real Prior::GetLowerBound(size_t i) const { return marginals[i].kind == Marginal::Uniform ? marginals[i].a : -kInf; }
Real Prior::GetUpperBound(size_t i) const { return marginals[i].kind == Marginal::Uniform ? marginals[i].b : kInf; }

bool Prior::EvaluateMarginalMean(size_t i, Real& mean) const
{
	const Marginal& m = marginals[i];
	mean = (m.kind == Marginal::Uniform) ? 0.5 * (m.a + m.b) : m.a;
	return true;
}

bool Prior::EvaluateMarginalVariance(size_t i, Real& var) const
{
	const Marginal& m = marginals[i];
	var = (m.kind == Marginal::Uniform) ? (m.b - m.a) * (m.b - m.a) / 12.0 : m.b * m.b;
	return true;
}

} // namespace bcm3
