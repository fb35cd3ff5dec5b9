#include "VariableSet.h"

namespace bcm3 {

bool VariableSet::LoadFromXML(const std::string& filename)
{
	XmlNode root;
	std::string err;
	if (!LoadXmlFile(filename, root, err)) return false;
	const XmlNode* node = root.child("prior");
	if (!node) node = root.child("variableset");
	if (!node) return false; // "Incorrect prior XML format"
	return LoadFromNode(*node);
}

bool VariableSet::LoadFromNode(const XmlNode& prior_node)
{
	for (const XmlNode& var : prior_node.children) {
		if (var.name != "variable") continue;
		if (!var.has("name")) return false;
		const std::string name = var.get("name");
		const long repeat = var.get_int("repeat", 1);
		const bool logspace = var.get_bool("logspace", false);
		const bool logistic = var.get_bool("logistic", false);
		for (long i = 0; i < repeat; i++) {
			ETransform t = logspace ? Transform_Log10 : (logistic ? Transform_Logit : Transform_None);
			AddVariable(repeat > 1 ? name + "_" + std::to_string(i) : name, t);
		}
	}
	return true;
}

void VariableSet::AddVariable(const std::string& name, ETransform transform)
{
	variables.push_back(name);
	transforms.push_back(transform);
}

size_t VariableSet::GetVariableIndex(const std::string& name) const
{
	for (size_t i = 0; i < variables.size(); i++)
		if (variables[i] == name) return i;
	return std::numeric_limits<size_t>::max();
}

Real VariableSet::TransformVariable(size_t i, Real x) const
{
	switch (transforms[i]) {
	case Transform_Log:
		return exp(x);
	case Transform_Log10:
		return exp(x * 2.3025850929940459); // bcm3::fastpow10, MathFunctions.h:13
	case Transform_Logit:
		if (x > 0) {
			Real z = exp(-x);
			return 1.0 / (1.0 + z);
		} else {
			Real z = exp(x);
			return z / (1.0 + z);
		}
	default:
		return x;
	}
}

} // namespace bcm3
