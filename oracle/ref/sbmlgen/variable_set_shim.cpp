// TEST INFRASTRUCTURE ONLY -- the part of bcm3::VariableSet the code generator touches (src/sampler/VariableSet.cpp:71-124),
// stated without Boost: the reference's own VariableSet.cpp cannot be compiled here (LoadFromXML uses boost::property_tree).
// The class declaration is the reference's header, included from where it lies.
#include "VariableSet.h"

namespace bcm3 {

VariableSet::VariableSet() {}
VariableSet::~VariableSet() {}

bool VariableSet::LoadFromXML(const std::string&) { return false; } // prior.xml is not read by the generator tool

// VariableSet.cpp:71-82
void VariableSet::AddVariable(const std::string& name, bool logspace, bool logistic)
{
	variables.push_back(name);
	transforms.push_back(logspace ? Transform_Log10 : (logistic ? Transform_Logit : Transform_None));
}

// VariableSet.cpp:84-95
size_t VariableSet::GetVariableIndex(const std::string& name, bool log_error) const
{
	for (size_t vi = 0; vi < variables.size(); vi++)
		if (variables[vi] == name) return vi;
	if (log_error) LOGERROR("Could not find variable \"%s\"", name.c_str());
	return std::numeric_limits<size_t>::max();
}

// VariableSet.cpp:97-124 (not reached by GenerateCode; the generator only needs names and indices)
Real VariableSet::TransformVariable(size_t i, Real x) const
{
	switch (transforms[i]) {
	case Transform_Log: return exp(x);
	case Transform_Log10: return bcm3::fastpow10(x);
	case Transform_Logit: {
		if (x > 0) {
			const Real z = exp(-x);
			return 1.0 / (1.0 + z);
		}
		const Real z = exp(x);
		return z / (1.0 + z);
	}
	default: return x;
	}
}

} // namespace bcm3
