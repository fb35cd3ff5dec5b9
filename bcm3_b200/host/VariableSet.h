// Mirror of bcm3::VariableSet (src/sampler/VariableSet.{h,cpp}): variable names + transforms, prior.xml loading.
#pragma once

#include "Types.h"
#include "Xml.h"

namespace bcm3 {

class VariableSet {
public:
	enum ETransform { Transform_None = 0, Transform_Log = 1, Transform_Log10 = 2, Transform_Logit = 3 };

	bool LoadFromXML(const std::string& filename);  // VariableSet.cpp:16-69
	bool LoadFromNode(const XmlNode& prior_node);
	void AddVariable(const std::string& name, ETransform transform = Transform_None);

	size_t GetNumVariables() const { return variables.size(); }
	const std::string& GetVariableName(size_t i) const { return variables[i]; }
	size_t GetVariableIndex(const std::string& name) const; // size_t max when missing (VariableSet.cpp:84-95)
	ETransform GetTransform(size_t i) const { return transforms[i]; }
	Real TransformVariable(size_t i, Real x) const; // VariableSet.cpp:97-124

private:
	std::vector<std::string> variables;
	std::vector<ETransform> transforms;
};

} // namespace bcm3
