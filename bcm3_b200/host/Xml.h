// Minimal XML reader for prior.xml / likelihood.xml (elements + attributes; no text, CDATA or entities beyond
// the five predefined ones). Stands in for boost::property_tree::read_xml, which the reference uses
// (src/sampler/VariableSet.cpp:16-69, src/likelihoods/LikelihoodFactory.cpp:31-46).
#pragma once

#include <map>
#include <string>
#include <vector>

namespace bcm3 {

struct XmlNode {
	std::string name;
	std::map<std::string, std::string> attr;
	std::vector<XmlNode> children;
	const XmlNode* child(const std::string& n) const;
	bool has(const std::string& a) const { return attr.count(a) != 0; }
	std::string get(const std::string& a, const std::string& def = "") const;
	double get_real(const std::string& a, double def) const;
	long get_int(const std::string& a, long def) const;
	bool get_bool(const std::string& a, bool def) const;
};

// Parses `text`; returns false and fills `error` on malformed input. `root` gets a synthetic node whose children
// are the top-level elements.
bool ParseXml(const std::string& text, XmlNode& root, std::string& error);
bool LoadXmlFile(const std::string& filename, XmlNode& root, std::string& error);

} // namespace bcm3
