#include "Xml.h"

#include <cctype>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace bcm3 {

const XmlNode* XmlNode::child(const std::string& n) const
{
	for (const auto& c : children)
		if (c.name == n) return &c;
	return nullptr;
}
std::string XmlNode::get(const std::string& a, const std::string& def) const
{
	auto it = attr.find(a);
	return it == attr.end() ? def : it->second;
}
double XmlNode::get_real(const std::string& a, double def) const
{
	auto it = attr.find(a);
	return it == attr.end() ? def : strtod(it->second.c_str(), nullptr);
}
long XmlNode::get_int(const std::string& a, long def) const
{
	auto it = attr.find(a);
	return it == attr.end() ? def : strtol(it->second.c_str(), nullptr, 10);
}
bool XmlNode::get_bool(const std::string& a, bool def) const
{
	auto it = attr.find(a);
	if (it == attr.end()) return def;
	return it->second == "true" || it->second == "1";
}

namespace {

struct Parser {
	const std::string& s;
	size_t p = 0;
	std::string err;
	explicit Parser(const std::string& t) : s(t) {}
	void ws() { while (p < s.size() && isspace((unsigned char)s[p])) p++; }
	bool starts(const char* t) const { return s.compare(p, strlen(t), t) == 0; }
	bool skip_misc()
	{
		for (;;) {
			ws();
			if (starts("<?")) {
				size_t e = s.find("?>", p);
				if (e == std::string::npos) return fail("unterminated declaration");
				p = e + 2;
			} else if (starts("<!--")) {
				size_t e = s.find("-->", p);
				if (e == std::string::npos) return fail("unterminated comment");
				p = e + 3;
			} else {
				return true;
			}
		}
	}
	bool fail(const std::string& m) { err = m + " at offset " + std::to_string(p); return false; }
	static std::string unescape(const std::string& v)
	{
		std::string o;
		for (size_t i = 0; i < v.size(); i++) {
			if (v[i] == '&') {
				struct { const char* e; char c; } tab[] = { { "&lt;", '<' }, { "&gt;", '>' }, { "&amp;", '&' }, { "&quot;", '"' }, { "&apos;", '\'' } };
				bool hit = false;
				for (auto& t : tab) {
					if (v.compare(i, strlen(t.e), t.e) == 0) { o += t.c; i += strlen(t.e) - 1; hit = true; break; }
				}
				if (!hit) o += v[i];
			} else o += v[i];
		}
		return o;
	}
	bool name(std::string& out)
	{
		size_t b = p;
		while (p < s.size() && (isalnum((unsigned char)s[p]) || s[p] == '_' || s[p] == '-' || s[p] == ':' || s[p] == '.')) p++;
		if (p == b) return fail("expected a name");
		out = s.substr(b, p - b);
		return true;
	}
	bool element(XmlNode& n)
	{
		if (p >= s.size() || s[p] != '<') return fail("expected '<'");
		p++;
		if (!name(n.name)) return false;
		for (;;) {
			ws();
			if (p >= s.size()) return fail("unterminated tag");
			if (s[p] == '/') {
				if (p + 1 < s.size() && s[p + 1] == '>') { p += 2; return true; }
				return fail("bad '/'");
			}
			if (s[p] == '>') { p++; break; }
			std::string an;
			if (!name(an)) return false;
			ws();
			if (p >= s.size() || s[p] != '=') return fail("expected '='");
			p++;
			ws();
			if (p >= s.size() || (s[p] != '"' && s[p] != '\'')) return fail("expected a quoted value");
			char q = s[p++];
			size_t e = s.find(q, p);
			if (e == std::string::npos) return fail("unterminated attribute value");
			n.attr[an] = unescape(s.substr(p, e - p));
			p = e + 1;
		}
		// children until the closing tag; character data is ignored
		for (;;) {
			size_t lt = s.find('<', p);
			if (lt == std::string::npos) return fail("missing closing tag for " + n.name);
			p = lt;
			if (starts("<!--") || starts("<?")) {
				if (!skip_misc()) return false;
				continue;
			}
			if (starts("</")) {
				p += 2;
				std::string cn;
				if (!name(cn)) return false;
				if (cn != n.name) return fail("mismatched closing tag " + cn);
				ws();
				if (p >= s.size() || s[p] != '>') return fail("expected '>'");
				p++;
				return true;
			}
			XmlNode c;
			if (!element(c)) return false;
			n.children.push_back(std::move(c));
		}
	}
};

} // namespace

bool ParseXml(const std::string& text, XmlNode& root, std::string& error)
{
	Parser ps(text);
	root = XmlNode();
	for (;;) {
		if (!ps.skip_misc()) { error = ps.err; return false; }
		if (ps.p >= text.size()) break;
		XmlNode n;
		if (!ps.element(n)) { error = ps.err; return false; }
		root.children.push_back(std::move(n));
	}
	if (root.children.empty()) { error = "no root element"; return false; }
	return true;
}

bool LoadXmlFile(const std::string& filename, XmlNode& root, std::string& error)
{
	std::ifstream f(filename);
	if (!f) { error = "cannot open " + filename; return false; }
	std::stringstream ss;
	ss << f.rdbuf();
	return ParseXml(ss.str(), root, error);
}

} // namespace bcm3
