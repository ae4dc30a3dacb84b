#include "hemo_xml.h"
#include <cctype>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace hemo {
namespace xml {

const Node* Node::firstChild(const std::string& n) const {
  for (auto& c : children) if (c->name == n) return c.get();
  return nullptr;
}
Node* Node::firstChild(const std::string& n) {
  for (auto& c : children) if (c->name == n) return c.get();
  return nullptr;
}
Node* Node::addChild(const std::string& n, const std::string& t) {
  children.emplace_back(new Node());
  Node* c = children.back().get();
  c->name = n; c->text = t; c->parent = this;
  return c;
}

namespace {

struct Parser {
  const std::string& s;
  size_t i = 0;
  explicit Parser(const std::string& s_) : s(s_) {}

  [[noreturn]] void fail(const std::string& what) const {
    size_t line = 1;
    for (size_t k = 0; k < i && k < s.size(); k++) if (s[k] == '\n') line++;
    throw std::runtime_error("XML parse error (line " + std::to_string(line) + "): " + what);
  }
  bool starts(const char* lit) const { return s.compare(i, std::char_traits<char>::length(lit), lit) == 0; }
  void skipSpace() { while (i < s.size() && std::isspace((unsigned char)s[i])) i++; }
  void skipUntil(const char* lit) {
    const size_t p = s.find(lit, i);
    if (p == std::string::npos) fail(std::string("unterminated construct, expected ") + lit);
    i = p + std::char_traits<char>::length(lit);
  }
  static bool nameChar(char c) { return std::isalnum((unsigned char)c) || c == '_' || c == '-' || c == '.' || c == ':'; }
  std::string name() {
    const size_t b = i;
    while (i < s.size() && nameChar(s[i])) i++;
    if (i == b) fail("expected a name");
    return s.substr(b, i - b);
  }
  static std::string unescape(const std::string& t) {
    std::string o; o.reserve(t.size());
    for (size_t k = 0; k < t.size(); k++) {
      if (t[k] != '&') { o += t[k]; continue; }
      const size_t e = t.find(';', k);
      if (e == std::string::npos) { o += t[k]; continue; }
      const std::string ent = t.substr(k + 1, e - k - 1);
      if (ent == "lt") o += '<'; else if (ent == "gt") o += '>'; else if (ent == "amp") o += '&';
      else if (ent == "quot") o += '"'; else if (ent == "apos") o += '\'';
      else { o += t.substr(k, e - k + 1); }
      k = e;
    }
    return o;
  }
  // skips comments, processing instructions and doctype; returns false at end of input
  bool skipMisc() {
    for (;;) {
      skipSpace();
      if (i >= s.size()) return false;
      if (starts("<!--")) { i += 4; skipUntil("-->"); continue; }
      if (starts("<?")) { i += 2; skipUntil("?>"); continue; }
      if (starts("<!DOCTYPE")) { skipUntil(">"); continue; }
      return true;
    }
  }
  void element(Node* parent) {
    if (s[i] != '<') fail("expected '<'");
    i++;
    Node* n = parent->addChild(name());
    for (;;) {                                              // attributes
      skipSpace();
      if (i >= s.size()) fail("unterminated start tag");
      if (s[i] == '/' || s[i] == '>') break;
      const std::string an = name();
      skipSpace();
      if (i >= s.size() || s[i] != '=') fail("expected '=' after attribute name");
      i++; skipSpace();
      if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) fail("expected a quoted attribute value");
      const char q = s[i++];
      const size_t e = s.find(q, i);
      if (e == std::string::npos) fail("unterminated attribute value");
      n->attributes.emplace_back(an, unescape(s.substr(i, e - i)));
      i = e + 1;
    }
    if (s[i] == '/') { if (i + 1 >= s.size() || s[i+1] != '>') fail("expected '/>'"); i += 2; return; }
    i++;                                                     // '>'
    for (;;) {                                              // content
      const size_t lt = s.find('<', i);
      if (lt == std::string::npos) fail("unterminated element <" + n->name + ">");
      n->text += unescape(s.substr(i, lt - i));
      i = lt;
      if (starts("<!--")) { i += 4; skipUntil("-->"); continue; }
      if (starts("<![CDATA[")) { i += 9; const size_t e = s.find("]]>", i); if (e == std::string::npos) fail("unterminated CDATA"); n->text += s.substr(i, e - i); i = e + 3; continue; }
      if (starts("<?")) { i += 2; skipUntil("?>"); continue; }
      if (starts("</")) {
        i += 2;
        const std::string cn = name();
        if (cn != n->name) fail("mismatched end tag </" + cn + "> for <" + n->name + ">");
        skipSpace();
        if (i >= s.size() || s[i] != '>') fail("expected '>'");
        i++;
        break;
      }
      element(n);
    }
    // trim the character data (tinyxml2's GetText() of "<a> 1 </a>" is then read through operator>>)
    size_t b = 0, e = n->text.size();
    while (b < e && std::isspace((unsigned char)n->text[b])) b++;
    while (e > b && std::isspace((unsigned char)n->text[e-1])) e--;
    n->text = n->text.substr(b, e - b);
  }
};

void write(const Node& n, int depth, std::ostringstream& o) {
  const std::string ind((size_t)depth*2, ' ');
  o << ind << '<' << n.name;
  for (auto& a : n.attributes) o << ' ' << a.first << "=\"" << a.second << '"';
  if (n.children.empty() && n.text.empty()) { o << "/>\n"; return; }
  o << '>';
  if (n.children.empty()) { o << n.text << "</" << n.name << ">\n"; return; }
  o << '\n';
  if (!n.text.empty()) o << ind << "  " << n.text << '\n';
  for (auto& c : n.children) write(*c, depth + 1, o);
  o << ind << "</" << n.name << ">\n";
}

}  // namespace

std::unique_ptr<Node> parse(const std::string& content) {
  std::unique_ptr<Node> root(new Node());
  Parser p(content);
  while (p.skipMisc()) p.element(root.get());
  if (root->children.empty()) throw std::runtime_error("XML parse error: no element found");
  return root;
}

std::unique_ptr<Node> parseFile(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open " + path);
  std::stringstream ss; ss << f.rdbuf();
  return parse(ss.str());
}

std::string serialize(const Node& root) {
  std::ostringstream o;
  o << "<?xml version=\"1.0\" ?>\n";
  for (auto& c : root.children) write(*c, 0, o);
  return o.str();
}

}  // namespace xml
}  // namespace hemo
