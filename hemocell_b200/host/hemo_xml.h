// Minimal XML reader behind hemo::Config (the reference reads config.xml / <CELL>.xml through
// tinyxml2: config/config.h:37-78).  Elements, text, comments, declarations; attributes are parsed
// and kept; no entities beyond the five predefined ones, no DTD.
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace hemo {
namespace xml {

struct Node {
  std::string name;
  std::string text;                                       // concatenated character data of this element
  std::vector<std::pair<std::string, std::string>> attributes;
  std::vector<std::unique_ptr<Node>> children;
  Node* parent = nullptr;
  const Node* firstChild(const std::string& n) const;
  Node* firstChild(const std::string& n);
  Node* addChild(const std::string& n, const std::string& text = "");
};

// throws std::runtime_error with line information on malformed input
std::unique_ptr<Node> parse(const std::string& content);     // returns a synthetic root holding the top-level elements
std::unique_ptr<Node> parseFile(const std::string& path);
std::string serialize(const Node& root);                     // root = synthetic root

}  // namespace xml
}  // namespace hemo
