// yaml_lite.hpp — the subset of YAML the reference's input files use, for the C++ host (no yaml library in this image).
//
// What serde_yaml reads for this path (simulation parameters, scene files, split-patterns.yaml, export job files):
// block mappings and block sequences by indentation (including "- key: value" items, nested "- - x" sequences and
// sequences at the indentation of their key), flow sequences "[a, b]" and flow mappings "{a: b}", plain / single- /
// double-quoted scalars, "#" comments, "---" document markers.  Not supported (and not used by those files): anchors,
// tags, multi-line scalars, multiple documents.  tests/test_native_host.py compares the parse of every shipped file — and,
// in the build container, of every YAML file of the reference — with PyYAML.
#pragma once
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace yaml_lite {

struct Node {
  enum Kind { Null, Scalar, Seq, Map } kind = Null;
  std::string scalar;
  bool quoted = false;
  std::vector<Node> seq;
  std::vector<std::pair<std::string, Node>> map;

  bool is_null() const { return kind == Null; }
  const Node* find(const std::string& key) const {
    if (kind != Map) return nullptr;
    for (auto& kv : map) if (kv.first == key) return &kv.second;
    return nullptr;
  }
  Node* find(const std::string& key) { return const_cast<Node*>(static_cast<const Node*>(this)->find(key)); }
  const Node& at(const std::string& key) const {
    const Node* n = find(key);
    if (!n) throw std::runtime_error("missing field `" + key + "`");
    return *n;
  }
  std::string as_string() const {
    if (kind == Null) return "null";
    if (kind != Scalar) throw std::runtime_error("expected a scalar");
    return scalar;
  }
  double as_double() const {
    if (kind != Scalar) throw std::runtime_error("expected a number");
    char* end = nullptr;
    const double v = std::strtod(scalar.c_str(), &end);
    if (end == scalar.c_str() || *end != '\0') throw std::runtime_error("not a number: " + scalar);
    return v;
  }
  bool as_bool() const {
    if (kind == Scalar && !quoted) {
      if (scalar == "true" || scalar == "True" || scalar == "TRUE") return true;
      if (scalar == "false" || scalar == "False" || scalar == "FALSE") return false;
    }
    throw std::runtime_error("expected a boolean, got " + (kind == Scalar ? scalar : std::string("a collection")));
  }
};

namespace detail {

struct Line { int indent; std::string text; };

inline std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r')) a++;
  while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r')) b--;
  return s.substr(a, b - a);
}

// remove a trailing comment: '#' at the start of the text or after whitespace, outside quotes
inline std::string strip_comment(const std::string& s) {
  char q = 0;
  for (size_t i = 0; i < s.size(); i++) {
    const char c = s[i];
    if (q) { if (c == q) q = 0; continue; }
    if (c == '"' || c == '\'') { if (i == 0 || s[i - 1] == ' ' || s[i - 1] == '[' || s[i - 1] == ',' || s[i - 1] == ':' || s[i - 1] == '{') q = c; continue; }
    if (c == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return s.substr(0, i);
  }
  return s;
}

inline bool null_word(const std::string& t) { return t.empty() || t == "~" || t == "null" || t == "Null" || t == "NULL"; }

// position of the ':' that ends a mapping key in `t` (followed by a space or the end of the text), or npos
inline size_t key_colon(const std::string& t) {
  char q = 0;
  int depth = 0;
  for (size_t i = 0; i < t.size(); i++) {
    const char c = t[i];
    if (q) { if (c == q) q = 0; continue; }
    if ((c == '"' || c == '\'') && i == 0) { q = c; continue; }
    if (c == '[' || c == '{') depth++;
    if (c == ']' || c == '}') depth--;
    if (c == ':' && depth == 0 && (i + 1 == t.size() || t[i + 1] == ' ')) return i;
  }
  return std::string::npos;
}

inline std::string unquote(const std::string& t, bool* quoted) {
  *quoted = false;
  if (t.size() >= 2 && ((t.front() == '"' && t.back() == '"') || (t.front() == '\'' && t.back() == '\''))) {
    *quoted = true;
    return t.substr(1, t.size() - 2);
  }
  return t;
}

Node parse_flow(const std::string& t);

inline Node parse_scalar_or_flow(const std::string& raw) {
  const std::string t = trim(raw);
  Node n;
  if (!t.empty() && (t[0] == '[' || t[0] == '{')) return parse_flow(t);
  bool q = false;
  const std::string s = unquote(t, &q);
  if (!q && null_word(s)) return n;
  n.kind = Node::Scalar; n.scalar = s; n.quoted = q;
  return n;
}

// split the inside of a flow collection at top-level commas
inline std::vector<std::string> split_flow(const std::string& inner) {
  std::vector<std::string> out;
  int depth = 0;
  char q = 0;
  std::string cur;
  for (char c : inner) {
    if (q) { cur += c; if (c == q) q = 0; continue; }
    if (c == '"' || c == '\'') { q = c; cur += c; continue; }
    if (c == '[' || c == '{') depth++;
    if (c == ']' || c == '}') depth--;
    if (c == ',' && depth == 0) { out.push_back(cur); cur.clear(); continue; }
    cur += c;
  }
  if (!trim(cur).empty() || !out.empty()) out.push_back(cur);
  return out;
}

inline Node parse_flow(const std::string& t) {
  Node n;
  if (t.size() < 2 || (t[0] == '[' && t.back() != ']') || (t[0] == '{' && t.back() != '}'))
    throw std::runtime_error("unterminated flow collection: " + t);
  const std::string inner = t.substr(1, t.size() - 2);
  if (t[0] == '[') {
    n.kind = Node::Seq;
    for (auto& part : split_flow(inner)) n.seq.push_back(parse_scalar_or_flow(part));
  } else {
    n.kind = Node::Map;
    for (auto& part : split_flow(inner)) {
      const std::string p = trim(part);
      const size_t c = key_colon(p);
      if (c == std::string::npos) throw std::runtime_error("flow mapping entry without ':': " + p);
      bool q;
      n.map.push_back({unquote(trim(p.substr(0, c)), &q), parse_scalar_or_flow(p.substr(c + 1))});
    }
  }
  return n;
}

struct Parser {
  std::vector<Line> lines;
  size_t pos = 0;

  Node block(int indent) {
    if (pos >= lines.size() || lines[pos].indent < indent) return Node();
    const std::string& first = lines[pos].text;
    if (first == "-" || first.compare(0, 2, "- ") == 0) return sequence(lines[pos].indent);
    if (key_colon(first) != std::string::npos) return mapping(lines[pos].indent);
    Node n = parse_scalar_or_flow(first);  // a bare scalar / flow collection on its own line
    pos++;
    return n;
  }

  Node sequence(int indent) {
    Node n;
    n.kind = Node::Seq;
    while (pos < lines.size() && lines[pos].indent == indent && (lines[pos].text == "-" || lines[pos].text.compare(0, 2, "- ") == 0)) {
      const std::string rest = lines[pos].text.size() > 1 ? lines[pos].text.substr(2) : std::string();
      const std::string r = trim(rest);
      if (r.empty()) {
        pos++;
        n.seq.push_back(pos < lines.size() && lines[pos].indent > indent ? block(lines[pos].indent) : Node());
      } else {
        // the item starts on this line: continue as a virtual line indented past the dash
        int lead = 0;  // spaces between "- " and the content
        while (size_t(lead) < rest.size() && rest[size_t(lead)] == ' ') lead++;
        lines[pos].indent = indent + 2 + lead;
        lines[pos].text = r;
        n.seq.push_back(block(lines[pos].indent));
      }
    }
    return n;
  }

  Node mapping(int indent) {
    Node n;
    n.kind = Node::Map;
    while (pos < lines.size() && lines[pos].indent == indent) {
      const std::string t = lines[pos].text;
      if (t == "-" || t.compare(0, 2, "- ") == 0) break;  // a sequence at the indentation of the parent key
      const size_t c = key_colon(t);
      if (c == std::string::npos) throw std::runtime_error("expected `key: value`, got: " + t);
      bool q;
      const std::string key = unquote(trim(t.substr(0, c)), &q);
      const std::string val = trim(t.substr(c + 1));
      pos++;
      if (!val.empty()) {
        n.map.push_back({key, parse_scalar_or_flow(val)});
      } else if (pos < lines.size() && lines[pos].indent > indent) {
        n.map.push_back({key, block(lines[pos].indent)});
      } else if (pos < lines.size() && lines[pos].indent == indent && (lines[pos].text == "-" || lines[pos].text.compare(0, 2, "- ") == 0)) {
        n.map.push_back({key, sequence(indent)});
      } else {
        n.map.push_back({key, Node()});
      }
    }
    return n;
  }
};

}  // namespace detail

inline Node parse(const std::string& text) {
  detail::Parser p;
  std::istringstream in(text);
  std::string raw;
  while (std::getline(in, raw)) {
    const std::string nc = detail::strip_comment(raw);
    const std::string t = detail::trim(nc);
    if (t.empty() || t == "---" || t == "...") continue;
    int indent = 0;
    while (indent < int(nc.size()) && nc[size_t(indent)] == ' ') indent++;
    p.lines.push_back({indent, t});
  }
  if (p.lines.empty()) return Node();
  Node n = p.block(p.lines[0].indent);
  if (p.pos != p.lines.size()) throw std::runtime_error("yaml: could not parse line: " + p.lines[p.pos].text);
  return n;
}

inline Node parse_file(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot read " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  return parse(ss.str());
}

// JSON rendering of a parse (strings only: the comparison with PyYAML is done on the text of the scalars)
inline void to_json(const Node& n, std::string& out) {
  auto esc = [&](const std::string& s) {
    out += '"';
    for (char c : s) {
      if (c == '"' || c == '\\') { out += '\\'; out += c; }
      else if (c == '\n') out += "\\n";
      else if (c == '\t') out += "\\t";
      else out += c;
    }
    out += '"';
  };
  switch (n.kind) {
    case Node::Null: out += "null"; break;
    case Node::Scalar: esc(n.scalar); break;
    case Node::Seq:
      out += '[';
      for (size_t i = 0; i < n.seq.size(); i++) { if (i) out += ','; to_json(n.seq[i], out); }
      out += ']';
      break;
    case Node::Map:
      out += '{';
      for (size_t i = 0; i < n.map.size(); i++) { if (i) out += ','; esc(n.map[i].first); out += ':'; to_json(n.map[i].second, out); }
      out += '}';
      break;
  }
}

}  // namespace yaml_lite
