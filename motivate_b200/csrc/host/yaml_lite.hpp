// yaml_lite.hpp — the YAML subset that the reference's files use (serde_yaml 0.7 output and the hand-written
// config/*.yaml): block mappings, block sequences (also at the indentation of their parent key), flow sequences
// `[a, b]` of scalars or of flow sequences, plain / quoted scalars, `#` comments, a leading `---`.
// Not a general YAML parser: anchors, tags, multi-line scalars and flow mappings are rejected with an error.
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
    std::vector<Node> seq;
    std::vector<std::pair<std::string, Node>> map;       // insertion order kept

    bool is_map() const { return kind == Map; }
    bool is_seq() const { return kind == Seq; }
    const Node* find(const std::string& k) const {
        for (auto& kv : map) if (kv.first == k) return &kv.second;
        return nullptr;
    }
    const Node& at(const std::string& k) const {
        const Node* n = find(k);
        if (!n) throw std::runtime_error("There was an error parsing the file: missing field `" + k + "`");
        return *n;
    }
    const std::string& str() const {
        if (kind != Scalar) throw std::runtime_error("There was an error parsing the file: expected a scalar");
        return scalar;
    }
    double as_double() const {                            // decimal -> f64, like serde_yaml
        char* end = nullptr;
        const double v = std::strtod(str().c_str(), &end);
        if (end == scalar.c_str() || *end) throw std::runtime_error("There was an error parsing the file: not a number: " + scalar);
        return v;
    }
    long long as_int() const {
        char* end = nullptr;
        const long long v = std::strtoll(str().c_str(), &end, 10);
        if (end == scalar.c_str() || *end) throw std::runtime_error("There was an error parsing the file: not an integer: " + scalar);
        return v;
    }
    bool as_bool() const {
        const std::string& s = str();
        if (s == "true" || s == "True") return true;
        if (s == "false" || s == "False") return false;
        throw std::runtime_error("There was an error parsing the file: not a boolean: " + s);
    }
};

namespace detail {

struct Line { int indent; std::string text; int no; };

inline std::string trim(const std::string& s) {
    size_t b = s.find_first_not_of(" \t\r"), e = s.find_last_not_of(" \t\r");
    return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
}

inline std::string strip_comment(const std::string& s) {   // a '#' outside quotes, at line start or after a space
    char q = 0;
    for (size_t i = 0; i < s.size(); ++i) {
        const char c = s[i];
        if (q) { if (c == q) q = 0; }
        else if (c == '"' || c == '\'') q = c;
        else if (c == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return s.substr(0, i);
    }
    return s;
}

inline std::string unquote(const std::string& t) {
    if (t.size() >= 2 && ((t.front() == '"' && t.back() == '"') || (t.front() == '\'' && t.back() == '\''))) return t.substr(1, t.size() - 2);
    return t;
}

inline Node parse_flow(const std::string& t, size_t& i);
inline Node parse_scalar_or_flow(const std::string& raw) {
    const std::string t = trim(raw);
    Node n;
    if (t.empty() || t == "~" || t == "null") return n;
    if (t[0] == '[') { size_t i = 0; return parse_flow(t, i); }
    if (t[0] == '{') {
        if (t == "{}") { n.kind = Node::Map; return n; }
        throw std::runtime_error("yaml_lite: flow mappings are not supported: " + t);
    }
    if (t[0] == '&' || t[0] == '*' || t[0] == '!' || t[0] == '|' || t[0] == '>') throw std::runtime_error("yaml_lite: unsupported YAML feature: " + t);
    n.kind = Node::Scalar; n.scalar = unquote(t);
    return n;
}
inline Node parse_flow(const std::string& t, size_t& i) {   // t[i] == '['
    Node n; n.kind = Node::Seq;
    ++i;
    std::string cur;
    auto flush = [&] { const std::string c = trim(cur); if (!c.empty()) n.seq.push_back(parse_scalar_or_flow(c)); cur.clear(); };
    for (; i < t.size(); ++i) {
        const char c = t[i];
        if (c == '[') { n.seq.push_back(parse_flow(t, i)); cur.clear(); }
        else if (c == ']') { flush(); return n; }
        else if (c == ',') flush();
        else cur += c;
    }
    throw std::runtime_error("yaml_lite: unterminated flow sequence: " + t);
}

// position of the ": " / ":<eol>" that ends a mapping key, or npos
inline size_t key_colon(const std::string& t) {
    char q = 0;
    for (size_t i = 0; i < t.size(); ++i) {
        const char c = t[i];
        if (q) { if (c == q) q = 0; continue; }
        if (c == '"' || c == '\'') { q = c; continue; }
        if (c == '[' || c == '{') return std::string::npos;
        if (c == ':' && (i + 1 == t.size() || t[i + 1] == ' ')) return i;
    }
    return std::string::npos;
}

struct Parser {
    std::vector<Line> lines;
    size_t pos = 0;

    Node parse_block(int indent) {
        if (pos >= lines.size() || lines[pos].indent < indent) return Node();
        const int ind = lines[pos].indent;
        if (lines[pos].text.rfind("- ", 0) == 0 || lines[pos].text == "-") return parse_seq(ind);
        if (key_colon(lines[pos].text) != std::string::npos) return parse_map(ind);
        Node n = parse_scalar_or_flow(lines[pos].text);
        ++pos;
        return n;
    }
    Node parse_seq(int ind) {
        Node n; n.kind = Node::Seq;
        while (pos < lines.size() && lines[pos].indent == ind && (lines[pos].text.rfind("- ", 0) == 0 || lines[pos].text == "-")) {
            const std::string rest = lines[pos].text.size() > 1 ? trim(lines[pos].text.substr(2)) : std::string();
            if (rest.empty()) { ++pos; n.seq.push_back(parse_block(ind + 1)); continue; }
            if (rest.rfind("- ", 0) == 0 || rest == "-") {
                // "- - x": the item is itself a block sequence whose first entry shares the line
                lines[pos].indent = ind + 2; lines[pos].text = rest;
                n.seq.push_back(parse_seq(ind + 2));
            } else if (key_colon(rest) != std::string::npos) {
                // "- key: value" starts a mapping whose keys sit at indent + 2: rewrite the line and parse it as one
                lines[pos].indent = ind + 2; lines[pos].text = rest;
                n.seq.push_back(parse_map(ind + 2));
            } else { n.seq.push_back(parse_scalar_or_flow(rest)); ++pos; }
        }
        return n;
    }
    Node parse_map(int ind) {
        Node n; n.kind = Node::Map;
        while (pos < lines.size() && lines[pos].indent == ind) {
            const std::string& t = lines[pos].text;
            if (t.rfind("- ", 0) == 0 || t == "-") break;
            const size_t c = key_colon(t);
            if (c == std::string::npos) throw std::runtime_error("yaml_lite: line " + std::to_string(lines[pos].no) + ": expected `key: value`");
            const std::string key = unquote(trim(t.substr(0, c))), rest = trim(t.substr(c + 1));
            ++pos;
            Node v;
            if (!rest.empty()) v = parse_scalar_or_flow(rest);
            else if (pos < lines.size() && (lines[pos].indent > ind ||
                     (lines[pos].indent == ind && (lines[pos].text.rfind("- ", 0) == 0 || lines[pos].text == "-"))))
                v = parse_block(lines[pos].indent);         // nested block; a sequence may sit at the key's own indent
            n.map.emplace_back(key, std::move(v));
        }
        return n;
    }
};

}  // namespace detail

inline Node parse(const std::string& text) {
    detail::Parser p;
    std::istringstream in(text);
    std::string raw;
    int no = 0;
    while (std::getline(in, raw)) {
        ++no;
        const std::string s = detail::strip_comment(raw);
        const std::string t = detail::trim(s);
        if (t.empty() || t == "---" || t == "...") continue;
        const int indent = (int)s.find_first_not_of(' ');
        if (s.find('\t') != std::string::npos && s.find('\t') < (size_t)indent + 1) throw std::runtime_error("yaml_lite: tabs in indentation");
        p.lines.push_back({indent, t, no});
    }
    if (p.lines.empty()) return Node();
    Node n = p.parse_block(p.lines[0].indent);
    if (p.pos != p.lines.size()) throw std::runtime_error("yaml_lite: line " + std::to_string(p.lines[p.pos].no) + ": unexpected indentation");
    return n;
}

inline Node parse_file(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("Failed to open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return parse(ss.str());
}

}  // namespace yaml_lite
