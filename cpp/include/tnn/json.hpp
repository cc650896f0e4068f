// Minimal JSON value (parse + dump) for TConfig / checkpoint headers / CLI configs.
// The reference depends on nlohmann::json (cmake/Dependencies.cmake:6-57, network FetchContent); this
// framework vendors nothing from the network, so the 1 % of JSON it needs lives here.
// Objects keep insertion order (checkpoint headers diff cleanly); numbers are int64 or double.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace tnn {

class Json {
 public:
  enum Kind { Null, Bool, Int, Real, Str, Arr, Obj };
  using Array = std::vector<Json>;
  using Member = std::pair<std::string, Json>;
  using Object = std::vector<Member>;

  Json() = default;
  Json(std::nullptr_t) {}
  Json(bool b) : kind_(Bool), b_(b) {}
  Json(int v) : kind_(Int), i_(v) {}
  Json(unsigned v) : kind_(Int), i_(v) {}
  Json(long v) : kind_(Int), i_(v) {}
  Json(unsigned long v) : kind_(Int), i_((int64_t)v) {}
  Json(long long v) : kind_(Int), i_(v) {}
  Json(unsigned long long v) : kind_(Int), i_((int64_t)v) {}
  Json(double v) : kind_(Real), d_(v) {}
  Json(float v) : kind_(Real), d_(shortest_float(v)) {}
  Json(const char* s) : kind_(Str), s_(s) {}
  Json(std::string s) : kind_(Str), s_(std::move(s)) {}
  Json(Array a) : kind_(Arr), a_(std::make_shared<Array>(std::move(a))) {}
  static Json array() { Json j; j.kind_ = Arr; j.a_ = std::make_shared<Array>(); return j; }
  static Json object() { Json j; j.kind_ = Obj; j.o_ = std::make_shared<Object>(); return j; }
  template <class T>
  static Json from_vector(const std::vector<T>& v) {
    Json j = array();
    for (const auto& e : v) j.push_back(Json(e));
    return j;
  }

  Kind kind() const { return kind_; }
  bool is_null() const { return kind_ == Null; }
  bool is_bool() const { return kind_ == Bool; }
  bool is_number() const { return kind_ == Int || kind_ == Real; }
  bool is_int() const { return kind_ == Int; }
  bool is_string() const { return kind_ == Str; }
  bool is_array() const { return kind_ == Arr; }
  bool is_object() const { return kind_ == Obj; }

  bool as_bool() const { need(Bool); return b_; }
  int64_t as_int() const {
    if (kind_ == Real) return (int64_t)d_;
    if (kind_ == Bool) return b_ ? 1 : 0;
    need(Int);
    return i_;
  }
  double as_double() const {
    if (kind_ == Int) return (double)i_;
    need(Real);
    return d_;
  }
  const std::string& as_string() const { need(Str); return s_; }
  const Array& items() const { need(Arr); return *a_; }
  Array& items() { need(Arr); return *a_; }
  const Object& members() const { need(Obj); return *o_; }
  size_t size() const { return kind_ == Arr ? a_->size() : kind_ == Obj ? o_->size() : 0; }
  bool empty() const { return size() == 0; }

  void push_back(Json v) { if (kind_ == Null) *this = array(); need(Arr); a_->push_back(std::move(v)); }
  const Json& operator[](size_t i) const { need(Arr); return a_->at(i); }
  bool contains(const std::string& k) const { return kind_ == Obj && find(k) != nullptr; }
  const Json* find(const std::string& k) const {
    if (kind_ != Obj) return nullptr;
    for (const auto& m : *o_) if (m.first == k) return &m.second;
    return nullptr;
  }
  const Json& at(const std::string& k) const {
    const Json* p = find(k);
    if (!p) throw std::out_of_range("json: missing key '" + k + "'");
    return *p;
  }
  Json& operator[](const std::string& k) {
    if (kind_ == Null) *this = object();
    need(Obj);
    for (auto& m : *o_) if (m.first == k) return m.second;
    o_->emplace_back(k, Json());
    return o_->back().second;
  }
  // typed getters with defaults
  int64_t value(const std::string& k, int64_t def) const { const Json* p = find(k); return p && p->is_number() ? p->as_int() : def; }
  int value(const std::string& k, int def) const { return (int)value(k, (int64_t)def); }
  double value(const std::string& k, double def) const { const Json* p = find(k); return p && p->is_number() ? p->as_double() : def; }
  bool value(const std::string& k, bool def) const { const Json* p = find(k); return p && p->is_bool() ? p->as_bool() : def; }
  std::string value(const std::string& k, const std::string& def) const { const Json* p = find(k); return p && p->is_string() ? p->as_string() : def; }
  std::string value(const std::string& k, const char* def) const { return value(k, std::string(def)); }

  bool operator==(const Json& o) const { return dump() == o.dump(); }

  // ---------------------------------------------------------------- dump
  std::string dump(int indent = -1) const {
    std::string out;
    write(out, indent, 0);
    return out;
  }

  // ---------------------------------------------------------------- parse
  // Parses ONE value starting at text[pos]; pos is left just past it (checkpoint files append binary records).
  static Json parse(const std::string& text, size_t* pos_io = nullptr) {
    size_t pos = pos_io ? *pos_io : 0;
    Json j = parse_value(text, pos);
    if (pos_io) *pos_io = pos;
    else {
      skip_ws(text, pos);
      if (pos != text.size()) throw std::runtime_error("json: trailing characters at offset " + std::to_string(pos));
    }
    return j;
  }

 private:
  // the double nearest to the SHORTEST decimal that still reads back as the same float (1e-5f dumps as 1e-05, not 9.99999974e-06)
  static double shortest_float(float v) {
    for (int p = 1; p <= 9; ++p) {
      char b[40];
      snprintf(b, sizeof b, "%.*g", p, (double)v);
      if (strtof(b, nullptr) == v) return strtod(b, nullptr);
    }
    return (double)v;
  }
  Kind kind_ = Null;
  bool b_ = false;
  int64_t i_ = 0;
  double d_ = 0;
  std::string s_;
  std::shared_ptr<Array> a_;
  std::shared_ptr<Object> o_;

  void need(Kind k) const {
    if (kind_ != k) throw std::runtime_error(std::string("json: wrong type (have ") + std::to_string((int)kind_) + ", want " + std::to_string((int)k) + ")");
  }
  static void escape(std::string& out, const std::string& s) {
    out.push_back('"');
    for (unsigned char c : s) {
      switch (c) {
        case '"': out += "\\\""; break;
        case '\\': out += "\\\\"; break;
        case '\n': out += "\\n"; break;
        case '\r': out += "\\r"; break;
        case '\t': out += "\\t"; break;
        default:
          if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); out += b; }
          else out.push_back((char)c);
      }
    }
    out.push_back('"');
  }
  void write(std::string& out, int indent, int depth) const {
    auto nl = [&](int d) { if (indent >= 0) { out.push_back('\n'); out.append((size_t)(indent * d), ' '); } };
    switch (kind_) {
      case Null: out += "null"; break;
      case Bool: out += b_ ? "true" : "false"; break;
      case Int: out += std::to_string(i_); break;
      case Real: {
        char b[40];
        snprintf(b, sizeof b, "%.17g", d_);
        // shortest representation that round-trips
        for (int p = 1; p < 17; ++p) { char t[40]; snprintf(t, sizeof t, "%.*g", p, d_); if (strtod(t, nullptr) == d_) { strcpy(b, t); break; } }
        out += b;
        if (!strpbrk(b, ".eEni")) out += ".0";
        break;
      }
      case Str: escape(out, s_); break;
      case Arr:
        if (a_->empty()) { out += "[]"; break; }
        out.push_back('[');
        for (size_t i = 0; i < a_->size(); ++i) { if (i) out.push_back(','); nl(depth + 1); (*a_)[i].write(out, indent, depth + 1); }
        nl(depth);
        out.push_back(']');
        break;
      case Obj:
        if (o_->empty()) { out += "{}"; break; }
        out.push_back('{');
        for (size_t i = 0; i < o_->size(); ++i) {
          if (i) out.push_back(',');
          nl(depth + 1);
          escape(out, (*o_)[i].first);
          out += indent >= 0 ? ": " : ":";
          (*o_)[i].second.write(out, indent, depth + 1);
        }
        nl(depth);
        out.push_back('}');
        break;
    }
  }
  static void skip_ws(const std::string& t, size_t& p) { while (p < t.size() && (t[p] == ' ' || t[p] == '\n' || t[p] == '\r' || t[p] == '\t')) ++p; }
  static Json parse_value(const std::string& t, size_t& p) {
    skip_ws(t, p);
    if (p >= t.size()) throw std::runtime_error("json: unexpected end");
    const char c = t[p];
    if (c == '{') {
      Json j = object();
      ++p;
      skip_ws(t, p);
      if (p < t.size() && t[p] == '}') { ++p; return j; }
      while (true) {
        skip_ws(t, p);
        if (p >= t.size() || t[p] != '"') throw std::runtime_error("json: expected key at " + std::to_string(p));
        std::string k = parse_string(t, p);
        skip_ws(t, p);
        if (p >= t.size() || t[p] != ':') throw std::runtime_error("json: expected ':' at " + std::to_string(p));
        ++p;
        j.o_->emplace_back(std::move(k), parse_value(t, p));
        skip_ws(t, p);
        if (p < t.size() && t[p] == ',') { ++p; continue; }
        if (p < t.size() && t[p] == '}') { ++p; return j; }
        throw std::runtime_error("json: expected ',' or '}' at " + std::to_string(p));
      }
    }
    if (c == '[') {
      Json j = array();
      ++p;
      skip_ws(t, p);
      if (p < t.size() && t[p] == ']') { ++p; return j; }
      while (true) {
        j.a_->push_back(parse_value(t, p));
        skip_ws(t, p);
        if (p < t.size() && t[p] == ',') { ++p; continue; }
        if (p < t.size() && t[p] == ']') { ++p; return j; }
        throw std::runtime_error("json: expected ',' or ']' at " + std::to_string(p));
      }
    }
    if (c == '"') return Json(parse_string(t, p));
    if (t.compare(p, 4, "true") == 0) { p += 4; return Json(true); }
    if (t.compare(p, 5, "false") == 0) { p += 5; return Json(false); }
    if (t.compare(p, 4, "null") == 0) { p += 4; return Json(); }
    // number
    size_t q = p;
    bool real = false;
    if (q < t.size() && (t[q] == '-' || t[q] == '+')) ++q;
    while (q < t.size() && (isdigit((unsigned char)t[q]) || t[q] == '.' || t[q] == 'e' || t[q] == 'E' || t[q] == '-' || t[q] == '+')) {
      if (t[q] == '.' || t[q] == 'e' || t[q] == 'E') real = true;
      ++q;
    }
    if (q == p) throw std::runtime_error("json: unexpected character '" + std::string(1, c) + "' at " + std::to_string(p));
    const std::string num = t.substr(p, q - p);
    p = q;
    if (real) return Json(strtod(num.c_str(), nullptr));
    errno = 0;
    const long long v = strtoll(num.c_str(), nullptr, 10);
    if (errno == ERANGE) return Json((double)strtod(num.c_str(), nullptr));
    return Json(v);
  }
  static std::string parse_string(const std::string& t, size_t& p) {
    std::string out;
    ++p;  // opening quote
    while (p < t.size()) {
      char c = t[p++];
      if (c == '"') return out;
      if (c != '\\') { out.push_back(c); continue; }
      if (p >= t.size()) break;
      char e = t[p++];
      switch (e) {
        case 'n': out.push_back('\n'); break;
        case 't': out.push_back('\t'); break;
        case 'r': out.push_back('\r'); break;
        case 'b': out.push_back('\b'); break;
        case 'f': out.push_back('\f'); break;
        case 'u': {
          unsigned cp = (unsigned)strtoul(t.substr(p, 4).c_str(), nullptr, 16);
          p += 4;
          if (cp < 0x80) out.push_back((char)cp);
          else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
          else { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
          break;
        }
        default: out.push_back(e);
      }
    }
    throw std::runtime_error("json: unterminated string");
  }
};

}  // namespace tnn
