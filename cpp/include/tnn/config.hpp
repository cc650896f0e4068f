// TConfig: the universal typed key/value config object of layers, optimizers, schedulers, nodes and graphs
// (reference: include/common/config.hpp:34-103, src/common/config.cpp:89-256).  Serialises to the same
// type-tagged JSON: {"name": .., "type": .., "parameters": {key: {"type": "size_t", "value": 256}, ...}}.
// Nested configs travel as "json" values (TConfig::to_json()), exactly as the reference stores them.
#pragma once
#include <string>
#include <type_traits>
#include <vector>

#include "tnn/json.hpp"

namespace tnn {

class TConfig {
 public:
  std::string name;
  std::string type;

  TConfig() = default;
  TConfig(std::string n, std::string t) : name(std::move(n)), type(std::move(t)) {}

  // ---- setters (type tag chosen from the C++ type, reference names)
  TConfig& set(const std::string& k, bool v) { return put(k, "bool", Json(v)); }
  TConfig& set(const std::string& k, int v) { return put(k, "int", Json(v)); }
  TConfig& set(const std::string& k, long v) { return put(k, "long", Json(v)); }
  TConfig& set(const std::string& k, long long v) { return put(k, "long long", Json(v)); }
  TConfig& set(const std::string& k, unsigned v) { return put(k, "unsigned int", Json(v)); }
  TConfig& set(const std::string& k, unsigned long v) { return put(k, "size_t", Json(v)); }
  TConfig& set(const std::string& k, unsigned long long v) { return put(k, "size_t", Json(v)); }
  TConfig& set(const std::string& k, float v) { return put(k, "float", Json(v)); }
  TConfig& set(const std::string& k, double v) { return put(k, "double", Json(v)); }
  TConfig& set(const std::string& k, const char* v) { return put(k, "string", Json(v)); }
  TConfig& set(const std::string& k, const std::string& v) { return put(k, "string", Json(v)); }
  TConfig& set(const std::string& k, const Json& v) { return put(k, "json", v); }
  TConfig& set(const std::string& k, const TConfig& v) { return put(k, "json", v.to_json()); }
  TConfig& set(const std::string& k, const std::vector<TConfig>& v) {
    Json a = Json::array();
    for (const auto& c : v) a.push_back(c.to_json());
    return put(k, "json", a);
  }
  TConfig& set(const std::string& k, const std::vector<size_t>& v) { return put(k, "vector<size_t>", Json::from_vector(v)); }
  TConfig& set(const std::string& k, const std::vector<int>& v) { return put(k, "vector<int>", Json::from_vector(v)); }
  TConfig& set(const std::string& k, const std::vector<float>& v) { return put(k, "vector<float>", Json::from_vector(v)); }
  TConfig& set(const std::string& k, const std::vector<double>& v) { return put(k, "vector<double>", Json::from_vector(v)); }
  TConfig& set(const std::string& k, const std::vector<std::string>& v) { return put(k, "vector<string>", Json::from_vector(v)); }

  bool has(const std::string& k) const { return find(k) != nullptr; }
  bool remove(const std::string& k) {
    for (size_t i = 0; i < params_.size(); ++i)
      if (params_[i].key == k) { params_.erase(params_.begin() + (long)i); return true; }
    return false;
  }
  std::vector<std::string> keys() const {
    std::vector<std::string> out;
    for (const auto& p : params_) out.push_back(p.key);
    return out;
  }
  const Json* raw(const std::string& k) const { const P* p = find(k); return p ? &p->value : nullptr; }
  std::string type_of(const std::string& k) const { const P* p = find(k); return p ? p->type : std::string(); }

  // ---- typed getters (numeric types convert freely; missing key -> default)
  template <class T>
  T get(const std::string& k, const T& def = T()) const {
    const P* p = find(k);
    if (!p) return def;
    return convert<T>(p->value);
  }
  TConfig get_config(const std::string& k) const {
    const P* p = find(k);
    if (!p || !p->value.is_object()) return TConfig();
    return from_json(p->value);
  }
  std::vector<TConfig> get_config_array(const std::string& k) const {
    std::vector<TConfig> out;
    const P* p = find(k);
    if (p && p->value.is_array())
      for (const auto& e : p->value.items()) out.push_back(from_json(e));
    return out;
  }

  // ---- JSON
  Json to_json() const {
    Json j = Json::object();
    j["name"] = name;
    j["type"] = type;
    Json params = Json::object();
    for (const auto& p : params_) {
      Json tv = Json::object();
      tv["type"] = p.type;
      tv["value"] = p.value;
      params[p.key] = tv;
    }
    j["parameters"] = params;
    return j;
  }
  std::string dump(int indent = 4) const { return to_json().dump(indent); }
  static TConfig from_json(const Json& j) {
    TConfig c;
    c.name = j.value("name", "");
    c.type = j.value("type", "");
    const Json* ps = j.find("parameters");
    if (ps && ps->is_object())
      for (const auto& m : ps->members()) {
        const Json& tv = m.second;
        if (tv.is_object() && tv.contains("type") && tv.contains("value"))
          c.params_.push_back(P{m.first, normalise(tv.at("type").as_string()), upgrade(tv.at("type").as_string(), tv.at("value"))});
        else
          c.params_.push_back(P{m.first, "json", tv});
      }
    return c;
  }
  static TConfig from_string(const std::string& s) { return from_json(Json::parse(s)); }
  bool operator==(const TConfig& o) const { return to_json() == o.to_json(); }

 private:
  struct P {
    std::string key, type;
    Json value;
  };
  std::vector<P> params_;

  const P* find(const std::string& k) const {
    for (const auto& p : params_) if (p.key == k) return &p;
    return nullptr;
  }
  TConfig& put(const std::string& k, const char* t, Json v) {
    for (auto& p : params_)
      if (p.key == k) { p.type = t; p.value = std::move(v); return *this; }
    params_.push_back(P{k, t, std::move(v)});
    return *this;
  }
  // type names written by the round-1 Python host
  static std::string normalise(const std::string& t) {
    if (t == "config" || t == "config_array") return "json";
    if (t == "size_t_array") return "vector<size_t>";
    if (t == "float_array") return "vector<float>";
    if (t == "string_array") return "vector<string>";
    if (t == "int64") return "long long";
    return t;
  }
  static Json upgrade(const std::string&, const Json& v) { return v; }

  template <class T>
  static T convert(const Json& v) {
    if constexpr (std::is_same_v<T, bool>) return v.is_bool() ? v.as_bool() : v.as_int() != 0;
    else if constexpr (std::is_same_v<T, std::string>) return v.as_string();
    else if constexpr (std::is_same_v<T, Json>) return v;
    else if constexpr (std::is_floating_point_v<T>) return (T)v.as_double();
    else if constexpr (std::is_integral_v<T>) return (T)v.as_int();
    else if constexpr (std::is_same_v<T, std::vector<size_t>> || std::is_same_v<T, std::vector<int>> ||
                       std::is_same_v<T, std::vector<long>> || std::is_same_v<T, std::vector<int64_t>>) {
      T out;
      for (const auto& e : v.items()) out.push_back((typename T::value_type)e.as_int());
      return out;
    } else if constexpr (std::is_same_v<T, std::vector<float>> || std::is_same_v<T, std::vector<double>>) {
      T out;
      for (const auto& e : v.items()) out.push_back((typename T::value_type)e.as_double());
      return out;
    } else if constexpr (std::is_same_v<T, std::vector<std::string>>) {
      T out;
      for (const auto& e : v.items()) out.push_back(e.as_string());
      return out;
    } else {
      static_assert(sizeof(T) == 0, "TConfig::get: unsupported type");
    }
  }
};

using LayerConfig = TConfig;
using NodeConfig = TConfig;
using GraphConfig = TConfig;
using OptimizerConfig = TConfig;
using SchedulerConfig = TConfig;

}  // namespace tnn
