// Legacy (NCHW) zoo: config-tree rewrite of the NHWC zoo + NCHW batch adapter.  See tnn/legacy.hpp.
#include "tnn/legacy.hpp"

#include <sys/stat.h>

namespace tnn {
namespace legacy {
namespace {

const char* twin(const std::string& t) {
  if (t == "conv2d") return "legacy_conv2d";
  if (t == "batchnorm") return "legacy_batchnorm";
  if (t == "maxpool2d") return "legacy_maxpool2d";
  if (t == "avgpool2d") return "legacy_avgpool2d";
  if (t == "dense") return "legacy_dense";
  return nullptr;
}
int64_t param_int(const Json& layer, const char* key, int64_t def) {
  const Json* ps = layer.find("parameters");
  const Json* p = ps ? ps->find(key) : nullptr;
  if (!p) return def;
  if (p->is_object() && p->contains("value")) return p->at("value").is_number() ? p->at("value").as_int() : def;
  return p->is_number() ? p->as_int() : def;
}
Json with_param_value(const Json& layer, const char* key, int64_t v) {
  Json out = Json::object();
  for (const auto& m : layer.members()) {
    if (m.first != "parameters") {
      out[m.first] = m.second;
      continue;
    }
    Json ps = Json::object();
    for (const auto& p : m.second.members()) {
      if (p.first == key && p.second.is_object()) {
        Json tv = Json::object();
        for (const auto& f : p.second.members()) tv[f.first] = f.first == "value" ? Json((long long)v) : f.second;
        ps[p.first] = tv;
      } else {
        ps[p.first] = p.second;
      }
    }
    out["parameters"] = ps;
  }
  return out;
}
bool is_layer(const Json& j) { return j.is_object() && j.contains("type") && j.contains("parameters"); }

Json convert(const Json& node) {
  if (node.is_array()) {
    Json out = Json::array();
    for (const Json& item : node.items()) {
      Json conv = convert(item);
      const Json* prev = out.size() ? &out[out.size() - 1] : nullptr;
      if (is_layer(conv) && conv.value("type", "") == "flatten" && prev && is_layer(*prev) && prev->value("type", "") == "legacy_conv2d" &&
          param_int(conv, "start_dim", 1) == 1 && param_int(conv, "end_dim", -1) == 2) {
        // NHWC patch flatten (H, W) -> NCHW: flatten dims 2..3, then [N, E, T] -> [N, T, E]
        Json f = with_param_value(with_param_value(conv, "start_dim", 2), "end_dim", 3);
        const std::string fname = f.value("name", "flatten");
        out.push_back(f);
        Json t = Json::object();
        t["name"] = fname + "_to_tokens";
        t["type"] = "transpose";
        t["parameters"] = Json::object();
        out.push_back(t);
        continue;
      }
      out.push_back(conv);
    }
    return out;
  }
  if (node.is_object()) {
    Json out = Json::object();
    for (const auto& m : node.members()) out[m.first] = convert(m.second);
    if (is_layer(out))
      if (const char* t = twin(out.value("type", ""))) out["type"] = t;
    return out;
  }
  return node;
}

}  // namespace

Json to_legacy_config(const Json& nhwc_config) { return convert(nhwc_config); }
std::unique_ptr<Layer> to_legacy(Layer& nhwc_model) {
  LayerFactory::register_defaults();
  return LayerFactory::create(TConfig::from_json(to_legacy_config(nhwc_model.get_config().to_json())));
}
std::unique_ptr<Layer> ExampleModels::create(const std::string& name, DType_t io) {
  std::unique_ptr<Sequential> m = tnn::ExampleModels::create(name, io);
  return to_legacy(*m);
}
Graph load_or_create_model(const std::string& name, const std::string& path, const Device& dev, uint64_t seed) {
  struct stat st;
  if (!path.empty() && stat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode)) return Graph::load_state(path, dev);
  return graph_from_sequential(ExampleModels::create(name), dev, seed, name);
}

Shape NCHWLoader::get_data_shape() const {
  Shape s = inner_->get_data_shape();
  return s.size() == 3 ? Shape{s[2], s[0], s[1]} : s;
}
bool NCHWLoader::get_batch(size_t batch_size, Tensor& x, Tensor& y) {
  Tensor nhwc;
  if (!inner_->get_batch(batch_size, nhwc, y)) return false;
  if (nhwc->shape().size() != 4) {
    x = nhwc;
    return true;
  }
  const size_t N = nhwc->dimension(0), H = nhwc->dimension(1), W = nhwc->dimension(2), C = nhwc->dimension(3);
  x = make_tensor(DType_t::FP32, {N, C, H, W});
  const float* s = nhwc->data_as<float>();
  float* d = x->data_as<float>();
  for (size_t n = 0; n < N; ++n)
    for (size_t p = 0; p < H * W; ++p)
      for (size_t c = 0; c < C; ++c) d[(n * C + c) * H * W + p] = s[(n * H * W + p) * C + c];
  return true;
}
DataLoaderPair DataLoaderFactory::create(const std::string& name, const std::string& path) {
  DataLoaderPair p = tnn::DataLoaderFactory::create(name, path);
  DataLoaderPair out;
  if (p.train) out.train = std::make_unique<NCHWLoader>(std::move(p.train));
  if (p.val) out.val = std::make_unique<NCHWLoader>(std::move(p.val));
  return out;
}

}  // namespace legacy
}  // namespace tnn
