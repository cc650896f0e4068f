// Graph IR, parameter slabs, executor, checkpoint files.
#include <algorithm>
#include <cstring>
#include <fstream>
#include <queue>
#include <set>
#include <sstream>

#include "kernels/simt_ops.cuh"
#include "tnn/graph.hpp"
#include "tnn/stepstate.hpp"

namespace tnn {

// ===================================================================================== builder
IONode& GraphBuilder::io(const std::string& uid) {
  for (auto& n : io_nodes_)
    if (n.uid() == uid) return n;
  io_nodes_.emplace_back(uid);
  return io_nodes_.back();
}
OpNode& GraphBuilder::add_layer(std::unique_ptr<Layer> layer, const std::string& uid) {
  const std::string u = uid.empty() ? "op" + std::to_string(op_nodes_.size()) + "_" + layer->name() : uid;
  for (auto& n : op_nodes_)
    if (n->uid() == u) throw std::invalid_argument("GraphBuilder: duplicate op node '" + u + "'");
  op_nodes_.push_back(std::make_unique<OpNode>(u, std::move(layer)));
  return *op_nodes_.back();
}
GraphBuilder& GraphBuilder::add_edge(const Vec<std::string>& inputs, const Vec<std::string>& outputs, const std::string& op_uid) {
  bool found = false;
  for (auto& n : op_nodes_) found |= n->uid() == op_uid;
  if (!found) throw std::invalid_argument("GraphBuilder: unknown op node '" + op_uid + "'");
  for (auto& u : inputs) io(u);
  for (auto& u : outputs) io(u);
  edges_.push_back(Edge{inputs, outputs, op_uid});
  return *this;
}
Vec<Edge> GraphBuilder::topo_sort() const {
  // Kahn over edges: an edge is ready when every producer-side io node it consumes has been produced (or is a graph input)
  std::map<std::string, int> producer_of;          // io uid -> edge index that writes it
  for (size_t i = 0; i < edges_.size(); ++i)
    for (auto& o : edges_[i].consumers) producer_of[o] = (int)i;
  Vec<int> indeg(edges_.size(), 0);
  Vec<Vec<int>> succ(edges_.size());
  for (size_t i = 0; i < edges_.size(); ++i)
    for (auto& in : edges_[i].producers) {
      auto it = producer_of.find(in);
      if (it != producer_of.end()) {
        ++indeg[i];
        succ[(size_t)it->second].push_back((int)i);
      }
    }
  std::queue<int> q;
  for (size_t i = 0; i < edges_.size(); ++i)
    if (!indeg[i]) q.push((int)i);
  Vec<Edge> order;
  while (!q.empty()) {
    const int e = q.front();
    q.pop();
    order.push_back(edges_[(size_t)e]);
    for (int s : succ[(size_t)e])
      if (--indeg[(size_t)s] == 0) q.push(s);
  }
  if (order.size() != edges_.size()) throw std::runtime_error("GraphBuilder: the graph has a cycle");
  return order;
}
Graph GraphBuilder::compile(const Device& dev, uint64_t seed) { return Graph(std::move(*this), dev, seed); }
// the reference signature (graph_builder.hpp compile(allocator)): the parameter / gradient slabs come from `allocator`
Graph GraphBuilder::compile(IAllocator& allocator, uint64_t seed) { return Graph(std::move(*this), allocator.device(), seed, &allocator); }

// ===================================================================================== context
GraphContext::GraphContext(const Vec<Layer*>& layers, const Device& dev, IAllocator* slabs) : dev_(dev), layers_(layers) {
  if (slabs && slabs->device() != dev) throw std::invalid_argument("GraphContext: the slab allocator serves " + slabs->device().name() + ", the graph lives on " + dev.name());
  auto slab = [&](DType_t dt, size_t n) { return slabs ? make_tensor(*slabs, dt, {n}) : make_tensor(dt, {n}, dev); };
  size_t off = 0, boff = 0;
  auto aligned = [](size_t n) { return (n + ALIGN_ELEMS - 1) / ALIGN_ELEMS * ALIGN_ELEMS; };
  for (Layer* l : layers)
    for (const auto& d : l->param_descriptors()) {
      const size_t n = shape_numel(d.shape);
      if (d.trainable) {
        entries_.push_back(Entry{l, d.name, off, n, true});
        off += aligned(n);
      } else {
        entries_.push_back(Entry{l, d.name, boff, n, false});
        boff += aligned(n);
      }
    }
  numel_ = off;
  const size_t pn = std::max(off, ALIGN_ELEMS), bn = std::max(boff, ALIGN_ELEMS);
  param_slab_ = slab(DType_t::FP32, pn);
  grad_slab_ = slab(DType_t::FP32, pn);
  buffer_slab_ = slab(DType_t::FP32, bn);
  param_slab_->zero();
  grad_slab_->zero();
  buffer_slab_->zero();
  if (dev.is_gpu()) {
    lp_slab_ = slab(DType_t::BF16, pn);
    lp_slab_->zero();
  }
  for (const auto& e : entries_) {
    Shape shape;
    for (const auto& d : e.layer->param_descriptors())
      if (d.name == e.name) shape = d.shape;
    Tensor src = e.layer->param(e.name);
    if (e.trainable) {
      Tensor pv = param_slab_->span(e.offset, shape);
      pv->copy_from(*src);
      Tensor gv = grad_slab_->span(e.offset, shape);
      Tensor lv = lp_slab_ ? lp_slab_->span(e.offset, shape) : nullptr;
      e.layer->bind(e.name, pv, gv, lv);
    } else {
      Tensor bv = buffer_slab_->span(e.offset, shape);
      bv->copy_from(*src);
      e.layer->bind(e.name, bv, nullptr, nullptr);
    }
  }
  refresh_lp();
}
size_t GraphContext::slab_bytes(const Vec<Layer*>& layers, bool gpu) {
  size_t off = 0, boff = 0;
  auto aligned = [](size_t n) { return (n + ALIGN_ELEMS - 1) / ALIGN_ELEMS * ALIGN_ELEMS; };
  for (Layer* l : layers)
    for (const auto& d : l->param_descriptors()) (d.trainable ? off : boff) += aligned(shape_numel(d.shape));
  const size_t pn = std::max(off, ALIGN_ELEMS), bn = std::max(boff, ALIGN_ELEMS);
  auto r256 = [](size_t b) { return (b + 255) / 256 * 256; };
  return 2 * r256(pn * 4) + r256(bn * 4) + (gpu ? r256(pn * 2) : 0);
}
size_t GraphContext::num_parameters() const {
  size_t n = 0;
  for (const auto& e : entries_)
    if (e.trainable) n += e.numel;
  return n;
}
void GraphContext::zero_grads() { grad_slab_->zero(); }
void GraphContext::refresh_lp() {
  if (lp_slab_) {
    dev_.set_current();
    cast_f32_to_bf16(param_slab_->data_as<float>(), lp_slab_->data(), (long long)param_slab_->size(), current_flow(dev_).stream());
    for (Layer* l : layers_) l->mark_lp_fresh();
  } else {
    for (Layer* l : layers_) l->mark_params_dirty();
  }
}
void GraphContext::mark_lp_fresh() {
  for (Layer* l : layers_) {
    if (lp_slab_) l->mark_lp_fresh();
    else l->mark_params_dirty();
  }
}
Vec<std::pair<std::string, Tensor>> GraphContext::parameters() const {
  Vec<std::pair<std::string, Tensor>> out;
  std::set<Layer*> seen;
  for (Layer* l : layers_) {
    if (!seen.insert(l).second) continue;
    for (int pass = 0; pass < 2; ++pass)
      for (const auto& e : entries_)
        if (e.layer == l && e.trainable == (pass == 0)) out.emplace_back(l->name() + "." + e.name, l->param(e.name));
  }
  return out;
}

// ===================================================================================== graph
Graph::Graph(GraphBuilder&& b, const Device& dev, uint64_t seed, IAllocator* slabs) : name_(b.name_), dev_(dev) {
  edges_ = b.topo_sort();
  io_nodes_ = std::move(b.io_nodes_);
  op_nodes_ = std::move(b.op_nodes_);
  Vec<Layer*> flat;
  for (auto& n : op_nodes_) {
    if (!n->layer()->initialized() || n->layer()->device() != dev) n->layer()->init(dev, seed);
    auto all = n->layer()->all_layers();
    flat.insert(flat.end(), all.begin(), all.end());
  }
  ctx_ = std::make_unique<GraphContext>(flat, dev, slabs);
}
void Graph::set_training(bool f) {
  for (auto& n : op_nodes_) n->layer()->set_training(f);
}
void Graph::set_seed(uint64_t s) {
  for (auto& n : op_nodes_) n->layer()->set_seed(s);
}
Vec<Layer*> Graph::get_layers() {
  Vec<Layer*> v;
  for (auto& n : op_nodes_) v.push_back(n->layer());
  return v;
}
OpNode& Graph::op(const std::string& uid) {
  for (auto& n : op_nodes_)
    if (n->uid() == uid) return *n;
  throw std::out_of_range("Graph: no op node '" + uid + "'");
}
GraphConfig Graph::get_config() const {
  GraphConfig c(name_, "graph");
  Json ops = Json::array(), ios = Json::array(), es = Json::array();
  for (auto& n : op_nodes_) ops.push_back(n->get_config().to_json());
  for (auto& n : io_nodes_) ios.push_back(n.get_config().to_json());
  for (auto& e : edges_) {
    Json j = Json::object();
    j["producers"] = Json::from_vector(e.producers);
    j["consumers"] = Json::from_vector(e.consumers);
    j["op_node"] = e.op;
    es.push_back(j);
  }
  c.set("op_nodes", ops).set("io_nodes", ios).set("edges", es);
  return c;
}
Graph Graph::create_from_config(const GraphConfig& cfg, const Device& dev, uint64_t seed) {
  GraphBuilder b(cfg.name.empty() ? "graph" : cfg.name);
  const Json* ios = cfg.raw("io_nodes");
  const Json* ops = cfg.raw("op_nodes");
  const Json* es = cfg.raw("edges");
  if (!ios || !ops || !es) throw std::runtime_error("Graph::create_from_config: io_nodes / op_nodes / edges missing");
  for (const auto& j : ios->items()) b.io(NodeConfig::from_json(j).get<std::string>("uid"));
  for (const auto& j : ops->items()) {
    const NodeConfig nc = NodeConfig::from_json(j);
    b.add_layer(LayerFactory::create(nc.get_config("layer_config")), nc.get<std::string>("uid"));
  }
  for (const auto& j : es->items()) {
    Vec<std::string> in, out;
    for (const auto& u : j.at("producers").items()) in.push_back(u.as_string());
    for (const auto& u : j.at("consumers").items()) out.push_back(u.as_string());
    b.add_edge(in, out, j.at("op_node").as_string());
  }
  return b.compile(dev, seed);
}

static const char kTrailerMagic[8] = {'T', 'N', 'N', 'T', 'R', 'L', 'R', '1'};
void Graph::save_state(std::ostream& os, const std::map<std::string, Tensor>* tensors, const Json* meta) const {
  os << get_config().to_json().dump(4);
  for (auto& p : ctx_->parameters()) p.second->save(os);
  if (tensors || meta) {
    // optional trailer (the reference format ends with the last parameter record): magic, JSON meta, named tensor records
    os.write(kTrailerMagic, 8);
    const std::string m = (meta ? *meta : Json::object()).dump();
    const uint64_t ml = m.size(), nt = tensors ? tensors->size() : 0;
    os.write(reinterpret_cast<const char*>(&ml), 8);
    os.write(m.data(), (std::streamsize)ml);
    os.write(reinterpret_cast<const char*>(&nt), 8);
    if (tensors)
      for (auto& kv : *tensors) {
        const uint64_t kl = kv.first.size();
        os.write(reinterpret_cast<const char*>(&kl), 8);
        os.write(kv.first.data(), (std::streamsize)kl);
        kv.second->save(os);
      }
  }
}
void Graph::save_state(const std::string& path, const std::map<std::string, Tensor>* tensors, const Json* meta) const {
  std::ofstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("Graph::save_state: cannot open " + path);
  save_state(f, tensors, meta);
}
void Graph::load_state_tensors(std::istream& is) {
  for (auto& p : ctx_->parameters()) load_into(is, *p.second);
  ctx_->refresh_lp();
}
Graph Graph::load_state(std::istream& is, const Device& dev, std::map<std::string, Tensor>* tensors, Json* meta) {
  std::stringstream ss;
  ss << is.rdbuf();
  const std::string data = ss.str();
  size_t pos = 0;
  const Json j = Json::parse(data, &pos);
  Graph g = create_from_config(GraphConfig::from_json(j), dev);
  std::istringstream rest(data.substr(pos), std::ios::binary);
  g.load_state_tensors(rest);
  char magic[8];
  rest.read(magic, 8);
  if (rest && std::memcmp(magic, kTrailerMagic, 8) == 0) {
    uint64_t ml = 0, nt = 0;
    rest.read(reinterpret_cast<char*>(&ml), 8);
    std::string m(ml, '\0');
    rest.read(&m[0], (std::streamsize)ml);
    if (meta) *meta = Json::parse(m);
    rest.read(reinterpret_cast<char*>(&nt), 8);
    for (uint64_t i = 0; i < nt && rest; ++i) {
      uint64_t kl = 0;
      rest.read(reinterpret_cast<char*>(&kl), 8);
      std::string k(kl, '\0');
      rest.read(&k[0], (std::streamsize)kl);
      Tensor t = load(rest, dev.is_gpu() ? static_cast<IAllocator&>(PoolAllocator::instance(dev)) : HostAllocator());
      if (tensors) (*tensors)[k] = t;
    }
  }
  return g;
}
Graph Graph::load_state(const std::string& path, const Device& dev, std::map<std::string, Tensor>* tensors, Json* meta) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("Graph::load_state: cannot open " + path);
  return load_state(f, dev, tensors, meta);
}

Graph graph_from_sequential(std::unique_ptr<Layer> model, const Device& dev, uint64_t seed, const std::string& name, IAllocator* slabs) {
  GraphBuilder b(name.empty() ? model->name() : name);
  b.io("input");
  b.io("output");
  OpNode& n = b.add_layer(std::move(model), "model");
  b.add_edge({"input"}, {"output"}, n);
  return Graph(std::move(b), dev, seed, slabs);
}

// ===================================================================================== executor
GraphExecutor::GraphExecutor(Graph& g, IAllocator* backing) : g_(g) {
  IAllocator& back = backing ? *backing : static_cast<IAllocator&>(DeviceAllocator::instance(g.device()));
  ws_ = std::make_unique<DELAllocatorV2>(back, g.device().is_gpu() ? ((size_t)256 << 20) : ((size_t)16 << 20));
  for (Layer* l : g_.get_layers()) l->set_allocator(ws_.get());
}
GraphExecutor::~GraphExecutor() {
  for (Layer* l : g_.get_layers()) {
    for (Layer* s : l->all_layers()) s->clear_caches();
    l->set_allocator(nullptr);
  }
}
void GraphExecutor::forward(const InputPack& inputs, OutputPack& outputs, size_t mb_id) {
  std::map<std::string, Tensor> vals(inputs.begin(), inputs.end());
  std::set<std::string> produced, consumed;
  for (const Edge& e : g_.edges()) {
    Layer* layer = g_.op(e.op).layer();
    Vec<Tensor> ins;
    for (auto& u : e.producers) {
      auto it = vals.find(u);
      if (it == vals.end()) throw std::runtime_error("GraphExecutor: io node '" + u + "' has no value");
      ins.push_back(it->second);
      consumed.insert(u);
    }
    Vec<Tensor> outs = ins.size() == 1 ? Vec<Tensor>{layer->forward(ins[0], mb_id)} : layer->forward(ins, mb_id);
    if (outs.size() < e.consumers.size()) outs = layer->forward(ins, mb_id);
    for (size_t i = 0; i < e.consumers.size(); ++i) {
      vals[e.consumers[i]] = outs.at(i);
      produced.insert(e.consumers[i]);
    }
    ws_->flip();
  }
  for (auto& u : produced)
    if (!consumed.count(u)) outputs[u] = vals[u];
}
void GraphExecutor::backward(const InputPack& grads, OutputPack& grad_inputs, size_t mb_id) {
  std::map<std::string, Tensor> g(grads.begin(), grads.end());
  std::set<std::string> produced;
  for (const Edge& e : g_.edges())
    for (auto& u : e.consumers) produced.insert(u);
  const Vec<Edge>& edges = g_.edges();
  for (size_t k = edges.size(); k-- > 0;) {
    const Edge& e = edges[k];
    Layer* layer = g_.op(e.op).layer();
    Vec<Tensor> gouts;
    bool any = false;
    for (auto& u : e.consumers) {
      auto it = g.find(u);
      gouts.push_back(it == g.end() ? nullptr : it->second);
      any |= it != g.end() && it->second;
    }
    if (!any) continue;
    Vec<Tensor> gins;
    if (gouts.size() == 1 && e.producers.size() == 1) {
      // graph inputs that nothing produces need no data gradient unless the caller asked for grad_inputs
      BwdAux aux;
      gins = {layer->backward_impl(gouts[0], mb_id, aux)};
    } else {
      gins = layer->backward(gouts, mb_id);
    }
    for (size_t i = 0; i < e.producers.size() && i < gins.size(); ++i) {
      if (!gins[i]) continue;
      Tensor& slot = g[e.producers[i]];
      if (!slot) slot = gins[i];
      else {
        Tensor sum = gins[i]->clone(ws_.get());
        sum->add(*slot);
        slot = sum;
      }
    }
    ws_->flip();
  }
  if (g_.device().is_gpu()) side_flow::join(g_.device());       // weight gradients issued on the side flow are complete from here on
  for (auto& kv : g)
    if (!produced.count(kv.first) && !grads.count(kv.first) && kv.second) grad_inputs[kv.first] = kv.second;
}
Tensor GraphExecutor::forward(const Tensor& x, size_t mb_id) {
  OutputPack out;
  forward(InputPack{{"input", x}}, out, mb_id);
  if (out.empty()) throw std::runtime_error("GraphExecutor: the graph produced no output");
  auto it = out.find("output");
  return it != out.end() ? it->second : out.begin()->second;
}
Tensor GraphExecutor::backward(const Tensor& dy, size_t mb_id) {
  OutputPack gin;
  backward(InputPack{{"output", dy}}, gin, mb_id);
  auto it = gin.find("input");
  return it != gin.end() ? it->second : nullptr;
}

}  // namespace tnn
