// Graph IR: IONode / OpNode / Edge / GraphBuilder / Graph / GraphContext / GraphExecutor
// (reference: include/nn/node.hpp, io_node.hpp, op_node.hpp, edge.hpp, graph_builder.hpp:17-136,
//  graph_context.hpp:19-89, graph.hpp:18-191, graph_executor.hpp:15-163).
// GraphContext packs ALL trainable parameters into one fp32 slab with a gradient slab (zeroed by one memset / by the
// optimizer kernel) and, on the GPU, a bf16 compute-copy slab beside it; buffers (BatchNorm running statistics) live
// in a fourth slab so they serialise with the parameters.
#pragma once
#include <iosfwd>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "tnn/layers.hpp"

namespace tnn {

class IONode {
 public:
  explicit IONode(std::string uid) : uid_(std::move(uid)) {}
  const std::string& uid() const { return uid_; }
  NodeConfig get_config() const { NodeConfig c("", "io_node"); c.set("uid", uid_); return c; }

 private:
  std::string uid_;
};

class OpNode {
 public:
  OpNode(std::string uid, std::unique_ptr<Layer> layer) : uid_(std::move(uid)), layer_(std::move(layer)) {}
  const std::string& uid() const { return uid_; }
  Layer* layer() { return layer_.get(); }
  const Layer* layer() const { return layer_.get(); }
  NodeConfig get_config() const {
    NodeConfig c("", "op_node");
    c.set("uid", uid_).set("name", layer_->name()).set("layer_config", layer_->get_config());
    return c;
  }

 private:
  std::string uid_;
  std::unique_ptr<Layer> layer_;
};

struct Edge {
  Vec<std::string> producers;     // io-node uids consumed by the op
  Vec<std::string> consumers;     // io-node uids produced by the op
  std::string op;                 // op-node uid
};

class Graph;
class GraphBuilder {
 public:
  explicit GraphBuilder(std::string name = "graph") : name_(std::move(name)) {}
  IONode& io(const std::string& uid);
  OpNode& add_layer(std::unique_ptr<Layer> layer, const std::string& uid = "");
  GraphBuilder& add_edge(const Vec<std::string>& inputs, const Vec<std::string>& outputs, const std::string& op_uid);
  GraphBuilder& add_edge(const Vec<std::string>& inputs, const Vec<std::string>& outputs, const OpNode& op) { return add_edge(inputs, outputs, op.uid()); }
  // Kahn topological order of the edges; throws on a cycle
  Vec<Edge> topo_sort() const;
  Graph compile(const Device& dev, uint64_t seed = 0);
  Graph compile(IAllocator& allocator, uint64_t seed = 0);

 private:
  friend class Graph;
  std::string name_;
  Vec<IONode> io_nodes_;
  Vec<std::unique_ptr<OpNode>> op_nodes_;
  Vec<Edge> edges_;
};

class GraphContext {
 public:
  // `slabs`: where the four slabs come from (default: the device allocator).  A SymmetricAllocator puts them into
  // peer-addressable / switch-reducible memory (data_parallel.hpp); the allocation order is params, grads, buffers, bf16.
  GraphContext(const Vec<Layer*>& layers, const Device& dev, IAllocator* slabs = nullptr);
  // bytes the slabs of these layers take from `slabs` (each rounded up to 256 bytes): sizes a symmetric slab up front
  static size_t slab_bytes(const Vec<Layer*>& layers, bool gpu);
  const Device& device() const { return dev_; }
  Tensor& param_slab() { return param_slab_; }
  Tensor& grad_slab() { return grad_slab_; }
  Tensor& lp_slab() { return lp_slab_; }          // null on the CPU
  Tensor& buffer_slab() { return buffer_slab_; }
  size_t numel() const { return numel_; }
  size_t num_parameters() const;
  void zero_grads();
  void refresh_lp();                              // bf16 copies <- fp32 masters (one cast kernel over the slab)
  void mark_lp_fresh();                           // the optimizer kernel just rewrote the bf16 slab
  // (qualified name, tensor) in registration order: per layer, trainable parameters then buffers
  Vec<std::pair<std::string, Tensor>> parameters() const;
  const Vec<Layer*>& layers() const { return layers_; }
  struct Entry {
    Layer* layer;
    std::string name;
    size_t offset, numel;
    bool trainable;
  };
  const Vec<Entry>& entries() const { return entries_; }
  static constexpr size_t ALIGN_ELEMS = 64;       // 256 bytes of fp32

 private:
  Device dev_;
  Vec<Layer*> layers_;
  Vec<Entry> entries_;
  Tensor param_slab_, grad_slab_, lp_slab_, buffer_slab_;
  size_t numel_ = 0;
};

class Graph {
 public:
  Graph(GraphBuilder&& b, const Device& dev, uint64_t seed, IAllocator* slabs = nullptr);
  Graph(Graph&&) = default;
  const std::string& name() const { return name_; }
  void set_name(const std::string& n) { name_ = n; }
  GraphContext& context() { return *ctx_; }
  const Device& device() const { return dev_; }
  void set_training(bool f);
  void set_seed(uint64_t s);
  Vec<Layer*> get_layers();
  OpNode& op(const std::string& uid);
  const Vec<Edge>& edges() const { return edges_; }
  const Vec<IONode>& io_nodes() const { return io_nodes_; }
  GraphConfig get_config() const;                 // reference schema (graph.hpp:78-117)
  static Graph create_from_config(const GraphConfig& cfg, const Device& dev, uint64_t seed = 0);
  // checkpoint: pretty JSON config immediately followed by one tensor record per parameter (graph.hpp:119-183);
  // `trailer` (optional) appends named sections after the records (optimizer / scheduler state)
  void save_state(std::ostream& os, const std::map<std::string, Tensor>* tensors = nullptr, const Json* meta = nullptr) const;
  void save_state(const std::string& path, const std::map<std::string, Tensor>* tensors = nullptr, const Json* meta = nullptr) const;
  static Graph load_state(std::istream& is, const Device& dev, std::map<std::string, Tensor>* tensors = nullptr, Json* meta = nullptr);
  static Graph load_state(const std::string& path, const Device& dev, std::map<std::string, Tensor>* tensors = nullptr, Json* meta = nullptr);
  void load_state_tensors(std::istream& is);

 private:
  std::string name_;
  Device dev_;
  Vec<IONode> io_nodes_;
  Vec<std::unique_ptr<OpNode>> op_nodes_;
  Vec<Edge> edges_;
  std::unique_ptr<GraphContext> ctx_;
};

using InputPack = std::map<std::string, Tensor>;
using OutputPack = std::map<std::string, Tensor>;

// Runs the compiled edge list.  Owns the step WORKSPACE (a DELAllocatorV2) that every layer allocates its activations
// from and flips it between layers, as the reference executor does (graph_executor.hpp:17-28, sequential.cpp:33-44).
class GraphExecutor {
 public:
  explicit GraphExecutor(Graph& g, IAllocator* backing = nullptr);
  ~GraphExecutor();
  void forward(const InputPack& inputs, OutputPack& outputs, size_t mb_id = 0);
  void backward(const InputPack& grads, OutputPack& grad_inputs, size_t mb_id = 0);
  Tensor forward(const Tensor& x, size_t mb_id = 0);          // single-input / single-output convenience
  Tensor backward(const Tensor& dy, size_t mb_id = 0);
  DELAllocatorV2& workspace() { return *ws_; }
  Graph& graph() { return g_; }

 private:
  Graph& g_;
  std::unique_ptr<DELAllocatorV2> ws_;
};

// input -> [model] -> output, the shape every zoo model uses (example_models.hpp:49-84)
Graph graph_from_sequential(std::unique_ptr<Layer> model, const Device& dev, uint64_t seed = 0, const std::string& name = "", IAllocator* slabs = nullptr);

}  // namespace tnn
