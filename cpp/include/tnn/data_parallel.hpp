// Data-parallel training in ONE process: a replica per GPU, a host thread per replica, and the fused data-parallel step
// kernel (csrc/kernels/optim.cu k_dp_step) instead of "all-reduce, then optimizer": cross-GPU flag barrier -> every rank
// reduces its 1/W slice of the gradients from all peers over NVLink peer loads -> Adam / SGD on the slice (optimizer state
// sharded W ways) -> broadcast fp32 + bf16 parameters into every replica -> barrier -> zero the gradients.
// (Reference: ParallelMode_t::DATA exists only as routing, SURVEY F4 -- there is no gradient reduction to mirror.)
// The Python host runs the same kernel with one process per GPU (parallel/dp.py, NVLS multicast through symmetric memory);
// inside one process the peers' device pointers are directly usable after cudaDeviceEnablePeerAccess.
// TNN_DP_NVLS=1 (or symmetric = true) puts every replica's parameter / gradient / bf16 slabs and flag words into the
// framework's own SymmetricMemory group (cuMemCreate + peer mappings + cuMulticast window, runtime/symmetric_memory.hpp):
// the kernel then pulls the gradient SUM from the switch (multimem.ld_reduce) and broadcasts the new parameters with
// multimem.st instead of W - 1 peer loads / stores per element.
#pragma once
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "tnn/train.hpp"

namespace tnn {

// Shard layout of the fused step: rank r owns elements [r * chunk, min(n, (r + 1) * chunk)) of a flat state vector.
// gather: W shard buffers (each >= its owned length) -> one vector of `total` (>= n) elements, zero beyond n;
// scatter: the inverse, every shard buffer `chunk` long and zero padded.  (Host-side helpers of save / load_checkpoint.)
std::vector<float> dp_gather_shards(const std::vector<std::vector<float>>& shards, size_t n, size_t chunk, size_t total);
std::vector<std::vector<float>> dp_scatter_shards(const std::vector<float>& full, size_t n, size_t chunk, size_t world);

class DataParallel {
 public:
  using ModelFactory = std::function<std::unique_ptr<Layer>()>;
  // `world` replicas on GPU 0..world-1; the same seed initialises every replica identically
  DataParallel(const ModelFactory& make_model, int world, uint64_t seed, const Optimizer& optimizer_proto, const std::string& loss = "logsoftmax_crossentropy",
               bool use_graph = true, int symmetric = -1 /* -1: TNN_DP_NVLS, 0: plain peer memory, 1: SymmetricMemory (+ NVLS) */);
  ~DataParallel();
  // xs[r] / ys[r]: replica r's batch (host, ideally pinned).  Returns {mean loss over replicas, total #correct}.
  std::pair<double, double> step(const Vec<Tensor>& xs, const Vec<Tensor>& ys);
  int world() const { return (int)ranks_.size(); }
  bool symmetric() const { return symm_ != nullptr; }               // slabs live in the own VMM group
  bool multicast() const { return g_mc_ != nullptr; }               // ... and the NVLS window is bound
  Graph& replica(int r) { return *ranks_.at((size_t)r)->graph; }
  size_t num_parameters() const { return n_; }
  // max over replicas r > 0 of max |param_r - param_0| (replicas must stay bit-identical)
  double max_replica_divergence();
  bool captured() const;
  // Checkpoint / resume.  The file is an ordinary model checkpoint (replica 0's graph + the "opt.*" trailer of
  // save_checkpoint, train.hpp): the optimizer moments, which live sharded W ways across the replicas, are gathered into
  // full-length tensors on save and scattered to whatever world size loads them, so a file written by W replicas resumes on
  // W' replicas or on one GPU (load_checkpoint) and the other way round.  Replicas must be idle (between step() calls).
  void save_checkpoint(const std::string& path);
  void load_checkpoint(const std::string& path);
  uint64_t step_count() const;

 private:
  struct Rank {
    Device dev{DeviceType::GPU, 0};
    std::unique_ptr<Graph> graph;
    std::unique_ptr<Optimizer> opt;
    std::unique_ptr<Loss> loss;
    std::unique_ptr<TrainStep> step;
    Tensor m, v, flags, sync;
  };
  void launch_fused_step(int r);
  // declared before ranks_: the replicas' slabs are freed through these allocators
  std::shared_ptr<SymmetricMemory> symm_;
  std::vector<std::unique_ptr<SymmetricAllocator>> symm_alloc_;
  const float* g_mc_ = nullptr;
  float* p_mc_ = nullptr;
  void* pb_mc_ = nullptr;
  std::vector<std::unique_ptr<Rank>> ranks_;
  size_t n_ = 0, chunk_ = 0;
  bool adam_ = true;
  std::vector<float*> g_peers_, p_peers_;
  std::vector<void*> pb_peers_;
  std::vector<uint32_t*> flag_peers_;
};

}  // namespace tnn
