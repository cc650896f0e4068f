// DataParallel: see tnn/data_parallel.hpp.
#include "tnn/data_parallel.hpp"

#include <cmath>
#include <thread>

#include "kernels/simt_ops.cuh"
#include "tnn/stepstate.hpp"

namespace tnn {

DataParallel::DataParallel(const ModelFactory& make_model, int world, uint64_t seed, const Optimizer& proto, const std::string& loss, bool use_graph, int symmetric) {
  auto& dm = DeviceManager::getInstance();
  if (world < 1 || world > dm.num_gpus()) throw std::invalid_argument("DataParallel: " + std::to_string(world) + " replicas requested, " + std::to_string(dm.num_gpus()) + " GPU(s) visible");
  adam_ = proto.name() == "adam";
  if (symmetric < 0) {
    const char* e = std::getenv("TNN_DP_NVLS");
    symmetric = e && e[0] == '1';
  }
  constexpr size_t kFlagWords = 64;
  if (symmetric && world > 1) {
    // one VMM slab per replica, sized for its four parameter slabs + the flag words; explicitly requested, so a box that
    // cannot provide it is an error, not a silent fall-back to peer loads
    std::vector<int> devs;
    for (int r = 0; r < world; ++r) devs.push_back(dm.getGPU(r).id());
    const std::unique_ptr<Layer> probe = make_model();
    const size_t need = GraphContext::slab_bytes(probe->all_layers(), true) + 4096;
    symm_ = std::make_shared<SymmetricMemory>(devs, need, /*want_multicast=*/true);
    for (int r = 0; r < world; ++r) symm_alloc_.push_back(std::make_unique<SymmetricAllocator>(symm_, r));
  }
  for (int r = 0; r < world; ++r) {
    auto rk = std::make_unique<Rank>();
    rk->dev = dm.getGPU(r);
    rk->dev.set_current();
    rk->graph = std::make_unique<Graph>(graph_from_sequential(make_model(), rk->dev, seed, "dp_replica" + std::to_string(r), symm_ ? symm_alloc_[(size_t)r].get() : nullptr));
    rk->opt = proto.clone();
    rk->loss = LossFactory::create(loss);
    ranks_.push_back(std::move(rk));
  }
  n_ = ranks_[0]->graph->context().numel();
  chunk_ = ((n_ + (size_t)world - 1) / (size_t)world + 1023) / 1024 * 1024;
  for (int r = 0; r < world; ++r) {                       // every pair of replicas may address each other's slabs
    TNN_CUDA(cudaSetDevice(ranks_[(size_t)r]->dev.id()));
    for (int q = 0; q < world; ++q) {
      if (q == r) continue;
      const cudaError_t e = cudaDeviceEnablePeerAccess(ranks_[(size_t)q]->dev.id(), 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) throw std::runtime_error(std::string("DataParallel: no peer access ") + cudaGetErrorString(e));
      cudaGetLastError();
    }
  }
  for (auto& rk : ranks_) {
    Rank& k = *rk;
    k.dev.set_current();
    GraphContext& ctx = k.graph->context();
    if (ctx.numel() != n_) throw std::logic_error("DataParallel: replicas differ in size");
    k.m = make_tensor(DType_t::FP32, {chunk_}, k.dev);
    k.m->zero();
    if (adam_) {
      k.v = make_tensor(DType_t::FP32, {chunk_}, k.dev);
      k.v->zero();
    }
    k.flags = symm_ ? make_tensor(*symm_alloc_[g_peers_.size()], DType_t::INT32_T, {kFlagWords}) : make_tensor(DType_t::INT32_T, {kFlagWords}, k.dev);
    k.flags->zero();
    k.sync = make_tensor(DType_t::INT32_T, {4}, k.dev);
    k.sync->zero();
    g_peers_.push_back(ctx.grad_slab()->data_as<float>());
    p_peers_.push_back(ctx.param_slab()->data_as<float>());
    pb_peers_.push_back(ctx.lp_slab()->data());
    flag_peers_.push_back(reinterpret_cast<uint32_t*>(k.flags->data()));
    StepState::get(k.dev);                                  // created before the replica threads start
    current_flow(k.dev).synchronize();
  }
  if (symm_) {
    // "symmetric" = one offset addresses the same slab in every replica (and in the multicast window)
    for (int r = 1; r < world; ++r)
      if (symm_->offset_of(r, g_peers_[(size_t)r]) != symm_->offset_of(0, g_peers_[0]) || symm_->offset_of(r, p_peers_[(size_t)r]) != symm_->offset_of(0, p_peers_[0]) ||
          symm_->offset_of(r, pb_peers_[(size_t)r]) != symm_->offset_of(0, pb_peers_[0]))
        throw std::logic_error("DataParallel: replicas allocated their slabs at different offsets of the symmetric group");
    if (symm_->has_multicast()) {
      g_mc_ = static_cast<const float*>(symm_->multicast(0, g_peers_[0]));
      p_mc_ = static_cast<float*>(symm_->multicast(0, p_peers_[0]));
      pb_mc_ = symm_->multicast(0, pb_peers_[0]);
    }
  }
  for (int r = 0; r < world; ++r) {
    Rank& k = *ranks_[(size_t)r];
    k.step = std::make_unique<TrainStep>(*k.graph, *k.loss, *k.opt, nullptr, 1, use_graph);
    k.step->update_override = [this, r] { launch_fused_step(r); };
  }
}
DataParallel::~DataParallel() {
  for (auto& rk : ranks_) {
    rk->dev.set_current();
    cudaDeviceSynchronize();
  }
}

void DataParallel::launch_fused_step(int r) {
  Rank& k = *ranks_[(size_t)r];
  k.dev.set_current();
  dp_fused_step(world(), r, (long long)n_, (long long)chunk_, g_peers_.data(), p_peers_.data(), pb_peers_.data(), flag_peers_.data(), g_mc_, p_mc_, pb_mc_,
                k.m->data_as<float>(), adam_ ? k.v->data_as<float>() : nullptr, reinterpret_cast<uint32_t*>(k.sync->data()), StepState::get(k.dev).dev(), adam_ ? 1 : 0,
                /*ctas=*/64, 0, (long long)n_, 0, current_flow(k.dev).stream());
  k.graph->context().mark_lp_fresh();                       // the kernel rewrote the bf16 copies of every replica
}

std::pair<double, double> DataParallel::step(const Vec<Tensor>& xs, const Vec<Tensor>& ys) {
  const size_t W = ranks_.size();
  if (xs.size() != W || ys.size() != W) throw std::invalid_argument("DataParallel::step: one batch per replica expected");
  std::vector<float> loss(W, 0.f), correct(W, 0.f);
  std::vector<std::string> errors(W);
  std::vector<std::thread> th;
  // one host thread per replica: the fused kernels of all replicas must be in flight together (they barrier on each
  // other), and a replica's eager warm-up steps may synchronise its own device
  for (size_t r = 0; r < W; ++r)
    th.emplace_back([&, r] {
      try {
        Rank& k = *ranks_[r];
        k.dev.set_current();
        Tensor st = (*k.step)(xs[r], ys[r]);
        float h[2];
        st->copy_to_host(h, 8);
        loss[r] = h[0];
        correct[r] = h[1];
      } catch (const std::exception& e) {
        errors[r] = e.what();
      }
    });
  for (auto& t : th) t.join();
  for (size_t r = 0; r < W; ++r)
    if (!errors[r].empty()) throw std::runtime_error("DataParallel replica " + std::to_string(r) + ": " + errors[r]);
  double l = 0, c = 0;
  for (size_t r = 0; r < W; ++r) {
    l += loss[r];
    c += correct[r];
  }
  return {l / (double)W, c};
}

double DataParallel::max_replica_divergence() {
  ranks_[0]->dev.set_current();
  const std::vector<float> ref = to_float_vector(*ranks_[0]->graph->context().param_slab());
  double worst = 0;
  for (size_t r = 1; r < ranks_.size(); ++r) {
    ranks_[r]->dev.set_current();
    const std::vector<float> p = to_float_vector(*ranks_[r]->graph->context().param_slab());
    for (size_t i = 0; i < n_; ++i) worst = std::max(worst, (double)std::fabs(p[i] - ref[i]));
  }
  return worst;
}
std::vector<float> dp_gather_shards(const std::vector<std::vector<float>>& shards, size_t n, size_t chunk, size_t total) {
  if (total < n) throw std::invalid_argument("dp_gather_shards: total < n");
  std::vector<float> full(total, 0.f);
  for (size_t r = 0; r < shards.size(); ++r) {
    const size_t lo = r * chunk, hi = std::min(n, lo + chunk);
    if (lo >= hi) continue;
    if (shards[r].size() < hi - lo) throw std::invalid_argument("dp_gather_shards: shard " + std::to_string(r) + " is shorter than the range it owns");
    std::copy(shards[r].begin(), shards[r].begin() + (std::ptrdiff_t)(hi - lo), full.begin() + (std::ptrdiff_t)lo);
  }
  return full;
}
std::vector<std::vector<float>> dp_scatter_shards(const std::vector<float>& full, size_t n, size_t chunk, size_t world) {
  if (full.size() < n || chunk * world < n) throw std::invalid_argument("dp_scatter_shards: the shards do not cover n");
  std::vector<std::vector<float>> shards(world, std::vector<float>(chunk, 0.f));
  for (size_t r = 0; r < world; ++r) {
    const size_t lo = r * chunk, hi = std::min(n, lo + chunk);
    if (lo < hi) std::copy(full.begin() + (std::ptrdiff_t)lo, full.begin() + (std::ptrdiff_t)hi, shards[r].begin());
  }
  return shards;
}
uint64_t DataParallel::step_count() const { return ranks_[0]->opt->step_count; }
void DataParallel::save_checkpoint(const std::string& path) {
  const size_t W = ranks_.size();
  std::vector<std::vector<float>> ms(W), vs(W);
  for (size_t r = 0; r < W; ++r) {
    Rank& k = *ranks_[r];
    k.dev.set_current();
    current_flow(k.dev).synchronize();
    ms[r] = to_float_vector(*k.m);
    if (adam_) vs[r] = to_float_vector(*k.v);
  }
  GraphContext& ctx = ranks_[0]->graph->context();
  const size_t total = ctx.param_slab()->size();
  std::map<std::string, Tensor> tensors;
  // the names a single-device optimizer uses for the same state (Adam: m, v; SGD: velocity)
  tensors[adam_ ? "opt.m" : "opt.velocity"] = make_tensor_from(dp_gather_shards(ms, n_, chunk_, total), {total});
  if (adam_) tensors["opt.v"] = make_tensor_from(dp_gather_shards(vs, n_, chunk_, total), {total});
  Optimizer& opt = *ranks_[0]->opt;
  Json meta = Json::object();
  meta["step"] = (long long)opt.step_count;
  meta["opt.config"] = opt.get_config().to_json();
  meta["opt.step"] = (long long)opt.step_count;
  meta["opt.lr"] = opt.get_learning_rate();
  meta["dp.world"] = (long long)W;
  ranks_[0]->dev.set_current();
  ranks_[0]->graph->save_state(path, &tensors, &meta);
}
void DataParallel::load_checkpoint(const std::string& path) {
  std::map<std::string, Tensor> tensors;
  Json meta;
  Graph host = Graph::load_state(path, getHost(), &tensors, &meta);
  GraphContext& src = host.context();
  if (src.numel() != n_) throw std::runtime_error("DataParallel::load_checkpoint: " + path + " holds " + std::to_string(src.numel()) + " parameters, the replicas " + std::to_string(n_));
  auto full = [&](const char* a, const char* b) -> std::vector<float> {
    auto it = tensors.find(a);
    if (it == tensors.end() && b) it = tensors.find(b);
    if (it == tensors.end()) return std::vector<float>(n_, 0.f);        // a checkpoint without optimizer state: fresh moments
    std::vector<float> v = to_float_vector(*it->second);
    if (v.size() < n_) throw std::runtime_error(std::string("DataParallel::load_checkpoint: optimizer tensor ") + a + " is too short");
    return v;
  };
  const auto ms = dp_scatter_shards(adam_ ? full("opt.m", nullptr) : full("opt.velocity", "opt.m"), n_, chunk_, ranks_.size());
  const auto vs = adam_ ? dp_scatter_shards(full("opt.v", nullptr), n_, chunk_, ranks_.size()) : std::vector<std::vector<float>>();
  for (size_t r = 0; r < ranks_.size(); ++r) {
    Rank& k = *ranks_[r];
    k.dev.set_current();
    GraphContext& ctx = k.graph->context();
    ctx.param_slab()->copy_from(*src.param_slab());
    ctx.buffer_slab()->copy_from(*src.buffer_slab());
    ctx.grad_slab()->zero();
    ctx.refresh_lp();
    k.m->copy_from_host(ms[r].data(), ms[r].size() * 4);
    if (adam_) k.v->copy_from_host(vs[r].data(), vs[r].size() * 4);
    if (meta.is_object()) {
      k.opt->step_count = (uint64_t)meta.value("opt.step", (int64_t)0);
      k.opt->set_learning_rate(meta.value("opt.lr", k.opt->get_learning_rate()));
    }
    current_flow(k.dev).synchronize();
  }
}
bool DataParallel::captured() const {
  for (auto& rk : ranks_)
    if (!rk->step->captured()) return false;
  return true;
}

}  // namespace tnn
