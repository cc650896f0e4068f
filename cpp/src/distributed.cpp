// Distributed runtime: messages, communicators, partitioners, worker, coordinator, distributed trainer.
#include <chrono>
#include <filesystem>
#include <iostream>
#include <sstream>

#include "runtime/planner.hpp"
#include "tnn/distributed.hpp"
#include "tnn/example_models.hpp"
#include "tnn/stepstate.hpp"

namespace tnn {

const char* command_name(CommandType c) {
  static const char* n[] = {"_START", "FORWARD_JOB", "BACKWARD_JOB", "BACKWARD_COMPLETE", "UPDATE_PARAMETERS", "TRAIN_MODE", "EVAL_MODE", "SHUTDOWN",
                            "CONFIG_TRANSFER", "CONFIG_RECEIVED", "LOAD_PARAMS", "PARAMS_LOADED", "SEND_PARAMS", "PARAMS_TRANSFER", "SAVE_TO_FILE",
                            "STATUS_REQUEST", "STATUS_RESPONSE", "PARAMETERS_UPDATED", "HEALTH_CHECK", "ERROR_REPORT", "JOB_FAILURE", "BARRIER_SYNC",
                            "CHECKPOINT_REQUEST", "CHECKPOINT_COMPLETE", "UPDATE_LOAD", "REPORT_LOAD", "LOAD_REPORT", "START_PROFILING", "PROFILING_STARTED",
                            "PRINT_PROFILING", "PROFILING_PRINTED", "CLEAR_PROFILING", "PROFILING_CLEARED", "REPORT_PROFILING", "PROFILING_REPORTED",
                            "HANDSHAKE", "HANDSHAKE_ACK"};
  return (size_t)c < sizeof(n) / sizeof(n[0]) ? n[(size_t)c] : "?";
}

// ===================================================================================== endpoints
static std::mutex& reg_mu() { static std::mutex* m = new std::mutex(); return *m; }
static std::map<std::string, Communicator*>& registry() { static auto* r = new std::map<std::string, Communicator*>(); return *r; }
Endpoint Endpoint::in_process(Communicator* c) {
  Endpoint e;
  e.type = IN_PROCESS;
  std::ostringstream k;
  k << "inproc-" << (const void*)c;
  e.key = k.str();
  std::lock_guard<std::mutex> g(reg_mu());
  registry()[e.key] = c;
  return e;
}
Communicator* Endpoint::communicator() const {
  std::lock_guard<std::mutex> g(reg_mu());
  auto it = registry().find(key);
  return it == registry().end() ? nullptr : it->second;
}
Json Endpoint::to_json() const {
  Json j = Json::object();
  static const char* names[] = {"NONE", "IN_PROCESS", "TCP", "ROCE"};
  j["type"] = names[type];
  Json p = Json::object();
  if (type == IN_PROCESS) p["key"] = key;
  else if (type != NONE) {
    p["host"] = host;
    p["port"] = port;
  }
  j["parameters"] = p;
  return j;
}
Endpoint Endpoint::from_json(const Json& j) {
  Endpoint e;
  if (!j.is_object()) return e;
  const std::string t = j.value("type", "NONE");
  e.type = t == "IN_PROCESS" ? IN_PROCESS : t == "TCP" ? TCP : t == "ROCE" ? ROCE : NONE;
  const Json* p = j.find("parameters");
  if (p && p->is_object()) {
    e.host = p->value("host", "");
    e.port = p->value("port", 0);
    e.key = p->value("key", "");
  }
  return e;
}
Json StageConfig::to_json() const {
  Json j = Json::object();
  j["model_config"] = model_config;
  j["optimizer_config"] = optimizer_config;
  j["scheduler_config"] = scheduler_config;
  j["next_stage_endpoint"] = next_stage_endpoint.to_json();
  j["prev_stage_endpoint"] = prev_stage_endpoint.to_json();
  j["coordinator_endpoint"] = coordinator_endpoint.to_json();
  j["stage_id"] = stage_id;
  j["seed"] = (long long)seed;
  j["device"] = device;
  return j;
}
StageConfig StageConfig::from_json(const Json& j) {
  StageConfig s;
  if (const Json* p = j.find("model_config")) s.model_config = *p;
  if (const Json* p = j.find("optimizer_config")) s.optimizer_config = *p;
  if (const Json* p = j.find("scheduler_config")) s.scheduler_config = *p;
  if (const Json* p = j.find("next_stage_endpoint")) s.next_stage_endpoint = Endpoint::from_json(*p);
  if (const Json* p = j.find("prev_stage_endpoint")) s.prev_stage_endpoint = Endpoint::from_json(*p);
  if (const Json* p = j.find("coordinator_endpoint")) s.coordinator_endpoint = Endpoint::from_json(*p);
  s.stage_id = j.value("stage_id", 0);
  s.seed = (uint64_t)j.value("seed", (int64_t)123456);
  s.device = j.value("device", "");
  return s;
}

// ===================================================================================== message
std::vector<uint8_t> Message::serialize() const {
  Writer w;
  w.prim<uint16_t>((uint16_t)command);
  w.prim<uint32_t>((uint32_t)kind);
  switch (kind) {
    case JOB: {
      w.prim<uint64_t>(job.mb_id);
      Tensor h = job.data->is_gpu() ? job.data->to_host() : job.data;      // host staging only on this (TCP) path
      std::vector<uint64_t> shape(h->shape().begin(), h->shape().end());
      w.tensor((uint32_t)h->data_type(), shape, h->data(), h->nbytes());
      break;
    }
    case STRING: w.str(str); break;
    case BOOL: w.prim<uint8_t>(flag ? 1 : 0); break;
    case PROFILER: w.str(profiler ? profiler->serialize() : Profiler().serialize()); break;
    case STAGE: w.str(stage.to_json().dump()); break;
    default: break;
  }
  return std::move(w.buffer());
}
Message Message::deserialize(const uint8_t* data, size_t n, bool swap, const Device* land_on) {
  Reader r(data, n, swap);
  Message m;
  m.command = (CommandType)r.prim<uint16_t>();
  m.kind = (Kind)r.prim<uint32_t>();
  switch (m.kind) {
    case JOB: {
      m.job.mb_id = r.prim<uint64_t>();
      const uint32_t dt = r.prim<uint32_t>();
      const uint64_t nd = r.prim<uint64_t>();
      if (dt >= (uint32_t)DType_t::UNKNOWN || nd > 16) throw std::runtime_error("Message: bad tensor header");
      Shape shape;
      for (uint64_t i = 0; i < nd; ++i) shape.push_back((size_t)r.prim<uint64_t>());
      Tensor t = make_tensor((DType_t)dt, shape, getHost());
      const uint8_t* raw = r.bytes(t->nbytes());
      std::memcpy(t->data(), raw, t->nbytes());
      if (swap && get_dtype_size((DType_t)dt) > 1) {
        const size_t es = get_dtype_size((DType_t)dt);
        uint8_t* p = static_cast<uint8_t*>(t->data());
        for (size_t i = 0; i < t->size(); ++i) std::reverse(p + i * es, p + (i + 1) * es);
      }
      m.job.data = (land_on && land_on->is_gpu()) ? t->to_device(*land_on) : t;
      break;
    }
    case STRING: m.str = r.str(); break;
    case BOOL: m.flag = r.prim<uint8_t>() != 0; break;
    case PROFILER: m.profiler = std::make_shared<Profiler>(Profiler::deserialize(r.str())); break;
    case STAGE: m.stage = StageConfig::from_json(Json::parse(r.str())); break;
    default: break;
  }
  return m;
}

// ===================================================================================== communicators
Communicator::~Communicator() {
  std::lock_guard<std::mutex> g(reg_mu());
  for (auto it = registry().begin(); it != registry().end();)
    it = it->second == this ? registry().erase(it) : std::next(it);
}
void Communicator::enqueue(Message m) {
  {
    std::lock_guard<std::mutex> g(mu_);
    const size_t q = std::min((size_t)m.command, (size_t)CommandType::_COUNT - 1);
    queues_[q].push_back(std::move(m));
    ++received_messages;
  }
  cv_.notify_all();
}
size_t Communicator::pending() const {
  std::lock_guard<std::mutex> g(mu_);
  size_t n = 0;
  for (auto& q : queues_) n += q.size();
  return n;
}
bool Communicator::dequeue(Message* out, int timeout_ms, const std::vector<CommandType>& types) {
  std::unique_lock<std::mutex> lk(mu_);
  auto pick = [&]() -> bool {
    if (types.empty()) {
      for (auto& q : queues_)
        if (!q.empty()) {
          *out = std::move(q.front());
          q.pop_front();
          return true;
        }
      return false;
    }
    std::vector<CommandType> ts = types;
    std::sort(ts.begin(), ts.end());
    for (CommandType t : ts) {
      auto& q = queues_[(size_t)t];
      if (!q.empty()) {
        *out = std::move(q.front());
        q.pop_front();
        return true;
      }
    }
    return false;
  };
  if (pick()) return true;
  if (timeout_ms == 0) return false;
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms < 0 ? 1 << 30 : timeout_ms);
  while (true) {
    if (cv_.wait_until(lk, deadline) == std::cv_status::timeout) return pick();
    if (pick()) return true;
  }
}
void Communicator::send_message(const Message& m, const Endpoint& to) {
  ++sent_messages;
  if (to.type == Endpoint::IN_PROCESS) {
    Communicator* peer = to.communicator();
    if (!peer) throw DistributedError(name_ + ": in-process peer " + to.key + " is gone");
    Message copy = m;               // tensors travel BY REFERENCE (device tensors stay on their device)
    copy.sender = name_;
    peer->enqueue(std::move(copy));
    return;
  }
  if (to.type == Endpoint::NONE) throw DistributedError(name_ + ": send to an empty endpoint");
  send_remote(m, to);
}

TCPCommunicator::TCPCommunicator(const std::string& host, int port, int spp, const Device* land_on, const std::string& name)
    : Communicator(name), host_(host), bus_(std::make_unique<TcpBus>(port, spp, (size_t)4 << 20, host)) {
  if (land_on) {
    land_ = *land_on;
    has_land_ = true;
  }
}
TCPCommunicator::~TCPCommunicator() { stop(); }
void TCPCommunicator::start() {
  if (running_.exchange(true)) return;
  bus_->start();
  pump_ = std::thread([this] {
    while (running_) {
      std::string from;
      std::vector<uint8_t> buf;
      if (!bus_->recv(200, &from, &buf)) continue;
      const int64_t t0 = Profiler::now_ns();
      try {
        Message m = Message::deserialize(buf.data(), buf.size(), false, has_land_ ? &land_ : nullptr);
        m.sender = from;
        GlobalProfiler::add_event(EventType::COMMUNICATION, t0, Profiler::now_ns(), std::string("recv ") + command_name(m.command), name_);
        enqueue(std::move(m));
      } catch (const std::exception& e) {
        std::cerr << "[" << name_ << "] dropped a malformed message from " << from << ": " << e.what() << std::endl;
      }
    }
  });
}
void TCPCommunicator::stop() {
  if (!running_.exchange(false)) return;
  if (pump_.joinable()) pump_.join();
  bus_->stop();
}
void TCPCommunicator::connect(const Endpoint& e) {
  if (e.type == Endpoint::TCP || e.type == Endpoint::ROCE) bus_->connect(e.host, e.port);
}
void TCPCommunicator::send_remote(const Message& m, const Endpoint& to) {
  const std::string key = to.host + ":" + std::to_string(to.port);
  if (!bus_->has_peer(key)) bus_->connect(to.host, to.port);
  const int64_t t0 = Profiler::now_ns();
  const std::vector<uint8_t> buf = m.serialize();
  bus_->send(key, buf.data(), buf.size());
  GlobalProfiler::add_event(EventType::COMMUNICATION, t0, Profiler::now_ns(), std::string("send ") + command_name(m.command), name_);
}

// ===================================================================================== partitioners
Vec<Tensor> Partitioner::partition_input(const Tensor& x, size_t parts) {
  Vec<Tensor> out;
  const size_t B = x->dimension(0);
  size_t cur = 0;
  for (size_t i = 0; i < parts; ++i) {
    const size_t n = i + 1 == parts ? B - cur : B / parts;
    if (n) out.push_back(x->slice_batch(cur, n));
    cur += n;
  }
  return out;
}
Vec<std::unique_ptr<Sequential>> Partitioner::split(const Vec<Layer*>& layers, const Vec<SeqPartition>& parts, const std::string& base) {
  Vec<std::unique_ptr<Sequential>> out;
  for (size_t s = 0; s < parts.size(); ++s) {
    Vec<std::unique_ptr<Layer>> ls;
    for (size_t i = parts[s].start; i < parts[s].start + parts[s].length; ++i) ls.push_back(LayerFactory::create(layers[i]->get_config()));
    out.push_back(std::make_unique<Sequential>(std::move(ls), base + "_stage" + std::to_string(s)));
  }
  return out;
}
static Vec<SeqPartition> to_parts(const std::vector<std::pair<int64_t, int64_t>>& v) {
  Vec<SeqPartition> p;
  for (auto& e : v) p.push_back(SeqPartition{(size_t)e.first, (size_t)e.second});
  return p;
}
Vec<SeqPartition> NaivePipelinePartitioner::partition_model(const Vec<Layer*>& layers, const Shape&) {
  std::vector<int64_t> props(props_.begin(), props_.end());
  return to_parts(planner::split_by_proportions((int64_t)layers.size(), props, true));
}
Vec<SeqPartition> NaiveDataPartitioner::partition_model(const Vec<Layer*>& layers, const Shape&) {
  return Vec<SeqPartition>(props_.size(), SeqPartition{0, layers.size()});
}
static void layer_costs(const Vec<Layer*>& layers, const Shape& in, std::vector<double>* flops, std::vector<double>* out_bytes) {
  Shape s = in;
  for (Layer* l : layers) {
    flops->push_back(3.0 * (double)l->forward_flops(s) + (double)shape_numel(s));     // fwd + bwd, plus a memory-bound floor
    s = l->output_shape(s);
    out_bytes->push_back(2.0 * (double)shape_numel(s) * 2.0);                          // bf16 activation forward + gradient back
  }
}
Vec<SeqPartition> WeightedPartitioner::partition_model(const Vec<Layer*>& layers, const Shape& in) {
  std::vector<double> fl, by;
  layer_costs(layers, in, &fl, &by);
  return to_parts(planner::min_bottleneck_split(fl, (int64_t)n_, {}));
}
Vec<SeqPartition> FTDPartitioner::partition_model(const Vec<Layer*>& layers, const Shape& in) {
  std::vector<double> fl, by;
  layer_costs(layers, in, &fl, &by);
  std::vector<double> cost, boundary;
  for (size_t i = 0; i < fl.size(); ++i) {
    cost.push_back(fl[i] / fl_);
    boundary.push_back(by[i] / bw_ + delay_);
  }
  return to_parts(planner::min_bottleneck_split(cost, (int64_t)n_, boundary));
}

// ===================================================================================== worker
static double process_cpu_seconds() {
  timespec ts{};
  return clock_gettime(CLOCK_PROCESS_CPUTIME_ID, &ts) == 0 ? (double)ts.tv_sec + ts.tv_nsec * 1e-9 : -1.0;
}
Worker::Worker(std::unique_ptr<Communicator> comm, const Device& dev, const std::string& name) : comm_(std::move(comm)), dev_(dev), name_(name) {
  namespace fs = std::filesystem;
  std::error_code ec;
  auto add_root = [&](const fs::path& p) {
    if (p.empty()) return;
    fs::path r = fs::weakly_canonical(p, ec);
    if (!ec && !r.empty()) io_roots_.push_back(r.string());
  };
  cpu_t0_ = std::max(0.0, process_cpu_seconds());
  add_root(fs::current_path(ec));
  add_root(fs::temp_directory_path(ec));
  if (const char* extra = std::getenv("TNN_WORKER_IO_ROOT")) {
    std::stringstream ss(extra);
    std::string item;
    while (std::getline(ss, item, ':')) add_root(item);
  }
}
LoadTracker Worker::load_tracker() const {
  LoadTracker t;
  t.avg_forward_time_ = (float)(fwd_ns_ / 1e6 / (double)std::max<uint64_t>(fwd_jobs_, 1));
  t.avg_backward_time_ = (float)(bwd_ns_ / 1e6 / (double)std::max<uint64_t>(bwd_jobs_, 1));
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall_t0_).count(), cpu = process_cpu_seconds();
  t.avg_cpu_utilization_ = (cpu >= 0 && wall > 0) ? (float)((cpu - cpu_t0_) / wall) : -1.f;
  {
    std::ifstream f("/proc/self/status");               // VmHWM: peak resident set
    std::string line;
    while (std::getline(f, line))
      if (line.rfind("VmHWM:", 0) == 0) t.max_memory_usage_ = (float)(atol(line.c_str() + 6) / 1024.0);
  }
  if (dev_.is_gpu()) {
    size_t free_b = 0, total_b = 0;
    dev_.set_current();
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) t.max_device_memory_usage_ = (float)((total_b - free_b) / 1048576.0);
    else cudaGetLastError();
  }
  return t;
}
std::string Worker::checked_path(const std::string& path) const {
  namespace fs = std::filesystem;
  std::error_code ec;
  // weakly_canonical resolves symlinks of the existing prefix and normalises "." / ".." of the rest
  const std::string real = fs::weakly_canonical(fs::absolute(path, ec), ec).string();
  if (!ec)
    for (const std::string& root : io_roots_) {
      std::string r = root;
      while (r.size() > 1 && r.back() == '/') r.pop_back();
      if (real == r || (real.size() > r.size() && real.compare(0, r.size(), r) == 0 && (r == "/" || real[r.size()] == '/'))) return real;
    }
  std::string roots;
  for (const std::string& r : io_roots_) roots += (roots.empty() ? "" : ", ") + r;
  throw std::runtime_error("worker " + name_ + ": path '" + path + "' is outside the allowed I/O roots [" + roots + "] (extend with TNN_WORKER_IO_ROOT)");
}
Worker::~Worker() { stop(); }
void Worker::start(bool background) {
  comm_->start();
  stop_ = false;
  if (background) thread_ = std::thread([this] { run(); });
  else run();
}
void Worker::stop() {
  stop_ = true;
  if (thread_.joinable()) {
    comm_->enqueue(Message(CommandType::SHUTDOWN));
    thread_.join();
  }
  comm_->stop();
}
void Worker::reply(const Message& m, const Endpoint* to) { comm_->send_message(m, to ? *to : cfg_.coordinator_endpoint); }
void Worker::set_config(const StageConfig& cfg) {
  cfg_ = cfg;
  if (dev_.is_gpu()) dev_.set_current();
  std::unique_ptr<Layer> model = LayerFactory::create(TConfig::from_json(cfg.model_config));
  exec_.reset();
  graph_ = std::make_unique<Graph>(graph_from_sequential(std::move(model), dev_, cfg.seed, "stage" + std::to_string(cfg.stage_id)));
  exec_ = std::make_unique<GraphExecutor>(*graph_);
  opt_.reset();
  sched_.reset();
  if (cfg.optimizer_config.is_object() && !cfg.optimizer_config.empty()) {
    opt_ = OptimizerFactory::create_from_config(TConfig::from_json(cfg.optimizer_config));
    opt_->attach(graph_->context());
    if (cfg.scheduler_config.is_object() && !cfg.scheduler_config.empty()) sched_ = SchedulerFactory::create_from_config(TConfig::from_json(cfg.scheduler_config), opt_.get());
  }
  for (const Endpoint* e : {&cfg.coordinator_endpoint, &cfg.next_stage_endpoint, &cfg.prev_stage_endpoint})
    if (*e) comm_->connect(*e);
}
void Worker::run() {
  while (!stop_) {
    Message m;
    if (!comm_->dequeue(&m, 200)) continue;
    try {
      if (fault_inject) fault_inject(m);
      process(m);
    } catch (const std::exception& e) {
      // workers report exceptions as ERROR_REPORT strings (worker.hpp:317-328); the coordinator aborts the batch on them
      try {
        if (cfg_.coordinator_endpoint) reply(Message(CommandType::ERROR_REPORT, name_ + ": " + command_name(m.command) + ": " + e.what()));
      } catch (...) {
      }
    }
  }
}
void Worker::process(Message& m) {
  if (dev_.is_gpu()) dev_.set_current();
  switch (m.command) {
    case CommandType::FORWARD_JOB: {
      const int64_t t0 = Profiler::now_ns();
      Tensor x = m.job.data->device() == dev_ ? m.job.data : m.job.data->to_device(dev_);     // GPU -> GPU peer copy for in-process neighbours
      Tensor y = exec_->forward(x, m.job.mb_id);
      if (dev_.is_gpu()) current_flow(dev_).synchronize();       // the consumer runs on another flow / device
      const int64_t t1 = Profiler::now_ns();
      GlobalProfiler::add_event(EventType::COMPUTE, t0, t1, "forward mb " + std::to_string(m.job.mb_id), name_);
      busy_ns_ += t1 - t0;
      fwd_ns_ += t1 - t0;
      ++fwd_jobs_;
      ++jobs_;
      reply(Message(CommandType::FORWARD_JOB, Job{y, m.job.mb_id}), cfg_.next_stage_endpoint ? &cfg_.next_stage_endpoint : nullptr);
      break;
    }
    case CommandType::BACKWARD_JOB: {
      const int64_t t0 = Profiler::now_ns();
      Tensor g = m.job.data->device() == dev_ ? m.job.data : m.job.data->to_device(dev_);
      Tensor dx = exec_->backward(g, m.job.mb_id);
      if (dev_.is_gpu()) current_flow(dev_).synchronize();
      const int64_t t1 = Profiler::now_ns();
      GlobalProfiler::add_event(EventType::COMPUTE, t0, t1, "backward mb " + std::to_string(m.job.mb_id), name_);
      busy_ns_ += t1 - t0;
      bwd_ns_ += t1 - t0;
      ++bwd_jobs_;
      if (cfg_.prev_stage_endpoint && dx) reply(Message(CommandType::BACKWARD_JOB, Job{dx, m.job.mb_id}), &cfg_.prev_stage_endpoint);
      else reply(Message(CommandType::BACKWARD_COMPLETE));
      break;
    }
    case CommandType::UPDATE_PARAMETERS: {
      const int64_t t0 = Profiler::now_ns();
      if (opt_) {
        opt_->update();
        if (!opt_->fused_zero_grad) opt_->zero_grads();
        if (sched_) sched_->step();
      } else if (graph_ && dev_.is_gpu()) {
        StepState::get(dev_).advance(nullptr);
      }
      if (dev_.is_gpu()) current_flow(dev_).synchronize();
      GlobalProfiler::add_event(EventType::COMPUTE, t0, Profiler::now_ns(), "update", name_);
      reply(Message(CommandType::PARAMETERS_UPDATED, true));
      break;
    }
    case CommandType::TRAIN_MODE: if (graph_) graph_->set_training(true); break;
    case CommandType::EVAL_MODE: if (graph_) graph_->set_training(false); break;
    case CommandType::CONFIG_TRANSFER:
      set_config(m.stage);
      reply(Message(CommandType::CONFIG_RECEIVED, true));
      break;
    case CommandType::SHUTDOWN: stop_ = true; break;
    case CommandType::START_PROFILING:
      GlobalProfiler::get_profiler().clear();
      GlobalProfiler::init_start_time();
      reply(Message(CommandType::PROFILING_STARTED, true));
      break;
    case CommandType::REPORT_PROFILING: reply(Message(CommandType::PROFILING_REPORTED, GlobalProfiler::get_profiler())); break;
    case CommandType::CLEAR_PROFILING:
      GlobalProfiler::get_profiler().clear();
      reply(Message(CommandType::PROFILING_CLEARED, true));
      break;
    case CommandType::PRINT_PROFILING:
      for (const auto& e : GlobalProfiler::get_profiler().events()) std::cout << "[" << name_ << "] " << e.name << " " << (e.end_ns - e.start_ns) / 1e6 << " ms\n";
      reply(Message(CommandType::PROFILING_PRINTED, true));
      break;
    case CommandType::SAVE_TO_FILE: graph_->save_state(checked_path(m.str + ".stage" + std::to_string(cfg_.stage_id))); break;
    case CommandType::LOAD_PARAMS: {
      const std::string path = checked_path(m.str);
      std::ifstream f(path, std::ios::binary);
      if (!f) throw std::runtime_error("cannot open " + path);
      std::stringstream ss;
      ss << f.rdbuf();
      const std::string data = ss.str();
      size_t pos = 0;
      Json::parse(data, &pos);
      std::istringstream rest(data.substr(pos), std::ios::binary);
      graph_->load_state_tensors(rest);
      reply(Message(CommandType::PARAMS_LOADED, true));
      break;
    }
    case CommandType::SEND_PARAMS: {
      Tensor slab = (m.kind == Message::STRING && m.str == "grads") ? graph_->context().grad_slab() : graph_->context().param_slab();
      reply(Message(CommandType::PARAMS_TRANSFER, Job{slab->clone(), (uint64_t)cfg_.stage_id}));
      break;
    }
    case CommandType::PARAMS_TRANSFER: {
      constexpr uint64_t GRAD_SLAB_TAG = (1ull << 62) | 0x6772;
      Tensor dst = m.job.mb_id == GRAD_SLAB_TAG ? graph_->context().grad_slab() : graph_->context().param_slab();
      dst->copy_from(*m.job.data);
      if (m.job.mb_id != GRAD_SLAB_TAG) graph_->context().refresh_lp();
      reply(Message(CommandType::PARAMS_LOADED, true));
      break;
    }
    case CommandType::HEALTH_CHECK:
    case CommandType::STATUS_REQUEST: {
      Json s = Json::object();
      s["name"] = name_;
      s["stage"] = cfg_.stage_id;
      s["configured"] = graph_ != nullptr;
      s["jobs"] = (long long)jobs_;
      s["device"] = dev_.name();
      reply(Message(CommandType::STATUS_RESPONSE, s.dump()));
      break;
    }
    case CommandType::CHECKPOINT_REQUEST: {
      const std::string p = checked_path(m.str + ".stage" + std::to_string(cfg_.stage_id));
      save_checkpoint(*graph_, p, opt_.get(), sched_.get(), opt_ ? opt_->step_count : 0);
      reply(Message(CommandType::CHECKPOINT_COMPLETE, p));
      break;
    }
    case CommandType::BARRIER_SYNC:
      if (dev_.is_gpu()) current_flow(dev_).synchronize();
      reply(Message(CommandType::BARRIER_SYNC, true));
      break;
    case CommandType::REPORT_LOAD: {
      Json s = Json::object();
      s["stage"] = cfg_.stage_id;
      s["avg_forward_backward_ms"] = busy_ns_ / 1e6 / (double)std::max<uint64_t>(jobs_, 1);
      s["jobs"] = (long long)jobs_;
      const Json lt = load_tracker().to_json();
      for (const auto& kv : lt.members()) s[kv.first] = kv.second;
      reply(Message(CommandType::LOAD_REPORT, s.dump()));
      break;
    }
    case CommandType::HANDSHAKE: reply(Message(CommandType::HANDSHAKE_ACK, name_)); break;
    default: throw std::runtime_error(std::string("worker does not handle ") + command_name(m.command));
  }
}

// ===================================================================================== coordinator
Coordinator::Coordinator(CoordinatorConfig cfg, std::unique_ptr<Communicator> comm) : cfg_(std::move(cfg)), comm_(std::move(comm)) {}
Coordinator::~Coordinator() { comm_->stop(); }
void Coordinator::initialize() {
  comm_->start();
  Vec<Layer*> layers = cfg_.model->layers();
  const size_t n = cfg_.worker_endpoints.size();
  parts_ = cfg_.partitioner->partition_model(layers, cfg_.input_shape);
  if (parts_.size() != n) throw DistributedError("partitioner produced " + std::to_string(parts_.size()) + " stages for " + std::to_string(n) + " workers");
  auto stage_models = Partitioner::split(layers, parts_, cfg_.model->name());
  const Endpoint me = comm_->endpoint();
  const bool pipeline = cfg_.parallel_mode == ParallelMode_t::PIPELINE;
  stages_.clear();
  for (size_t s = 0; s < n; ++s) {
    StageConfig sc;
    sc.model_config = stage_models[s]->get_config().to_json();
    sc.optimizer_config = cfg_.optimizer_config.to_json();
    sc.scheduler_config = cfg_.scheduler_config.type.empty() ? Json::object() : cfg_.scheduler_config.to_json();
    sc.coordinator_endpoint = me;
    if (pipeline && s + 1 < n) sc.next_stage_endpoint = cfg_.worker_endpoints[s + 1];
    if (pipeline && s > 0) sc.prev_stage_endpoint = cfg_.worker_endpoints[s - 1];
    sc.stage_id = (int)s;
    sc.seed = cfg_.seed;
    stages_.push_back(sc);
  }
}
void Coordinator::deploy_stages() {
  for (size_t s = 0; s < stages_.size(); ++s) {
    comm_->connect(cfg_.worker_endpoints[s]);
    comm_->send_message(Message(CommandType::CONFIG_TRANSFER, stages_[s]), cfg_.worker_endpoints[s]);
  }
  join(CommandType::CONFIG_RECEIVED, stages_.size(), cfg_.timeout_deploy_ms, "deploy_stages");
}
void Coordinator::broadcast(const Message& m) {
  for (auto& e : cfg_.worker_endpoints) comm_->send_message(m, e);
}
void Coordinator::stop() { broadcast(Message(CommandType::SHUTDOWN)); }
void Coordinator::set_training(bool f) { broadcast(Message(f ? CommandType::TRAIN_MODE : CommandType::EVAL_MODE)); }
void Coordinator::forward(const Tensor& x, uint64_t mb, int stage) { comm_->send_message(Message(CommandType::FORWARD_JOB, Job{x, mb}), cfg_.worker_endpoints[(size_t)stage]); }
void Coordinator::backward(const Tensor& g, uint64_t mb, int stage) {
  const size_t s = stage < 0 ? stages_.size() - 1 : (size_t)stage;
  comm_->send_message(Message(CommandType::BACKWARD_JOB, Job{g, mb}), cfg_.worker_endpoints[s]);
}
Vec<Message> Coordinator::join(CommandType type, size_t n, int timeout_ms, const std::string& what) {
  Vec<Message> got;
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms);
  while (got.size() < n) {
    const int left = (int)std::chrono::duration_cast<std::chrono::milliseconds>(deadline - std::chrono::steady_clock::now()).count();
    if (left <= 0) throw DistributedError(what + ": timed out with " + std::to_string(got.size()) + "/" + std::to_string(n) + " " + command_name(type));
    Message m;
    if (!comm_->dequeue(&m, std::min(left, 500), {type, CommandType::ERROR_REPORT})) continue;
    if (m.command == CommandType::ERROR_REPORT) throw DistributedError("worker error during " + what + ": " + m.str);
    got.push_back(std::move(m));
  }
  return got;
}
void Coordinator::update_parameters() {
  if (cfg_.parallel_mode == ParallelMode_t::DATA && stages_.size() > 1) {
    // gradient reduction over the bus: gather every replica's gradient slab, sum, scatter (the reference has none)
    constexpr uint64_t GRAD_SLAB_TAG = (1ull << 62) | 0x6772;
    broadcast(Message(CommandType::SEND_PARAMS, "grads"));
    Vec<Message> got = join(CommandType::PARAMS_TRANSFER, stages_.size(), cfg_.timeout_update_ms, "gather replica gradients");
    Tensor total = got[0].job.data->is_gpu() ? got[0].job.data->to_host() : got[0].job.data->clone();
    for (size_t i = 1; i < got.size(); ++i) {
      Tensor h = got[i].job.data->is_gpu() ? got[i].job.data->to_host() : got[i].job.data;
      total->add(*h);
    }
    broadcast(Message(CommandType::PARAMS_TRANSFER, Job{total, GRAD_SLAB_TAG}));
    join(CommandType::PARAMS_LOADED, stages_.size(), cfg_.timeout_update_ms, "scatter summed gradients");
  }
  broadcast(Message(CommandType::UPDATE_PARAMETERS));
  join(CommandType::PARAMETERS_UPDATED, stages_.size(), cfg_.timeout_update_ms, "update_parameters");
}
std::pair<double, double> Coordinator::async_train_batch(const Vec<Tensor>& xs, const Vec<Tensor>& ys, Loss& criterion) {
  const size_t M = xs.size(), S = stages_.size();
  const bool data_mode = cfg_.parallel_mode == ParallelMode_t::DATA;
  for (size_t i = 0; i < M; ++i) forward(xs[i], i, data_mode ? (int)(i % S) : 0);
  double loss_total = 0, corrects = 0;
  size_t outputs = 0, completes = 0;
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(cfg_.timeout_job_ms);
  while (outputs < M || completes < M) {
    const int left = (int)std::chrono::duration_cast<std::chrono::milliseconds>(deadline - std::chrono::steady_clock::now()).count();
    if (left <= 0)
      throw DistributedError("async_train_batch: timed out with " + std::to_string(outputs) + "/" + std::to_string(M) + " outputs and " +
                             std::to_string(completes) + "/" + std::to_string(M) + " BACKWARD_COMPLETE (a lost microbatch deadlocks the reference; here it raises)");
    Message m;
    if (!comm_->dequeue(&m, std::min(left, 500), {CommandType::FORWARD_JOB, CommandType::BACKWARD_COMPLETE, CommandType::ERROR_REPORT})) continue;
    if (m.command == CommandType::ERROR_REPORT) throw DistributedError("worker error during async_train_batch: " + m.str);
    if (m.command == CommandType::BACKWARD_COMPLETE) {
      ++completes;
      continue;
    }
    // loss + gradient on the device the logits live on, scaled by 1/M (coordinator.hpp:188-207)
    const Tensor& out = m.job.data;
    if (out->is_gpu()) out->device().set_current();
    Tensor y = ys[m.job.mb_id]->device() == out->device() ? ys[m.job.mb_id] : ys[m.job.mb_id]->to_device(out->device());
    Tensor stats, grad;
    criterion.loss_and_grad(out, y, stats, &grad, 1.f / (float)M);
    float st[2];
    stats->copy_to_host(st, 8);
    loss_total += st[0];
    corrects += st[1];
    backward(grad, m.job.mb_id, data_mode ? (int)(m.job.mb_id % S) : -1);
    ++outputs;
  }
  return {loss_total, corrects};
}
std::pair<double, double> Coordinator::async_val_batch(const Tensor& x, const Tensor& y, Loss& criterion) {
  forward(x, 0, 0);
  Vec<Message> got = join(CommandType::FORWARD_JOB, 1, cfg_.timeout_job_ms, "async_val_batch");
  const Tensor& out = got[0].job.data;
  Tensor yy = y->device() == out->device() ? y : y->to_device(out->device());
  Tensor stats;
  criterion.loss_and_grad(out, yy, stats, nullptr, 1.f);
  float st[2];
  stats->copy_to_host(st, 8);
  return {st[0], st[1]};
}
void Coordinator::start_profiling() {
  GlobalProfiler::init_start_time();
  broadcast(Message(CommandType::START_PROFILING));
  join(CommandType::PROFILING_STARTED, stages_.size(), cfg_.timeout_profile_ms, "start_profiling");
}
Profiler Coordinator::fetch_profiling() {
  broadcast(Message(CommandType::REPORT_PROFILING));
  Vec<Message> got = join(CommandType::PROFILING_REPORTED, stages_.size(), cfg_.timeout_profile_ms, "fetch_profiling");
  Profiler merged;
  merged.init_start_time();
  std::vector<const void*> seen;                       // in-process workers share the process-global profiler
  for (auto& m : got) {
    if (!m.profiler) continue;
    const std::string sig = m.profiler->serialize();
    bool dup = false;
    for (auto& other : got)
      if (&other != &m && &other < &m && other.profiler && other.profiler->serialize() == sig) dup = true;
    if (!dup) merged.merge(*m.profiler);
  }
  merged.sort();
  return merged;
}
void Coordinator::print_profiling() {
  broadcast(Message(CommandType::PRINT_PROFILING));
  join(CommandType::PROFILING_PRINTED, stages_.size(), cfg_.timeout_profile_ms, "print_profiling");
}
void Coordinator::clear_profiling() {
  broadcast(Message(CommandType::CLEAR_PROFILING));
  join(CommandType::PROFILING_CLEARED, stages_.size(), cfg_.timeout_profile_ms, "clear_profiling");
}
Vec<Json> Coordinator::health_check(int timeout_ms) {
  broadcast(Message(CommandType::HEALTH_CHECK));
  Vec<Json> out;
  for (auto& m : join(CommandType::STATUS_RESPONSE, stages_.size(), timeout_ms, "health_check")) out.push_back(Json::parse(m.str));
  return out;
}
Vec<Json> Coordinator::report_load(int timeout_ms) {
  broadcast(Message(CommandType::REPORT_LOAD));
  Vec<Json> out;
  for (auto& m : join(CommandType::LOAD_REPORT, stages_.size(), timeout_ms, "report_load")) out.push_back(Json::parse(m.str));
  std::sort(out.begin(), out.end(), [](const Json& a, const Json& b) { return a.value("stage", 0) < b.value("stage", 0); });
  return out;
}
void Coordinator::barrier(int timeout_ms) {
  broadcast(Message(CommandType::BARRIER_SYNC));
  join(CommandType::BARRIER_SYNC, stages_.size(), timeout_ms, "barrier");
}
Vec<std::string> Coordinator::checkpoint(const std::string& path, int timeout_ms) {
  broadcast(Message(CommandType::CHECKPOINT_REQUEST, path));
  Vec<std::string> out;
  for (auto& m : join(CommandType::CHECKPOINT_COMPLETE, stages_.size(), timeout_ms, "checkpoint")) out.push_back(m.str);
  std::sort(out.begin(), out.end());
  return out;
}

// ===================================================================================== distributed trainer
void split_microbatches(const Tensor& x, const Tensor& y, size_t M, Vec<Tensor>* xs, Vec<Tensor>* ys) {
  const size_t B = x->dimension(0);
  M = std::max<size_t>(1, std::min(M, B));
  size_t cur = 0;
  for (size_t i = 0; i < M; ++i) {
    const size_t n = B / M + (i < B % M ? 1 : 0);
    xs->push_back(x->slice_batch(cur, n));
    ys->push_back(y->slice_batch(cur, n));
    cur += n;
  }
}
TrainHistory train_model(Coordinator& co, BaseDataLoader& train, BaseDataLoader* val, Loss& criterion, const TrainingConfig& cfg,
                         const std::function<void(const std::string&)>& log_in) {
  auto log = log_in ? log_in : [](const std::string& s) { std::cout << s << std::endl; };
  TrainHistory hist;
  co.start_profiling();
  uint64_t global_step = 0;
  for (size_t epoch = 0; epoch < cfg.epochs; ++epoch) {
    train.shuffle();
    train.reset();
    co.set_training(true);
    Tensor x, y;
    double loss_sum = 0, correct = 0;
    size_t batches = 0, seen = 0;
    const auto t0 = std::chrono::steady_clock::now();
    while (train.get_batch(cfg.batch_size, x, y)) {
      Vec<Tensor> xs, ys;
      split_microbatches(x, y, cfg.num_microbatches, &xs, &ys);
      const auto r = co.async_train_batch(xs, ys, criterion);
      loss_sum += r.first;
      correct += r.second;
      ++batches;
      seen += y->size();
      ++global_step;
      if (global_step % cfg.gradient_accumulation_steps == 0) co.update_parameters();
      if (cfg.max_steps > 0 && (long)global_step >= cfg.max_steps) break;
    }
    hist.train_loss.push_back(loss_sum / (double)std::max<size_t>(batches, 1));
    char b[200];
    snprintf(b, sizeof b, "epoch %zu: loss %.4f acc %.2f%% in %.2fs", epoch + 1, hist.train_loss.back(), 100.0 * correct / (double)std::max<size_t>(seen, 1),
             std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    log(b);
    if (val) {
      co.set_training(false);
      val->reset();
      double vl = 0, vc = 0;
      size_t vb = 0, vn = 0;
      while (val->get_batch(cfg.batch_size, x, y)) {
        const auto r = co.async_val_batch(x, y, criterion);
        vl += r.first;
        vc += r.second;
        ++vb;
        vn += y->size();
      }
      hist.val_acc.push_back(vc / (double)std::max<size_t>(vn, 1));
      snprintf(b, sizeof b, "validation: loss %.4f acc %.2f%%", vl / (double)std::max<size_t>(vb, 1), 100.0 * hist.val_acc.back());
      log(b);
    }
    if (cfg.max_steps > 0 && (long)global_step >= cfg.max_steps) break;
  }
  return hist;
}

}  // namespace tnn
