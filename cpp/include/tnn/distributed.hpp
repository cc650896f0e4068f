// Distributed runtime: wire format, communicators, partitioners, Worker, Coordinator, distributed trainer
// (reference: include/distributed/{command_type,message,job,stage_config,endpoint,communicator,worker,coordinator,train}.hpp,
//  include/partitioner/*.hpp).  Same Message / CommandType / Job / StageConfig protocol as the reference, serialised with
// the native archiver byte for byte like the Python host (tnn_b200/distributed/message.py), so C++ and Python peers
// interoperate.  B200-first differences: IN_PROCESS endpoints hand DEVICE tensors over by reference --
// a pipeline of workers on the GPUs of one node moves activations GPU to GPU (peer copies over NVLink), never through
// host serialisation; the TCP transport (csrc/runtime/tcp_bus.hpp) remains for control and for peers on other hosts.
#pragma once
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <thread>

#include "runtime/tcp_bus.hpp"
#include "tnn/common.hpp"
#include "tnn/profiler.hpp"
#include "tnn/train.hpp"

namespace tnn {

enum class CommandType : uint16_t {
  _START = 0, FORWARD_JOB = 1, BACKWARD_JOB = 2, BACKWARD_COMPLETE = 3, UPDATE_PARAMETERS = 4, TRAIN_MODE = 5, EVAL_MODE = 6, SHUTDOWN = 7,
  CONFIG_TRANSFER = 8, CONFIG_RECEIVED = 9, LOAD_PARAMS = 10, PARAMS_LOADED = 11, SEND_PARAMS = 12, PARAMS_TRANSFER = 13, SAVE_TO_FILE = 14,
  STATUS_REQUEST = 15, STATUS_RESPONSE = 16, PARAMETERS_UPDATED = 17, HEALTH_CHECK = 18, ERROR_REPORT = 19, JOB_FAILURE = 20, BARRIER_SYNC = 21,
  CHECKPOINT_REQUEST = 22, CHECKPOINT_COMPLETE = 23, UPDATE_LOAD = 24, REPORT_LOAD = 25, LOAD_REPORT = 26, START_PROFILING = 27,
  PROFILING_STARTED = 28, PRINT_PROFILING = 29, PROFILING_PRINTED = 30, CLEAR_PROFILING = 31, PROFILING_CLEARED = 32, REPORT_PROFILING = 33,
  PROFILING_REPORTED = 34, HANDSHAKE = 35, HANDSHAKE_ACK = 36, _COUNT = 37
};
const char* command_name(CommandType c);
enum class ParallelMode_t { DATA = 0, PIPELINE = 1 };

class Communicator;
struct Endpoint {
  enum Type { NONE, IN_PROCESS, TCP, ROCE } type = NONE;
  std::string host;
  int port = 0;
  std::string key;                    // IN_PROCESS registry key
  static Endpoint tcp(const std::string& host, int port) { Endpoint e; e.type = TCP; e.host = host; e.port = port; return e; }
  static Endpoint roce(const std::string& host, int port, const std::string& = "", int = -1) { Endpoint e; e.type = ROCE; e.host = host; e.port = port; return e; }
  static Endpoint in_process(Communicator* c);
  Communicator* communicator() const;
  std::string id() const { return type == IN_PROCESS ? key : host + ":" + std::to_string(port); }
  explicit operator bool() const { return type != NONE; }
  Json to_json() const;
  static Endpoint from_json(const Json& j);
};

struct Job {
  Tensor data;
  uint64_t mb_id = 0;
};
struct StageConfig {
  Json model_config = Json::object(), optimizer_config = Json::object(), scheduler_config = Json::object();
  Endpoint next_stage_endpoint, prev_stage_endpoint, coordinator_endpoint;
  int stage_id = 0;
  uint64_t seed = 123456;             // worker.hpp:98-130 seeds every stage with 123456
  std::string device;
  Json to_json() const;
  static StageConfig from_json(const Json& j);
};

struct Message {
  enum Kind { NONE = 0, JOB = 1, STRING = 2, BOOL = 3, PROFILER = 4, STAGE = 5 };
  CommandType command = CommandType::_START;
  Kind kind = NONE;
  Job job;
  std::string str;
  bool flag = false;
  std::shared_ptr<Profiler> profiler;
  StageConfig stage;
  std::string sender;                 // filled in by the receiving communicator
  Message() = default;
  explicit Message(CommandType c) : command(c) {}
  Message(CommandType c, Job j) : command(c), kind(JOB), job(std::move(j)) {}
  Message(CommandType c, const std::string& s) : command(c), kind(STRING), str(s) {}
  Message(CommandType c, const char* s) : command(c), kind(STRING), str(s) {}
  Message(CommandType c, bool b) : command(c), kind(BOOL), flag(b) {}
  Message(CommandType c, const Profiler& p) : command(c), kind(PROFILER), profiler(std::make_shared<Profiler>(p)) {}
  Message(CommandType c, const StageConfig& s) : command(c), kind(STAGE), stage(s) {}
  // archiver encoding: u16 command, u32 variant index, payload (Job = u64 mb_id + {u32 dtype, shape, raw}); device tensors
  // are staged to the host here -- only the TCP transport serialises
  std::vector<uint8_t> serialize() const;
  static Message deserialize(const uint8_t* data, size_t n, bool swap = false, const Device* land_on = nullptr);
};

// ------------------------------------------------------------------------------------- communicators
class Communicator {
 public:
  explicit Communicator(std::string name) : name_(std::move(name)) {}
  virtual ~Communicator();
  virtual void start() {}
  virtual void stop() {}
  virtual Endpoint endpoint() = 0;
  virtual void connect(const Endpoint&) {}
  void send_message(const Message& m, const Endpoint& to);
  // highest-priority message (lowest CommandType value first: FORWARD before BACKWARD before ...), optionally restricted to `types`
  bool dequeue(Message* out, int timeout_ms, const std::vector<CommandType>& types = {});
  void enqueue(Message m);                                 // delivery entry point (transports + in-process peers)
  size_t pending() const;
  const std::string& name() const { return name_; }
  uint64_t sent_messages = 0, received_messages = 0;

 protected:
  virtual void send_remote(const Message&, const Endpoint&) { throw std::runtime_error(name_ + ": transport cannot reach a remote endpoint"); }
  std::string name_;
  mutable std::mutex mu_;
  std::condition_variable cv_;
  std::deque<Message> queues_[(size_t)CommandType::_COUNT];
};
class InProcessCommunicator : public Communicator {
 public:
  explicit InProcessCommunicator(const std::string& name = "inproc") : Communicator(name) {}
  Endpoint endpoint() override { return Endpoint::in_process(this); }
};
class TCPCommunicator : public Communicator {
 public:
  TCPCommunicator(const std::string& host = "127.0.0.1", int port = 0, int sockets_per_peer = 4, const Device* land_on = nullptr, const std::string& name = "tcp");
  ~TCPCommunicator() override;
  void start() override;
  void stop() override;
  Endpoint endpoint() override { return Endpoint::tcp(host_, bus_->port()); }
  void connect(const Endpoint& e) override;
  TcpBus& bus() { return *bus_; }

 protected:
  void send_remote(const Message& m, const Endpoint& to) override;
  std::string host_;
  std::unique_ptr<TcpBus> bus_;
  std::thread pump_;
  std::atomic<bool> running_{false};
  Device land_;
  bool has_land_ = false;
};

// One-sided rendezvous transport (the RoCE backend of the reference, include/distributed/roce_communicator.hpp:39-391,
// on one node): frame-compatible with tnn_b200/distributed/rendezvous.py, so Python and C++ peers interoperate.
//   * small / control messages: INLINE frames on the TCP bus; large host messages: PREPARE(len) -> READY_TO_WRITE(offset,
//     segment) -> the sender memcpy's the serialised message into the receiver's registered shared-memory slab -> DONE
//     (at most 16 in flight per peer, PREPAREs queue while the slab is full, inline fallback across hosts);
//   * GPU jobs: DEVICE rendezvous.  Every communicator that lands on a GPU registers a device slab (cudaMalloc +
//     cudaIpcGetMemHandle); G_PREPARE(nbytes) -> G_READY(offset, ipc handle) -> the sender copies the tensor device to
//     device straight into the peer's block (NVLink between the GPUs of a node; the GPUDirect-RDMA write of the
//     reference) -> G_DONE(metadata).  The delivered tensor IS the block (zero copy); the block returns to the slab
//     when the tensor dies and the consumer stream has passed that point.
class RoCECommunicator : public TCPCommunicator {
 public:
  RoCECommunicator(const std::string& host = "127.0.0.1", int port = 0, int sockets_per_peer = 4, const Device* land_on = nullptr,
                   const std::string& name = "roce", size_t device_slab_bytes = 256u << 20, size_t inline_bytes = 16u << 10,
                   size_t host_slab_bytes = 64u << 20);
  ~RoCECommunicator() override;
  void start() override;
  void stop() override;
  Endpoint endpoint() override { return Endpoint::roce(host_, bus_->port()); }
  struct Stats {
    uint64_t inline_sent = 0, device_rendezvous_sent = 0, device_rendezvous_received = 0, device_fallback_host = 0, slab_peak = 0;
    uint64_t rendezvous_sent = 0, rendezvous_received = 0, fallback_inline = 0, host_slab_peak = 0;     // host (shared-memory) slab
  };
  Stats stats() const;
  struct Impl;

 protected:
  void send_remote(const Message& m, const Endpoint& to) override;

 private:
  std::shared_ptr<Impl> impl_;      // shared with the deleters of delivered tensors (they may outlive the communicator)
};

// ------------------------------------------------------------------------------------- partitioners
struct SeqPartition {
  size_t start, length;
};
class Partitioner {
 public:
  virtual ~Partitioner() = default;
  virtual Vec<SeqPartition> partition_model(const Vec<Layer*>& layers, const Shape& input_shape) = 0;
  // batch rows per replica (DATA mode)
  virtual Vec<Tensor> partition_input(const Tensor& x, size_t parts);
  // builds one Sequential per partition from the layers' CONFIGS (fresh, uninitialised layers, as the reference's split())
  static Vec<std::unique_ptr<Sequential>> split(const Vec<Layer*>& layers, const Vec<SeqPartition>& parts, const std::string& base_name);
};
class NaivePipelinePartitioner : public Partitioner {     // split the layer COUNT by integer proportions (>= 1 layer each)
 public:
  explicit NaivePipelinePartitioner(Vec<size_t> proportions) : props_(std::move(proportions)) {}
  Vec<SeqPartition> partition_model(const Vec<Layer*>& layers, const Shape&) override;

 private:
  Vec<size_t> props_;
};
class NaiveDataPartitioner : public Partitioner {         // every replica holds the whole model; the batch is split
 public:
  explicit NaiveDataPartitioner(Vec<size_t> proportions) : props_(std::move(proportions)) {}
  Vec<SeqPartition> partition_model(const Vec<Layer*>& layers, const Shape&) override;

 private:
  Vec<size_t> props_;
};
// exact min-bottleneck contiguous split by forward FLOPs (the reference leaves this one commented out)
class WeightedPartitioner : public Partitioner {
 public:
  explicit WeightedPartitioner(size_t stages) : n_(stages) {}
  Vec<SeqPartition> partition_model(const Vec<Layer*>& layers, const Shape& input_shape) override;

 private:
  size_t n_;
};
// FLOPs / Transfer / Delay: stage cost = flops / flops_per_s + boundary activation bytes / link_bytes_per_s + delay
// (declared but never defined in the reference, ftd_partitioner.hpp:29-50)
class FTDPartitioner : public Partitioner {
 public:
  FTDPartitioner(size_t stages, double flops_per_s = 1.0e15, double link_bytes_per_s = 7.0e11, double delay_s = 5e-6)
      : n_(stages), fl_(flops_per_s), bw_(link_bytes_per_s), delay_(delay_s) {}
  Vec<SeqPartition> partition_model(const Vec<Layer*>& layers, const Shape& input_shape) override;

 private:
  size_t n_;
  double fl_, bw_, delay_;
};

// ------------------------------------------------------------------------------------- worker
class Worker {
 public:
  Worker(std::unique_ptr<Communicator> comm, const Device& dev, const std::string& name = "worker");
  ~Worker();
  void start(bool background = true);      // event loop (worker.hpp:64-89)
  void stop();
  void run();                              // blocking loop
  Communicator& comm() { return *comm_; }
  const Device& device() const { return dev_; }
  void set_config(const StageConfig& cfg); // rebuild the stage from JSON (worker.hpp:98-130)
  Graph* graph() { return graph_.get(); }
  std::function<void(Message&)> fault_inject;     // tests: drop / delay / throw
  // Directories a peer may name in SAVE_TO_FILE / LOAD_PARAMS / CHECKPOINT_REQUEST (the bus is unauthenticated): the working
  // directory, the system temp directory and TNN_WORKER_IO_ROOT (':'-separated), as in the Python worker.  Throws on a path
  // that resolves (symlinks and ".." included) outside every root.
  std::string checked_path(const std::string& path) const;
  // the stage's LoadTracker record (load_tracker.hpp:10-21): what REPORT_LOAD answers with
  LoadTracker load_tracker() const;
  const std::vector<std::string>& io_roots() const { return io_roots_; }

 private:
  void process(Message& m);
  void reply(const Message& m, const Endpoint* to = nullptr);
  std::unique_ptr<Communicator> comm_;
  Device dev_;
  std::string name_;
  std::unique_ptr<Graph> graph_;
  std::unique_ptr<GraphExecutor> exec_;
  std::unique_ptr<Optimizer> opt_;
  std::unique_ptr<Scheduler> sched_;
  StageConfig cfg_;
  std::thread thread_;
  std::atomic<bool> stop_{false};
  uint64_t jobs_ = 0, fwd_jobs_ = 0, bwd_jobs_ = 0;
  int64_t busy_ns_ = 0, fwd_ns_ = 0, bwd_ns_ = 0;
  double cpu_t0_ = 0;                                 // process CPU seconds / wall clock at construction
  std::chrono::steady_clock::time_point wall_t0_ = std::chrono::steady_clock::now();
  std::vector<std::string> io_roots_;
};

// ------------------------------------------------------------------------------------- coordinator
struct CoordinatorConfig {
  ParallelMode_t parallel_mode = ParallelMode_t::PIPELINE;
  std::unique_ptr<Sequential> model;
  Shape input_shape;                        // with batch dimension (FLOPs / transfer estimates)
  TConfig optimizer_config, scheduler_config;
  std::unique_ptr<Partitioner> partitioner;
  Vec<Endpoint> worker_endpoints;           // in stage order
  uint64_t seed = 123456;
  int timeout_deploy_ms = 60000, timeout_job_ms = 120000, timeout_update_ms = 60000, timeout_profile_ms = 30000;
};
class Coordinator {
 public:
  Coordinator(CoordinatorConfig cfg, std::unique_ptr<Communicator> comm);
  ~Coordinator();
  void initialize();                        // partition + StageConfigs (coordinator.hpp:398-445)
  void deploy_stages();                     // connect, CONFIG_TRANSFER, join(CONFIG_RECEIVED) (coordinator.hpp:368-395)
  void start() { initialize(); deploy_stages(); }
  void stop();
  void set_training(bool f);
  void forward(const Tensor& x, uint64_t mb_id, int stage = 0);
  void backward(const Tensor& grad, uint64_t mb_id, int stage = -1);
  void update_parameters();
  Vec<Message> join(CommandType type, size_t n, int timeout_ms, const std::string& what);
  // semi-asynchronous micro-batched step; returns (sum of microbatch mean losses / M, #correct) (coordinator.hpp:165-223)
  std::pair<double, double> async_train_batch(const Vec<Tensor>& xs, const Vec<Tensor>& ys, Loss& criterion);
  std::pair<double, double> async_val_batch(const Tensor& x, const Tensor& y, Loss& criterion);
  void start_profiling();
  Profiler fetch_profiling();
  void print_profiling();
  void clear_profiling();
  Vec<Json> health_check(int timeout_ms = 5000);
  // REPORT_LOAD -> one record per stage: {stage, jobs, avg_forward_backward_ms} + the LoadTracker fields
  Vec<Json> report_load(int timeout_ms = 10000);
  void barrier(int timeout_ms = 30000);
  Vec<std::string> checkpoint(const std::string& path, int timeout_ms = 60000);
  size_t num_stages() const { return stages_.size(); }
  const Vec<SeqPartition>& partitions() const { return parts_; }
  Communicator& comm() { return *comm_; }

 private:
  void broadcast(const Message& m);
  CoordinatorConfig cfg_;
  std::unique_ptr<Communicator> comm_;
  Vec<StageConfig> stages_;
  Vec<SeqPartition> parts_;
};

struct DistributedError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
// batch -> M microbatches by rows
void split_microbatches(const Tensor& x, const Tensor& y, size_t M, Vec<Tensor>* xs, Vec<Tensor>* ys);
TrainHistory train_model(Coordinator& co, BaseDataLoader& train, BaseDataLoader* val, Loss& criterion, const TrainingConfig& cfg,
                         const std::function<void(const std::string&)>& log = nullptr);

}  // namespace tnn
