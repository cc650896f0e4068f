// tcp_coordinator: pipeline-parallel training driven over the message bus (reference: examples/tcp_coordinator.cpp:27-108).
// Endpoints come from the environment like the reference's: WORKER_ENDPOINTS="host:port,host:port" (or WORKER1_HOST / WORKER1_PORT ...);
// with --in-process N the stages are N local workers (one per GPU when GPUs are visible): activations then move GPU to GPU.
//   tcp_coordinator [--in-process N] [--proportions 1,2] [--partitioner naive|weighted|ftd] [--synthetic] [--samples N]
#include <cstring>
#include <iostream>
#include <sstream>

#include "tnn/distributed.hpp"
#include "tnn/example_models.hpp"

using namespace tnn;

int main(int argc, char** argv) {
  int in_process = 0;
  std::string proportions, part = "naive";
  size_t samples = 2048;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--in-process") && i + 1 < argc) in_process = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--proportions") && i + 1 < argc) proportions = argv[++i];
    else if (!strcmp(argv[i], "--partitioner") && i + 1 < argc) part = argv[++i];
    else if (!strcmp(argv[i], "--samples") && i + 1 < argc) samples = (size_t)atol(argv[++i]);
  }
  try {
    TrainingConfig cfg = TrainingConfig::load_from_env();
    ExampleModels::register_defaults();
    Vec<std::unique_ptr<Worker>> local;
    Vec<Endpoint> eps;
    std::unique_ptr<Communicator> comm;
    if (in_process > 0) {
      const int ngpu = DeviceManager::getInstance().num_gpus();
      for (int i = 0; i < in_process; ++i) {
        const Device dev = ngpu > 0 && cfg.device_type != "CPU" ? getGPU(i % ngpu) : getHost();
        auto w = std::make_unique<Worker>(std::make_unique<InProcessCommunicator>("worker" + std::to_string(i)), dev, "worker" + std::to_string(i));
        eps.push_back(w->comm().endpoint());
        w->start(true);
        local.push_back(std::move(w));
      }
      comm = std::make_unique<InProcessCommunicator>("coordinator");
    } else {
      const char* list = getenv("WORKER_ENDPOINTS");
      std::string s = list ? list : "";
      if (s.empty())
        for (int i = 1; i <= 8; ++i) {
          const char* h = getenv(("WORKER" + std::to_string(i) + "_HOST").c_str());
          const char* p = getenv(("WORKER" + std::to_string(i) + "_PORT").c_str());
          if (h && p) s += (s.empty() ? "" : ",") + std::string(h) + ":" + p;
        }
      std::stringstream ss(s);
      std::string item;
      while (std::getline(ss, item, ',')) {
        const size_t c = item.rfind(':');
        if (c != std::string::npos) eps.push_back(Endpoint::tcp(item.substr(0, c), atoi(item.c_str() + c + 1)));
      }
      if (eps.empty()) throw std::runtime_error("no workers: set WORKER_ENDPOINTS=host:port,... or pass --in-process N");
      const char* ch = getenv("COORDINATOR_HOST");
      const char* cp = getenv("COORDINATOR_PORT");
      comm = std::make_unique<TCPCommunicator>(ch ? ch : "127.0.0.1", cp ? atoi(cp) : 0, 4, nullptr, "coordinator");
    }
    CoordinatorConfig cc;
    cc.parallel_mode = ParallelMode_t::PIPELINE;
    cc.model = ExampleModels::create(cfg.model_name);
    const Shape s = ExampleModels::input_shape(cfg.model_name);
    cc.input_shape = {cfg.batch_size / std::max<size_t>(cfg.num_microbatches, 1)};
    cc.input_shape.insert(cc.input_shape.end(), s.begin(), s.end());
    // examples/roce_coordinator.cpp:111: Adam eps 1e-5, wd 1e-4
    cc.optimizer_config = OptimizerFactory::create_adam(cfg.lr_initial, 0.9, 0.999, 1e-5, 1e-4)->get_config();
    Vec<size_t> props;
    std::stringstream ps(proportions);
    std::string tok;
    while (std::getline(ps, tok, ',')) props.push_back((size_t)atol(tok.c_str()));
    if (props.size() != eps.size()) props.assign(eps.size(), 1);
    if (part == "weighted") cc.partitioner = std::make_unique<WeightedPartitioner>(eps.size());
    else if (part == "ftd") cc.partitioner = std::make_unique<FTDPartitioner>(eps.size());
    else cc.partitioner = std::make_unique<NaivePipelinePartitioner>(props);
    cc.worker_endpoints = eps;
    Coordinator co(std::move(cc), std::move(comm));
    co.start();
    std::cout << "[coordinator] " << co.num_stages() << " stages:";
    for (auto& p : co.partitions()) std::cout << " [" << p.start << "," << p.start + p.length << ")";
    std::cout << std::endl;
    const size_t classes = ExampleModels::num_classes(cfg.model_name);
    SyntheticDataLoader train(s, classes, samples, 1), val(s, classes, std::max<size_t>(samples / 8, cfg.batch_size), 2);
    auto crit = LossFactory::create_logsoftmax_crossentropy();
    TrainHistory h = train_model(co, train, &val, *crit, cfg);
    Profiler prof = co.fetch_profiling();
    prof.log("logs_profiler.log");
    std::cout << "[coordinator] done: final loss " << (h.train_loss.empty() ? 0.0 : h.train_loss.back()) << ", " << prof.events().size() << " profiler events" << std::endl;
    co.stop();
    for (auto& w : local) w->stop();
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "[tcp_coordinator] error: " << e.what() << std::endl;
    return 1;
  }
}
