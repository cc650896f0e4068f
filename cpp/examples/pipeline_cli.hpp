// Shared bodies of tcp_worker / roce_worker and tcp_coordinator / roce_coordinator (reference: examples/tcp_worker.cpp:39-141,
// examples/roce_worker.cpp, examples/tcp_coordinator.cpp:27-108, examples/roce_coordinator.cpp): the two families differ only
// in the transport -- the packetised TCP bus, or the one-sided rendezvous transport (tnn::RoCECommunicator: GPU jobs are
// written device to device into the peer's CUDA-IPC slab).
#pragma once
#include <cstring>
#include <iostream>
#include <sstream>

#include "tnn/distributed.hpp"
#include "tnn/example_models.hpp"

namespace tnn {
namespace cli {

inline std::unique_ptr<Communicator> make_transport(bool roce, const std::string& host, int port, const Device* land, const std::string& name) {
  if (roce) return std::make_unique<RoCECommunicator>(host, port, 4, land, name);
  return std::make_unique<TCPCommunicator>(host, port, 4, land, name);
}

//   <tcp|roce>_worker <port> [--gpu [index]] [--host 127.0.0.1] [--io-threads N] [--num-threads N]
inline int worker_main(int argc, char** argv, bool roce) {
  const std::string tag = roce ? "roce" : "tcp";
  int port = 0, gpu = -1;
  std::string host = "127.0.0.1";
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--gpu")) gpu = (i + 1 < argc && isdigit((unsigned char)argv[i + 1][0])) ? atoi(argv[++i]) : 0;
    else if (!strcmp(argv[i], "--host") && i + 1 < argc) host = argv[++i];
    else if ((!strcmp(argv[i], "--io-threads") || !strcmp(argv[i], "--num-threads")) && i + 1 < argc) ++i;
    else if (isdigit((unsigned char)argv[i][0])) port = atoi(argv[i]);
    else if (!strcmp(argv[i], "--help")) {
      std::cout << "usage: " << tag << "_worker <port> [--gpu [index]] [--host ADDR]\n";
      return 0;
    }
  }
  try {
    const Device dev = gpu >= 0 && DeviceManager::getInstance().has_gpu() ? getGPU(gpu) : getHost();
    const std::string name = tag + "-worker-" + std::to_string(port);
    Worker w(make_transport(roce, host, port, &dev, name), dev, name);
    std::cout << "[" << tag << "_worker] listening on " << host << ":" << port << " device " << dev.name() << std::endl;
    w.start(false);       // blocks until SHUTDOWN
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "[" << tag << "_worker] error: " << e.what() << std::endl;
    return 1;
  }
}

//   <tcp|roce>_coordinator [--in-process N] [--proportions 1,2] [--partitioner naive|weighted|ftd] [--samples N] [--gpu [index]]
// Endpoints come from the environment like the reference's: WORKER_ENDPOINTS="host:port,host:port" (or WORKER1_HOST / WORKER1_PORT ...);
// with --in-process N the stages are N local workers (one per GPU when GPUs are visible): activations then move GPU to GPU.
inline int coordinator_main(int argc, char** argv, bool roce) {
  const std::string tag = roce ? "roce" : "tcp";
  int in_process = 0, coord_gpu = -1;
  std::string proportions, part = "naive";
  size_t samples = 2048;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--in-process") && i + 1 < argc) in_process = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--proportions") && i + 1 < argc) proportions = argv[++i];
    else if (!strcmp(argv[i], "--partitioner") && i + 1 < argc) part = argv[++i];
    else if (!strcmp(argv[i], "--samples") && i + 1 < argc) samples = (size_t)atol(argv[++i]);
    else if (!strcmp(argv[i], "--gpu")) coord_gpu = (i + 1 < argc && isdigit((unsigned char)argv[i + 1][0])) ? atoi(argv[++i]) : 0;
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
        if (c != std::string::npos) eps.push_back(roce ? Endpoint::roce(item.substr(0, c), atoi(item.c_str() + c + 1)) : Endpoint::tcp(item.substr(0, c), atoi(item.c_str() + c + 1)));
      }
      if (eps.empty()) throw std::runtime_error("no workers: set WORKER_ENDPOINTS=host:port,... or pass --in-process N");
      const char* ch = getenv("COORDINATOR_HOST");
      const char* cp = getenv("COORDINATOR_PORT");
      // the coordinator feeds stage 0 and receives the last stage's outputs: with --gpu its tensors live on a GPU and, over the
      // rendezvous transport, reach the workers device to device
      const Device cdev = coord_gpu >= 0 && DeviceManager::getInstance().has_gpu() ? getGPU(coord_gpu) : getHost();
      comm = make_transport(roce, ch ? ch : "127.0.0.1", cp ? atoi(cp) : 0, cdev.is_gpu() ? &cdev : nullptr, "coordinator");
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
    std::cerr << "[" << tag << "_coordinator] error: " << e.what() << std::endl;
    return 1;
  }
}

}  // namespace cli
}  // namespace tnn
