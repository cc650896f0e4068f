// op_bench: operator micro-benchmarks of the C++ host (reference: benchmarks/{gemm,dense,conv2d,batchnorm,attention,copy,
// grad_clear,serialization,tcp_communicator,roce_communicator,pipeline}_benchmark.cpp).  Times forward (+ backward) of one layer with CUDA events on the layer's
// flow (wall clock on the CPU), prints one JSON line per case.
//   op_bench [gemm|dense|conv2d|batchnorm|attention|copy|grad_clear|serialization|tcp|roce|pipeline|all] [--device CPU|GPU:i] [--iters N]
#include <chrono>
#include <cstring>
#include <iostream>
#include <sstream>
#include <thread>

#include "tnn/distributed.hpp"
#include "tnn/example_models.hpp"
#include "tnn/train.hpp"
#include "tnn/utils.hpp"

using namespace tnn;

static double time_ms(const Device& dev, int iters, const std::function<void()>& fn) {
  for (int i = 0; i < 2; ++i) fn();
  if (!dev.is_gpu()) return benchmark(fn, iters, 0);
  Flow& flow = current_flow(dev);
  Event e0(true), e1(true);
  flow.synchronize();
  flow.record(e0);
  for (int i = 0; i < iters; ++i) fn();
  flow.record(e1);
  e1.synchronize();
  return Event::elapsed_ms(e0, e1) / iters;
}
static void layer_case(const char* bench, const char* shape_name, std::unique_ptr<Layer> layer, const Shape& xs, const Device& dev, int iters) {
  layer->init(dev, 1);
  layer->set_training(true);
  Tensor x = make_tensor(DType_t::FP32, xs);
  x->fill_random_normal(0.0, 1.0, 3);
  if (dev.is_gpu()) x = x->to_device(dev);
  Tensor y = layer->forward(x, 0);
  Tensor dy = create_like(*y);
  dy->fill(0.01);
  const double fwd = time_ms(dev, iters, [&] { layer->forward(x, 0); });
  const double both = time_ms(dev, iters, [&] {
    layer->forward(x, 0);
    layer->backward(dy, 0);
  });
  const double flops = (double)layer->forward_flops(xs);
  std::cout << "{\"bench\": \"" << bench << "\", \"shape\": \"" << shape_name << "\", \"device\": \"" << dev.name() << "\", \"fwd_ms\": " << fwd
            << ", \"fwd_bwd_ms\": " << both << ", \"fwd_tflops\": " << (fwd > 0 ? flops / fwd / 1e9 : 0.0) << "}" << std::endl;
}

int main(int argc, char** argv) {
  std::string what = "all", device;
  int iters = 10;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--device") && i + 1 < argc) device = argv[++i];
    else if (!strcmp(argv[i], "--iters") && i + 1 < argc) iters = atoi(argv[++i]);
    else what = argv[i];
  }
  try {
    auto& dm = DeviceManager::getInstance();
    const Device dev = device.empty() ? (dm.has_gpu() ? dm.getGPU(0) : dm.getHost()) : dm.getDevice(device);
    const bool big = dev.is_gpu();
    auto want = [&](const char* n) { return what == "all" || what == n; };
    if (want("gemm")) {
      // gemm_benchmark.cpp: C[M,N] = A[M,K] x B[K,N] per dtype through tnn::matmul (tcgen05 kernels on the GPU, packed AVX2 on the CPU)
      const size_t M = big ? 4096 : 256, N = big ? 4096 : 256, K = big ? 4096 : 256;
      for (DType_t dt : big ? std::vector<DType_t>{DType_t::BF16, DType_t::FP16, DType_t::FP32} : std::vector<DType_t>{DType_t::FP32}) {
        Tensor a = make_tensor(DType_t::FP32, {M, K}), b = make_tensor(DType_t::FP32, {K, N});
        a->fill_random_normal(0.0, 1.0, 1);
        b->fill_random_normal(0.0, 1.0, 2);
        if (dev.is_gpu()) {
          a = a->to_device(dev);
          b = b->to_device(dev);
        }
        if (dt != DType_t::FP32) {
          a = dtype_cast(*a, dt);
          b = dtype_cast(*b, dt);
        }
        Tensor c;
        const double ms = time_ms(dev, iters, [&] { c = matmul(*a, *b); });
        std::cout << "{\"bench\": \"gemm\", \"dtype\": \"" << dtype_to_string(dt) << "\", \"shape\": \"" << M << "x" << N << "x" << K << "\", \"device\": \"" << dev.name()
                  << "\", \"ms\": " << ms << ", \"tflops\": " << 2.0 * M * N * K / ms / 1e9 << "}" << std::endl;
      }
    }
    for (const char* transport : {"tcp", "roce"}) {
      if (!want(transport)) continue;
      // tcp_communicator_benchmark.cpp / roce_communicator_benchmark.cpp: two endpoints bounce FORWARD_JOB messages for a
      // fixed number of round trips; bandwidth = payload bytes received / elapsed (both endpoints in this process, loopback)
      const bool roce = !strcmp(transport, "roce");
      auto make = [&](const char* nm) -> std::unique_ptr<Communicator> {
        if (roce) return std::make_unique<RoCECommunicator>("127.0.0.1", 0, 4, nullptr, nm);
        return std::make_unique<TCPCommunicator>("127.0.0.1", 0, 4, nullptr, nm);
      };
      auto a = make("bench_a"), b = make("bench_b");
      a->start();
      b->start();
      a->connect(b->endpoint());
      b->connect(a->endpoint());
      const size_t n = big ? (size_t)(32u << 20) : (size_t)(1u << 20);       // fp32 elements: 128 MiB (the reference's message) / 4 MiB
      Tensor t = make_tensor(DType_t::FP32, {n});
      t->fill_random_uniform(-1.0, 1.0, 5);
      const int trips = std::max(2, iters);
      std::thread echo([&] {
        Message m;
        for (int i = 0; i < trips; ++i) {
          if (!b->dequeue(&m, 60000, {CommandType::FORWARD_JOB})) return;
          b->send_message(m, a->endpoint());
        }
      });
      const auto t0 = std::chrono::steady_clock::now();
      bool ok = true;
      for (int i = 0; i < trips && ok; ++i) {
        a->send_message(Message(CommandType::FORWARD_JOB, Job{t, (uint64_t)i}), b->endpoint());
        Message back;
        ok = a->dequeue(&back, 60000, {CommandType::FORWARD_JOB}) && back.job.data->size() == n;
      }
      const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      echo.join();
      a->stop();
      b->stop();
      if (!ok) throw std::runtime_error(std::string(transport) + " ping-pong: message lost");
      std::cout << "{\"bench\": \"" << transport << "_communicator\", \"message_MB\": " << n * 4.0 / (1 << 20) << ", \"round_trips\": " << trips << ", \"seconds\": " << sec
                << ", \"MBps_per_direction\": " << trips * n * 4.0 / sec / 1e6 << "}" << std::endl;
    }
    if (want("pipeline")) {
      // pipeline_benchmark.cpp: a two-stage in-process pipeline (coordinator + two workers) against the same model on one device
      ExampleModels::register_defaults();
      const char* model = big ? "cifar10_resnet9" : "mnist_cnn";
      const Shape in = big ? Shape{32, 32, 3} : Shape{28, 28, 1};
      const size_t B = big ? 256 : 32, M = 4;
      SyntheticDataLoader data(in, 10, 4 * B, 5);
      Tensor x, y;
      data.get_batch(B, x, y);
      Vec<std::unique_ptr<Worker>> workers;
      Vec<Endpoint> eps;
      for (int i = 0; i < 2; ++i) {
        workers.push_back(std::make_unique<Worker>(std::make_unique<InProcessCommunicator>("w" + std::to_string(i)), dev, "w" + std::to_string(i)));
        workers.back()->start(true);
        eps.push_back(workers.back()->comm().endpoint());
      }
      CoordinatorConfig cc;
      cc.model = ExampleModels::create(model);
      Shape full{B / M};
      full.insert(full.end(), in.begin(), in.end());
      cc.input_shape = full;
      cc.optimizer_config = OptimizerFactory::create_sgd(0.01, 0.9)->get_config();
      cc.partitioner = std::make_unique<NaivePipelinePartitioner>(Vec<size_t>{1, 1});
      cc.worker_endpoints = eps;
      Coordinator co(std::move(cc), std::make_unique<InProcessCommunicator>("coord"));
      co.start();
      auto crit = LossFactory::create_logsoftmax_crossentropy();
      Vec<Tensor> xs, ys;
      split_microbatches(x, y, M, &xs, &ys);
      auto one = [&] {
        co.async_train_batch(xs, ys, *crit);
        co.update_parameters();
      };
      for (int i = 0; i < 2; ++i) one();
      const double pp_ms = benchmark(one, iters, 0);
      co.stop();
      for (auto& w : workers) w->stop();
      Graph g = load_or_create_model(model, "", dev, 1);
      auto opt = OptimizerFactory::create_sgd(0.01, 0.9);
      TrainStep step(g, *crit, *opt, nullptr, 1, false);
      for (int i = 0; i < 2; ++i) step(x, y);
      const double one_ms = benchmark([&] {
        Tensor st = step(x, y);
        float h[2];
        st->copy_to_host(h, 8);
      }, iters, 0);
      std::cout << "{\"bench\": \"pipeline\", \"model\": \"" << model << "\", \"device\": \"" << dev.name() << "\", \"batch\": " << B << ", \"microbatches\": " << M
                << ", \"two_stage_ms\": " << pp_ms << ", \"single_device_ms\": " << one_ms << "}" << std::endl;
    }
    if (want("dense")) {
      layer_case("dense", big ? "4096x4096x4096" : "256x512x512", std::make_unique<DenseLayer>(big ? 4096 : 512, big ? 4096 : 512), big ? Shape{4096, 4096} : Shape{256, 512}, dev, iters);
      if (big) layer_case("dense", "gpt2_fc1_32768x768x3072", std::make_unique<DenseLayer>(768, 3072), {32768, 768}, dev, iters);
    }
    if (want("conv2d")) {
      layer_case("conv2d", big ? "256x32x32x128_3x3" : "8x16x16x16_3x3", std::make_unique<Conv2DLayer>(big ? 128 : 16, big ? 128 : 16, 3, 3, 1, 1, 1, 1),
                 big ? Shape{256, 32, 32, 128} : Shape{8, 16, 16, 16}, dev, iters);
      if (big) layer_case("conv2d", "256x8x8x512_3x3", std::make_unique<Conv2DLayer>(512, 512, 3, 3, 1, 1, 1, 1), {256, 8, 8, 512}, dev, iters);
    }
    if (want("batchnorm"))
      layer_case("batchnorm", big ? "256x32x32x128" : "8x16x16x16", std::make_unique<BatchNormLayer>(big ? 128 : 16, 1e-5, 0.1, true, true),
                 big ? Shape{256, 32, 32, 128} : Shape{8, 16, 16, 16}, dev, iters);
    if (want("attention"))
      layer_case("attention", big ? "B16_L512_E768_H12" : "B2_L32_E128_H2", std::make_unique<AttentionBlock>(big ? 768 : 128, big ? 12 : 2, true),
                 big ? Shape{16, 512, 768} : Shape{2, 32, 128}, dev, iters);
    if (want("copy")) {
      const size_t n = big ? (256u << 20) : (16u << 20);
      Tensor a = make_tensor(DType_t::FP32, {n / 4}, dev), b = make_tensor(DType_t::FP32, {n / 4}, dev);
      a->fill(1.0);
      const double ms = time_ms(dev, iters, [&] { b->copy_from(*a); });
      std::cout << "{\"bench\": \"copy\", \"bytes\": " << n << ", \"device\": \"" << dev.name() << "\", \"ms\": " << ms << ", \"GBps\": " << 2.0 * n / ms / 1e6 << "}" << std::endl;
    }
    if (want("grad_clear")) {
      ExampleModels::register_defaults();
      Graph g = load_or_create_model("cifar10_resnet9", "", dev, 1);
      const double ms = time_ms(dev, iters, [&] { g.context().zero_grads(); });
      std::cout << "{\"bench\": \"grad_clear\", \"parameters\": " << g.context().num_parameters() << ", \"device\": \"" << dev.name() << "\", \"ms\": " << ms << "}" << std::endl;
    }
    if (want("serialization")) {
      Tensor t = make_tensor(DType_t::FP32, {big ? (size_t)(8u << 20) : (size_t)(1u << 20)});
      t->fill_random_uniform(-1.0, 1.0, 9);
      Job job;
      job.mb_id = 7;
      job.data = t;
      Message msg(CommandType::FORWARD_JOB, std::move(job));
      std::vector<uint8_t> wire;
      const double ser = benchmark([&] { wire = msg.serialize(); }, iters, 1);
      const double de = benchmark([&] { Message back = Message::deserialize(wire.data(), wire.size()); }, iters, 1);
      std::cout << "{\"bench\": \"serialization\", \"bytes\": " << wire.size() << ", \"serialize_ms\": " << ser << ", \"deserialize_ms\": " << de
                << ", \"serialize_GBps\": " << wire.size() / ser / 1e6 << "}" << std::endl;
    }
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "[op_bench] error: " << e.what() << std::endl;
    return 1;
  }
}
