// usage_tour: the snippets of docs/USAGE.md as one runnable program (tensor, LayerBuilder / Sequential, data augmentation,
// an in-process two-stage pipeline).  Reference counterparts: docs/tensor_usage.md, docs/sequential_usage.en.md,
// docs/data_augmentation.md, docs/pipeline_architecture.md.
//   usage_tour [--device CPU|GPU:0]
#include <cstring>
#include <iostream>
#include <sstream>

#include "tnn/augmentation.hpp"
#include "tnn/distributed.hpp"
#include "tnn/example_models.hpp"
#include "tnn/layer_builder.hpp"
#include "tnn/train.hpp"

using namespace tnn;

// ---- a custom augmentation: invert the colours of an image with probability p (docs/USAGE.md section 3)
class ColorInversion : public AugmentationStrategy {
 public:
  explicit ColorInversion(double p) : p_(p) {}
  std::string name() const override { return "color_inversion"; }
  void apply(float* img, size_t H, size_t W, size_t C, std::mt19937_64& rng) const override {
    if (!hit(rng, p_)) return;
    for (size_t i = 0; i < H * W * C; ++i) img[i] = 1.0f - img[i];
  }

 private:
  double p_;
};

static void tensors(const Device& dev) {
  // [1] creation: dtype + shape (+ device); NHWC for images, [B, L, E] for sequences
  Tensor a = make_tensor<float>({2, 3});                       // host FP32
  Tensor b = make_tensor_from({1, 2, 3, 4, 5, 6}, {2, 3});
  a->fill(0.5);
  // [2] arithmetic: in place, or operators returning a new tensor
  a->add(*b);
  a->mul_scalar(2.0);
  Tensor c = a + b;
  std::cout << "c = " << c->head(6) << "  sum " << c->sum() << " mean " << c->mean() << " max " << c->max() << "\n";
  // [3] dtype / device: casts keep the shape; to_device / to_host copy
  Tensor h = dtype_cast(*c, DType_t::BF16);
  Tensor d = c->to_device(dev);                                // no-op copy on the CPU
  std::cout << "bf16 copy holds " << h->size() << " elements of " << get_dtype_size(h->data_type()) << " bytes; on " << d->device().name() << "\n";
  // [4] views share storage: a batch slice, a flat span, a reshape
  Tensor x = make_tensor<float>({4, 2, 2, 1});
  x->fill_random_uniform(0.0, 1.0, 7);
  Tensor first_two = x->slice_batch(0, 2);
  first_two->zero();
  std::cout << "after zeroing the view, sum of the first half: " << x->slice_batch(0, 2)->sum() << "\n";
  // [5] matrix product of 2-D tensors ([M, K] x [N, K]^T; tcgen05 kernels on the GPU)
  Tensor A = make_tensor_from({1, 2, 3, 4}, {2, 2}), B = make_tensor_from({1, 0, 0, 1}, {2, 2});
  std::cout << "A x I = " << matmul(*A->to_device(dev), *B->to_device(dev))->to_host()->head(4) << "\n";
  // [6] serialisation: {u32 dtype, u64 ndims, u64 dims[], raw bytes} -- the record checkpoints are made of
  std::stringstream ss;
  c->save(ss);
  Tensor back = load(ss, HostAllocator());
  std::cout << "round trip max |diff| = " << max_abs_diff(*back, *c) << "\n";
}

static void sequential(const Device& dev) {
  // LayerBuilder tracks the running shape, so layers only name what they add
  LayerBuilder b({32, 32, 3});                                 // batchless NHWC input
  b.conv2d(16, 3, 3, 1, 1, 1, 1, false, "conv1").batchnorm(1e-5, 0.1, true, true, "bn1").maxpool2d(2, 2, 2, 2, 0, 0, "pool1");
  b.basic_residual_block(16, 32, 2, "res1");
  b.avgpool2d(8, 8, 1, 1, 0, 0, "gap").flatten(1, -1, "flat").dense(10, true, "head");
  std::cout << "shape after the builder: [" << b.get_batchless_current_shape()[0] << "]\n";
  Graph g = graph_from_sequential(b.build_sequential("tiny_resnet"), dev, /*seed=*/1);
  auto opt = OptimizerFactory::create_sgd(0.05, 0.9);
  auto crit = LossFactory::create_logsoftmax_crossentropy();
  TrainStep step(g, *crit, *opt);                              // GPU: eager warm-up, then one captured CUDA graph per call
  SyntheticDataLoader data({32, 32, 3}, 10, 128, 3);
  data.set_pinned(dev.is_gpu());
  Tensor x, y;
  float first = 0, last = 0;
  for (int i = 0; i < 8 && data.get_batch(16, x, y); ++i) {
    float st[2];
    step(x, y)->copy_to_host(st, sizeof st);                   // {mean loss, #correct}
    if (i == 0) first = st[0];
    last = st[0];
  }
  std::cout << "tiny_resnet: loss " << first << " -> " << last << " over 8 steps; " << g.context().num_parameters() << " parameters\n";
  // configuration round trip: every layer (and the whole graph) is a TConfig
  const std::string json = g.get_config().to_json().dump();
  std::cout << "graph config is " << json.size() << " bytes of JSON\n";
}

static void augmentation() {
  auto aug = AugmentationBuilder(/*seed=*/42).horizontal_flip(0.5).random_crop(1.0, 4).brightness(0.3, 0.15).custom_augmentation(std::make_unique<ColorInversion>(0.25))
                 .normalize({0.5f, 0.5f, 0.5f}, {0.25f, 0.25f, 0.25f}).build();
  std::cout << "pipeline:";
  for (const auto& n : aug->names()) std::cout << " " << n;
  std::cout << "\n";
  Tensor batch = make_tensor<float>({8, 32, 32, 3});
  batch->fill_random_uniform(0.0, 1.0, 1);
  aug->apply(*batch, /*batch_index=*/0);                       // host FP32 [B, H, W, C], in place, deterministic per (seed, batch, image)
  std::cout << "augmented batch mean " << batch->mean() << "\n";
  aug->remove_augmentation("color_inversion");
  aug->remove_augmentation(0);                                 // by index: horizontal_flip
  std::cout << "after removals: " << aug->size() << " strategies\n";
  SyntheticDataLoader loader({32, 32, 3}, 10, 64, 3);
  loader.set_augmentation(aug);                                // loaders apply it to every batch they hand out
}

static void pipeline(const Device& dev) {
  // two stages in this process; tcp_worker / roce_worker are the same Worker behind a socket
  ExampleModels::register_defaults();
  Vec<std::unique_ptr<Worker>> workers;
  Vec<Endpoint> eps;
  for (int i = 0; i < 2; ++i) {
    workers.push_back(std::make_unique<Worker>(std::make_unique<InProcessCommunicator>("stage" + std::to_string(i)), dev, "stage" + std::to_string(i)));
    workers.back()->start(/*background=*/true);
    eps.push_back(workers.back()->comm().endpoint());
  }
  CoordinatorConfig cc;
  cc.model = ExampleModels::create("mnist_cnn");
  cc.input_shape = {8, 28, 28, 1};                             // one micro-batch, used for the FLOPs-weighted split
  cc.optimizer_config = OptimizerFactory::create_sgd(0.05, 0.9)->get_config();
  cc.partitioner = std::make_unique<NaivePipelinePartitioner>(Vec<size_t>{1, 1});
  cc.worker_endpoints = eps;
  Coordinator co(std::move(cc), std::make_unique<InProcessCommunicator>("coordinator"));
  co.start();                                                  // partitions the model, ships each stage as JSON (CONFIG_TRANSFER)
  auto crit = LossFactory::create_logsoftmax_crossentropy();
  SyntheticDataLoader data({28, 28, 1}, 10, 128, 5);
  Tensor x, y;
  for (int i = 0; i < 3 && data.get_batch(32, x, y); ++i) {
    Vec<Tensor> xs, ys;
    split_microbatches(x, y, 4, &xs, &ys);
    const auto res = co.async_train_batch(xs, ys, *crit);      // all forwards injected, backward per micro-batch as its output returns
    co.update_parameters();
    std::cout << "pipeline batch " << i << ": loss " << res.first << "\n";
  }
  for (const auto& l : co.report_load()) std::cout << "stage " << l.value("stage", -1) << ": " << l.value("avg_forward_time", 0.0) << " ms fwd, " << l.value("avg_backward_time", 0.0) << " ms bwd per job\n";
  co.stop();
  for (auto& w : workers) w->stop();
}

int main(int argc, char** argv) {
  std::string device;
  for (int i = 1; i < argc; ++i)
    if (!strcmp(argv[i], "--device") && i + 1 < argc) device = argv[++i];
  try {
    auto& dm = DeviceManager::getInstance();
    const Device dev = device.empty() ? (dm.has_gpu() ? dm.getGPU(0) : dm.getHost()) : dm.getDevice(device);
    std::cout << "== tensors\n";
    tensors(dev);
    std::cout << "== LayerBuilder / Sequential / TrainStep\n";
    sequential(dev);
    std::cout << "== data augmentation\n";
    augmentation();
    std::cout << "== pipeline\n";
    pipeline(dev);
    std::cout << "usage tour complete\n";
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "[usage_tour] error: " << e.what() << std::endl;
    return 1;
  }
}
