// graph_test: rebuild a zoo model layer by layer through LayerFactory configs into a GraphBuilder, compile it and train
// it through the GraphExecutor API (reference: examples/graph_test.cpp:20-121).
//   graph_test [--config cfg.json] [--model NAME] [--steps N] [--batch B]
#include <cstring>
#include <iostream>

#include "tnn/example_models.hpp"
#include "tnn/train.hpp"

using namespace tnn;

int main(int argc, char** argv) {
  try {
    std::string config_path, model;
    size_t steps = 8, batch = 0;
    for (int i = 1; i < argc; ++i) {
      if (!strcmp(argv[i], "--config") && i + 1 < argc) config_path = argv[++i];
      else if (!strcmp(argv[i], "--model") && i + 1 < argc) model = argv[++i];
      else if (!strcmp(argv[i], "--steps") && i + 1 < argc) steps = (size_t)atol(argv[++i]);
      else if (!strcmp(argv[i], "--batch") && i + 1 < argc) batch = (size_t)atol(argv[++i]);
      else if (!strcmp(argv[i], "--help")) {
        std::cout << "usage: graph_test [--config cfg.json] [--model NAME] [--steps N] [--batch B]\n";
        return 0;
      }
    }
    ExampleModels::register_defaults();
    LayerFactory::register_defaults();
    TrainingConfig cfg = config_path.empty() ? TrainingConfig::load_from_env() : TrainingConfig::load_from_json(config_path);
    if (!model.empty()) cfg.model_name = model;
    if (batch) cfg.batch_size = batch;
    const Device dev = cfg.device();
    // every top-level layer of the zoo model becomes its own op node, recreated from its config
    std::unique_ptr<Sequential> zoo = ExampleModels::create(cfg.model_name);
    GraphBuilder builder(cfg.model_name + "_graph");
    builder.io("input");
    std::string current = "input";
    size_t i = 0;
    for (Layer* l : zoo->layers()) {
      OpNode& op = builder.add_layer(LayerFactory::create(l->get_config()));
      const std::string out = "layer_output_" + std::to_string(i++);
      builder.io(out);
      builder.add_edge({current}, {out}, op);
      current = out;
    }
    Graph graph = builder.compile(dev, 1234);
    std::cout << "[graph_test] " << graph.edges().size() << " edges, " << graph.context().num_parameters() << " parameters on " << dev.name() << std::endl;
    GraphExecutor executor(graph);
    DataLoaderPair data = DataLoaderFactory::create(cfg.dataset_name.empty() ? "synthetic" : cfg.dataset_name, cfg.dataset_path);
    if (cfg.dataset_name.empty())
      data.train = std::make_unique<SyntheticDataLoader>(ExampleModels::input_shape(cfg.model_name), ExampleModels::num_classes(cfg.model_name),
                                                         cfg.batch_size * steps, 1);
    data.train->set_seed(123456);
    auto criterion = LossFactory::create_logsoftmax_crossentropy();
    auto optimizer = OptimizerFactory::create_adam(cfg.lr_initial, 0.9, 0.999, 1e-3, 3e-4, false);
    optimizer->attach(graph.context());
    Tensor x, y;
    size_t step = 0;
    while (step < steps && data.train->get_batch(cfg.batch_size, x, y)) {
      Tensor dx = x->to_device(dev), dl = y->to_device(dev);
      InputPack in = {{"input", dx}};
      OutputPack out = {{current, nullptr}};
      executor.forward(in, out);
      Tensor pred = out[current];
      float loss = 0;
      criterion->compute_loss(pred, dl, loss);
      const size_t correct = compute_class_corrects(pred, dl);
      std::cout << "Loss: " << loss << ", Accuracy: " << 100.0 * (double)correct / (double)pred->dimension(0) << "%" << std::endl;
      Tensor grad = create_like(*pred);
      criterion->compute_gradient(pred, dl, grad);
      InputPack gout = {{current, grad}};
      OutputPack gin = {{"input", nullptr}};
      executor.backward(gout, gin);
      optimizer->update();
      optimizer->zero_grads();
      ++step;
    }
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "[graph_test] error: " << e.what() << std::endl;
    return 1;
  }
}
