// msequential_example: parallel Sequential branches joined by an n-ary layer (reference:
// examples/msequential_example.cpp:105-272): a two-branch basic case and three heterogeneous branches of different
// depth, forward + backward, with the memory-aware branch order printed.
#include <iostream>

#include "tnn/layer_builder.hpp"
#include "tnn/layers.hpp"

using namespace tnn;

static void run(MSequential& ms, const Tensor& x, const char* title) {
  std::cout << "=== " << title << " ===" << std::endl;
  ms.set_training(true);
  std::cout << "Forward Pass" << std::endl;
  Tensor y = ms.forward(x, 0);
  std::cout << "Output shape: " << shape_to_string(y->shape()) << std::endl;
  const auto v = to_float_vector(*y);
  std::cout << "First 5 output values:";
  for (size_t i = 0; i < 5 && i < v.size(); ++i) std::cout << ' ' << v[i];
  std::cout << "\nBranch order chosen by the block (largest transient memory first):";
  for (size_t i : ms.order()) std::cout << ' ' << i;
  std::cout << "\nBackward Pass" << std::endl;
  Tensor dy = create_like(*y);
  dy->fill(1.0);
  Tensor dx = ms.backward(dy, 0);
  size_t with_grads = 0;
  for (Layer* l : ms.all_layers()) with_grads += l->grads().size();
  std::cout << "Input gradient shape: " << shape_to_string(dx->shape()) << "; parameter tensors with gradients: " << with_grads << std::endl;
  std::cout << "Backward pass completed successfully.\n" << std::endl;
}

int main() {
  try {
    auto& dm = DeviceManager::getInstance();
    const Device dev = dm.has_gpu() ? dm.getGPU(0) : dm.getHost();
    std::cout << "[msequential_example] device " << dev.name() << std::endl;
    {
      Vec<std::unique_ptr<Layer>> branches;
      branches.push_back(LayerBuilder({16}).dense(24, true, "b1_fc1").activation("relu", "b1_relu").dense(8, true, "b1_fc2").build_sequential("branch_1"));
      branches.push_back(LayerBuilder({16}).dense(16, true, "b2_fc1").activation("tanh", "b2_tanh").dense(8, true, "b2_fc2").build_sequential("branch_2"));
      MSequential ms(std::move(branches), std::make_unique<NAryLayer>(0, "join_add"), "mseq_basic");
      ms.init(dev, 1);
      Tensor x = make_tensor(DType_t::FP32, {4, 16});
      x->fill_random_normal(0.0, 1.0, 11);
      run(ms, dev.is_gpu() ? x->to_device(dev) : x, "MSequential Basic Example");
    }
    {
      Vec<std::unique_ptr<Layer>> branches;
      branches.push_back(LayerBuilder({104}).dense(56, true, "branch1_dense").build_sequential("shallow_branch"));
      branches.push_back(LayerBuilder({104}).dense(200, true, "branch2_dense1").activation("relu", "branch2_relu").dense(56, true, "branch2_dense2").build_sequential("medium_branch"));
      branches.push_back(LayerBuilder({104}).dense(400, true, "branch3_dense1").activation("relu", "branch3_relu1").dense(200, true, "branch3_dense2")
                             .activation("relu", "branch3_relu2").dense(56, true, "branch3_dense3").build_sequential("deep_branch"));
      MSequential ms(std::move(branches), std::make_unique<NAryLayer>(2, "join_mul"), "mseq_hetero");
      ms.init(dev, 2);
      Tensor x = make_tensor(DType_t::FP32, {8, 104});
      x->fill_random_normal(0.0, 1.0, 12);
      run(ms, dev.is_gpu() ? x->to_device(dev) : x, "MSequential Heterogeneous Branches Example");
    }
    std::cout << "=== Example Complete ===" << std::endl;
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "[msequential_example] error: " << e.what() << std::endl;
    return 1;
  }
}
