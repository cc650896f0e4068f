// device_manager_example: enumerate devices, create flows, move tensors between them (reference:
// examples/device_manager_example.cpp:10-91).
#include <iostream>

#include "tnn/device.hpp"
#include "tnn/tensor.hpp"

using namespace tnn;

int main() {
  try {
    auto& dm = DeviceManager::getInstance();
    std::cout << "devices:" << std::endl;
    for (const Device& d : dm.devices()) {
      std::cout << "  " << d.name() << "  total " << (double)d.total_memory() / (1 << 30) << " GiB, free " << (double)d.free_memory() / (1 << 30) << " GiB" << std::endl;
    }
    Tensor h = make_tensor(DType_t::FP32, {1024, 1024});
    h->fill_random_uniform(-1.0, 1.0, 7);
    std::cout << "host tensor " << shape_to_string(h->shape()) << " sum " << h->sum() << std::endl;
    if (!dm.has_gpu()) {
      std::cout << "no GPU visible: host-only run" << std::endl;
      return 0;
    }
    const Device gpu = dm.getGPU(0);
    Flow& f0 = dm.getFlow(gpu, 0);
    Flow& f1 = dm.getFlow(gpu, 1);
    Tensor d = h->to_device(gpu);
    d->mul_scalar(2.0);
    Event e(true);
    f0.record(e);
    f1.wait(e);                                   // cross-flow dependency
    f1.synchronize();
    Tensor back = d->to_host();
    std::cout << gpu.name() << " round trip: sum " << back->sum() << " (expected " << 2.0 * h->sum() << "), flows " << f0.handle() << " and " << f1.handle() << std::endl;
    for (int i = 1; i < dm.num_gpus(); ++i) {
      Tensor p = d->to_device(dm.getGPU(i));      // peer copy
      std::cout << "  copy to " << dm.getGPU(i).name() << ": sum " << p->sum() << std::endl;
    }
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "[device_manager_example] error: " << e.what() << std::endl;
    return 1;
  }
}
