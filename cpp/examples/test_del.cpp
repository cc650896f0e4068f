// test_del: exercise the double-ended workspace allocator (reference: examples/test_del.cpp:8-35): allocate from both
// ends, flip, free, and print the slab statistics.
#include <iostream>

#include "tnn/allocator.hpp"
#include "tnn/device.hpp"
#include "tnn/tensor.hpp"

using namespace tnn;

int main() {
  try {
    auto& dm = DeviceManager::getInstance();
    const Device dev = dm.has_gpu() ? dm.getGPU(0) : dm.getHost();
    auto del = DELAllocatorV2::create(dev);
    del->reserve((size_t)8 << 20);
    std::cout << "[test_del] device " << dev.name() << ", slabs " << del->num_slabs() << std::endl;
    Tensor a = make_tensor(*del, DType_t::FP32, {256, 256});          // front end
    del->flip();
    Tensor b = make_tensor(*del, DType_t::FP32, {128, 128});          // back end
    del->flip();
    Tensor c = make_tensor(*del, DType_t::FP32, {64, 64});            // front again
    std::cout << "a " << a->data() << " (" << a->nbytes() << " B)\n" << "b " << b->data() << " (" << b->nbytes() << " B)\n" << "c " << c->data() << " ("
              << c->nbytes() << " B)" << std::endl;
    const bool front_then_back = static_cast<char*>(b->data()) > static_cast<char*>(a->data()) && static_cast<char*>(c->data()) < static_cast<char*>(b->data());
    a->fill(1.0);
    b->fill(2.0);
    c->fill(3.0);
    std::cout << "sums " << a->sum() << ' ' << b->sum() << ' ' << c->sum() << "; double-ended placement " << (front_then_back ? "ok" : "UNEXPECTED") << std::endl;
    a.reset();
    c.reset();
    Tensor d = make_tensor(*del, DType_t::FP32, {256, 256});          // reuses the front of the slab
    std::cout << "d " << d->data() << " after freeing a, c; slabs " << del->num_slabs() << std::endl;
    return front_then_back ? 0 : 1;
  } catch (const std::exception& e) {
    std::cerr << "[test_del] error: " << e.what() << std::endl;
    return 1;
  }
}
