// Device / Flow / Event / DeviceManager: the execution resources of the host runtime
// (reference: include/device/device.hpp:19-54, device_manager.hpp:16-83, flow.hpp:18-72, task.hpp:28-145).
// A Flow IS a cudaStream_t (non-blocking, never the legacy default stream, so every Flow can be captured into a
// CUDA graph); work is launched directly on it -- there is no Task object between a layer and its kernel launch.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace tnn {

#define TNN_CUDA(x)                                                                                          \
  do {                                                                                                       \
    cudaError_t e__ = (x);                                                                                   \
    if (e__ != cudaSuccess)                                                                                  \
      throw std::runtime_error(std::string("CUDA error ") + cudaGetErrorString(e__) + " at " + __FILE__ + ":" + \
                               std::to_string(__LINE__) + " (" #x ")");                                      \
  } while (0)

enum class DeviceType { CPU, GPU, NULL_DEVICE };

class Device {
 public:
  Device() = default;
  Device(DeviceType t, int id) : type_(t), id_(id) {}
  DeviceType device_type() const { return type_; }
  int id() const { return id_; }
  bool is_gpu() const { return type_ == DeviceType::GPU; }
  bool is_cpu() const { return type_ == DeviceType::CPU; }
  std::string name() const { return type_ == DeviceType::GPU ? "GPU:" + std::to_string(id_) : type_ == DeviceType::CPU ? "CPU:0" : "NULL"; }
  void set_current() const {
    if (is_gpu()) TNN_CUDA(cudaSetDevice(id_));
  }
  size_t total_memory() const;
  size_t free_memory() const;
  bool operator==(const Device& o) const { return type_ == o.type_ && id_ == o.id_; }
  bool operator!=(const Device& o) const { return !(*this == o); }

 private:
  DeviceType type_ = DeviceType::NULL_DEVICE;
  int id_ = 0;
};

class Event {
 public:
  explicit Event(bool timing = false);
  ~Event();
  Event(const Event&) = delete;
  Event& operator=(const Event&) = delete;
  cudaEvent_t get() const { return ev_; }
  void synchronize() const;
  static float elapsed_ms(const Event& a, const Event& b);

 private:
  cudaEvent_t ev_ = nullptr;
};

using flowHandle_t = size_t;
constexpr flowHandle_t defaultFlowHandle = 0;

// A stream of work on one device.  CPU flows execute inline (every operation is complete when it returns).
class Flow {
 public:
  Flow(Device dev, flowHandle_t id);
  ~Flow();
  Flow(const Flow&) = delete;
  Flow& operator=(const Flow&) = delete;
  const Device& device() const { return dev_; }
  flowHandle_t handle() const { return id_; }
  cudaStream_t stream() const { return stream_; }
  void synchronize() const;
  void record(Event& e) const;
  void wait(const Event& e) const;
  void wait(const Flow& other) const;     // everything queued on `other` so far happens before later work here

 private:
  Device dev_;
  flowHandle_t id_;
  cudaStream_t stream_ = nullptr;
};

class DeviceManager {
 public:
  static DeviceManager& getInstance();
  const Device& getHost() const { return host_; }
  const Device& getGPU(int i) const;
  int num_gpus() const { return (int)gpus_.size(); }
  bool has_gpu() const { return !gpus_.empty(); }
  // "CPU" | "CPU:0" | "GPU" | "GPU:3" | "CUDA:1"
  Device getDevice(const std::string& spec) const;
  Device getDevice(DeviceType t, int id = 0) const;
  // flows are owned by the manager: (device, handle) -> Flow
  Flow& getFlow(const Device& d, flowHandle_t h = defaultFlowHandle);
  std::vector<Device> devices() const;

 private:
  DeviceManager();
  Device host_{DeviceType::CPU, 0};
  std::vector<Device> gpus_;
  std::mutex mu_;
  std::vector<std::unique_ptr<Flow>> flows_;
};

inline const Device& getHost() { return DeviceManager::getInstance().getHost(); }
inline const Device& getGPU(int i = 0) { return DeviceManager::getInstance().getGPU(i); }

}  // namespace tnn
