// tnn_tests: unit tests of the C++ host (reference strategy, SURVEY 4: gtest executables per area -- tensor_test,
// archiver_test, device_manager_test, layer tests with CPU-vs-GPU parity, partitioner_test ...).  No network deps, so
// the harness is 30 lines of macros instead of GoogleTest.  GPU cases skip when no device is visible.
//   tnn_tests [filter-substring]
#include <atomic>
#include <cmath>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <sstream>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include "runtime/archiver.hpp"
#include "tnn/common.hpp"
#include "tnn/data_parallel.hpp"
#include "tnn/distributed.hpp"
#include "tnn/example_models.hpp"
#include "tnn/matrix.hpp"
#include "tnn/hardware_info.hpp"
#include "tnn/inference.hpp"
#include "tnn/legacy.hpp"
#include "tnn/stepstate.hpp"
#include "tnn/train.hpp"
#include "tnn/utils.hpp"

using namespace tnn;

struct TestCase {
  const char* name;
  std::function<void()> fn;
};
static std::vector<TestCase>& registry() {
  static std::vector<TestCase> r;
  return r;
}
struct Registrar {
  Registrar(const char* n, std::function<void()> f) { registry().push_back({n, std::move(f)}); }
};
struct Skip : std::exception {
  std::string why;
  explicit Skip(std::string w) : why(std::move(w)) {}
};
#define TEST(name)                                    \
  static void test_##name();                          \
  static Registrar reg_##name(#name, test_##name);    \
  static void test_##name()
#define CHECK(cond)                                                                                      \
  do {                                                                                                   \
    if (!(cond)) throw std::runtime_error(std::string("CHECK failed: " #cond " at line ") + std::to_string(__LINE__)); \
  } while (0)
#define CHECK_NEAR(a, b, tol)                                                                                              \
  do {                                                                                                                     \
    const double a__ = (a), b__ = (b);                                                                                     \
    if (!(std::fabs(a__ - b__) <= (tol)))                                                                                  \
      throw std::runtime_error("CHECK_NEAR failed: " #a " = " + std::to_string(a__) + ", " #b " = " + std::to_string(b__) + \
                               " (tol " + std::to_string((double)(tol)) + ") at line " + std::to_string(__LINE__));         \
  } while (0)
static Device need_gpu() {
  if (!DeviceManager::getInstance().has_gpu()) throw Skip("no GPU");
  return getGPU(0);
}

// ------------------------------------------------------------------------------------- json / config / dtype
TEST(json_round_trip) {
  const std::string s = R"({"a": 1, "b": [1.5, true, null, "x\ny"], "c": {"d": -7, "e": 1e-05}})";
  Json j = Json::parse(s);
  CHECK(j.at("a").as_int() == 1 && j.at("b")[0].as_double() == 1.5 && j.at("b")[1].as_bool() && j.at("b")[2].is_null());
  CHECK(j.at("b")[3].as_string() == "x\ny" && j.at("c").at("d").as_int() == -7);
  CHECK(Json::parse(j.dump(4)) == j && Json::parse(j.dump()) == j);
  size_t pos = 0;
  const std::string two = s + "BINARY";
  Json::parse(two, &pos);
  CHECK(two.substr(pos) == "BINARY");                  // checkpoint header parsing stops at the end of the value
}
TEST(tconfig_reference_schema) {
  TConfig c("fc", "dense");
  c.set("input_features", (size_t)512).set("use_bias", true).set("epsilon", 1e-5f).set("activation", "relu").set("dims", std::vector<size_t>{1, 2, 3});
  const Json j = c.to_json();
  CHECK(j.at("parameters").at("input_features").at("type").as_string() == "size_t");
  CHECK(j.at("parameters").at("dims").at("type").as_string() == "vector<size_t>");
  TConfig seq("m", "sequential");
  seq.set("layers", std::vector<TConfig>{c, c});
  CHECK(seq.to_json().at("parameters").at("layers").at("type").as_string() == "json");
  TConfig back = TConfig::from_string(seq.dump());
  CHECK(back == seq && back.get_config_array("layers")[1].get<size_t>("input_features") == 512);
  CHECK(back.get_config_array("layers")[0].get<std::vector<size_t>>("dims").size() == 3);
  CHECK(back.get<int>("missing", 7) == 7);
}
TEST(dtype_helpers) {
  CHECK(get_dtype_size(DType_t::BF16) == 2 && get_dtype_size(DType_t::FP64) == 8 && string_to_dtype("float") == DType_t::FP32);
  CHECK(dtype_eps(DType_t::BF16) == 1e-2 && dtype_eps(DType_t::FP16) == 1e-3);
  CHECK_NEAR(bf16_to_float(float_to_bf16(3.14159f)), 3.140625, 1e-6);
  CHECK_NEAR(fp16_to_float(float_to_fp16(0.1f)), 0.0999755859375, 1e-9);
  CHECK(fp16_to_float(float_to_fp16(70000.f)) == INFINITY && fp16_to_float(float_to_fp16(1e-7f)) > 0);
  CHECK(bf16_to_float(float_to_bf16(1.00390625f)) == 1.0f);     // round to nearest even at the tie
}

// ------------------------------------------------------------------------------------- archiver (wire format)
TEST(archiver_byte_layout) {
  Writer w;
  w.prim<uint16_t>(0x1234);
  w.str("hello");
  std::vector<uint64_t> v{1, 2, 3};
  w.prim<uint64_t>(v.size());
  for (auto e : v) w.prim<uint64_t>(e);
  CHECK(w.size() == 2 + 8 + 5 + 8 + 24);
  Reader r(w.buffer().data(), w.size(), false);
  CHECK(r.prim<uint16_t>() == 0x1234 && r.str() == "hello" && r.prim<uint64_t>() == 3);
  PacketHeader h;
  h.endianness = 1;            // little-endian sender
  h.packet_length = 77;
  h.msg_length = 1234567;
  h.msg_serial_id = 42;
  h.packet_offset = 3;
  h.total_packets = 9;
  Writer hw;
  h.write(hw);
  CHECK(hw.size() == PacketHeader::kBytes);
  PacketHeader g = PacketHeader::read(hw.buffer().data(), hw.size());
  CHECK(g.msg_length == 1234567 && g.total_packets == 9 && g.packet_offset == 3 && g.msg_serial_id == 42);
  bool threw = false;
  try {
    Reader bad(w.buffer().data(), 4, false);
    bad.prim<uint64_t>();
  } catch (const std::out_of_range&) { threw = true; }
  CHECK(threw);
}

// ------------------------------------------------------------------------------------- common value types
TEST(logger_level_names) {
  CHECK(Logger::level_from_string("debug", LogLevel::INFO) == LogLevel::DEBUG && Logger::level_from_string("WARNING", LogLevel::INFO) == LogLevel::WARN);
  CHECK(Logger::level_from_string(nullptr, LogLevel::ERROR) == LogLevel::ERROR && Logger::level_from_string("loud", LogLevel::INFO) == LogLevel::INFO);
}
TEST(common_variant_blob_endian_dispatch) {
  using V = Variant<std::monostate, bool, std::string, int64_t>;
  static_assert(V::index_of<std::string>() == 2 && V::size() == 4, "index_of follows the declaration order (the wire index)");
  V v;
  CHECK(v.index() == 0 && v.is<std::monostate>());
  v = std::string("payload");
  CHECK(v.index() == 2 && v.is<std::string>() && v.get<std::string>() == "payload" && v.get_if<bool>() == nullptr);
  bool threw = false;
  try {
    v.get<int64_t>();
  } catch (const std::runtime_error&) {
    threw = true;
  }
  CHECK(threw);
  v.set_index(3);                                       // read side: index first, then the value is filled in
  CHECK(v.is<int64_t>() && v.get<int64_t>() == 0);
  v.get<int64_t>() = 41;
  CHECK(v.visit([](auto& x) -> int64_t {
    if constexpr (std::is_same_v<std::decay_t<decltype(x)>, int64_t>) return x + 1;
    else return -1;
  }) == 42);
  V w = v;
  CHECK(w == v);
  threw = false;
  try {
    v.set_index(9);
  } catch (const std::out_of_range&) {
    threw = true;
  }
  CHECK(threw);

  float data[4] = {1, 2, 3, 4};
  auto blob = make_blob(data, 4);
  CHECK(blob.count == 4 && blob.bytes() == 16 && !blob.device.is_gpu() && blob.ptr[2] == 3.f);

  CHECK(bswap<uint32_t>(0x11223344u) == 0x44332211u && bswap<uint16_t>(0xABCD) == 0xCDAB && bswap<uint8_t>(7) == 7);
  uint32_t words[2] = {0x01020304u, 0xA0B0C0D0u};
  bswap_buffer(words, 2, 4);
  CHECK(words[0] == 0x04030201u && words[1] == 0xD0C0B0A0u);
  CHECK(host_endianness() == (host_is_little_endian() ? Endianness::LITTLE : Endianness::BIG));

  size_t sz = 0;
  TNN_DISPATCH_DTYPE(DType_t::BF16, T, sz = sizeof(T));
  CHECK(sz == 2);
  TNN_DISPATCH_ANY_DTYPE(DType_t::INT64_T, T, sz = sizeof(T));
  CHECK(sz == 8);
  CHECK(dispatch_float_dtype(DType_t::FP64, [](auto t) { return sizeof(t); }) == 8);
  threw = false;
  try {
    dispatch_float_dtype(DType_t::INT32_T, [](auto) { return 0; });
  } catch (const std::runtime_error&) {
    threw = true;
  }
  CHECK(threw);
  CHECK(string_to_profiler_type("cumulative") == ProfilerType::CUMULATIVE && std::string(profiler_type_to_string(ProfilerType::NORMAL)) == "NORMAL");
  LoadTracker lt;
  lt.avg_forward_time_ = 1.5f;
  lt.max_memory_usage_ = 123.f;
  const LoadTracker back = LoadTracker::from_json(Json::parse(lt.to_json().dump()));
  CHECK(back.avg_forward_time_ == 1.5f && back.max_memory_usage_ == 123.f && back.avg_cpu_utilization_ == -1.f);
}
TEST(trainer_layer_profiler_and_memory_table) {
  // TrainingConfig::profiler_type / print_layer_profiling / print_layer_memory_usage
  ExampleModels::register_defaults();
  Graph g = load_or_create_model("mnist_cnn", "", getHost(), 11);
  auto opt = OptimizerFactory::create_sgd(0.01, 0.9);
  auto crit = LossFactory::create_logsoftmax_crossentropy();
  SyntheticDataLoader train({28, 28, 1}, 10, 64, 5);
  for (const char* mode : {"NORMAL", "CUMULATIVE"}) {
    TrainingConfig cfg;
    cfg.epochs = 1;
    cfg.batch_size = 16;
    cfg.progress_print_interval = 2;
    cfg.profiler_type = mode;
    cfg.print_layer_profiling = true;
    cfg.print_layer_memory_usage = true;
    cfg.snapshot_dir = "/tmp/tnn_tests_snap";
    std::vector<std::string> lines;
    train_model(g, train, nullptr, *opt, *crit, nullptr, cfg, [&](const std::string& s) { lines.push_back(s); });
    size_t tables = 0, mem = 0;
    for (const auto& l : lines) {
      if (l.find("layer profile") != std::string::npos) {
        ++tables;
        CHECK(l.find("fwd ms") != std::string::npos && l.find("sum") != std::string::npos);
      }
      if (l.rfind("layer memory usage:", 0) == 0) {
        ++mem;
        CHECK(l.find("activation workspace") != std::string::npos && l.find("parameters") != std::string::npos);
      }
    }
    CHECK(mem == 2);
    CHECK(tables == (std::string(mode) == "NORMAL" ? 2u : 1u));     // 4 steps, interval 2 -> two interval tables; one cumulative table
    CHECK(!LayerProfiler::instance().enabled());                    // switched off again when training returns
  }
  // a cumulative table has every top-level layer with 4 forward and 4 backward calls
  const auto rows = LayerProfiler::instance().rows();
  CHECK(!rows.empty());
  for (const auto& r : rows) CHECK(r.fwd_calls == 4 && r.bwd_calls == 4 && r.fwd_ns > 0);
}

// ------------------------------------------------------------------------------------- allocators
TEST(offset_free_list_best_fit_and_coalescing) {
  OffsetFreeList fl(1 << 20, 256);
  const uint64_t a = fl.allocate(1000), b = fl.allocate(5000), c = fl.allocate(300), d = fl.allocate(70000);
  CHECK(a == 0 && b == 1024 && c == 1024 + 5120 && d == c + 512);
  CHECK(fl.block_size(a) == 1024 && fl.block_size(b) == 5120 && fl.in_use() == 1024 + 5120 + 512 + 70144);
  CHECK(fl.release(b) && !fl.release(b) && !fl.release(12345));       // double free / foreign offset are refused
  // two holes now: [1024, 6144) and the tail; best fit puts a 4 KB request into the small hole, not the tail
  CHECK(fl.num_free_ranges() == 2);
  const uint64_t e = fl.allocate(4096);
  CHECK(e == 1024);
  CHECK(fl.release(e) && fl.release(a));                              // a + hole coalesce into [0, 6144)
  CHECK(fl.num_free_ranges() == 2 && fl.allocate(6144) == 0);
  CHECK(fl.release(0) && fl.release(c) && fl.release(d));
  CHECK(fl.num_free_ranges() == 1 && fl.largest_free() == (1u << 20) && fl.in_use() == 0 && fl.num_live() == 0);
  CHECK(fl.allocate((1 << 20) + 1) == OffsetFreeList::kNone);
  CHECK(fl.peak() >= 70144);
}
TEST(fd_mailbox_passes_descriptors_between_processes) {
  // the transport of the POSIX-fd shareable handles (cuMemExportToShareableHandle): SCM_RIGHTS over Unix datagram sockets
  char tmpl[] = "/tmp/tnn_fdbox_XXXXXX";
  CHECK(mkdtemp(tmpl) != nullptr);
  const std::string dir = tmpl;
  const pid_t child = fork();
  if (child == 0) {                                     // child: receives a descriptor, writes through it, answers with a tag
    int rc = 1;
    try {
      FdMailbox box(dir, "b.sock");
      int fd = -1;
      const uint64_t tag = box.recv(&fd, 20000);
      if (fd >= 0 && tag == 0x1234 && write(fd, "through the passed fd", 21) == 21) rc = 0;
      if (fd >= 0) close(fd);
      box.send(dir + "/a.sock", 0x5678 + (uint64_t)rc, -1, 20000);
    } catch (...) {
      rc = 2;
    }
    _exit(rc);
  }
  {
    FdMailbox box(dir, "a.sock");
    const std::string file = dir + "/payload";
    const int fd = open(file.c_str(), O_CREAT | O_RDWR, 0600);
    CHECK(fd >= 0);
    box.send(dir + "/b.sock", 0x1234, fd, 20000);       // retries until the child has bound its mailbox
    close(fd);
    int none = 7;
    CHECK(box.recv(&none, 20000) == 0x5678 && none == -1);
    std::ifstream f(file);
    std::string got((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    CHECK(got == "through the passed fd");
    unlink(file.c_str());
  }
  int st = 0;
  waitpid(child, &st, 0);
  CHECK(WIFEXITED(st) && WEXITSTATUS(st) == 0);
  rmdir(dir.c_str());
  bool timed_out = false;
  try {
    FdMailbox lonely("/tmp", "tnn_fdbox_lonely.sock");
    int fd;
    lonely.recv(&fd, 50);
  } catch (const std::runtime_error&) {
    timed_out = true;
  }
  CHECK(timed_out);
}
// ---- a memfd-backed stand-in for the CUDA VMM / multicast driver calls: "physical allocations" are memfds, exporting one
// dup()s its descriptor, mapping is mmap(MAP_SHARED) into a reserved range.  With it the whole SymmetricMemory group protocol
// (sizes, mappings, symmetric offsets, descriptor exchange between processes, phase ordering, teardown) runs on a CPU box;
// what stays unverified without a GPU is only that the real driver accepts the same call sequence.
namespace fake_vmm {
struct Handle {
  int fd = -1;
  size_t size = 0;
  bool multicast = false;
  int devices_added = 0, binds = 0;
};
static std::mutex mu;
static std::map<CUmemGenericAllocationHandle, Handle> handles;
static std::map<CUdeviceptr, size_t> reserved;
static uint64_t next_handle = 1;
static int access_calls = 0;
static bool refuse_multicast = false;                   // the object cannot be created although the devices advertise support
static CUmemGenericAllocationHandle add(int fd, size_t size, bool mc) {
  std::lock_guard<std::mutex> g(mu);
  Handle h;
  h.fd = fd;
  h.size = size;
  h.multicast = mc;
  handles[next_handle] = h;
  return next_handle++;
}
static int new_memfd(size_t size) {
  const int fd = memfd_create("tnn_fake_vmm", MFD_CLOEXEC);
  if (fd >= 0 && ftruncate(fd, (off_t)size) != 0) {
    close(fd);
    return -1;
  }
  return fd;
}
static const VmmDriver& table(int device_count) {
  static VmmDriver d;
  d.fake_device_count = device_count;
  d.GetErrorString = [](CUresult, const char** s) -> CUresult {
    *s = "fake driver error";
    return CUDA_SUCCESS;
  };
  d.DeviceGet = [](CUdevice* dev, int ordinal) -> CUresult {
    *dev = ordinal;
    return CUDA_SUCCESS;
  };
  d.DeviceGetAttribute = [](int* v, CUdevice_attribute, CUdevice) -> CUresult {
    *v = 1;
    return CUDA_SUCCESS;
  };
  d.MemGetAllocationGranularity = [](size_t* g, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) -> CUresult {
    *g = 64 << 10;
    return CUDA_SUCCESS;
  };
  d.MulticastGetGranularity = [](size_t* g, const CUmulticastObjectProp*, CUmulticastGranularity_flags) -> CUresult {
    *g = 128 << 10;
    return CUDA_SUCCESS;
  };
  d.MemCreate = [](CUmemGenericAllocationHandle* h, size_t size, const CUmemAllocationProp* p, unsigned long long) -> CUresult {
    if (p->requestedHandleTypes != CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR || size % (64 << 10)) return CUDA_ERROR_INVALID_VALUE;
    const int fd = new_memfd(size);
    if (fd < 0) return CUDA_ERROR_OUT_OF_MEMORY;
    *h = add(fd, size, false);
    return CUDA_SUCCESS;
  };
  d.MulticastCreate = [](CUmemGenericAllocationHandle* h, const CUmulticastObjectProp* p) -> CUresult {
    if (p->numDevices < 2 || p->size % (128 << 10)) return CUDA_ERROR_INVALID_VALUE;
    if (refuse_multicast) return CUDA_ERROR_NOT_SUPPORTED;
    const int fd = new_memfd(p->size);
    if (fd < 0) return CUDA_ERROR_OUT_OF_MEMORY;
    *h = add(fd, p->size, true);
    return CUDA_SUCCESS;
  };
  d.MemRelease = [](CUmemGenericAllocationHandle h) -> CUresult {
    std::lock_guard<std::mutex> g(mu);
    auto it = handles.find(h);
    if (it == handles.end()) return CUDA_ERROR_INVALID_VALUE;
    close(it->second.fd);
    handles.erase(it);
    return CUDA_SUCCESS;
  };
  d.MemExportToShareableHandle = [](void* out, CUmemGenericAllocationHandle h, CUmemAllocationHandleType, unsigned long long) -> CUresult {
    std::lock_guard<std::mutex> g(mu);
    auto it = handles.find(h);
    if (it == handles.end()) return CUDA_ERROR_INVALID_VALUE;
    *static_cast<int*>(out) = dup(it->second.fd);
    return CUDA_SUCCESS;
  };
  d.MemImportFromShareableHandle = [](CUmemGenericAllocationHandle* h, void* os_handle, CUmemAllocationHandleType) -> CUresult {
    const int fd = dup((int)(uintptr_t)os_handle);      // the caller closes its descriptor afterwards, like with the real driver
    struct stat st {};
    if (fd < 0 || fstat(fd, &st) != 0) return CUDA_ERROR_INVALID_VALUE;
    *h = add(fd, (size_t)st.st_size, false);
    return CUDA_SUCCESS;
  };
  d.MemAddressReserve = [](CUdeviceptr* ptr, size_t size, size_t, CUdeviceptr, unsigned long long) -> CUresult {
    void* p = mmap(nullptr, size, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return CUDA_ERROR_OUT_OF_MEMORY;
    std::lock_guard<std::mutex> g(mu);
    reserved[(CUdeviceptr)(uintptr_t)p] = size;
    *ptr = (CUdeviceptr)(uintptr_t)p;
    return CUDA_SUCCESS;
  };
  d.MemAddressFree = [](CUdeviceptr ptr, size_t size) -> CUresult {
    std::lock_guard<std::mutex> g(mu);
    if (!reserved.erase(ptr)) return CUDA_ERROR_INVALID_VALUE;
    munmap(reinterpret_cast<void*>((uintptr_t)ptr), size);
    return CUDA_SUCCESS;
  };
  d.MemMap = [](CUdeviceptr ptr, size_t size, size_t offset, CUmemGenericAllocationHandle h, unsigned long long) -> CUresult {
    std::lock_guard<std::mutex> g(mu);
    auto it = handles.find(h);
    if (it == handles.end() || offset != 0 || size > it->second.size || !reserved.count(ptr)) return CUDA_ERROR_INVALID_VALUE;
    void* p = mmap(reinterpret_cast<void*>((uintptr_t)ptr), size, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, it->second.fd, 0);
    return p == MAP_FAILED ? CUDA_ERROR_OUT_OF_MEMORY : CUDA_SUCCESS;
  };
  d.MemUnmap = [](CUdeviceptr ptr, size_t size) -> CUresult {
    void* p = mmap(reinterpret_cast<void*>((uintptr_t)ptr), size, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_FIXED, -1, 0);
    return p == MAP_FAILED ? CUDA_ERROR_INVALID_VALUE : CUDA_SUCCESS;
  };
  d.MemSetAccess = [](CUdeviceptr, size_t, const CUmemAccessDesc* desc, size_t count) -> CUresult {
    if (!desc || count == 0) return CUDA_ERROR_INVALID_VALUE;
    ++access_calls;
    return CUDA_SUCCESS;
  };
  d.MulticastAddDevice = [](CUmemGenericAllocationHandle h, CUdevice) -> CUresult {
    std::lock_guard<std::mutex> g(mu);
    auto it = handles.find(h);
    if (it == handles.end()) return CUDA_ERROR_INVALID_VALUE;
    ++it->second.devices_added;
    return CUDA_SUCCESS;
  };
  d.MulticastBindMem = [](CUmemGenericAllocationHandle mc, size_t mc_off, CUmemGenericAllocationHandle mem, size_t mem_off, size_t size, unsigned long long) -> CUresult {
    std::lock_guard<std::mutex> g(mu);
    auto a = handles.find(mc), b = handles.find(mem);
    if (a == handles.end() || b == handles.end() || mc_off || mem_off || size > a->second.size || size > b->second.size) return CUDA_ERROR_INVALID_VALUE;
    if (a->second.devices_added == 0) return CUDA_ERROR_INVALID_VALUE;         // bind before add
    ++a->second.binds;
    return CUDA_SUCCESS;
  };
  d.MulticastUnbind = [](CUmemGenericAllocationHandle, CUdevice, size_t, size_t) -> CUresult { return CUDA_SUCCESS; };
  return d;
}
struct Install {
  explicit Install(int devices) { VmmDriver::install_for_testing(&table(devices)); }
  ~Install() { VmmDriver::install_for_testing(nullptr); }
};
}  // namespace fake_vmm

TEST(symmetric_memory_group_in_one_process_over_a_fake_driver) {
  if (DeviceManager::getInstance().has_gpu()) throw Skip("GPU present: the real driver is used");
  fake_vmm::Install fake(4);
  {
    const auto sup = SymmetricMemory::probe({0, 1, 2});
    CHECK(sup.vmm && sup.multicast);
    CHECK(!SymmetricMemory::probe({0}).multicast && !SymmetricMemory::probe({0, 7}).vmm);
    auto group = std::make_shared<SymmetricMemory>(std::vector<int>{0, 1, 2}, (size_t)300 << 10, true);
    CHECK(group->world() == 3 && group->my_rank() == -1 && group->granularity() == (128u << 10) && group->bytes() == (384u << 10));
    CHECK(group->has_multicast() && fake_vmm::handles.size() == 4);       // three slabs + the multicast object
    for (const auto& kv : fake_vmm::handles)
      if (kv.second.multicast) CHECK(kv.second.devices_added == 3 && kv.second.binds == 3);
    std::vector<std::unique_ptr<SymmetricAllocator>> al;
    std::vector<int*> blk, blk2;
    for (int r = 0; r < 3; ++r) {
      al.push_back(std::make_unique<SymmetricAllocator>(group, r));
      CHECK(al.back()->device() == Device(DeviceType::GPU, r));
      blk.push_back(static_cast<int*>(al.back()->allocate(1000)));
      blk2.push_back(static_cast<int*>(al.back()->allocate(5000)));
    }
    for (int r = 0; r < 3; ++r) {
      CHECK(group->offset_of(r, blk[(size_t)r]) == 0 && group->offset_of(r, blk2[(size_t)r]) == 1024);     // symmetric offsets
      blk2[(size_t)r][7] = 100 + r;
    }
    // rank 0's view of rank 2's block is rank 2's block
    CHECK(static_cast<int*>(al[0]->peer(blk2[0], 2)) == blk2[2] && static_cast<int*>(al[0]->peer(blk2[0], 2))[7] == 102);
    CHECK(al[1]->multicast(blk2[1]) == static_cast<char*>(group->multicast_base()) + 1024);
    bool threw = false;
    try {
      group->offset_of(0, blk[1]);                        // another rank's pointer
    } catch (const std::out_of_range&) {
      threw = true;
    }
    CHECK(threw);
    threw = false;
    try {
      al[0]->allocate(group->bytes());                    // larger than what is left
    } catch (const std::bad_alloc&) {
      threw = true;
    }
    CHECK(threw);
    CHECK(al[0]->bytes_in_use() == 1024 + 5120 && al[0]->bytes_reserved() == group->bytes());
    const int fd = group->export_fd(1);
    CHECK(fd >= 0);
    close(fd);
    for (int r = 0; r < 3; ++r) {
      al[(size_t)r]->deallocate(blk[(size_t)r]);
      al[(size_t)r]->deallocate(blk2[(size_t)r]);
    }
    CHECK(group->free_list(2).in_use() == 0 && group->free_list(2).num_free_ranges() == 1);
    // a model's parameter slabs inside the group (what DataParallel does with TNN_DP_NVLS=1)
    LayerBuilder lb({16});
    lb.dense(8, true, "d");
    CHECK(GraphContext::slab_bytes(lb.build_sequential("m")->all_layers(), true) < group->bytes());
  }
  CHECK(fake_vmm::handles.empty() && fake_vmm::reserved.empty());         // teardown released every handle and range
  {
    // the window is refused at creation: the group still forms, on peer mappings alone
    fake_vmm::refuse_multicast = true;
    SymmetricMemory group({0, 1}, (size_t)64 << 10, true);
    fake_vmm::refuse_multicast = false;
    CHECK(!group.has_multicast() && group.multicast_base() == nullptr && group.multicast_error().find("cuMulticastCreate") != std::string::npos);
    int* a = static_cast<int*>(group.allocate(0, 64));
    CHECK(group.multicast(0, a) == nullptr && group.peer(0, a, 1) == group.base(1) && fake_vmm::handles.size() == 2);
  }
  CHECK(fake_vmm::handles.empty() && fake_vmm::reserved.empty());
}
TEST(symmetric_memory_rendezvous_between_processes_over_a_fake_driver) {
  if (DeviceManager::getInstance().has_gpu()) throw Skip("GPU present: the real driver is used");
  // three processes, one rank each: physical handles and the multicast handle cross as SCM_RIGHTS descriptors; because the
  // fake "allocations" are memfds the slabs really are shared, so every rank can check what its peers wrote
  char tmpl[] = "/tmp/tnn_symm_XXXXXX";
  CHECK(mkdtemp(tmpl) != nullptr);
  const std::string dir = tmpl;
  const int W = 3;
  fake_vmm::Install fake(1);
  std::vector<pid_t> kids;
  for (int r = 0; r < W; ++r) {
    const pid_t pid = fork();
    if (pid == 0) {
      int rc = 1;
      try {
        if (r == 1) usleep(200 * 1000);                 // a late rank: the others retry their sends / hold early datagrams
        auto m = SymmetricMemory::rendezvous(r, W, /*device=*/0, (size_t)100 << 10, dir, true, 20000);
        rc = 2;
        if (m->world() == W && m->my_rank() == r && m->has_multicast() && m->bytes() == (128u << 10)) {
          int* mine = static_cast<int*>(m->allocate(r, 4096));
          rc = 3;
          if (m->offset_of(r, mine) == 0) {
            mine[0] = 1000 + r;
            __sync_synchronize();
            FdMailbox box(dir, "t." + std::to_string(r));
            for (int q = 0; q < W; ++q)
              if (q != r) box.send(dir + "/t." + std::to_string(q), (uint64_t)r, -1, 20000);
            for (int i = 0; i < W - 1; ++i) {
              int f = -1;
              box.recv(&f, 20000);
            }
            rc = 0;
            for (int q = 0; q < W; ++q)
              if (static_cast<int*>(m->peer(r, mine, q))[0] != 1000 + q) rc = 4;          // peers' writes through MY mappings
            if (m->multicast(r, mine) != m->multicast_base()) rc = 5;
            bool refused = false;
            try {
              m->allocate((r + 1) % W, 16);             // only the local rank allocates in a multi-process group
            } catch (const std::invalid_argument&) {
              refused = true;
            }
            if (!refused) rc = 6;
            // nobody unmaps before everybody has checked
            for (int q = 0; q < W; ++q)
              if (q != r) box.send(dir + "/t." + std::to_string(q), 100 + (uint64_t)r, -1, 20000);
            for (int i = 0; i < W - 1; ++i) {
              int f = -1;
              box.recv(&f, 20000);
            }
          }
        }
      } catch (const std::exception& e) {
        fprintf(stderr, "rank %d: %s\n", r, e.what());
        rc = 9;
      }
      _exit(rc);
    }
    kids.push_back(pid);
  }
  for (int r = 0; r < W; ++r) {
    int st = 0;
    waitpid(kids[(size_t)r], &st, 0);
    if (!(WIFEXITED(st) && WEXITSTATUS(st) == 0)) throw std::runtime_error("rank " + std::to_string(r) + " exited with " + std::to_string(WIFEXITED(st) ? WEXITSTATUS(st) : -1));
  }
  // every mailbox unlinked itself
  CHECK(rmdir(dir.c_str()) == 0);
}
TEST(symmetric_memory_reports_unsupported_on_a_cpu_box) {
  if (DeviceManager::getInstance().has_gpu()) throw Skip("GPU present");
  const auto sup = SymmetricMemory::probe({0, 1});
  CHECK(!sup.vmm && !sup.multicast && !sup.why.empty());
  bool threw = false;
  try {
    SymmetricMemory m({0, 1}, 1 << 20);
  } catch (const std::runtime_error&) {
    threw = true;
  }
  CHECK(threw);
}
TEST(graph_slabs_come_from_the_given_allocator) {
  // GraphBuilder::compile(allocator) (the reference signature): parameter / gradient / buffer slabs are blocks of that allocator
  DELAllocatorV2 arena(HostAllocator(), 4 << 20, 256);
  LayerBuilder lb({16});
  lb.dense(32, true, "d1").activation("relu", "r").dense(4, true, "d2");
  auto model = lb.build_sequential("tiny");
  const size_t expect = GraphContext::slab_bytes(model->all_layers(), false);
  GraphBuilder b("g");
  b.io("input");
  b.io("output");
  OpNode& n = b.add_layer(std::move(model), "model");
  b.add_edge({"input"}, {"output"}, n);
  {
    Graph g = b.compile(arena, 3);
    CHECK(arena.bytes_in_use() == expect && expect >= (16 * 32 + 32 + 32 * 4 + 4) * 4 * 2);
    CHECK(g.context().param_slab()->data() != nullptr && g.context().numel() >= 16 * 32 + 32 + 32 * 4 + 4);
  }
  CHECK(arena.bytes_in_use() == 0);                     // the slabs went back to it
}
TEST(del_allocator_double_ended) {
  DELAllocatorV2 a(HostAllocator(), 1 << 20, 256);
  void* l1 = a.allocate(1000);
  a.flip();
  void* r1 = a.allocate(1000);
  a.flip();
  void* l2 = a.allocate(5000);
  CHECK(a.num_slabs() == 1);
  CHECK((char*)l2 == (char*)l1 + 1024);                       // left end bumps upward
  CHECK((char*)r1 > (char*)l2);                               // right end comes from the top of the slab
  CHECK(((uintptr_t)l1 & 255) == 0 && ((uintptr_t)r1 & 255) == 0);
  a.deallocate(l1);                                           // interior hole on the left
  void* l3 = a.allocate(512);
  CHECK(l3 == l1);                                            // first fit reuses it
  a.deallocate(l2);
  a.deallocate(l3);
  a.deallocate(r1);
  CHECK(a.bytes_in_use() == 0);
  // same sequence again -> same addresses (what CUDA-graph capture of a step relies on)
  void* l1b = a.allocate(1000);
  a.flip();
  void* r1b = a.allocate(1000);
  CHECK(l1b == l1 && r1b == r1);
  a.deallocate(l1b);
  a.deallocate(r1b);
  a.reserve(8 << 20);
  CHECK(a.num_slabs() == 1 && a.bytes_reserved() >= (8u << 20));
  void* big = a.allocate(6 << 20);
  CHECK(big != nullptr && a.num_slabs() == 1);
  a.deallocate(big);
  a.clear();
  CHECK(a.num_slabs() == 0);
}
TEST(pool_allocator_reuse_rule) {
  PoolAllocator p(HostAllocator());
  void* a = p.allocate(10000);
  p.deallocate(a);
  void* b = p.allocate(4000);          // 10240 <= 4 * 4096: reused
  CHECK(b == a);
  p.deallocate(b);
  void* c = p.allocate(1000);          // 10240 > 4 * 1024: fresh block
  CHECK(c != a);
  p.deallocate(c);
  CHECK(p.bytes_in_use() == 0 && p.bytes_reserved() > 0);
  p.release_cached();
  CHECK(p.bytes_reserved() == 0);
}

// ------------------------------------------------------------------------------------- tensors
TEST(tensor_basics_cpu) {
  Tensor t = make_tensor<float>({2, 3, 4});
  CHECK(t->size() == 24 && t->stride(0) == 12 && t->stride(1) == 4 && t->dims() == 3);
  t->fill(2.0);
  t->at<float>({1, 2, 3}) = 5.f;
  CHECK_NEAR(t->sum(), 2.0 * 23 + 5, 1e-6);
  CHECK_NEAR(t->max(), 5.0, 0);
  Tensor u = t->clone();
  u->mul_scalar(0.5);
  t->add(*u);
  CHECK_NEAR(t->at<float>({0, 0, 0}), 3.0, 1e-6);
  Tensor v = t->span(12, {3, 4});
  CHECK_NEAR(v->at<float>({2, 3}), 7.5, 1e-6);
  v->fill(1.0);                                               // views share storage
  CHECK_NEAR(t->at<float>({1, 0, 0}), 1.0, 0);
  Tensor sum = u + u;
  CHECK_NEAR(sum->at<float>({0, 0, 0}), 2.0, 1e-6);
  t->fill_random_uniform(-1, 1, 7);
  CHECK(t->min() >= -1 && t->max() <= 1 && std::fabs(t->mean()) < 0.6);
  Tensor b = dtype_cast(*t, DType_t::BF16);
  CHECK(max_abs_diff(*b, *t) < 1e-2);
}
TEST(matmul_every_float_dtype_cpu) {
  // a [5, 7] x b [4, 7]^T and the MN-major spellings of the same product, FP32 (packed SGEMM) and FP64 / FP16 / BF16 (host loop)
  Tensor a = make_tensor<float>({5, 7}), b = make_tensor<float>({4, 7});
  a->fill_random_uniform(-1, 1, 1);
  b->fill_random_uniform(-1, 1, 2);
  Tensor at = make_tensor<float>({7, 5}), bt = make_tensor<float>({7, 4});
  for (size_t m = 0; m < 5; ++m)
    for (size_t k = 0; k < 7; ++k) at->at<float>({k, m}) = a->at<float>({m, k});
  for (size_t n = 0; n < 4; ++n)
    for (size_t k = 0; k < 7; ++k) bt->at<float>({k, n}) = b->at<float>({n, k});
  Tensor ref = matmul(*a, *b);
  double want = 0;
  for (size_t k = 0; k < 7; ++k) want += (double)a->at<float>({2, k}) * b->at<float>({3, k});
  CHECK_NEAR(ref->at<float>({2, 3}), want, 1e-5);
  CHECK(max_abs_diff(*matmul(*at, *b, true, false), *ref) < 1e-5);
  CHECK(max_abs_diff(*matmul(*a, *bt, false, true), *ref) < 1e-5);
  CHECK(max_abs_diff(*matmul(*at, *bt, true, true, 2.0), *(ref + ref)) < 1e-5);
  for (DType_t dt : {DType_t::FP64, DType_t::FP16, DType_t::BF16}) {
    Tensor y = matmul(*dtype_cast(*a, dt), *dtype_cast(*bt, dt), false, true);
    CHECK(y->data_type() == dt && max_abs_diff(*y, *ref) < (dt == DType_t::FP64 ? 1e-6 : dt == DType_t::FP16 ? 2e-2 : 1e-1));
  }
}
TEST(legacy_matrix_value_semantics) {
  // the legacy 2-D value type (include/matrix/matrix.hpp of the reference): deep copies, +, -, scalar * and /, product, reshape / resize
  Matrix<float> a(2, 3), b(2, 3);
  for (size_t r = 0; r < 2; ++r)
    for (size_t c = 0; c < 3; ++c) {
      a(r, c) = (float)(r * 3 + c);
      b(r, c) = 1.f;
    }
  Matrix<float> s = a + b, d = a - b, k = a * 2.f, h = a / 2.f, k2 = 3.f * a;
  CHECK_NEAR(s(1, 2), 6.0, 0);
  CHECK_NEAR(d(0, 0), -1.0, 0);
  CHECK_NEAR(k(1, 1), 8.0, 0);
  CHECK_NEAR(h(1, 0), 1.5, 0);
  CHECK_NEAR(k2(0, 2), 6.0, 0);
  Matrix<float> copy = a;                                     // value semantics: the copy owns its storage
  copy(0, 0) = 42.f;
  CHECK_NEAR(a(0, 0), 0.0, 0);
  a += b;
  a -= b;
  CHECK_NEAR(a(1, 2), 5.0, 0);
  Matrix<float> bt(3, 2);
  for (size_t r = 0; r < 3; ++r)
    for (size_t c = 0; c < 2; ++c) bt(r, c) = (float)(r + c);
  Matrix<float> p = a * bt;                                   // [2, 3] x [3, 2]
  CHECK(p.rows() == 2 && p.cols() == 2);
  CHECK_NEAR(p(0, 0), 0 * 0 + 1 * 1 + 2 * 2, 1e-5);
  CHECK_NEAR(p(1, 1), 3 * 1 + 4 * 2 + 5 * 3, 1e-5);
  Matrix<float> r = a.reshape(3, 2);
  CHECK(r.rows() == 3 && r(2, 1) == 5.f);
  bool threw = false;
  try { a.reshape(4, 2); } catch (const std::invalid_argument&) { threw = true; }
  CHECK(threw);
  threw = false;
  try { a + bt; } catch (const std::invalid_argument&) { threw = true; }
  CHECK(threw);
  threw = false;
  try { a / 0.f; } catch (const std::invalid_argument&) { threw = true; }
  CHECK(threw);
  a.resize(4, 4);
  CHECK(a.rows() == 4 && a.size() == 16 && a(3, 3) == 0.f);
  Matrix<float> e;
  CHECK(e.rows() == 0 && e.size() == 0 && (e + e).size() == 0);
  Matrix<double> dd(2, 2);
  dd.fill(0.5);
  Matrix<double> pd = dd * dd;
  CHECK_NEAR(pd(1, 1), 0.5, 1e-12);
  Matrix<float> u(8, 8);
  u.fill_random_uniform(0.25f, 7);
  CHECK(u.tensor()->min() >= -0.25 && u.tensor()->max() <= 0.25);
}
TEST(tensor_save_load_record) {
  Tensor t = make_tensor<float>({3, 5});
  t->fill_random_normal(0, 1, 3);
  std::stringstream ss;
  t->save(ss);
  CHECK(ss.str().size() == 4 + 8 + 16 + 60);                  // {u32 dtype, u64 ndims, u64 dims[], raw}
  uint32_t dt;
  std::memcpy(&dt, ss.str().data(), 4);
  CHECK(dt == 5);                                             // FP32
  Tensor u = load(ss, HostAllocator());
  CHECK(u->shape() == t->shape() && max_abs_diff(*t, *u) == 0);
}
TEST(tensor_gpu_round_trip) {
  const Device g = need_gpu();
  Tensor h = make_tensor<float>({1000});
  h->fill_random_uniform(-2, 2, 11);
  Tensor d = h->to_device(g);
  CHECK(d->is_gpu());
  d->mul_scalar(2.0);
  d->add_scalar(1.0);
  Tensor back = d->to_host();
  for (size_t i = 0; i < 1000; i += 97) CHECK_NEAR(back->data_as<float>()[i], 2 * h->data_as<float>()[i] + 1, 1e-5);
  CHECK_NEAR(d->sum(), 2 * h->sum() + 1000, 1e-1);
  Tensor b = dtype_cast(*d, DType_t::BF16);
  CHECK(max_abs_diff(*b, *back) < 5e-2);
}

TEST(gpu_tensor_every_float_dtype) {
  // FP16 / FP32 (tf32 tensor cores) / FP64 / BF16 on the device: casts, arithmetic, reductions and matmul (gemm_xt.cu)
  const Device g = need_gpu();
  Tensor a = make_tensor<float>({200, 136}), b = make_tensor<float>({72, 136});
  a->fill_random_uniform(-1, 1, 5);
  b->fill_random_uniform(-1, 1, 6);
  Tensor ref = matmul(*a, *b);                                 // CPU fp32
  Tensor bt = make_tensor<float>({136, 72});
  for (size_t n = 0; n < 72; ++n)
    for (size_t k = 0; k < 136; ++k) bt->at<float>({k, n}) = b->at<float>({n, k});
  Tensor ad = a->to_device(g), bd = b->to_device(g), btd = bt->to_device(g);
  for (DType_t dt : {DType_t::FP16, DType_t::FP32, DType_t::FP64, DType_t::BF16}) {
    const double tol = dt == DType_t::FP64 ? 1e-4 : dt == DType_t::FP32 ? 2e-2 : dt == DType_t::FP16 ? 5e-2 : 3e-1;
    Tensor x = dtype_cast(*ad, dt), w = dtype_cast(*bd, dt), wt = dtype_cast(*btd, dt);
    CHECK(x->is_gpu() && x->data_type() == dt);
    CHECK(max_abs_diff(*dtype_cast(*x, DType_t::FP32), *a) < (dt == DType_t::BF16 ? 1e-2 : 1e-3));      // device cast round trip
    Tensor y = matmul(*x, *w);
    CHECK(y->data_type() == dt && max_abs_diff(*y, *ref) < tol);
    CHECK(max_abs_diff(*matmul(*x, *wt, false, true), *ref) < tol);                                      // MN-major B
    Tensor s2 = x->clone();
    s2->add(*x);
    s2->mul_scalar(0.5);
    CHECK(max_abs_diff(*s2, *x) < 1e-6);
    CHECK_NEAR(x->sum(), a->sum(), dt == DType_t::FP64 || dt == DType_t::FP32 ? 5e-2 : 5.0);
  }
}

// ------------------------------------------------------------------------------------- layers: finite differences (CPU)
static double loss_of(Layer& l, const Tensor& x, const Tensor& w) {       // L = sum(w * f(x))
  Tensor y = l.forward(x, 0);
  double s = 0;
  for (size_t i = 0; i < y->size(); ++i) s += (double)y->data_as<float>()[i] * w->data_as<float>()[i];
  return s;
}
// analytic derivative vs central AND one-sided differences: at a ReLU / max-pool kink the analytic value equals one of the
// one-sided derivatives, not their average
static bool deriv_ok(double l0, double lp, double lm, double eps, double ana, double tol, double* shown) {
  const double c = (lp - lm) / (2 * eps), f = (lp - l0) / eps, b = (l0 - lm) / eps;
  *shown = c;
  for (double num : {c, f, b})
    if (std::fabs(num - ana) <= tol * (1 + std::fabs(num))) return true;
  return false;
}
static void gradcheck(std::unique_ptr<Layer> layer, const Shape& xs, double tol = 2e-2, bool check_params = true, float eps = 1e-2f) {
  Layer& l = *layer;
  l.init(getHost(), 3);
  l.set_training(true);
  Tensor x = make_tensor<float>(xs);
  x->fill_random_normal(0, 1, 5);
  Tensor y0 = l.forward(x, 0);
  Tensor w = make_tensor<float>(y0->shape());
  w->fill_random_normal(0, 1, 9);
  for (Layer* s : l.all_layers())
    for (auto& g : s->grads()) g.second->zero();
  Tensor dx = l.backward(w, 0);
  const double l0 = loss_of(l, x, w);
  l.clear_caches();
  if (dx) {
    for (size_t i = 0; i < x->size(); i += std::max<size_t>(1, x->size() / 17)) {
      const float keep = x->data_as<float>()[i];
      x->data_as<float>()[i] = keep + eps;
      const double lp = loss_of(l, x, w);
      l.clear_caches();
      x->data_as<float>()[i] = keep - eps;
      const double lm = loss_of(l, x, w);
      l.clear_caches();
      x->data_as<float>()[i] = keep;
      const double ana = dx->data_as<float>()[i];
      double num;
      if (!deriv_ok(l0, lp, lm, eps, ana, tol, &num))
        throw std::runtime_error(l.type() + ": dx[" + std::to_string(i) + "] numeric " + std::to_string(num) + " vs analytic " + std::to_string(ana));
    }
  }
  if (!check_params) return;
  for (Layer* s : l.all_layers())
    for (auto& g : s->grads()) {
      Tensor p = s->param(g.first);
      for (size_t i = 0; i < p->size(); i += std::max<size_t>(1, p->size() / 7)) {
        const float keep = p->data_as<float>()[i];
        p->data_as<float>()[i] = keep + eps;
        const double lp = loss_of(l, x, w);
        l.clear_caches();
        p->data_as<float>()[i] = keep - eps;
        const double lm = loss_of(l, x, w);
        l.clear_caches();
        p->data_as<float>()[i] = keep;
        const double ana = g.second->data_as<float>()[i];
        double num;
        if (!deriv_ok(l0, lp, lm, eps, ana, tol, &num))
          throw std::runtime_error(s->name() + "." + g.first + "[" + std::to_string(i) + "] numeric " + std::to_string(num) + " vs analytic " + std::to_string(ana));
      }
    }
}
TEST(gradcheck_dense) { gradcheck(std::make_unique<DenseLayer>(7, 5), {4, 7}); }
TEST(gradcheck_conv2d) {
  gradcheck(std::make_unique<Conv2DLayer>(3, 4, 3, 3, 1, 1, 1, 1), {2, 6, 6, 3});
  gradcheck(std::make_unique<Conv2DLayer>(2, 3, 3, 3, 2, 2, 1, 1, false), {2, 8, 8, 2});
  gradcheck(std::make_unique<Conv2DLayer>(2, 3, 3, 2, 2, 1, 1, 0), {1, 7, 6, 2});        // anisotropic kernel / stride / pad
}
TEST(gradcheck_batchnorm) {
  gradcheck(std::make_unique<BatchNormLayer>(4, 1e-5, 0.1, true, false), {3, 5, 5, 4});
  gradcheck(std::make_unique<BatchNormLayer>(4, 1e-5, 0.1, true, true), {3, 5, 5, 4}, 6e-2);
}
TEST(gradcheck_pools_acts_norms) {
  gradcheck(std::make_unique<MaxPool2DLayer>(2, 2), {2, 6, 6, 3});
  gradcheck(std::make_unique<AvgPool2DLayer>(3, 3, 2, 2, 1, 1), {2, 7, 7, 2});
  for (const char* a : {"relu", "leaky_relu", "sigmoid", "tanh", "elu", "gelu", "softmax"}) gradcheck(std::make_unique<ActivationLayer>(a), {3, 9});
  gradcheck(std::make_unique<LayerNormLayer>(8), {5, 8});
  gradcheck(std::make_unique<GroupNormLayer>(2, 6), {2, 4, 4, 6});
  gradcheck(std::make_unique<SliceLayer>(1, 1, 2), {3, 5, 4});
  gradcheck(std::make_unique<TransposeLayer>(), {2, 3, 5});
  gradcheck(std::make_unique<ClassTokenLayer>(4), {2, 3, 4});
  gradcheck(std::make_unique<PositionalEmbeddingLayer>(4, 6), {2, 6, 4});
}
TEST(gradcheck_attention_and_blocks) {
  gradcheck(std::make_unique<AttentionBlock>(8, 2, true), {2, 5, 8}, 3e-2);
  gradcheck(std::make_unique<AttentionBlock>(8, 2, false), {1, 4, 8}, 3e-2);
  LayerBuilder b({6, 6, 4});
  b.wide_residual_block(4, 8, 2, 0.0, "w");
  gradcheck(b.build_sequential("wrn_block"), {2, 6, 6, 4}, 6e-2, true, 2e-3f);
  LayerBuilder r({6, 6, 4});
  r.basic_residual_block(4, 4, 1, "r");
  gradcheck(r.build_sequential("res_block"), {2, 6, 6, 4}, 6e-2, true, 2e-3f);
  LayerBuilder g({5, 8});
  g.gpt_block(8, 2, 16, 0.0, true, "gelu", "h0");
  gradcheck(g.build_sequential("gpt"), {2, 5, 8}, 4e-2);
}
TEST(msequential_and_nary) {
  Vec<std::unique_ptr<Layer>> seqs;
  for (int i = 0; i < 2; ++i) {
    LayerBuilder b({6});
    b.dense(4, true, "fc" + std::to_string(i));
    seqs.push_back(b.build_sequential("branch" + std::to_string(i)));
  }
  auto m = std::make_unique<MSequential>(std::move(seqs), std::make_unique<NAryLayer>(0, "join"), "m");
  gradcheck(std::move(m), {3, 6});
}

// ------------------------------------------------------------------------------------- factory / zoo / graph / checkpoint
TEST(factory_covers_reference_type_names) {
  const char* names[] = {"dense", "conv2d", "batchnorm", "maxpool2d", "avgpool2d", "groupnorm", "layer_norm", "dropout", "flatten", "activation", "embedding",
                         "pos_embedding", "class_token", "slice", "transpose", "identity", "add", "sub", "mul", "mbroadcast", "sdpa", "sequential", "msequential",
                         "residual_block", "attention_block", "flash_attention_block", "legacy_dense", "legacy_conv2d", "legacy_maxpool2d", "legacy_avgpool2d",
                         "legacy_batchnorm"};
  for (const char* n : names) CHECK(LayerFactory::has(n));
  // config round trip through the factory for a composite model
  auto wrn = ExampleModels::create("cifar100_wrn16_8");
  const TConfig cfg = wrn->get_config();
  auto again = LayerFactory::create(cfg);
  CHECK(again->get_config() == cfg);
  CHECK(again->output_shape({2, 32, 32, 3}) == Shape({2, 100}));
  auto legacy = LayerFactory::create(R"({"name":"lc","type":"legacy_conv2d","parameters":{"in_channels":{"type":"size_t","value":3},"out_channels":{"type":"size_t","value":4},
     "kernel_h":{"type":"size_t","value":3},"kernel_w":{"type":"size_t","value":3},"pad_h":{"type":"size_t","value":1},"pad_w":{"type":"size_t","value":1}}})");
  CHECK(legacy->type() == "legacy_conv2d" && legacy->output_shape({2, 3, 8, 8}) == Shape({2, 4, 8, 8}));
  legacy->init(getHost(), 1);
  Tensor x = make_tensor<float>({2, 3, 8, 8});
  x->fill_random_normal(0, 1, 1);
  CHECK(legacy->forward(x, 0)->shape() == Shape({2, 4, 8, 8}));
}
TEST(zoo_builds_every_model) {
  for (const auto& n : ExampleModels::available_models()) {
    auto m = ExampleModels::create(n);
    Shape in{2};
    const Shape s = ExampleModels::input_shape(n);
    in.insert(in.end(), s.begin(), s.end());
    const Shape out = m->output_shape(in);
    CHECK(out.back() == ExampleModels::num_classes(n));
  }
  CHECK(ExampleModels::available_models().size() >= 17);
}
TEST(graph_topology_and_cycle_detection) {
  GraphBuilder b("g");
  OpNode& a = b.add_layer(std::make_unique<DenseLayer>(4, 4, true, "a"), "a");
  OpNode& c = b.add_layer(std::make_unique<DenseLayer>(4, 2, true, "c"), "c");
  b.add_edge({"mid"}, {"out"}, c);                            // declared out of order
  b.add_edge({"in"}, {"mid"}, a);
  const auto order = b.topo_sort();
  CHECK(order[0].op == "a" && order[1].op == "c");
  GraphBuilder cyc("cyc");
  OpNode& p = cyc.add_layer(std::make_unique<IdentityLayer>("p"), "p");
  OpNode& q = cyc.add_layer(std::make_unique<IdentityLayer>("q"), "q");
  cyc.add_edge({"x"}, {"y"}, p).add_edge({"y"}, {"x"}, q);
  bool threw = false;
  try { cyc.topo_sort(); } catch (const std::runtime_error&) { threw = true; }
  CHECK(threw);
}
TEST(graph_context_slabs_and_checkpoint_round_trip) {
  Graph g = load_or_create_model("mnist_cnn", "", getHost(), 7);
  GraphContext& ctx = g.context();
  CHECK(ctx.num_parameters() > 1000 && ctx.param_slab()->size() >= ctx.num_parameters());
  // parameters live inside the slab
  Layer* conv1 = g.get_layers()[0]->children()[0];
  const char* base = static_cast<const char*>(ctx.param_slab()->data());
  const char* wp = static_cast<const char*>(conv1->param("weight")->data());
  CHECK(wp >= base && wp < base + ctx.param_slab()->nbytes());
  auto opt = OptimizerFactory::create_adam(1e-2, 0.9, 0.999, 1e-3, 3e-4);
  auto crit = LossFactory::create_logsoftmax_crossentropy();
  TrainStep step(g, *crit, *opt);
  SyntheticDataLoader data({28, 28, 1}, 10, 64, 3);
  Tensor x, y;
  data.get_batch(16, x, y);
  for (int i = 0; i < 3; ++i) step(x, y);
  const std::string path = "/tmp/tnn_tests_ckpt.bin";
  save_checkpoint(g, path, opt.get(), nullptr, 3);
  // header is the reference schema
  std::ifstream f(path, std::ios::binary);
  std::stringstream ss;
  ss << f.rdbuf();
  size_t pos = 0;
  const Json hdr = Json::parse(ss.str(), &pos);
  CHECK(hdr.at("type").as_string() == "graph" && hdr.at("parameters").at("op_nodes").at("type").as_string() == "json");
  CHECK(hdr.at("parameters").at("edges").at("value")[0].contains("producers"));
  auto opt2 = OptimizerFactory::create_adam(1e-2, 0.9, 0.999, 1e-3, 3e-4);
  uint64_t step_no = 0;
  Graph g2 = load_checkpoint(path, getHost(), opt2.get(), nullptr, &step_no);
  CHECK(step_no == 3 && opt2->step_count == 3);
  CHECK(max_abs_diff(*g2.context().param_slab(), *ctx.param_slab()) == 0);
  CHECK(max_abs_diff(*g2.context().buffer_slab(), *ctx.buffer_slab()) == 0);        // BatchNorm running statistics serialise too
  CHECK(max_abs_diff(*opt2->state_tensors()["m"], *opt->state_tensors()["m"]) == 0);
  // resumed training continues identically
  TrainStep step2(g2, *crit, *opt2);
  Tensor s1 = step(x, y), s2 = step2(x, y);
  CHECK_NEAR(s1->data_as<float>()[0], s2->data_as<float>()[0], 1e-6);
}
TEST(schedulers_closed_forms) {
  auto opt = OptimizerFactory::create_sgd(1.0);
  auto s = SchedulerFactory::create_step_lr(opt.get(), 2, 0.5);
  s->step(); CHECK_NEAR(opt->get_learning_rate(), 1.0, 0);
  s->step(); CHECK_NEAR(opt->get_learning_rate(), 0.5, 0);
  s->step(); s->step(); CHECK_NEAR(opt->get_learning_rate(), 0.25, 0);
  auto o2 = OptimizerFactory::create_sgd(1.0);
  auto c = SchedulerFactory::create_cosine_annealing_lr(o2.get(), 10, 0.0);
  for (int i = 0; i < 5; ++i) c->step();
  CHECK_NEAR(o2->get_learning_rate(), 0.5, 1e-9);
  auto o3 = OptimizerFactory::create_sgd(1.0);
  auto w = SchedulerFactory::create_linear_warmup(o3.get(), 4, 0.0);
  w->step(); CHECK_NEAR(o3->get_learning_rate(), 0.25, 1e-9);
  auto o4 = OptimizerFactory::create_sgd(1.0);
  auto r = SchedulerFactory::create_reduce_lr_on_plateau(o4.get(), 0.1, 1);
  r->step(1.0); r->step(1.0); r->step(1.0);
  CHECK_NEAR(o4->get_learning_rate(), 0.1, 1e-12);
  for (const char* t : {"no_op", "step_lr", "multi_step_lr", "exponential_lr", "cosine_annealing_lr", "cosine_annealing_warm_restarts", "linear_warmup",
                        "warmup_cosine_annealing", "reduce_lr_on_plateau", "polynomial_lr", "one_cycle_lr"}) {
    TConfig cfg(t, t);
    auto sc = SchedulerFactory::create_from_config(cfg, o4.get());
    CHECK(sc->get_config().type == t || std::string(t) == "no_op");
    auto sc2 = SchedulerFactory::create_from_config(sc->get_config(), o4.get());
    CHECK(sc2->name() == sc->name());
  }
}
TEST(losses_match_closed_forms) {
  Tensor p = make_tensor_from({2.f, 1.f, 0.f, 0.f, 0.f, 3.f}, {2, 3});
  Tensor y = make_tensor(DType_t::INT32_T, {2});
  y->data_as<int32_t>()[0] = 0;
  y->data_as<int32_t>()[1] = 1;
  auto ce = LossFactory::create("logsoftmax_crossentropy");
  Tensor stats, grad;
  ce->loss_and_grad(p, y, stats, &grad);
  const double l0 = std::log(std::exp(2.) + std::exp(1.) + 1) - 2, l1 = std::log(2 + std::exp(3.)) - 0;
  CHECK_NEAR(stats->data_as<float>()[0], (l0 + l1) / 2, 1e-5);
  CHECK_NEAR(stats->data_as<float>()[1], 1.0, 0);              // row 0 correct, row 1 wrong
  CHECK_NEAR(grad->data_as<float>()[0], (std::exp(2.) / (std::exp(2.) + std::exp(1.) + 1) - 1) / 2, 1e-6);
  float loss = 0;
  Tensor t = make_tensor_from({1.f, 1.f, 1.f, 1.f, 1.f, 1.f}, {2, 3});
  LossFactory::create("mse")->compute_loss(p, t, loss);
  CHECK_NEAR(loss, (1 + 0 + 1 + 1 + 1 + 4) / 6.0, 1e-6);
  LossFactory::create("mae")->compute_loss(p, t, loss);
  CHECK_NEAR(loss, (1 + 0 + 1 + 1 + 1 + 2) / 6.0, 1e-6);
  LossFactory::create("huber")->compute_loss(p, t, loss);
  CHECK_NEAR(loss, (0.5 + 0 + 0.5 + 0.5 + 0.5 + 1.5) / 6.0, 1e-6);
  CHECK(compute_class_corrects(p, y) == 1);
}
TEST(training_converges_on_cpu) {
  Graph g = load_or_create_model("mnist_cnn", "", getHost(), 11);
  auto opt = OptimizerFactory::create_adam(3e-3, 0.9, 0.999, 1e-3, 0.0);
  auto crit = LossFactory::create_logsoftmax_crossentropy();
  SyntheticDataLoader train({28, 28, 1}, 10, 256, 5), val({28, 28, 1}, 10, 128, 5);
  TrainingConfig cfg;
  cfg.epochs = 3;
  cfg.batch_size = 32;
  cfg.progress_print_interval = 0;
  cfg.snapshot_dir = "/tmp/tnn_tests_snap";
  TrainHistory h = train_model(g, train, &val, *opt, *crit, nullptr, cfg, [](const std::string&) {});
  CHECK(h.train_loss.size() == 3 && h.train_loss.back() < h.train_loss.front() * 0.8);
  CHECK(h.val_acc.back() > 0.3);
}

// ------------------------------------------------------------------------------------- distributed runtime
TEST(message_wire_format_round_trip) {
  Tensor t = make_tensor<float>({2, 3});
  t->fill_random_normal(0, 1, 4);
  Message m(CommandType::FORWARD_JOB, Job{t, 7});
  const auto buf = m.serialize();
  // u16 command, u32 variant index 1, u64 mb_id, u32 dtype FP32 = 5, u64 ndims, dims, raw
  uint16_t cmd; uint32_t var; uint64_t mb; uint32_t dt;
  std::memcpy(&cmd, buf.data(), 2); std::memcpy(&var, buf.data() + 2, 4); std::memcpy(&mb, buf.data() + 6, 8); std::memcpy(&dt, buf.data() + 14, 4);
  CHECK(cmd == 1 && var == 1 && mb == 7 && dt == 5 && buf.size() == 2 + 4 + 8 + 4 + 8 + 16 + 24);
  Message r = Message::deserialize(buf.data(), buf.size());
  CHECK(r.command == CommandType::FORWARD_JOB && r.job.mb_id == 7 && max_abs_diff(*r.job.data, *t) == 0);
  StageConfig sc;
  sc.stage_id = 2;
  sc.next_stage_endpoint = Endpoint::tcp("10.0.0.2", 5555);
  sc.model_config = ExampleModels::create("mnist_cnn")->get_config().to_json();
  const auto b2 = Message(CommandType::CONFIG_TRANSFER, sc).serialize();
  Message r2 = Message::deserialize(b2.data(), b2.size());
  CHECK(r2.kind == Message::STAGE && r2.stage.stage_id == 2 && r2.stage.next_stage_endpoint.port == 5555 && r2.stage.model_config == sc.model_config);
  Profiler p;
  p.init_start_time();
  p.add_event(EventType::COMPUTE, 10, 20, "fwd", "w0");
  const auto b3 = Message(CommandType::PROFILING_REPORTED, p).serialize();
  CHECK(Message::deserialize(b3.data(), b3.size()).profiler->events().size() == 1);
  const auto b4 = Message(CommandType::ERROR_REPORT, "boom").serialize();
  CHECK(Message::deserialize(b4.data(), b4.size()).str == "boom");
  // priority dequeue: FORWARD_JOB before BACKWARD_JOB before UPDATE_PARAMETERS regardless of arrival order
  InProcessCommunicator c("c");
  c.enqueue(Message(CommandType::UPDATE_PARAMETERS));
  c.enqueue(Message(CommandType::BACKWARD_JOB, Job{t, 1}));
  c.enqueue(Message(CommandType::FORWARD_JOB, Job{t, 2}));
  Message o;
  CHECK(c.dequeue(&o, 0) && o.command == CommandType::FORWARD_JOB);
  CHECK(c.dequeue(&o, 0) && o.command == CommandType::BACKWARD_JOB);
  CHECK(c.dequeue(&o, 0) && o.command == CommandType::UPDATE_PARAMETERS && !c.dequeue(&o, 0));
}
TEST(partitioners_cover_the_model_contiguously) {
  auto m = ExampleModels::create("cifar100_wrn16_8");
  const Vec<Layer*> layers = m->layers();
  auto check = [&](Vec<SeqPartition> p, size_t n) {
    CHECK(p.size() == n);
    size_t cur = 0;
    for (auto& s : p) {
      CHECK(s.start == cur && s.length >= 1);
      cur += s.length;
    }
    CHECK(cur == layers.size());
  };
  check(NaivePipelinePartitioner({1, 2, 1}).partition_model(layers, {64, 32, 32, 3}), 3);
  check(WeightedPartitioner(4).partition_model(layers, {64, 32, 32, 3}), 4);
  const auto ftd = FTDPartitioner(4).partition_model(layers, {64, 32, 32, 3});
  check(ftd, 4);
  // the FLOPs-aware split balances the six wide blocks instead of leaving most layers to the last stage
  Shape s{64, 32, 32, 3};
  std::vector<double> fl;
  for (Layer* l : layers) { fl.push_back((double)l->forward_flops(s)); s = l->output_shape(s); }
  double total = 0, worst = 0;
  for (double f : fl) total += f;
  for (auto& p : ftd) { double c = 0; for (size_t i = p.start; i < p.start + p.length; ++i) c += fl[i]; worst = std::max(worst, c); }
  CHECK(worst < 0.45 * total);
  auto parts = Partitioner::split(layers, ftd, "wrn");
  CHECK(parts.size() == 4 && parts[0]->size() == ftd[0].length);
}
static std::vector<double> single_device_losses(const std::string& model, const Vec<std::pair<Tensor, Tensor>>& batches, size_t M, double lr,
                                                const Device& dev = getHost()) {
  Graph g = load_or_create_model(model, "", dev, 123456);
  auto opt = OptimizerFactory::create_sgd(lr, 0.9);
  opt->attach(g.context());
  auto crit = LossFactory::create_logsoftmax_crossentropy();
  GraphExecutor ex(g);
  std::vector<double> out;
  for (auto& b : batches) {
    Vec<Tensor> xs, ys;
    split_microbatches(b.first, b.second, M, &xs, &ys);
    double loss = 0;
    for (size_t i = 0; i < xs.size(); ++i) {
      Tensor o = ex.forward(xs[i], i);
      Tensor stats, grad;
      crit->loss_and_grad(o, ys[i]->device() == o->device() ? ys[i] : ys[i]->to_device(o->device()), stats, &grad, 1.f / (float)xs.size());
      loss += (stats->is_gpu() ? stats->to_host() : stats)->data_as<float>()[0];
      ex.backward(grad, i);
    }
    opt->update();
    out.push_back(loss);
  }
  return out;
}
// transport: 0 in-process, 1 TCP bus, 2 one-sided rendezvous (RoCECommunicator); `dev`: where the stages run
static void pipeline_case(int transport, const Device& dev = getHost(), double tol = 1e-4) {
  const bool tcp = transport != 0;
  const size_t M = 2;
  SyntheticDataLoader data({28, 28, 1}, 10, 64, 3);
  Vec<std::pair<Tensor, Tensor>> batches;
  for (int i = 0; i < 3; ++i) {
    Tensor x, y;
    data.get_batch(16, x, y);
    batches.emplace_back(x, y);
  }
  Vec<std::unique_ptr<Worker>> workers;
  Vec<Endpoint> eps;
  for (int i = 0; i < 2; ++i) {
    std::unique_ptr<Communicator> c;
    if (transport == 2) c = std::make_unique<RoCECommunicator>("127.0.0.1", 0, 2, &dev, "w" + std::to_string(i), 32u << 20, 1024);   // stage-0 activations are exactly 16 KiB
    else if (tcp) c = std::make_unique<TCPCommunicator>("127.0.0.1", 0, 2, &dev, "w" + std::to_string(i));
    else c = std::make_unique<InProcessCommunicator>("w" + std::to_string(i));
    auto w = std::make_unique<Worker>(std::move(c), dev, "w" + std::to_string(i));
    w->start(true);
    eps.push_back(w->comm().endpoint());
    workers.push_back(std::move(w));
  }
  CoordinatorConfig cc;
  cc.model = ExampleModels::create("mnist_cnn");
  cc.input_shape = {8, 28, 28, 1};
  cc.optimizer_config = OptimizerFactory::create_sgd(0.05, 0.9)->get_config();
  cc.partitioner = std::make_unique<NaivePipelinePartitioner>(Vec<size_t>{1, 2});
  cc.worker_endpoints = eps;
  cc.timeout_job_ms = 30000;
  std::unique_ptr<Communicator> comm;
  if (transport == 2) comm = std::make_unique<RoCECommunicator>("127.0.0.1", 0, 2, dev.is_gpu() ? &dev : nullptr, "coord", 32u << 20);
  else if (tcp) comm = std::make_unique<TCPCommunicator>("127.0.0.1", 0, 2, nullptr, "coord");
  else comm = std::make_unique<InProcessCommunicator>("coord");
  Coordinator co(std::move(cc), std::move(comm));
  co.start();
  auto crit = LossFactory::create_logsoftmax_crossentropy();
  std::vector<double> losses;
  for (auto& b : batches) {
    Vec<Tensor> xs, ys;
    split_microbatches(b.first, b.second, M, &xs, &ys);
    losses.push_back(co.async_train_batch(xs, ys, *crit).first);
    co.update_parameters();
  }
  // the stages are re-created from config with the reference's fixed seed (123456): compare with a single-device run whose
  // layers were initialised the same way -- per-layer streams depend on (seed, layer name), not on the stage they live in
  // (on the GPU the reference run is a GPU run too: dropout masks come from the device Philox stream, keyed by seed / layer / micro-batch)
  const auto ref = single_device_losses("mnist_cnn", batches, M, 0.05, dev);
  {
    std::string all;
    for (size_t i = 0; i < ref.size(); ++i) all += " " + std::to_string(losses[i]) + "/" + std::to_string(ref[i]);
    for (size_t i = 0; i < ref.size(); ++i)
      if (!(std::fabs(losses[i] - ref[i]) <= tol)) throw std::runtime_error("pipeline losses differ from the single-device run (pipeline/reference):" + all);
  }
  if (transport == 2 && dev.is_gpu()) {
    // activations and gradients crossed through the device slabs, not through host staging
    for (auto& w : workers) {
      const auto st = static_cast<RoCECommunicator&>(w->comm()).stats();
      if (!(st.device_rendezvous_sent > 0 && st.device_rendezvous_received > 0 && st.device_fallback_host == 0))
        throw std::runtime_error(w->comm().name() + ": device rendezvous sent " + std::to_string(st.device_rendezvous_sent) + " received " +
                                 std::to_string(st.device_rendezvous_received) + " host fallbacks " + std::to_string(st.device_fallback_host) +
                                 " inline " + std::to_string(st.inline_sent));
    }
  }
  CHECK(co.health_check().size() == 2);
  {
    const auto loads = co.report_load();                // LoadTracker record per stage (load_tracker.hpp)
    CHECK(loads.size() == 2 && loads[0].value("stage", -1) == 0 && loads[1].value("stage", -1) == 1);
    for (const auto& l : loads) {
      const LoadTracker t = LoadTracker::from_json(l);
      CHECK(t.avg_forward_time_ > 0 && t.avg_backward_time_ > 0 && t.max_memory_usage_ > 1 && t.avg_cpu_utilization_ >= 0);
      CHECK(l.value("jobs", 0) > 0);
    }
  }
  co.barrier();
  Profiler prof = co.fetch_profiling();
  CHECK(!prof.events().empty());
  const auto files = co.checkpoint("/tmp/tnn_tests_pp_ckpt");
  CHECK(files.size() == 2);
  co.stop();
  for (auto& w : workers) w->stop();
}
TEST(worker_refuses_paths_outside_its_io_roots) {
  // SAVE_TO_FILE / LOAD_PARAMS / CHECKPOINT_REQUEST arrive over an unauthenticated bus: only the working directory, the
  // temp directory and TNN_WORKER_IO_ROOT may be named
  Worker w(std::make_unique<InProcessCommunicator>("w"), getHost(), "w");
  CHECK(w.io_roots().size() >= 2);
  CHECK(w.checked_path("/tmp/tnn_tests_ok/model.stage0").rfind("/tmp/", 0) == 0 || w.checked_path("/tmp/tnn_tests_ok/model.stage0").find("tnn_tests_ok") != std::string::npos);
  CHECK(!w.checked_path("relative_snapshot").empty());
  auto refused = [&](const std::string& p) {
    try {
      w.checked_path(p);
    } catch (const std::runtime_error&) {
      return true;
    }
    return false;
  };
  CHECK(refused("/etc/passwd"));
  CHECK(refused("/tmp/../etc/cron.d/x"));
  CHECK(refused("../../../../../../etc/shadow"));
  CHECK(refused("/tmpfoo/model"));                      // a root is a directory prefix, not a string prefix
}
TEST(in_process_pipeline_equals_single_device) { pipeline_case(0); }
TEST(tcp_pipeline_equals_single_device) { pipeline_case(1); }
TEST(roce_pipeline_equals_single_device) { pipeline_case(2); }
TEST(roce_host_slab_rendezvous_and_fallbacks) {
  // CPU tensors between two RoCECommunicators of one node: small -> inline frame, large -> PREPARE / READY_TO_WRITE / write
  // into the peer's shared-memory slab / DONE, larger than the slab -> inline fallback, more than the slab holds at once ->
  // the late PREPAREs wait for frees
  RoCECommunicator a("127.0.0.1", 0, 2, nullptr, "a", 0, 16u << 10, 4u << 20), b("127.0.0.1", 0, 2, nullptr, "b", 0, 16u << 10, 4u << 20);
  a.start();
  b.start();
  a.connect(b.endpoint());
  auto make = [](size_t n, uint64_t seed) {
    Tensor t = make_tensor<float>({n});
    t->fill_random_uniform(-1, 1, seed);
    return t;
  };
  Tensor small = make(16, 1), big = make(1u << 18, 2), huge = make(5u << 18, 3);        // 64 B, 1 MiB, 5 MiB
  const Tensor ts[3] = {small, big, huge};
  for (int i = 0; i < 3; ++i) a.send_message(Message(CommandType::FORWARD_JOB, Job{ts[i], (uint64_t)i}), b.endpoint());
  for (int i = 0; i < 3; ++i) {
    Message m;
    CHECK(b.dequeue(&m, 30000));
    CHECK(m.command == CommandType::FORWARD_JOB && m.job.mb_id < 3 && max_abs_diff(*m.job.data, *ts[m.job.mb_id]) == 0);
  }
  auto sa = a.stats(), sb = b.stats();
  CHECK(sa.inline_sent == 1 && sa.rendezvous_sent == 1 && sa.fallback_inline == 1);
  CHECK(sb.rendezvous_received == 1 && sb.host_slab_peak >= (1u << 20));
  b.send_message(Message(CommandType::BACKWARD_JOB, Job{big, 9}), a.endpoint());                // the other direction
  Message back;
  CHECK(a.dequeue(&back, 30000) && back.job.mb_id == 9 && max_abs_diff(*back.job.data, *big) == 0);
  for (int i = 0; i < 12; ++i) a.send_message(Message(CommandType::FORWARD_JOB, Job{big, (uint64_t)(100 + i)}), b.endpoint());
  std::vector<uint64_t> ids;
  for (int i = 0; i < 12; ++i) {
    Message m;
    CHECK(b.dequeue(&m, 30000));
    ids.push_back(m.job.mb_id);
  }
  std::sort(ids.begin(), ids.end());
  for (int i = 0; i < 12; ++i) CHECK(ids[i] == (uint64_t)(100 + i));
  CHECK(b.stats().rendezvous_received == 13 && b.stats().host_slab_peak <= (4u << 20));
  a.stop();
  b.stop();
}
TEST(lost_microbatch_and_worker_errors_are_detected) {
  auto w = std::make_unique<Worker>(std::make_unique<InProcessCommunicator>("w"), getHost(), "w");
  w->start(true);
  int drops = 0;
  w->fault_inject = [&](Message& m) {
    if (m.command == CommandType::BACKWARD_JOB && drops++ == 0) throw std::runtime_error("injected fault");
  };
  CoordinatorConfig cc;
  cc.model = ExampleModels::create("mnist_cnn");
  cc.input_shape = {8, 28, 28, 1};
  cc.optimizer_config = OptimizerFactory::create_sgd(0.05)->get_config();
  cc.partitioner = std::make_unique<NaivePipelinePartitioner>(Vec<size_t>{1});
  cc.worker_endpoints = {w->comm().endpoint()};
  cc.timeout_job_ms = 3000;
  Coordinator co(std::move(cc), std::make_unique<InProcessCommunicator>("coord"));
  co.start();
  SyntheticDataLoader data({28, 28, 1}, 10, 16, 3);
  Tensor x, y;
  data.get_batch(8, x, y);
  auto crit = LossFactory::create_logsoftmax_crossentropy();
  bool threw = false;
  try {
    co.async_train_batch({x}, {y}, *crit);
  } catch (const DistributedError& e) {
    threw = std::string(e.what()).find("injected fault") != std::string::npos;
  }
  CHECK(threw);
  co.stop();
  w->stop();
}

// ------------------------------------------------------------------------------------- GPU: layer parity with the CPU oracle
TEST(utils_threading_concurrent_map_logger) {
  std::vector<std::atomic<int>> hits(1000);
  parallel_for(0, 1000, [&](int64_t i) { hits[(size_t)i]++; }, 16);
  for (auto& h : hits) CHECK(h.load() == 1);
  std::atomic<long> sum{0};
  parallel_for_2d(20, 30, [&](int64_t i, int64_t j) { sum += i * 30 + j; });
  CHECK(sum.load() == 599L * 600 / 2);
  CHECK(ThreadWrapper(2).execute([] { return thread_limit(); }) == 2 && thread_limit() == 0);
  ConcurrentMap<std::string, int> m;
  CHECK(m.emplace("a", 1) && !m.emplace("a", 2) && m.try_get("a").value() == 1 && !m.try_get("b").has_value());
  parallel_for(0, 500, [&](int64_t) { m.update("n", [](int& v) { ++v; }); });
  CHECK(m.try_get("n").value() == 500 && m.contains("n") && m.erase("a") && m.size() == 1 && m.snapshot().count("n") == 1);
  const std::string path = "/tmp/tnn_cpp_logger_test.log", steps = "/tmp/tnn_cpp_steplog_test.jsonl";
  std::remove(path.c_str());
  std::remove(steps.c_str());
  Logger log("unit", LogLevel::WARN);
  log.set_file(path);
  log.info("dropped");
  log.warn("kept 1");
  log.error("kept 2");
  std::ifstream f(path);
  std::string all((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  CHECK(all.find("dropped") == std::string::npos && all.find("[warn] kept 1") != std::string::npos && all.find("[error] kept 2") != std::string::npos);
  {
    StepLog sl(steps);
    Json r = Json::object();
    r["step"] = 3;
    r["loss"] = 0.25;
    sl.write(r);
  }
  std::ifstream g(steps);
  std::string line;
  std::getline(g, line);
  CHECK(Json::parse(line).at("step").as_int() == 3);
  CHECK(get_shape_hash({2, 3, 4}) != get_shape_hash({4, 3, 2}) && get_shape_hash({2, 3}) == get_shape_hash({2, 3}));
  CHECK(benchmark([] {}, 3, 1) >= 0.0);
}

// ------------------------------------------------------------------------------------- augmentation + file loaders
TEST(augmentation_strategies_and_builder) {
  auto img = [](size_t B, size_t H, size_t W, size_t C, uint64_t seed) {
    Tensor t = make_tensor(DType_t::FP32, {B, H, W, C});
    t->fill_random_uniform(0.0, 1.0, seed);
    return t;
  };
  // flips applied twice with p = 1 are the identity; a single flip mirrors
  Tensor x = img(2, 6, 5, 3, 1), ref = x->clone();
  auto fl = AugmentationBuilder(3).horizontal_flip(1.0).build();
  fl->apply(*x);
  CHECK_NEAR(x->at<float>({0, 2, 0, 1}), ref->at<float>({0, 2, 4, 1}), 0.0);
  fl->apply(*x);
  for (size_t i = 0; i < x->size(); ++i) CHECK_NEAR(x->data_as<float>()[i], ref->data_as<float>()[i], 0.0);
  auto vf = AugmentationBuilder(3).vertical_flip(1.0).build();
  vf->apply(*x);
  CHECK_NEAR(x->at<float>({1, 0, 3, 2}), ref->at<float>({1, 5, 3, 2}), 0.0);
  // normalize: (x - mean) / std per channel
  x = ref->clone();
  AugmentationBuilder().normalize({0.5f, 0.4f, 0.3f}, {0.2f, 0.25f, 0.5f}).build()->apply(*x);
  CHECK_NEAR(x->at<float>({0, 1, 1, 1}), (ref->at<float>({0, 1, 1, 1}) - 0.4f) / 0.25f, 1e-6);
  // crop keeps the shape and only moves pixels; cutout zeroes a square; rotation by 0 degrees is the identity
  x = ref->clone();
  AugmentationBuilder(5).random_crop(1.0, 2).build()->apply(*x);
  CHECK(x->shape() == ref->shape());
  double zeros = 0;
  x = ref->clone();
  AugmentationBuilder(5).cutout(1.0, 3).build()->apply(*x);
  for (size_t i = 0; i < x->size(); ++i) zeros += x->data_as<float>()[i] == 0.f;
  CHECK(zeros >= 2 * 3 * 4 && zeros <= 2 * 3 * 16 + 4);            // 2 images x 3 channels x (clipped 2x2 .. 4x4 squares)
  x = ref->clone();
  AugmentationBuilder(5).rotation(1.0, 0.0).build()->apply(*x);
  for (size_t i = 0; i < x->size(); ++i) CHECK_NEAR(x->data_as<float>()[i], ref->data_as<float>()[i], 1e-6);
  // value-range strategies stay in [0, 1]; the whole chain is deterministic in (seed, batch index)
  auto chain = [&](uint64_t seed) {
    return AugmentationBuilder(seed).horizontal_flip(0.5).vertical_flip(0.2).rotation(0.5, 10).brightness(0.7, 0.3).contrast(0.7, 0.3).gaussian_noise(0.7, 0.1)
        .random_crop(1.0, 2).cutout(0.5, 2).build();
  };
  CHECK(chain(1)->size() == 8 && chain(1)->names().front() == "horizontal_flip");
  Tensor a = ref->clone(), b = ref->clone(), c = ref->clone();
  chain(9)->apply(*a, 4);
  chain(9)->apply(*b, 4);
  chain(9)->apply(*c, 5);
  double diff = 0;
  for (size_t i = 0; i < a->size(); ++i) {
    CHECK(a->data_as<float>()[i] >= 0.f && a->data_as<float>()[i] <= 1.f);
    CHECK_NEAR(a->data_as<float>()[i], b->data_as<float>()[i], 0.0);
    diff += std::fabs(a->data_as<float>()[i] - c->data_as<float>()[i]);
  }
  CHECK(diff > 1e-3);
  // a loader applies its augmentation to every image batch
  SyntheticDataLoader plain({8, 8, 3}, 4, 32, 1), aug({8, 8, 3}, 4, 32, 1);
  aug.set_augmentation(AugmentationBuilder(2).horizontal_flip(1.0).build());
  Tensor px, py, ax, ay;
  plain.get_batch(4, px, py);
  aug.get_batch(4, ax, ay);
  CHECK_NEAR(ax->at<float>({1, 3, 0, 2}), px->at<float>({1, 3, 7, 2}), 0.0);
}
TEST(token_csv_and_wifi_loaders) {
  const std::string dir = "/tmp/tnn_cpp_loader_test";
  mkdir(dir.c_str(), 0755);
  {                                               // 5000 uint16 tokens: t[i] = i % 1000
    std::ofstream f(dir + "/train.bin", std::ios::binary);
    for (uint32_t i = 0; i < 5000; ++i) {
      const uint16_t t = (uint16_t)(i % 1000);
      f.write(reinterpret_cast<const char*>(&t), 2);
    }
  }
  OpenWebTextDataLoader owt(64);
  CHECK(owt.load_data(dir + "/train.bin") && owt.num_tokens() == 5000 && owt.size() == 5000 / 64);
  Tensor x, y;
  CHECK(owt.get_batch(3, x, y));
  CHECK((x->shape() == Shape{3, 64}) && x->data_type() == DType_t::INT32_T && (y->shape() == Shape{3, 64}));
  for (size_t b = 0; b < 3; ++b)
    for (size_t i = 0; i + 1 < 64; ++i) {
      CHECK(y->data_as<int32_t>()[b * 64 + i] == x->data_as<int32_t>()[b * 64 + i + 1]);        // targets are the next tokens
      CHECK((x->data_as<int32_t>()[b * 64 + i] + 1) % 1000 == x->data_as<int32_t>()[b * 64 + i + 1]);
    }
  DataLoaderPair p = DataLoaderFactory::create("openwebtext", dir);
  CHECK(p.train && p.train->size() == 5000 / 1024);
  {                                               // CSV: 3 features, 2 targets
    std::ofstream f(dir + "/reg.csv");
    f << "a,b,c,t0,t1\n";
    for (int i = 0; i < 10; ++i) f << i << ',' << 2 * i << ',' << -i << ',' << 10 * i << ',' << i * i << "\n";
  }
  RegressionCSVDataLoader reg(2);
  CHECK(reg.load_data(dir + "/reg.csv") && reg.size() == 10 && reg.num_features() == 3);
  CHECK(reg.get_batch(4, x, y));
  CHECK((x->shape() == Shape{4, 3}) && (y->shape() == Shape{4, 2}));
  CHECK_NEAR(x->at<float>({3, 1}), 6.0, 0.0);
  CHECK_NEAR(y->at<float>({3, 1}), 9.0, 0.0);
  CHECK(reg.get_batch(100, x, y) && x->dimension(0) == 6 && !reg.get_batch(1, x, y));
  WiFiDataLoader wifi(true);
  CHECK(wifi.load_data(dir + "/reg.csv", 0, 3, 3, 5) && wifi.num_features() == 3 && wifi.num_targets() == 2);
  wifi.normalize_data();
  CHECK(wifi.is_normalized());
  CHECK_NEAR(wifi.feature_means()[0], 4.5, 1e-5);
  CHECK_NEAR(wifi.target_means()[0], 45.0, 1e-4);
  CHECK(wifi.get_batch(10, x, y));
  double m = 0, v = 0;
  for (size_t i = 0; i < 10; ++i) m += x->at<float>({i, 1});
  for (size_t i = 0; i < 10; ++i) v += x->at<float>({i, 1}) * x->at<float>({i, 1});
  CHECK_NEAR(m / 10, 0.0, 1e-5);
  CHECK_NEAR(v / 10, 1.0, 1e-4);
  WiFiDataLoader cls(false);
  CHECK(cls.load_data(dir + "/reg.csv", 0, 3, 3, 4));
  CHECK(cls.get_batch(5, x, y) && y->data_type() == DType_t::INT32_T && y->data_as<int32_t>()[2] == 20);
}

// ------------------------------------------------------------------------------------- GPT-2 inference engine / legacy zoo / hardware info
static std::unique_ptr<Sequential> tiny_gpt(size_t vocab = 97, size_t E = 128, size_t H = 2, size_t layers = 2, size_t seq = 48) {
  LayerBuilder b({seq});
  b.embedding(vocab, E, -1, "token_embed").positional_embedding(E, seq, "pos_embed");
  for (size_t i = 0; i < layers; ++i) b.gpt_block(E, H, 4 * E, 0.0, true, "gelu", "h" + std::to_string(i));
  b.layernorm(1e-5, true, "ln_f").dense(vocab, true, "head");
  return b.build_sequential("tiny_gpt");
}
static std::vector<int32_t> toy_ids(size_t n, size_t vocab, uint64_t seed) {
  std::vector<int32_t> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = (int32_t)(((i + 1) * 2654435761ull + seed * 40503ull) % vocab);
  return v;
}
TEST(gpt2_engine_cached_decode_equals_full_recompute_cpu) {
  auto m = tiny_gpt();
  m->init(getHost(), 5);
  GPT2Engine eng(*m, 2, 0, "bf16");
  const size_t B = 2, L = 9, V = eng.vocab();
  std::vector<int32_t> ids = toy_ids(B * L, V, 1);
  std::vector<float> lg;
  std::vector<int32_t> tok = eng.prefill(ids, B, L, &lg);
  // oracle: whole model over the whole sequence, logits of the last position
  std::vector<float> full = eng.full_forward_logits(ids, B, L);
  for (size_t b = 0; b < B; ++b)
    for (size_t v = 0; v < V; ++v) CHECK_NEAR(lg[b * V + v], full[(b * L + L - 1) * V + v], 2e-4);
  // three cached decode steps against a recompute over the grown sequence
  std::vector<std::vector<int32_t>> rows(B);
  for (size_t b = 0; b < B; ++b) rows[b].assign(ids.begin() + (long)(b * L), ids.begin() + (long)((b + 1) * L));
  for (int step = 0; step < 3; ++step) {
    for (size_t b = 0; b < B; ++b) rows[b].push_back(tok[b]);
    tok = eng.decode(tok, &lg);
    const size_t Ln = rows[0].size();
    std::vector<int32_t> grown;
    for (size_t b = 0; b < B; ++b) grown.insert(grown.end(), rows[b].begin(), rows[b].end());
    full = eng.full_forward_logits(grown, B, Ln);
    for (size_t b = 0; b < B; ++b) {
      size_t arg = 0;
      for (size_t v = 0; v < V; ++v) {
        CHECK_NEAR(lg[b * V + v], full[(b * Ln + Ln - 1) * V + v], 3e-4);
        if (full[(b * Ln + Ln - 1) * V + v] > full[(b * Ln + Ln - 1) * V + arg]) arg = v;
      }
      CHECK((size_t)tok[b] == arg);
    }
  }
  CHECK(eng.cur_len() == L + 3);
  // generate() == prefill + decode loop; decode_graph() on the CPU is that loop too
  size_t n = 0;
  std::vector<int32_t> gen = eng.generate(ids, B, L, 4, -1, &n);
  CHECK(n == 4 && gen.size() == B * 4);
  std::vector<int32_t> first = eng.prefill(ids, B, L);
  std::vector<int32_t> rest = eng.decode_graph(first, 3);
  for (size_t b = 0; b < B; ++b) {
    CHECK(gen[b * 4] == first[b]);
    for (size_t s = 0; s < 3; ++s) CHECK(gen[b * 4 + 1 + s] == rest[b * 3 + s]);
  }
  bool threw = false;
  try { eng.prefill(toy_ids(3 * L, V, 2), 3, L); } catch (const std::exception&) { threw = true; }
  CHECK(threw);                                   // batch above capacity is rejected
}
TEST(legacy_zoo_is_the_nchw_twin_of_the_zoo) {
  legacy::ExampleModels::register_defaults();
  std::unique_ptr<Sequential> nhwc = ExampleModels::create("cifar10_resnet9");
  std::unique_ptr<Layer> nchw = legacy::to_legacy(*nhwc);
  const std::string cfg = nchw->get_config().to_json().dump();
  CHECK(cfg.find("legacy_conv2d") != std::string::npos && cfg.find("legacy_batchnorm") != std::string::npos);
  CHECK(cfg.find("\"conv2d\"") == std::string::npos);
  // same parameters, same function: copy the weights across and compare outputs on transposed inputs
  nhwc->init(getHost(), 3);
  nchw->init(getHost(), 4);
  auto a = nhwc->all_layers(), b = nchw->all_layers();
  std::map<std::string, Layer*> by_name;
  for (Layer* l : a) by_name[l->name()] = l;
  size_t copied = 0;
  for (Layer* l : b) {
    auto it = by_name.find(l->name());
    if (it == by_name.end()) continue;
    for (auto& kv : l->params())
      if (it->second->has_param(kv.first) && it->second->param(kv.first)->size() == kv.second->size()) {
        kv.second->copy_from(*it->second->param(kv.first));
        ++copied;
      }
  }
  CHECK(copied > 10);
  nhwc->set_training(false);
  nchw->set_training(false);
  Tensor x = make_tensor(DType_t::FP32, {2, 32, 32, 3});
  x->fill_random_uniform(0.0, 1.0, 5);
  Tensor xt = make_tensor(DType_t::FP32, {2, 3, 32, 32});
  for (size_t n = 0; n < 2; ++n)
    for (size_t p = 0; p < 1024; ++p)
      for (size_t c = 0; c < 3; ++c) xt->data_as<float>()[(n * 3 + c) * 1024 + p] = x->data_as<float>()[(n * 1024 + p) * 3 + c];
  const auto y0 = to_float_vector(*nhwc->forward(x, 0)), y1 = to_float_vector(*nchw->forward(xt, 0));
  CHECK(y0.size() == y1.size() && y0.size() == 20);
  for (size_t i = 0; i < y0.size(); ++i) CHECK_NEAR(y0[i], y1[i], 1e-3 * std::max(1.0, std::fabs((double)y0[i])));
  // NCHW loader adapter
  legacy::NCHWLoader loader(std::make_unique<SyntheticDataLoader>(Shape{8, 6, 3}, 4, 16, 1));
  Tensor bx, by;
  CHECK(loader.get_batch(4, bx, by));
  CHECK((bx->shape() == Shape{4, 3, 8, 6}) && (loader.get_data_shape() == Shape{3, 8, 6}));
}
TEST(hardware_info_reports_the_host) {
  HardwareInfo hw;
  CHECK(hw.logical_cores >= 1 && hw.physical_cores >= 1 && hw.total_ram_kb > 0);
  const Json j = hw.to_json();
  CHECK(j.at("cpu").at("logical_cores").as_int() == hw.logical_cores && j.contains("gpus") && j.contains("numa"));
  CHECK(!hw.numa_topology().empty() && get_memory_usage_kb() > 0);
  CHECK(!ThreadAffinity::current().empty());
  CHECK(Json::parse(j.dump()) == j);
}

static double rel_l2(const TensorImpl& a, const TensorImpl& b) {
  const auto x = to_float_vector(a), y = to_float_vector(b);
  double num = 0, den = 0;
  for (size_t i = 0; i < x.size(); ++i) {
    num += ((double)x[i] - y[i]) * ((double)x[i] - y[i]);
    den += (double)x[i] * x[i];
  }
  return std::sqrt(num / std::max(den, 1e-30));
}
// `l2`: compare gradients by relative L2 norm instead of the largest element error.  Used for blocks that END in a ReLU:
// a pre-activation within bf16 rounding of zero flips its mask on one side only, which is an O(|dy|) difference in a
// handful of elements (visible to max-abs, irrelevant to training) -- single layers keep the strict max-abs check.
static void parity(const std::function<std::unique_ptr<Layer>()>& make, const Shape& xs, double tol, bool train = true, bool l2 = false) {
  const Device g = need_gpu();
  auto lc = make(), lg = make();
  lc->init(getHost(), 5);
  lg->init(g, 5);
  // same parameters on both sides
  auto cl = lc->all_layers(), gl = lg->all_layers();
  for (size_t i = 0; i < cl.size(); ++i)
    for (auto& p : cl[i]->params()) gl[i]->param(p.first)->copy_from(*p.second);
  for (Layer* l : gl) l->mark_params_dirty();
  lc->set_training(train);
  lg->set_training(train);
  Tensor x = make_tensor<float>(xs);
  x->fill_random_normal(0, 1, 21);
  // bf16 inputs on both sides so that the comparison measures the kernels, not the input rounding
  Tensor xb = dtype_cast(*dtype_cast(*x, DType_t::BF16), DType_t::FP32);
  Tensor yc = lc->forward(xb, 0);
  Tensor yg = lg->forward(xb->to_device(g), 0);
  const double scale = std::max(1.0, yc->max() - yc->min());
  const double d = max_abs_diff(*yc, *yg);
  if (!(d <= tol * scale)) throw std::runtime_error(lc->type() + ": forward differs by " + std::to_string(d) + " (scale " + std::to_string(scale) + ")");
  if (!train) return;
  Tensor w = make_tensor<float>(yc->shape());
  w->fill_random_normal(0, 1, 22);
  Tensor wb = dtype_cast(*dtype_cast(*w, DType_t::BF16), DType_t::FP32);
  Tensor dc = lc->backward(wb, 0);
  Tensor dg = lg->backward(wb->to_device(g), 0);
  side_flow::join(g);
  if (dc && dg && l2) {
    const double r = rel_l2(*dc, *dg);
    if (!(r <= tol)) throw std::runtime_error(lc->type() + ": dx relative L2 error " + std::to_string(r));
  } else if (dc && dg) {
    const double ds = std::max(1.0, dc->max() - dc->min()), dd = max_abs_diff(*dc, *dg);
    if (!(dd <= tol * ds)) throw std::runtime_error(lc->type() + ": dx differs by " + std::to_string(dd) + " (scale " + std::to_string(ds) + ")");
  }
  for (size_t i = 0; i < cl.size(); ++i)
    for (auto& gr : cl[i]->grads()) {
      if (l2) {
        const double r = rel_l2(*gr.second, *gl[i]->grad(gr.first));
        if (!(r <= 2 * tol)) throw std::runtime_error(cl[i]->name() + "." + gr.first + ": gradient relative L2 error " + std::to_string(r));
        continue;
      }
      const double gs = std::max(1.0, gr.second->max() - gr.second->min()), gd = max_abs_diff(*gr.second, *gl[i]->grad(gr.first));
      if (!(gd <= 3 * tol * gs)) throw std::runtime_error(cl[i]->name() + "." + gr.first + ": gradient differs by " + std::to_string(gd) + " (scale " + std::to_string(gs) + ")");
    }
}
TEST(gpu_parity_dense_conv) {
  parity([]() -> std::unique_ptr<Layer> { return std::make_unique<DenseLayer>(64, 48); }, {40, 64}, 2e-2);
  parity([]() -> std::unique_ptr<Layer> { return std::make_unique<DenseLayer>(512, 100); }, {256, 512}, 2e-2);            // classifier head: N % 8 != 0
  parity([]() -> std::unique_ptr<Layer> { return std::make_unique<Conv2DLayer>(16, 32, 3, 3, 1, 1, 1, 1); }, {4, 16, 16, 16}, 2e-2);
  parity([]() -> std::unique_ptr<Layer> { return std::make_unique<Conv2DLayer>(32, 64, 3, 3, 2, 2, 1, 1); }, {4, 16, 16, 32}, 2e-2);
  parity([]() -> std::unique_ptr<Layer> { return std::make_unique<Conv2DLayer>(32, 64, 1, 1, 2, 2, 0, 0, false); }, {4, 16, 16, 32}, 2e-2);
  parity([]() -> std::unique_ptr<Layer> { return std::make_unique<Conv2DLayer>(3, 16, 3, 3, 1, 1, 1, 1); }, {8, 32, 32, 3}, 2e-2);      // direct first-layer kernel
  parity([]() -> std::unique_ptr<Layer> { return std::make_unique<Conv2DLayer>(8, 16, 5, 5, 1, 1, 0, 0, false); }, {2, 12, 12, 8}, 2e-2);   // im2col path
  parity([]() -> std::unique_ptr<Layer> { return std::make_unique<Conv2DLayer>(8, 16, 3, 2, 2, 1, 1, 0); }, {2, 13, 10, 8}, 2e-2);         // anisotropic stride / padding
}
TEST(gpu_parity_norms_pools_acts) {
  parity([]() -> std::unique_ptr<Layer> { return std::make_unique<BatchNormLayer>(32, 1e-5, 0.1, true, true); }, {8, 8, 8, 32}, 3e-2);
  parity([]() -> std::unique_ptr<Layer> { return std::make_unique<BatchNormLayer>(64, 1e-5, 0.1, true, false); }, {4, 16, 16, 64}, 3e-2);
  // legacy_batchnorm: NCHW in its own layout on the GPU (norm_nchw.cu), vector (H*W % 8 == 0) and scalar planes
  parity([]() -> std::unique_ptr<Layer> { return std::make_unique<LegacyNCHWLayer>(std::make_unique<BatchNormLayer>(32, 1e-5, 0.1, true, false, "lbn"), "legacy_batchnorm"); }, {8, 32, 12, 12}, 3e-2);
  parity([]() -> std::unique_ptr<Layer> { return std::make_unique<LegacyNCHWLayer>(std::make_unique<BatchNormLayer>(24, 1e-5, 0.1, true, false, "lbn"), "legacy_batchnorm"); }, {5, 24, 7, 7}, 3e-2);
  parity([]() -> std::unique_ptr<Layer> { return std::make_unique<LayerNormLayer>(256); }, {6, 10, 256}, 3e-2);
  parity([]() -> std::unique_ptr<Layer> { return std::make_unique<GroupNormLayer>(4, 32); }, {2, 8, 8, 32}, 3e-2);
  parity([]() -> std::unique_ptr<Layer> { return std::make_unique<MaxPool2DLayer>(2, 2); }, {2, 8, 8, 16}, 1e-2);
  parity([]() -> std::unique_ptr<Layer> { return std::make_unique<AvgPool2DLayer>(8, 8, 1, 1); }, {4, 8, 8, 32}, 1e-2);
  for (const char* a : {"relu", "leaky_relu", "sigmoid", "tanh", "elu", "gelu"}) parity([a]() -> std::unique_ptr<Layer> { return std::make_unique<ActivationLayer>(a); }, {4, 64}, 2e-2);
}
TEST(gpu_parity_blocks) {
  std::string failures;
  auto run = [&](const char* what, const std::function<std::unique_ptr<Layer>()>& mk, const Shape& xs, double tol) {
    try {
      parity(mk, xs, tol, true, std::string(what).find("basic block") == 0);
    } catch (const Skip&) {
      throw;
    } catch (const std::exception& e) {
      failures += std::string("\n    ") + what + ": " + e.what();
    }
  };
  run("bn+relu -> conv3x3 (sequential)", []() -> std::unique_ptr<Layer> { LayerBuilder b({8, 8, 32}); b.batchnorm(-1, 0.1, true, true, "bn").conv2d(32, 3, 3, 1, 1, 1, 1, true, "c"); return b.build_sequential("s"); }, {4, 8, 8, 32}, 5e-2);
  run("wide block 32->32 stride 1 (identity shortcut)", []() -> std::unique_ptr<Layer> { LayerBuilder b({8, 8, 32}); b.wide_residual_block(32, 32, 1, 0.0, "w"); return b.build_sequential("wrn_block"); }, {4, 8, 8, 32}, 5e-2);
  run("wide block 32->64 stride 1 (1x1 shortcut)", []() -> std::unique_ptr<Layer> { LayerBuilder b({8, 8, 32}); b.wide_residual_block(32, 64, 1, 0.0, "w"); return b.build_sequential("wrn_block"); }, {4, 8, 8, 32}, 5e-2);
  run("wide block 32->64 stride 2", []() -> std::unique_ptr<Layer> { LayerBuilder b({16, 16, 32}); b.wide_residual_block(32, 64, 2, 0.0, "w"); return b.build_sequential("wrn_block"); }, {4, 16, 16, 32}, 5e-2);
  run("basic block 32->32", []() -> std::unique_ptr<Layer> { LayerBuilder b({8, 8, 32}); b.basic_residual_block(32, 32, 1, "r"); return b.build_sequential("res_block"); }, {4, 8, 8, 32}, 5e-2);
  run("basic block 32->64 stride 2", []() -> std::unique_ptr<Layer> { LayerBuilder b({8, 8, 32}); b.basic_residual_block(32, 64, 2, "r"); return b.build_sequential("res_block"); }, {4, 8, 8, 32}, 5e-2);
  run("attention block", []() -> std::unique_ptr<Layer> { return std::make_unique<AttentionBlock>(128, 2, true); }, {2, 64, 128}, 4e-2);
  run("gpt block", []() -> std::unique_ptr<Layer> { LayerBuilder b({64, 128}); b.gpt_block(128, 2, 512, 0.0, true, "gelu", "h0"); return b.build_sequential("gpt"); }, {2, 64, 128}, 5e-2);
  if (!failures.empty()) throw std::runtime_error(failures);
}
TEST(gpu_train_step_graph_matches_eager_and_learns) {
  const Device g = need_gpu();
  auto run = [&](bool use_graph) {
    Graph gr = load_or_create_model("cifar10_resnet9", "", g, 3);
    auto opt = OptimizerFactory::create_adam(1e-3, 0.9, 0.999, 1e-3, 3e-4);
    auto crit = LossFactory::create_logsoftmax_crossentropy();
    TrainStep step(gr, *crit, *opt, nullptr, 1, use_graph);
    SyntheticDataLoader data({32, 32, 3}, 10, 256, 9);
    data.set_pinned(true);
    Tensor x, y;
    data.get_batch(64, x, y);
    std::vector<float> losses;
    for (int i = 0; i < 12; ++i) {
      Tensor st = step(x, y);
      float h[2];
      st->copy_to_host(h, 8);
      losses.push_back(h[0]);
    }
    CHECK(use_graph == step.captured());
    return losses;
  };
  const auto eager = run(false), graph = run(true);
  CHECK(eager.back() < eager.front() * 0.7 && graph.back() < graph.front() * 0.7);
  for (size_t i = 0; i < eager.size(); ++i) CHECK_NEAR(eager[i], graph[i], 0.05 * std::max(1.f, eager[i]));
}

TEST(gpu_gpt2_engine_bf16_fp8_and_graph_decode) {
  const Device g = need_gpu();
  auto mc = tiny_gpt(211, 128, 2, 2, 160), mg = tiny_gpt(211, 128, 2, 2, 160);
  mc->init(getHost(), 21);
  mg->init(g, 21);
  {                                               // same weights on both devices
    auto a = mc->all_layers(), b = mg->all_layers();
    for (size_t i = 0; i < a.size(); ++i)
      for (auto& kv : b[i]->params()) kv.second->copy_from(*a[i]->param(kv.first));
    for (Layer* l : b) l->mark_params_dirty();
  }
  const size_t B = 3, L = 130, V = 211;
  std::vector<int32_t> ids = toy_ids(B * L, V, 3);
  GPT2Engine cpu(*mc, B), gpu(*mg, B);
  std::vector<float> lc, lg, l8;
  std::vector<int32_t> tc = cpu.prefill(ids, B, L, &lc), tg = gpu.prefill(ids, B, L, &lg);
  double num = 0, den = 0;
  for (size_t i = 0; i < lc.size(); ++i) { num += (lc[i] - lg[i]) * (lc[i] - lg[i]); den += lc[i] * lc[i]; }
  CHECK(std::sqrt(num / den) < 4e-2);             // bf16 kernels vs the fp32 oracle
  // eager decode == graph decode of the multi-kernel path (same kernels, position from device memory)
  std::vector<int32_t> eager_tok = tg, eager;
  std::vector<float> eager_l1;
  for (int s = 0; s < 5; ++s) { eager_tok = gpu.decode(eager_tok, s == 0 ? &eager_l1 : nullptr); eager.insert(eager.end(), eager_tok.begin(), eager_tok.end()); }
  setenv("TNN_DECODE_MEGA", "0", 1);
  gpu.prefill(ids, B, L);
  std::vector<int32_t> graph = gpu.decode_graph(tg, 5);
  CHECK(!gpu.uses_persistent_decode());
  for (size_t b = 0; b < B; ++b)
    for (size_t s = 0; s < 5; ++s) CHECK(graph[b * 5 + s] == eager[s * B + b]);
  CHECK(gpu.cur_len() == L + 5);
  // a second graph run after a new prefill replays the same captured graph
  gpu.prefill(ids, B, L);
  std::vector<int32_t> graph2 = gpu.decode_graph(tg, 5);
  CHECK(graph2 == graph);
  // the persistent whole-model decode kernel: one launch per token.  Same math, different reduction order: compare
  // the first step's logits with the eager path, then check the bookkeeping over several steps and a re-run.
  unsetenv("TNN_DECODE_MEGA");
  GPT2Engine mega(*mg, B);
  mega.prefill(ids, B, L);
  std::vector<int32_t> m1 = mega.decode_graph(tg, 1);
  CHECK(mega.uses_persistent_decode() && mega.cur_len() == L + 1);
  std::vector<float> ml = mega.last_logits();
  num = den = 0;
  for (size_t i = 0; i < ml.size(); ++i) { num += (ml[i] - eager_l1[i]) * (ml[i] - eager_l1[i]); den += eager_l1[i] * eager_l1[i]; }
  CHECK(std::sqrt(num / den) < 2e-2);
  size_t agree = 0;
  for (size_t b = 0; b < B; ++b) agree += m1[b] == eager[b];
  CHECK(agree + 1 >= B);
  mega.prefill(ids, B, L);
  std::vector<int32_t> m5 = mega.decode_graph(tg, 5);
  CHECK(mega.cur_len() == L + 5);
  for (size_t b = 0; b < B; ++b) CHECK(m5[b * 5] == m1[b]);
  mega.prefill(ids, B, L);
  CHECK(mega.decode_graph(tg, 5) == m5);
  // continuing from the cache the persistent kernel appended: eager decode of the next token agrees with a fresh run
  std::vector<int32_t> lastcol(B);
  for (size_t b = 0; b < B; ++b) lastcol[b] = m5[b * 5 + 4];
  std::vector<float> cont_l;
  mega.decode(lastcol, &cont_l);
  {
    std::vector<int32_t> grown;
    for (size_t b = 0; b < B; ++b) {
      grown.insert(grown.end(), ids.begin() + (long)(b * L), ids.begin() + (long)((b + 1) * L));
      grown.push_back(tg[b]);
      for (size_t s = 0; s < 5; ++s) grown.push_back(m5[b * 5 + s]);
    }
    std::vector<float> ref;
    GPT2Engine fresh(*mg, B);
    fresh.prefill(grown, B, L + 6, &ref);
    num = den = 0;
    for (size_t i = 0; i < ref.size(); ++i) { num += (ref[i] - cont_l[i]) * (ref[i] - cont_l[i]); den += ref[i] * ref[i]; }
    CHECK(std::sqrt(num / den) < 3e-2);
  }
  // block-scaled fp8 prefill against bf16
  GPT2Engine f8(*mg, B, 0, "fp8");
  f8.prefill(ids, B, L, &l8);
  num = den = 0;
  for (size_t i = 0; i < lg.size(); ++i) { num += (l8[i] - lg[i]) * (l8[i] - lg[i]); den += lg[i] * lg[i]; }
  CHECK(std::sqrt(num / den) < 0.12);
  CHECK(f8.precision() == "fp8");
  // the fp8 QKV epilogue wrote K / V straight into the caches: a decode step over them must agree with the bf16 engine
  {
    gpu.prefill(ids, B, L);
    std::vector<float> db, d8;
    gpu.decode(tg, &db);
    f8.decode(tg, &d8);
    num = den = 0;
    for (size_t i = 0; i < db.size(); ++i) { num += (d8[i] - db[i]) * (d8[i] - db[i]); den += db[i] * db[i]; }
    CHECK(std::sqrt(num / den) < 0.12);
  }
  {
    // persistent decode over the fp8 KV cache (opt-in: written by the QKV epilogue at prefill, appended per token)
    setenv("TNN_KV_FP8", "1", 1);
    GPT2Engine f8(*mg, B, 0, "fp8");
    unsetenv("TNN_KV_FP8");
    std::vector<float> db;
    gpu.prefill(ids, B, L);
    gpu.decode(tg, &db);
    f8.prefill(ids, B, L);
    std::vector<int32_t> g8 = f8.decode_graph(tg, 1);
    CHECK(f8.uses_persistent_decode() && f8.uses_fp8_kv_cache() && f8.cur_len() == L + 1);
    std::vector<float> l8kv = f8.last_logits();
    num = den = 0;
    for (size_t i = 0; i < db.size(); ++i) { num += (l8kv[i] - db[i]) * (l8kv[i] - db[i]); den += db[i] * db[i]; }
    CHECK(std::sqrt(num / den) < 0.15);
    // three more tokens through the fp8 cache, then an eager step over the (always complete) bf16 cache still works
    std::vector<int32_t> more = f8.decode_graph(g8, 3);
    CHECK(f8.cur_len() == L + 4 && more.size() == B * 3);
    std::vector<int32_t> lastc(B);
    for (size_t b = 0; b < B; ++b) lastc[b] = more[b * 3 + 2];
    std::vector<float> cont;
    f8.decode(lastc, &cont);
    for (float v : cont) CHECK(std::isfinite(v));
    // after the eager step the fp8 cache is stale: the next graph decode falls back to the bf16 cache
    f8.decode_graph(lastc, 1);
    CHECK(!f8.uses_fp8_kv_cache());
  }
}

TEST(gpu_pipeline_in_process_and_tcp_equal_single_device) {
  pipeline_case(0, need_gpu(), 5e-2);      // device tensors by reference
  pipeline_case(1, need_gpu(), 5e-2);      // host-staged through the TCP bus
}
TEST(gpu_roce_pipeline_moves_activations_device_to_device) {
  // two stages on the GPU behind RoCECommunicators: activations / gradients are written device to device into the
  // peer's registered device slab (same process: the slab pointer; other processes: its CUDA-IPC handle)
  pipeline_case(2, need_gpu(), 5e-2);
}
static void dp2_case(int symmetric) {
  need_gpu();
  if (DeviceManager::getInstance().num_gpus() < 2) throw Skip("needs 2 GPUs");
  // BatchNorm-free model: the two-replica run on half batches must follow the single-GPU run on the full batch
  auto make = []() -> std::unique_ptr<Layer> {
    LayerBuilder b({8, 8, 16});
    b.conv2d(32, 3, 3, 1, 1, 1, 1, true, "c1").activation("relu", "r1").conv2d(32, 3, 3, 2, 2, 1, 1, true, "c2").activation("relu", "r2").flatten(1, -1, "f")
        .dense(64, true, "d1").activation("relu", "r3").dense(10, true, "head");
    return b.build_sequential("dp_test_net");
  };
  const size_t B = 64;
  SyntheticDataLoader data({8, 8, 16}, 10, 4 * B, 5);
  data.set_pinned(true);
  Vec<Tensor> fx, fy;
  for (int k = 0; k < 4; ++k) {
    Tensor x, y;
    data.get_batch(B, x, y);
    fx.push_back(x);
    fy.push_back(y);
  }
  Adam proto(1e-3, 0.9, 0.999, 1e-8, 0.0, false);
  std::vector<float> single;
  {
    Graph g = graph_from_sequential(make(), getGPU(0), 77, "single");
    auto opt = proto.clone();
    auto crit = LossFactory::create_logsoftmax_crossentropy();
    TrainStep step(g, *crit, *opt, nullptr, 1, true);
    for (int i = 0; i < 10; ++i) {
      float h[2];
      step(fx[(size_t)i % 4], fy[(size_t)i % 4])->copy_to_host(h, 8);
      single.push_back(h[0]);
    }
  }
  DataParallel dp(make, 2, 77, proto, "logsoftmax_crossentropy", true, symmetric);
  CHECK(dp.symmetric() == (symmetric == 1));
  if (symmetric == 1) std::cout << "  symmetric group: multicast window " << (dp.multicast() ? "bound (NVLS)" : "unavailable (peer mappings only)") << std::endl;
  CHECK(dp.max_replica_divergence() == 0.0);
  for (int i = 0; i < 10; ++i) {
    Vec<Tensor> xs, ys;
    for (size_t r = 0; r < 2; ++r) {
      xs.push_back(fx[(size_t)i % 4]->slice_batch(r * B / 2, B / 2));
      ys.push_back(fy[(size_t)i % 4]->slice_batch(r * B / 2, B / 2));
    }
    const auto res = dp.step(xs, ys);
    CHECK_NEAR(res.first, single[(size_t)i], 0.03 * std::max(1.0, (double)single[(size_t)i]));
  }
  CHECK(dp.captured());
  CHECK(dp.max_replica_divergence() == 0.0);          // every replica received the same broadcast parameters
  CHECK(single.back() < single.front());
}
TEST(dp2_fused_step_matches_single_device) { dp2_case(0); }
TEST(data_parallel_shard_layout_gather_scatter) {
  // host-side layout of the sharded optimizer state (DataParallel::save_checkpoint / load_checkpoint): n = 10 elements,
  // 3 ranks, chunk 4 -> rank 2 owns only [8, 10); the gathered vector is padded to the slab size
  const size_t n = 10, chunk = 4, total = 16;
  std::vector<std::vector<float>> shards = {{0, 1, 2, 3}, {4, 5, 6, 7}, {8, 9, -1, -1}};
  const auto full = dp_gather_shards(shards, n, chunk, total);
  CHECK(full.size() == total);
  for (size_t i = 0; i < total; ++i) CHECK(full[i] == (i < n ? (float)i : 0.f));          // rank 2's padding is not state
  // the same state on 2 ranks (chunk 8) and back on 3
  const auto two = dp_scatter_shards(full, n, 8, 2);
  CHECK(two.size() == 2 && two[0].size() == 8 && two[1][0] == 8.f && two[1][1] == 9.f && two[1][2] == 0.f);
  CHECK(dp_gather_shards(two, n, 8, total) == full);
  const auto three = dp_scatter_shards(full, n, chunk, 3);
  CHECK(three[2][0] == 8.f && three[2][1] == 9.f && three[2][2] == 0.f && three[0] == shards[0] && three[1] == shards[1]);
  bool threw = false;
  try {
    dp_scatter_shards(full, n, 4, 2);                   // 2 x 4 < 10
  } catch (const std::invalid_argument&) {
    threw = true;
  }
  CHECK(threw);
  threw = false;
  try {
    dp_gather_shards({{0, 1}, {4, 5, 6, 7}, {8, 9}}, n, chunk, total);      // rank 0's buffer is shorter than its range
  } catch (const std::invalid_argument&) {
    threw = true;
  }
  CHECK(threw);
}
TEST(data_parallel_checkpoint_resumes_sharded_and_single) {
  // written after the round's GPU budget was spent: not yet run on hardware (needs 2 GPUs; the driver's GPU tier has one)
  need_gpu();
  if (DeviceManager::getInstance().num_gpus() < 2) throw Skip("needs 2 GPUs");
  auto make = []() -> std::unique_ptr<Layer> {
    LayerBuilder b({8, 8, 16});
    b.conv2d(32, 3, 3, 1, 1, 1, 1, true, "c1").activation("relu", "r1").flatten(1, -1, "f").dense(10, true, "head");
    return b.build_sequential("dp_ckpt_net");
  };
  const size_t B = 32;
  SyntheticDataLoader data({8, 8, 16}, 10, 2 * B, 5);
  data.set_pinned(true);
  Tensor x, y;
  data.get_batch(B, x, y);
  auto halves = [&](Vec<Tensor>* xs, Vec<Tensor>* ys) {
    for (size_t r = 0; r < 2; ++r) {
      xs->push_back(x->slice_batch(r * B / 2, B / 2));
      ys->push_back(y->slice_batch(r * B / 2, B / 2));
    }
  };
  Adam proto(1e-3, 0.9, 0.999, 1e-8, 0.0, false);
  const std::string path = "/tmp/tnn_tests_dp_ckpt";
  std::vector<double> cont;
  {
    DataParallel dp(make, 2, 5, proto);
    for (int i = 0; i < 6; ++i) {
      Vec<Tensor> xs, ys;
      halves(&xs, &ys);
      dp.step(xs, ys);
    }
    dp.save_checkpoint(path);
    CHECK(dp.step_count() == 6);
    for (int i = 0; i < 3; ++i) {
      Vec<Tensor> xs, ys;
      halves(&xs, &ys);
      cont.push_back(dp.step(xs, ys).first);
    }
  }
  {
    DataParallel dp(make, 2, 99, proto);               // different seed: everything must come from the file
    dp.load_checkpoint(path);
    CHECK(dp.step_count() == 6 && dp.max_replica_divergence() == 0.0);
    for (int i = 0; i < 3; ++i) {
      Vec<Tensor> xs, ys;
      halves(&xs, &ys);
      CHECK_NEAR(dp.step(xs, ys).first, cont[(size_t)i], 2e-3 * std::max(1.0, cont[(size_t)i]));
    }
  }
  {
    // the same file resumes single-GPU training: full-length moments, the single-device optimizer's tensor names
    auto opt = proto.clone();
    uint64_t step_no = 0;
    Graph g = load_checkpoint(path, getGPU(0), opt.get(), nullptr, &step_no);
    CHECK(step_no == 6 && opt->step_count == 6);
    auto crit = LossFactory::create_logsoftmax_crossentropy();
    TrainStep step(g, *crit, *opt, nullptr, 1, false);
    float h[2];
    step(x, y)->copy_to_host(h, 8);
    CHECK_NEAR(h[0], cont[0], 0.03 * std::max(1.0, cont[0]));
  }
}
// The two cases below exercise the framework's own VMM / multicast group on real GPUs.  They were written after the
// round's GPU budget was spent and have not run on hardware yet, so they are opt-in (TNN_TEST_NVLS=1) instead of being able
// to turn the round-end GPU suite red.
static void need_nvls_opt_in() {
  const char* e = std::getenv("TNN_TEST_NVLS");
  if (!(e && e[0] == '1')) throw Skip("opt-in: TNN_TEST_NVLS=1 (not yet run on hardware)");
}
TEST(own_symmetric_memory_data_parallel_step) {
  need_nvls_opt_in();
  dp2_case(1);
}
TEST(symmetric_memory_peer_and_multicast_mappings_on_devices) {
  need_nvls_opt_in();
  need_gpu();
  const int W = std::min(DeviceManager::getInstance().num_gpus(), 4);
  if (W < 2) throw Skip("needs 2 GPUs");
  std::vector<int> devs;
  for (int r = 0; r < W; ++r) devs.push_back(r);
  auto group = std::make_shared<SymmetricMemory>(devs, (size_t)8 << 20, true);
  CHECK(group->bytes() >= ((size_t)8 << 20) && group->bytes() % group->granularity() == 0);
  std::vector<std::unique_ptr<SymmetricAllocator>> al;
  std::vector<float*> blk;
  const size_t n = 1 << 16;
  for (int r = 0; r < W; ++r) {
    al.push_back(std::make_unique<SymmetricAllocator>(group, r));
    blk.push_back(static_cast<float*>(al.back()->allocate(n * 4)));
    CHECK(group->offset_of(r, blk.back()) == group->offset_of(0, blk[0]));
  }
  // rank r fills its block with r + 1 from ITS device; rank 0's device then reads every block through the peer mappings
  for (int r = 0; r < W; ++r) {
    std::vector<float> h(n, (float)(r + 1));
    TNN_CUDA(cudaSetDevice(r));
    TNN_CUDA(cudaMemcpy(blk[(size_t)r], h.data(), n * 4, cudaMemcpyHostToDevice));
    TNN_CUDA(cudaDeviceSynchronize());
  }
  TNN_CUDA(cudaSetDevice(0));
  for (int r = 0; r < W; ++r) {
    std::vector<float> h(n, 0.f);
    TNN_CUDA(cudaMemcpy(h.data(), al[0]->peer(blk[0], r), n * 4, cudaMemcpyDeviceToHost));
    CHECK(h[0] == (float)(r + 1) && h[n - 1] == (float)(r + 1));
  }
  std::cout << "  multicast window: " << (group->has_multicast() ? "bound" : "unavailable") << std::endl;
  if (group->has_multicast()) CHECK(al[1]->multicast(blk[1]) == static_cast<char*>(group->multicast_base()) + group->offset_of(0, blk[0]));
  for (int r = 0; r < W; ++r) al[(size_t)r]->deallocate(blk[(size_t)r]);
  CHECK(group->free_list(0).in_use() == 0);
}

int main(int argc, char** argv) {
  const std::string filter = argc > 1 ? argv[1] : "";
  int passed = 0, failed = 0, skipped = 0;
  for (auto& t : registry()) {
    if (!filter.empty() && std::string(t.name).find(filter) == std::string::npos) continue;
    try {
      t.fn();
      std::cout << "[ OK      ] " << t.name << std::endl;
      ++passed;
    } catch (const Skip& s) {
      std::cout << "[ SKIPPED ] " << t.name << " (" << s.why << ")" << std::endl;
      ++skipped;
    } catch (const std::exception& e) {
      std::cout << "[ FAILED  ] " << t.name << ": " << e.what() << std::endl;
      ++failed;
    }
  }
  std::cout << passed << " passed, " << failed << " failed, " << skipped << " skipped" << std::endl;
  return failed ? 1 : 0;
}
