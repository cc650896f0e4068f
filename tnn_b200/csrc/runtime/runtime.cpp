// pybind11 surface of the C++ host runtime: archiver (wire/checkpoint format), offset
// allocators, native data path, tokenizer, profiler event store.
#include <pybind11/pybind11.h>
namespace tnn {
void runtime_bind(pybind11::module_& m);
}

#include <pybind11/numpy.h>
#include <pybind11/stl.h>

#include <chrono>
#include <cmath>
#include <memory>
#include <mutex>

#include <cuda_runtime.h>

#include <cstring>

#include "allocator.hpp"
#include "cpu_ops.hpp"
#include "planner.hpp"
#include "archiver.hpp"
#include "cpu_math.hpp"
#include "data_loader.hpp"
#include "symmetric_memory.hpp"
#include "tcp_bus.hpp"
#include "tokenizer.hpp"

namespace py = pybind11;

namespace tnn {

// ------------------------------------------------------------------------------ profiler
// Reference: include/profiling/event.hpp:11-62, profiler.hpp:13-157.
struct ProfEvent {
  int type;  // 0 COMPUTE, 1 COMMUNICATION, 2 OTHER
  int64_t start_ns, end_ns;
  std::string name, source;
};
class Profiler {
 public:
  void init_start_time() {
    start_ns_ = now_ns();
  }
  static int64_t now_ns() {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(
               std::chrono::system_clock::now().time_since_epoch())
        .count();
  }
  void add_event(int type, int64_t s, int64_t e, const std::string& name, const std::string& src) {
    std::lock_guard<std::mutex> g(mu_);
    events_.push_back(ProfEvent{type, s, e, name, src});
  }
  // merge another profiler's events, re-basing them by the difference of the start times
  void merge(const Profiler& o) {
    std::lock_guard<std::mutex> g(mu_);
    const int64_t delta = start_ns_ - o.start_ns_;
    for (const auto& ev : o.events_) events_.push_back(ProfEvent{ev.type, ev.start_ns + delta, ev.end_ns + delta, ev.name, ev.source});
  }
  void sort() {
    std::lock_guard<std::mutex> g(mu_);
    std::sort(events_.begin(), events_.end(), [](const ProfEvent& a, const ProfEvent& b) { return a.start_ns < b.start_ns; });
  }
  void clear() {
    std::lock_guard<std::mutex> g(mu_);
    events_.clear();
  }
  py::bytes serialize() const {
    Writer w;
    w.prim<int64_t>(start_ns_);
    w.prim<uint64_t>(events_.size());
    for (const auto& e : events_) {
      w.prim<uint8_t>((uint8_t)e.type);
      w.prim<int64_t>(e.start_ns);
      w.prim<int64_t>(e.end_ns);
      w.str(e.name);
      w.str(e.source);
    }
    return py::bytes(reinterpret_cast<const char*>(w.buffer().data()), w.size());
  }
  static std::shared_ptr<Profiler> deserialize(const std::string& b) {
    auto p = std::make_shared<Profiler>();
    Reader r(reinterpret_cast<const uint8_t*>(b.data()), b.size());
    p->start_ns_ = r.prim<int64_t>();
    uint64_t n = r.prim<uint64_t>();
    for (uint64_t i = 0; i < n; ++i) {
      ProfEvent e;
      e.type = r.prim<uint8_t>();
      e.start_ns = r.prim<int64_t>();
      e.end_ns = r.prim<int64_t>();
      e.name = r.str();
      e.source = r.str();
      p->events_.push_back(std::move(e));
    }
    return p;
  }
  std::vector<py::tuple> events() const {
    std::vector<py::tuple> out;
    for (const auto& e : events_) out.push_back(py::make_tuple(e.type, e.start_ns, e.end_ns, e.name, e.source));
    return out;
  }
  int64_t start_ns() const { return start_ns_; }

 private:
  int64_t start_ns_ = 0;
  std::vector<ProfEvent> events_;
  mutable std::mutex mu_;
};

static const uint8_t* as_ptr(int64_t p) { return reinterpret_cast<const uint8_t*>(p); }

void runtime_bind(py::module_& m) {
  // ---------------------------------------------------------------- archiver
  py::class_<Sizer>(m, "Sizer")
      .def(py::init<>())
      .def("size", &Sizer::size)
      .def("u8", [](Sizer& s, uint8_t v) { s.prim(v); })
      .def("u16", [](Sizer& s, uint16_t v) { s.prim(v); })
      .def("u32", [](Sizer& s, uint32_t v) { s.prim(v); })
      .def("u64", [](Sizer& s, uint64_t v) { s.prim(v); })
      .def("i64", [](Sizer& s, int64_t v) { s.prim(v); })
      .def("f32", [](Sizer& s, float v) { s.prim(v); })
      .def("f64", [](Sizer& s, double v) { s.prim(v); })
      .def("boolean", [](Sizer& s, bool v) { s.prim<uint8_t>(v); })
      .def("string", [](Sizer& s, const std::string& v) { s.str(v); })
      .def("sizes", [](Sizer& s, const std::vector<uint64_t>& v) { s.sizes(v); })
      .def("raw", [](Sizer& s, const py::bytes& b) { s.bytes(nullptr, std::string(b).size()); })
      .def("tensor", [](Sizer& s, uint32_t dt, const std::vector<uint64_t>& shape, uint64_t nbytes) { s.tensor(dt, shape, nbytes); });
  py::class_<Writer>(m, "Writer")
      .def(py::init<size_t>(), py::arg("reserve") = 0)
      .def("size", &Writer::size)
      .def("u8", [](Writer& w, uint8_t v) { w.prim(v); })
      .def("u16", [](Writer& w, uint16_t v) { w.prim(v); })
      .def("u32", [](Writer& w, uint32_t v) { w.prim(v); })
      .def("u64", [](Writer& w, uint64_t v) { w.prim(v); })
      .def("i64", [](Writer& w, int64_t v) { w.prim(v); })
      .def("f32", [](Writer& w, float v) { w.prim(v); })
      .def("f64", [](Writer& w, double v) { w.prim(v); })
      .def("boolean", [](Writer& w, bool v) { w.prim<uint8_t>(v); })
      .def("string", [](Writer& w, const std::string& v) { w.str(v); })
      .def("sizes", [](Writer& w, const std::vector<uint64_t>& v) { w.sizes(v); })
      .def("raw", [](Writer& w, const py::bytes& b) { std::string s = b; w.bytes(s.data(), s.size()); })
      .def("raw_ptr", [](Writer& w, int64_t ptr, uint64_t n) { w.bytes(as_ptr(ptr), n); })
      .def("tensor", [](Writer& w, uint32_t dt, const std::vector<uint64_t>& shape, int64_t ptr, uint64_t nbytes) {
        w.tensor(dt, shape, as_ptr(ptr), nbytes);
      })
      .def("packet_header", [](Writer& w, uint8_t type, uint64_t packet_len, uint64_t msg_len, uint64_t serial,
                               uint64_t offset, uint32_t total) {
        PacketHeader h;
        h.type = type;
        h.endianness = host_is_little_endian() ? 1 : 0;
        h.packet_length = packet_len;
        h.msg_length = msg_len;
        h.msg_serial_id = serial;
        h.packet_offset = offset;
        h.total_packets = total;
        h.write(w);
      })
      .def("getvalue", [](const Writer& w) { return py::bytes(reinterpret_cast<const char*>(w.buffer().data()), w.size()); });
  struct PyReader {
    py::bytes data;   // keeps the buffer alive; the reader parses it in place (no copy)
    std::unique_ptr<Reader> r;
  };
  py::class_<PyReader>(m, "Reader")
      .def(py::init([](const py::bytes& b, bool swap) {
             auto p = std::make_unique<PyReader>();
             p->data = b;
             char* buf = nullptr;
             Py_ssize_t len = 0;
             if (PyBytes_AsStringAndSize(p->data.ptr(), &buf, &len) != 0) throw py::error_already_set();
             p->r = std::make_unique<Reader>(reinterpret_cast<const uint8_t*>(buf), (size_t)len, swap);
             return p;
           }),
           py::arg("data"), py::arg("swap") = false)
      .def("offset", [](PyReader& r) { return r.r->offset(); })
      .def("remaining", [](PyReader& r) { return r.r->remaining(); })
      .def("u8", [](PyReader& r) { return r.r->prim<uint8_t>(); })
      .def("u16", [](PyReader& r) { return r.r->prim<uint16_t>(); })
      .def("u32", [](PyReader& r) { return r.r->prim<uint32_t>(); })
      .def("u64", [](PyReader& r) { return r.r->prim<uint64_t>(); })
      .def("i64", [](PyReader& r) { return r.r->prim<int64_t>(); })
      .def("f32", [](PyReader& r) { return r.r->prim<float>(); })
      .def("f64", [](PyReader& r) { return r.r->prim<double>(); })
      .def("boolean", [](PyReader& r) { return r.r->prim<uint8_t>() != 0; })
      .def("string", [](PyReader& r) { return py::bytes(r.r->str()); })
      .def("sizes", [](PyReader& r) { return r.r->sizes(); })
      .def("raw", [](PyReader& r, uint64_t n) {
        const uint8_t* p = r.r->bytes(n);
        return py::bytes(reinterpret_cast<const char*>(p), n);
      })
      .def("raw_into", [](PyReader& r, int64_t dst, uint64_t n, uint32_t elem_size) {
        const uint8_t* p = r.r->bytes(n);
        uint8_t* d = reinterpret_cast<uint8_t*>(dst);
        std::memcpy(d, p, n);
        if (r.r->swap() && elem_size > 1)
          for (uint64_t i = 0; i + elem_size <= n; i += elem_size) std::reverse(d + i, d + i + elem_size);
      })
      .def("packet_header", [](PyReader& r) {
        const uint8_t* p = r.r->bytes(PacketHeader::kBytes);
        PacketHeader h = PacketHeader::read(p, PacketHeader::kBytes);
        return py::make_tuple(h.version, h.type, h.endianness, h.packet_length, h.msg_length, h.msg_serial_id,
                              h.packet_offset, h.total_packets, h.compression);
      });
  m.attr("PACKET_HEADER_BYTES") = (int)PacketHeader::kBytes;
  m.def("host_is_little_endian", &host_is_little_endian);

  // ---- CUDA-IPC device slabs: the "registered memory" of the one-sided rendezvous transport when both peers are
  // processes on one node (distributed/rendezvous.py).  The reference registers a host slab with the NIC
  // (include/device/ibv_allocator.hpp:31-266) and RDMA-writes into it; here the slab is cudaMalloc'd device memory,
  // exported with cudaIpcGetMemHandle, and the peer's write is a device-to-device copy over NVLink.
  struct DevGuard {
    int prev = -1;
    explicit DevGuard(int d) {
      cudaGetDevice(&prev);
      if (d != prev) cudaSetDevice(d);
    }
    ~DevGuard() {
      int cur = -1;
      cudaGetDevice(&cur);
      if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
  };
  auto ck = [](cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
  };
  m.def("cuda_ipc_alloc", [ck](int device, size_t bytes) {
    DevGuard g(device);
    void* p = nullptr;
    ck(cudaMalloc(&p, bytes), "cuda_ipc_alloc: cudaMalloc");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
      cudaFree(p);
      ck(e, "cuda_ipc_alloc: cudaIpcGetMemHandle");
    }
    return py::make_tuple((uint64_t)(uintptr_t)p, py::bytes(reinterpret_cast<const char*>(&h), sizeof(h)));
  });
  m.def("cuda_ipc_free", [](int device, uint64_t ptr) {
    DevGuard g(device);
    cudaFree(reinterpret_cast<void*>((uintptr_t)ptr));
  });
  m.def("cuda_ipc_open", [ck](int device, const std::string& handle) {
    if (handle.size() != sizeof(cudaIpcMemHandle_t)) throw std::runtime_error("cuda_ipc_open: handle must be 64 bytes");
    DevGuard g(device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle.data(), sizeof(h));
    void* p = nullptr;
    ck(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cuda_ipc_open: cudaIpcOpenMemHandle");
    return (uint64_t)(uintptr_t)p;
  });
  m.def("cuda_ipc_close", [](int device, uint64_t ptr) {
    DevGuard g(device);
    cudaIpcCloseMemHandle(reinterpret_cast<void*>((uintptr_t)ptr));
  });
  // ---- the framework's own symmetric NVLink memory (symmetric_memory.hpp): cuMemCreate slabs mapped into every rank,
  // an NVLS multicast window over them, handles exchanged as file descriptors.  parallel/dp.py uses it with TNN_DP_OWN_SYMM=1.
  py::class_<OffsetFreeList>(m, "OffsetFreeList")
      .def(py::init<uint64_t, uint64_t>(), py::arg("bytes"), py::arg("align") = 256)
      .def("allocate", [](OffsetFreeList& f, uint64_t n) -> py::object {
        const uint64_t off = f.allocate(n);
        if (off == OffsetFreeList::kNone) return py::none();
        return py::int_(off);
      })
      .def("release", &OffsetFreeList::release)
      .def("block_size", &OffsetFreeList::block_size)
      .def_property_readonly("in_use", &OffsetFreeList::in_use)
      .def_property_readonly("peak", &OffsetFreeList::peak)
      .def_property_readonly("capacity", &OffsetFreeList::capacity)
      .def_property_readonly("largest_free", &OffsetFreeList::largest_free)
      .def_property_readonly("num_free_ranges", &OffsetFreeList::num_free_ranges);
  py::class_<FdMailbox>(m, "FdMailbox")
      .def(py::init<const std::string&, const std::string&>(), py::arg("dir"), py::arg("name"))
      .def_property_readonly("path", &FdMailbox::path)
      .def("send", [](const FdMailbox& b, const std::string& to, uint64_t tag, int fd, int timeout_ms) {
        py::gil_scoped_release nogil;
        b.send(to, tag, fd, timeout_ms);
      }, py::arg("to"), py::arg("tag"), py::arg("fd") = -1, py::arg("timeout_ms") = 30000)
      .def("recv", [](const FdMailbox& b, int timeout_ms) {
        int fd = -1;
        uint64_t tag;
        {
          py::gil_scoped_release nogil;
          tag = b.recv(&fd, timeout_ms);
        }
        return py::make_tuple(tag, fd);
      }, py::arg("timeout_ms") = 30000);
  py::class_<SymmetricMemory>(m, "SymmetricMemory")
      .def_static("probe", [](const std::vector<int>& devices) {
        const auto s = SymmetricMemory::probe(devices);
        py::dict d;
        d["vmm"] = s.vmm;
        d["multicast"] = s.multicast;
        d["why"] = s.why;
        return d;
      })
      .def_static("rendezvous", [](int rank, int world, int device, size_t bytes, const std::string& dir, bool want_multicast, int timeout_ms) {
        py::gil_scoped_release nogil;
        return SymmetricMemory::rendezvous(rank, world, device, bytes, dir, want_multicast, timeout_ms);
      }, py::arg("rank"), py::arg("world"), py::arg("device"), py::arg("bytes"), py::arg("dir"), py::arg("want_multicast") = true, py::arg("timeout_ms") = 60000)
      .def(py::init([](const std::vector<int>& devices, size_t bytes, bool want_multicast) { return std::make_unique<SymmetricMemory>(devices, bytes, want_multicast); }),
           py::arg("devices"), py::arg("bytes"), py::arg("want_multicast") = true)
      .def_property_readonly("world", &SymmetricMemory::world)
      .def_property_readonly("my_rank", &SymmetricMemory::my_rank)
      .def_property_readonly("bytes", &SymmetricMemory::bytes)
      .def_property_readonly("granularity", &SymmetricMemory::granularity)
      .def_property_readonly("has_multicast", &SymmetricMemory::has_multicast)
      .def("base", [](const SymmetricMemory& s, int rank) { return (uint64_t)(uintptr_t)s.base(rank); })
      .def_property_readonly("multicast_base", [](const SymmetricMemory& s) { return (uint64_t)(uintptr_t)s.multicast_base(); })
      .def("allocate", [](SymmetricMemory& s, int rank, size_t n) { return (uint64_t)(uintptr_t)s.allocate(rank, n); })
      .def("deallocate", [](SymmetricMemory& s, int rank, uint64_t p) { s.deallocate(rank, reinterpret_cast<void*>((uintptr_t)p)); });
  // blocking device-to-device copy on a private stream of `device` (either pointer may live on a peer device)
  m.def("cuda_copy_d2d", [ck](int device, uint64_t dst, uint64_t src, size_t bytes) {
    if (bytes == 0) return;
    py::gil_scoped_release nogil;
    DevGuard g(device);
    static std::mutex mu;
    static cudaStream_t streams[64] = {};
    cudaStream_t s;
    {
      std::lock_guard<std::mutex> lk(mu);
      if (!streams[device & 63]) ck(cudaStreamCreateWithFlags(&streams[device & 63], cudaStreamNonBlocking), "cuda_copy_d2d: stream");
      s = streams[device & 63];
    }
    ck(cudaMemcpyAsync(reinterpret_cast<void*>((uintptr_t)dst), reinterpret_cast<const void*>((uintptr_t)src), bytes, cudaMemcpyDefault, s),
       "cuda_copy_d2d: cudaMemcpyAsync");
    ck(cudaStreamSynchronize(s), "cuda_copy_d2d: cudaStreamSynchronize");
  });

  // ---------------------------------------------------------------- CPU math / threading
  m.def("cpu_sgemm", [](int64_t a, int64_t b, int64_t c, int64_t M, int64_t N, int64_t K, bool ta, bool tb, double alpha,
                        double beta, int64_t lda, int64_t ldb, int64_t ldc, int threads) {
    py::gil_scoped_release rel;
    cpu_sgemm(reinterpret_cast<const float*>(a), reinterpret_cast<const float*>(b), reinterpret_cast<float*>(c), M, N, K, ta, tb,
              (float)alpha, (float)beta, lda, ldb, ldc, threads);
  }, py::arg("a"), py::arg("b"), py::arg("c"), py::arg("M"), py::arg("N"), py::arg("K"), py::arg("trans_a") = false,
     py::arg("trans_b") = false, py::arg("alpha") = 1.0, py::arg("beta") = 0.0, py::arg("lda") = 0, py::arg("ldb") = 0,
     py::arg("ldc") = 0, py::arg("threads") = 0);
  m.def("cpu_dot", [](int64_t a, int64_t b, int64_t n) {
    py::gil_scoped_release rel;
    return cpu_dot(reinterpret_cast<const float*>(a), reinterpret_cast<const float*>(b), n);
  });
  // ---------------------------------------------------------------- planner (graph order, slabs, schedules, partitions)
  m.def("plan_topo_order", &planner::topo_order, py::arg("inputs"), py::arg("outputs"));
  m.def("plan_slab", &planner::plan_slab, py::arg("sizes"), py::arg("align"));
  m.def("plan_buckets", &planner::plan_buckets, py::arg("offsets"), py::arg("bucket_elems"));
  m.def("plan_schedule_1f1b", [](int64_t stages, int64_t rank, int64_t mbs) {
    std::vector<std::pair<std::string, int64_t>> out;
    for (const auto& st : planner::schedule_1f1b(stages, rank, mbs)) out.emplace_back(std::string(1, st.op), st.mb);
    return out;
  }, py::arg("stages"), py::arg("rank"), py::arg("microbatches"));
  m.def("plan_schedule_gpipe", [](int64_t mbs) {
    std::vector<std::pair<std::string, int64_t>> out;
    for (const auto& st : planner::schedule_gpipe(mbs)) out.emplace_back(std::string(1, st.op), st.mb);
    return out;
  }, py::arg("microbatches"));
  m.def("plan_peak_in_flight", [](const std::vector<std::pair<std::string, int64_t>>& steps) {
    std::vector<planner::Step> s;
    for (const auto& p : steps) s.push_back({p.first.empty() ? 'F' : p.first[0], p.second});
    return planner::peak_in_flight(s);
  });
  m.def("plan_split_by_proportions", &planner::split_by_proportions, py::arg("total"), py::arg("proportions"),
        py::arg("at_least_one") = true);
  m.def("plan_min_bottleneck_split", &planner::min_bottleneck_split, py::arg("cost"), py::arg("n"),
        py::arg("boundary") = std::vector<double>());

  // CPU elementwise / reduction family (cpu_ops.hpp); pointers are raw data_ptr()s of contiguous fp32 / fp64 tensors
  m.def("cpu_binary", [](int64_t a, int64_t b, double s, int64_t o, int64_t n, int op, bool f64, int threads) {
    py::gil_scoped_release rel;
    if (f64) cpuops::binary<double>((const double*)a, (const double*)b, s, (double*)o, n, op, threads);
    else cpuops::binary<float>((const float*)a, (const float*)b, (float)s, (float*)o, n, op, threads);
  }, py::arg("a"), py::arg("b"), py::arg("scalar"), py::arg("out"), py::arg("n"), py::arg("op"), py::arg("f64") = false,
     py::arg("threads") = 0);
  m.def("cpu_unary", [](int64_t a, int64_t o, int64_t n, int op, double lo, double hi, bool f64, int threads) {
    py::gil_scoped_release rel;
    if (f64) cpuops::unary<double>((const double*)a, (double*)o, n, op, lo, hi, threads);
    else cpuops::unary<float>((const float*)a, (float*)o, n, op, (float)lo, (float)hi, threads);
  }, py::arg("a"), py::arg("out"), py::arg("n"), py::arg("op"), py::arg("lo") = 0.0, py::arg("hi") = 0.0,
     py::arg("f64") = false, py::arg("threads") = 0);
  m.def("cpu_ternary", [](int64_t a, int64_t b, int64_t c, int64_t o, int64_t n, int op, bool f64, int threads) {
    py::gil_scoped_release rel;
    if (f64) cpuops::ternary<double>((const double*)a, (const double*)b, (const double*)c, (double*)o, n, op, threads);
    else cpuops::ternary<float>((const float*)a, (const float*)b, (const float*)c, (float*)o, n, op, threads);
  }, py::arg("a"), py::arg("b"), py::arg("c"), py::arg("out"), py::arg("n"), py::arg("op"), py::arg("f64") = false,
     py::arg("threads") = 0);
  m.def("cpu_axpy", [](int64_t y, int64_t x, double alpha, int64_t n, bool f64, int threads) {
    py::gil_scoped_release rel;
    if (f64) cpuops::axpy<double>((double*)y, (const double*)x, alpha, n, threads);
    else cpuops::axpy<float>((float*)y, (const float*)x, (float)alpha, n, threads);
  }, py::arg("y"), py::arg("x"), py::arg("alpha"), py::arg("n"), py::arg("f64") = false, py::arg("threads") = 0);
  m.def("cpu_fill", [](int64_t y, double v, int64_t n, bool f64, int threads) {
    py::gil_scoped_release rel;
    if (f64) cpuops::fill<double>((double*)y, v, n, threads);
    else cpuops::fill<float>((float*)y, (float)v, n, threads);
  }, py::arg("y"), py::arg("value"), py::arg("n"), py::arg("f64") = false, py::arg("threads") = 0);
  m.def("cpu_reduce", [](int64_t a, int64_t b, double mu, int64_t n, int op, bool f64, int threads) {
    py::gil_scoped_release rel;
    return f64 ? cpuops::reduce<double>((const double*)a, (const double*)b, mu, n, op, threads)
               : cpuops::reduce<float>((const float*)a, (const float*)b, mu, n, op, threads);
  }, py::arg("a"), py::arg("b"), py::arg("mu"), py::arg("n"), py::arg("op"), py::arg("f64") = false, py::arg("threads") = 0);
  m.def("cpu_adam_step", [](int64_t p, int64_t g, int64_t mm, int64_t v, int64_t n, double lr, double b1, double b2, double eps,
                            double wd, bool decoupled, double grad_scale, int64_t t, bool zero_grad, int threads) {
    py::gil_scoped_release rel;
    cpuops::AdamArgs a{(float)lr, (float)b1, (float)b2, (float)eps, (float)wd, (float)grad_scale,
                       (float)(1.0 - std::pow(b1, (double)t)), (float)(1.0 - std::pow(b2, (double)t)), decoupled, zero_grad};
    cpuops::adam_step((float*)p, (float*)g, (float*)mm, (float*)v, n, a, threads);
  }, py::arg("p"), py::arg("g"), py::arg("m"), py::arg("v"), py::arg("n"), py::arg("lr"), py::arg("beta1"), py::arg("beta2"),
     py::arg("eps"), py::arg("weight_decay"), py::arg("decoupled"), py::arg("grad_scale"), py::arg("step"),
     py::arg("zero_grad") = false, py::arg("threads") = 0);
  m.def("cpu_sgd_step", [](int64_t p, int64_t g, int64_t vel, int64_t n, double lr, double momentum, double wd, double grad_scale,
                           bool zero_grad, int threads) {
    py::gil_scoped_release rel;
    cpuops::SgdArgs a{(float)lr, (float)momentum, (float)wd, (float)grad_scale, zero_grad};
    cpuops::sgd_step((float*)p, (float*)g, (float*)vel, n, a, threads);
  }, py::arg("p"), py::arg("g"), py::arg("velocity"), py::arg("n"), py::arg("lr"), py::arg("momentum"), py::arg("weight_decay"),
     py::arg("grad_scale"), py::arg("zero_grad") = false, py::arg("threads") = 0);
  m.def("parallel_for", [](int64_t n, int64_t grain, const py::function& fn, int threads) {
    // Python callables run under the GIL; this is for API parity and for callables that release it
    global_pool(threads).parallel_for(n, grain, [&](int64_t b, int64_t e) {
      py::gil_scoped_acquire acq;
      fn(b, e);
    });
  }, py::arg("n"), py::arg("grain"), py::arg("fn"), py::arg("threads") = 0, py::call_guard<py::gil_scoped_release>());
  m.def("pool_threads", [](int threads) { return global_pool(threads).size(); }, py::arg("threads") = 0);

  // ---------------------------------------------------------------- TCP message bus
  py::class_<TcpBus>(m, "TcpBus")
      .def(py::init<int, int, size_t, const std::string&, uint64_t>(), py::arg("listen_port"), py::arg("sockets_per_peer") = 4,
           py::arg("max_packet") = (size_t)(4u << 20), py::arg("bind_host") = std::string("127.0.0.1"),
           py::arg("max_message") = (uint64_t)(8ull << 30))
      .def("start", &TcpBus::start)
      .def("stop", [](TcpBus& b) { py::gil_scoped_release rel; b.stop(); })
      .def("port", &TcpBus::port)
      .def("connect", [](TcpBus& b, const std::string& host, int port, int timeout_ms) {
        py::gil_scoped_release rel;
        b.connect(host, port, timeout_ms);
      }, py::arg("host"), py::arg("port"), py::arg("timeout_ms") = 10000)
      .def("has_peer", &TcpBus::has_peer)
      .def("peers", &TcpBus::peers)
      .def("send", [](TcpBus& b, const std::string& key, const py::bytes& msg) {
        char* buf = nullptr;
        Py_ssize_t len = 0;
        if (PyBytes_AsStringAndSize(msg.ptr(), &buf, &len) != 0) throw py::error_already_set();
        py::gil_scoped_release rel;   // `msg` stays referenced by the caller for the duration of the call
        b.send(key, reinterpret_cast<const uint8_t*>(buf), (size_t)len);
      })
      .def("recv", [](TcpBus& b, int timeout_ms) -> py::object {
        std::string from;
        std::vector<uint8_t> msg;
        bool ok;
        {
          py::gil_scoped_release rel;
          ok = b.recv(timeout_ms, &from, &msg);
        }
        if (!ok) return py::none();
        return py::make_tuple(from, py::bytes(reinterpret_cast<const char*>(msg.data()), msg.size()));
      }, py::arg("timeout_ms") = -1)
      .def("stats", [](const TcpBus& b) {
        auto s = b.stats();
        py::dict d;
        d["bytes_sent"] = s.bytes_sent;
        d["bytes_received"] = s.bytes_received;
        d["messages_sent"] = s.messages_sent;
        d["messages_received"] = s.messages_received;
        d["packets_sent"] = s.packets_sent;
        d["bad_packets"] = s.bad_packets;
        return d;
      });

  // ---------------------------------------------------------------- allocators
  py::class_<DELAllocator>(m, "DELAllocator")
      .def(py::init<uint64_t, uint64_t>(), py::arg("slab_bytes") = (64ull << 20), py::arg("align") = 256)
      .def("allocate", [](DELAllocator& a, uint64_t n) {
        Block b = a.allocate(n);
        return py::make_tuple(b.slab, b.offset, b.size);
      })
      .def("free", &DELAllocator::free)
      .def("flip", &DELAllocator::flip)
      .def("set_side", &DELAllocator::set_side)
      .def("side", &DELAllocator::side)
      .def("reserve", &DELAllocator::reserve)
      .def("clear", &DELAllocator::clear)
      .def("num_slabs", &DELAllocator::num_slabs)
      .def("slab_size", &DELAllocator::slab_size)
      .def("total_allocated", &DELAllocator::total_allocated)
      .def("peak_allocated", &DELAllocator::peak_allocated)
      .def("reset_peak", &DELAllocator::reset_peak)
      .def("total_reserved", &DELAllocator::total_reserved);
  py::class_<PoolAllocator>(m, "PoolAllocatorCore")
      .def(py::init<>())
      .def("acquire", &PoolAllocator::acquire)
      .def("add", &PoolAllocator::add)
      .def("release", &PoolAllocator::release)
      .def("block_size", &PoolAllocator::block_size)
      .def("total_bytes", &PoolAllocator::total_bytes)
      .def("cached_bytes", &PoolAllocator::cached_bytes);

  // ---------------------------------------------------------------- data path
  py::class_<RecordImageDataset>(m, "RecordImageDataset")
      .def(py::init<int, int, int, int>(), py::arg("C"), py::arg("H"), py::arg("W"), py::arg("label_bytes"))
      .def("add_file", &RecordImageDataset::add_file)
      .def("add_records", [](RecordImageDataset& d, const py::bytes& b) {
        std::string s = b;
        d.add_records(reinterpret_cast<const uint8_t*>(s.data()), s.size() / d.record_bytes());
      })
      .def("add_mnist_csv", [](RecordImageDataset& d, const std::string& path) {
        size_t n = 0;
        std::vector<uint8_t> recs = parse_mnist_csv(path, &n);
        d.add_records(recs.data(), n);
        return n;
      })
      .def("size", &RecordImageDataset::size)
      .def("assemble", [](const RecordImageDataset& d, py::array_t<int64_t, py::array::c_style> idx, int64_t out_ptr,
                          int64_t label_ptr, const py::dict& aug, uint64_t seed, int nthreads) {
        AugmentParams a;
        auto geti = [&](const char* k, int dflt) { return aug.contains(k) ? aug[k].cast<int>() : dflt; };
        auto getf = [&](const char* k, float dflt) { return aug.contains(k) ? aug[k].cast<float>() : dflt; };
        a.random_crop_pad = geti("random_crop_pad", 0);
        a.hflip_p = getf("hflip_p", 0.f);
        a.vflip_p = getf("vflip_p", 0.f);
        a.brightness = getf("brightness", 0.f);
        a.contrast = getf("contrast", 0.f);
        a.noise_std = getf("noise_std", 0.f);
        a.cutout = geti("cutout", 0);
        if (aug.contains("mean") && aug.contains("std")) {
          a.normalize = 1;
          auto mv = aug["mean"].cast<std::vector<float>>(), sv = aug["std"].cast<std::vector<float>>();
          for (int i = 0; i < 3; ++i) {
            a.mean[i] = mv[i % mv.size()];
            a.stdv[i] = sv[i % sv.size()];
          }
        }
        const int B = (int)idx.size();
        py::gil_scoped_release rel;
        d.assemble(idx.data(), B, reinterpret_cast<float*>(out_ptr), reinterpret_cast<int32_t*>(label_ptr), a, seed,
                   nthreads);
      });
  py::class_<TokenDataset>(m, "TokenDataset")
      .def(py::init<const std::string&>())
      .def("num_tokens", &TokenDataset::num_tokens)
      .def("sample", [](const TokenDataset& d, int B, int L, uint64_t seed, int64_t x_ptr, int64_t y_ptr) {
        d.sample(B, L, seed, reinterpret_cast<int32_t*>(x_ptr), reinterpret_cast<int32_t*>(y_ptr));
      })
      .def("window", [](const TokenDataset& d, uint64_t start, int L, int64_t x_ptr) {
        d.window(start, L, reinterpret_cast<int32_t*>(x_ptr));
      });

  // ---------------------------------------------------------------- tokenizer
  py::class_<Tokenizer>(m, "Tokenizer")
      .def(py::init<>())
      .def("load", &Tokenizer::load)
      .def("save", &Tokenizer::save)
      .def("set_vocab", [](Tokenizer& t, const std::vector<py::bytes>& v) {
        std::vector<std::string> s;
        for (auto& b : v) s.emplace_back(std::string(b));
        t.set_vocab(std::move(s));
      })
      .def("vocab_size", &Tokenizer::vocab_size)
      .def("decode", [](const Tokenizer& t, const std::vector<int32_t>& ids) { return py::bytes(t.decode(ids)); })
      .def("decode_one", [](const Tokenizer& t, int32_t id) { return py::bytes(t.decode_one(id)); })
      .def("encode", [](const Tokenizer& t, const std::string& s) { return t.encode(s); });

  // ---------------------------------------------------------------- profiler
  py::class_<Profiler, std::shared_ptr<Profiler>>(m, "ProfilerCore")
      .def(py::init<>())
      .def("init_start_time", &Profiler::init_start_time)
      .def_static("now_ns", &Profiler::now_ns)
      .def("add_event", &Profiler::add_event)
      .def("merge", &Profiler::merge)
      .def("sort", &Profiler::sort)
      .def("clear", &Profiler::clear)
      .def("serialize", &Profiler::serialize)
      .def_static("deserialize", [](const py::bytes& b) { return Profiler::deserialize(std::string(b)); })
      .def("events", &Profiler::events)
      .def("start_ns", &Profiler::start_ns);
}

}  // namespace tnn
