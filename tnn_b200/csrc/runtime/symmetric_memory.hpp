// Symmetric NVLink memory: the framework's OWN slab for peer-addressable and switch-reducible (NVLS) buffers.
//
// The reference registers one host slab with the NIC and carves every message buffer out of it (include/device/
// ibv_allocator.hpp:45-61: ibv_reg_mr once, best fit + coalescing inside).  The NVLink analogue of "registered memory"
// is memory that (a) every GPU of the group can address and (b) the NVSwitch can reduce / broadcast in flight:
//
//   physical    cuMemCreate            one allocation per rank on its own GPU, shareable as a POSIX file descriptor
//   unicast     cuMemAddressReserve + cuMemMap + cuMemSetAccess(all devices of the group, read/write)
//   multicast   cuMulticastCreate(numDevices = world) -> cuMulticastAddDevice (every rank) -> cuMulticastBindMem (every
//               rank's physical allocation at the same offset) -> the multicast handle mapped like memory: a
//               `multimem.ld_reduce` on that window returns the SUM over all ranks' slabs, a `multimem.st` writes all
//               of them (csrc/kernels/optim.cu k_dp_step uses exactly these two).
//
// Two ways to form a group:
//   * SymmetricMemory(devices, bytes)                 all ranks in ONE process (the C++ host's DataParallel: one thread per GPU)
//   * SymmetricMemory::rendezvous(rank, world, ...)   one process per GPU: handles travel as file descriptors over
//                                                     Unix-domain sockets (SCM_RIGHTS), rank 0 creates the multicast object
// Blocks inside the slab come from OffsetFreeList (best fit, coalescing frees); ranks that allocate in the same order get
// the same offsets, which is what "symmetric" means to the kernels (one offset, W base pointers, one multicast base).
//
// The driver API is reached through cudaGetDriverEntryPoint: no link-time dependency on libcuda, and a CPU-only box
// reports "unsupported" instead of failing to load.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <fcntl.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace tnn {

// ------------------------------------------------------------------------------------------- offsets inside one slab
// Best fit over a sorted free list, `align`-byte blocks, coalescing frees (reference: IbvAllocator, ibv_allocator.hpp:63-140).
class OffsetFreeList {
 public:
  static constexpr uint64_t kNone = ~0ull;
  explicit OffsetFreeList(uint64_t bytes = 0, uint64_t align = 256) : align_(align) { reset(bytes); }
  void reset(uint64_t bytes) {
    free_.clear();
    live_.clear();
    in_use_ = 0;
    capacity_ = bytes;
    if (bytes) free_[0] = bytes;
  }
  void set_align(uint64_t a) { align_ = a ? a : 1; }
  uint64_t allocate(uint64_t n) {
    n = std::max<uint64_t>(align_, (n + align_ - 1) / align_ * align_);
    auto best = free_.end();
    for (auto it = free_.begin(); it != free_.end(); ++it)
      if (it->second >= n && (best == free_.end() || it->second < best->second)) best = it;
    if (best == free_.end()) return kNone;
    const uint64_t off = best->first, len = best->second;
    free_.erase(best);
    if (len > n) free_[off + n] = len - n;
    live_[off] = n;
    in_use_ += n;
    peak_ = std::max(peak_, in_use_);
    return off;
  }
  // false when `off` is not the start of a live block (double free / foreign offset)
  bool release(uint64_t off) {
    auto it = live_.find(off);
    if (it == live_.end()) return false;
    uint64_t lo = off, len = it->second;
    in_use_ -= len;
    live_.erase(it);
    auto nx = free_.lower_bound(lo);
    if (nx != free_.end() && nx->first == lo + len) {
      len += nx->second;
      nx = free_.erase(nx);
    }
    if (nx != free_.begin()) {
      auto pv = std::prev(nx);
      if (pv->first + pv->second == lo) {
        lo = pv->first;
        len += pv->second;
        free_.erase(pv);
      }
    }
    free_[lo] = len;
    return true;
  }
  uint64_t block_size(uint64_t off) const {
    auto it = live_.find(off);
    return it == live_.end() ? 0 : it->second;
  }
  uint64_t in_use() const { return in_use_; }
  uint64_t peak() const { return peak_; }
  uint64_t capacity() const { return capacity_; }
  uint64_t align() const { return align_; }
  size_t num_free_ranges() const { return free_.size(); }
  size_t num_live() const { return live_.size(); }
  uint64_t largest_free() const {
    uint64_t m = 0;
    for (auto& kv : free_) m = std::max(m, kv.second);
    return m;
  }

 private:
  std::map<uint64_t, uint64_t> free_, live_;             // offset -> length
  uint64_t in_use_ = 0, peak_ = 0, capacity_ = 0, align_ = 256;
};

// ------------------------------------------------------------------------------- file descriptors between processes
// One datagram = (tag, optional descriptor).  The shareable handle of a cuMemCreate allocation IS a file descriptor; it
// can only cross a process boundary as SCM_RIGHTS ancillary data of a Unix-domain socket.
class FdMailbox {
 public:
  // binds <dir>/<name>; the path is unlinked again by the destructor
  FdMailbox(const std::string& dir, const std::string& name) : path_(dir + "/" + name) {
    sock_ = ::socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC, 0);
    if (sock_ < 0) throw std::runtime_error("FdMailbox: socket: " + std::string(std::strerror(errno)));
    sockaddr_un a = addr_of(path_);
    ::unlink(path_.c_str());
    if (::bind(sock_, reinterpret_cast<sockaddr*>(&a), sizeof a) != 0) {
      const std::string why = std::strerror(errno);
      ::close(sock_);
      throw std::runtime_error("FdMailbox: bind " + path_ + ": " + why);
    }
  }
  ~FdMailbox() {
    if (sock_ >= 0) ::close(sock_);
    ::unlink(path_.c_str());
  }
  FdMailbox(const FdMailbox&) = delete;
  FdMailbox& operator=(const FdMailbox&) = delete;
  const std::string& path() const { return path_; }

  // send (tag, fd) to the mailbox bound at `to`; fd < 0 sends the tag alone.  Retries while the peer has not bound yet.
  void send(const std::string& to, uint64_t tag, int fd, int timeout_ms = 30000) const {
    sockaddr_un a = addr_of(to);
    msghdr msg{};
    iovec iov{&tag, sizeof tag};
    alignas(cmsghdr) char ctl[CMSG_SPACE(sizeof(int))];
    msg.msg_name = &a;
    msg.msg_namelen = sizeof a;
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    if (fd >= 0) {
      std::memset(ctl, 0, sizeof ctl);
      msg.msg_control = ctl;
      msg.msg_controllen = sizeof ctl;
      cmsghdr* c = CMSG_FIRSTHDR(&msg);
      c->cmsg_level = SOL_SOCKET;
      c->cmsg_type = SCM_RIGHTS;
      c->cmsg_len = CMSG_LEN(sizeof(int));
      std::memcpy(CMSG_DATA(c), &fd, sizeof(int));
    }
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms);
    for (;;) {
      if (::sendmsg(sock_, &msg, 0) == (ssize_t)sizeof tag) return;
      const int e = errno;
      const bool retry = e == ENOENT || e == ECONNREFUSED || e == EAGAIN || e == ENOBUFS || e == EINTR;
      if (!retry || std::chrono::steady_clock::now() > deadline)
        throw std::runtime_error("FdMailbox: send to " + to + ": " + std::strerror(e));
      std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
  }
  // next datagram: returns its tag, *fd = the received descriptor (caller closes) or -1
  uint64_t recv(int* fd, int timeout_ms = 30000) const {
    pollfd p{sock_, POLLIN, 0};
    for (;;) {
      const int r = ::poll(&p, 1, timeout_ms);
      if (r > 0) break;
      if (r == 0) throw std::runtime_error("FdMailbox: timed out waiting on " + path_);
      if (errno != EINTR) throw std::runtime_error("FdMailbox: poll: " + std::string(std::strerror(errno)));
    }
    uint64_t tag = 0;
    msghdr msg{};
    iovec iov{&tag, sizeof tag};
    alignas(cmsghdr) char ctl[CMSG_SPACE(sizeof(int))];
    std::memset(ctl, 0, sizeof ctl);
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    msg.msg_control = ctl;
    msg.msg_controllen = sizeof ctl;
    const ssize_t n = ::recvmsg(sock_, &msg, MSG_CMSG_CLOEXEC);
    if (n != (ssize_t)sizeof tag) throw std::runtime_error("FdMailbox: short datagram on " + path_);
    if (fd) *fd = -1;
    for (cmsghdr* c = CMSG_FIRSTHDR(&msg); c; c = CMSG_NXTHDR(&msg, c))
      if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) {
        int got;
        std::memcpy(&got, CMSG_DATA(c), sizeof(int));
        if (fd && *fd < 0) *fd = got;
        else ::close(got);
      }
    return tag;
  }

 private:
  static sockaddr_un addr_of(const std::string& path) {
    sockaddr_un a{};
    a.sun_family = AF_UNIX;
    if (path.size() >= sizeof a.sun_path) throw std::runtime_error("FdMailbox: path too long: " + path);
    std::memcpy(a.sun_path, path.c_str(), path.size() + 1);
    return a;
  }
  std::string path_;
  int sock_ = -1;
};

// ------------------------------------------------------------------------------------------------ driver entry points
struct VmmDriver {
  decltype(&cuMemCreate) MemCreate = nullptr;
  decltype(&cuMemRelease) MemRelease = nullptr;
  decltype(&cuMemAddressReserve) MemAddressReserve = nullptr;
  decltype(&cuMemAddressFree) MemAddressFree = nullptr;
  decltype(&cuMemMap) MemMap = nullptr;
  decltype(&cuMemUnmap) MemUnmap = nullptr;
  decltype(&cuMemSetAccess) MemSetAccess = nullptr;
  decltype(&cuMemGetAllocationGranularity) MemGetAllocationGranularity = nullptr;
  decltype(&cuMemExportToShareableHandle) MemExportToShareableHandle = nullptr;
  decltype(&cuMemImportFromShareableHandle) MemImportFromShareableHandle = nullptr;
  decltype(&cuMulticastCreate) MulticastCreate = nullptr;
  decltype(&cuMulticastAddDevice) MulticastAddDevice = nullptr;
  decltype(&cuMulticastBindMem) MulticastBindMem = nullptr;
  decltype(&cuMulticastUnbind) MulticastUnbind = nullptr;
  decltype(&cuMulticastGetGranularity) MulticastGetGranularity = nullptr;
  decltype(&cuDeviceGetAttribute) DeviceGetAttribute = nullptr;
  decltype(&cuDeviceGet) DeviceGet = nullptr;
  decltype(&cuGetErrorString) GetErrorString = nullptr;
  std::string why;                                     // empty: every entry point resolved
  // A table installed by install_for_testing() stands in for the driver (tests/ use a memfd-backed one: "physical
  // allocations" are memfds, "mapping" is mmap, so the whole group protocol runs on a CPU box).  `device_count` > 0 then
  // replaces cudaGetDeviceCount and the CUDA runtime calls (context creation, device synchronise) are skipped.
  int fake_device_count = 0;

  static const VmmDriver& get() {
    if (const VmmDriver* t = testing_slot()) return *t;
    static const VmmDriver d = load();
    return d;
  }
  static void install_for_testing(const VmmDriver* table) { testing_slot() = table; }
  bool fake() const { return fake_device_count > 0; }
  bool ok() const { return why.empty(); }
  std::string err(CUresult r) const {
    const char* s = nullptr;
    if (GetErrorString && GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
    return "CUresult " + std::to_string((int)r);
  }

 private:
  static const VmmDriver*& testing_slot() {
    static const VmmDriver* t = nullptr;
    return t;
  }
  template <class F>
  static bool one(const char* name, F* out, std::string* why) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    const cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
      cudaGetLastError();
      if (why->empty()) *why = std::string("driver entry point ") + name + " unavailable" + (e != cudaSuccess ? std::string(": ") + cudaGetErrorString(e) : "");
      return false;
    }
    *out = reinterpret_cast<F>(p);
    return true;
  }
  static VmmDriver load() {
    VmmDriver d;
    std::string w;
#define TNN_VMM_ENTRY(n) one("cu" #n, &d.n, &w)
    TNN_VMM_ENTRY(GetErrorString);
    TNN_VMM_ENTRY(DeviceGet);
    TNN_VMM_ENTRY(DeviceGetAttribute);
    TNN_VMM_ENTRY(MemCreate);
    TNN_VMM_ENTRY(MemRelease);
    TNN_VMM_ENTRY(MemAddressReserve);
    TNN_VMM_ENTRY(MemAddressFree);
    TNN_VMM_ENTRY(MemMap);
    TNN_VMM_ENTRY(MemUnmap);
    TNN_VMM_ENTRY(MemSetAccess);
    TNN_VMM_ENTRY(MemGetAllocationGranularity);
    TNN_VMM_ENTRY(MemExportToShareableHandle);
    TNN_VMM_ENTRY(MemImportFromShareableHandle);
    TNN_VMM_ENTRY(MulticastCreate);
    TNN_VMM_ENTRY(MulticastAddDevice);
    TNN_VMM_ENTRY(MulticastBindMem);
    TNN_VMM_ENTRY(MulticastUnbind);
    TNN_VMM_ENTRY(MulticastGetGranularity);
#undef TNN_VMM_ENTRY
    d.why = w;
    return d;
  }
};

// -------------------------------------------------------------------------------------------------- the symmetric slab
class SymmetricMemory {
 public:
  struct Support {
    bool vmm = false, multicast = false;
    std::string why;                                   // why `vmm` (or `multicast`) is false
  };
  // what the driver / devices of this box offer for a group over `devices` (CUDA ordinals)
  static Support probe(const std::vector<int>& devices) {
    Support s;
    const VmmDriver& D = VmmDriver::get();
    if (!D.ok()) {
      s.why = D.why;
      return s;
    }
    int n = D.fake_device_count;
    if (!D.fake() && (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)) {
      cudaGetLastError();
      s.why = "no CUDA device";
      return s;
    }
    s.vmm = s.multicast = !devices.empty();
    for (int d : devices) {
      CUdevice cd;
      int v = 0;
      if (d < 0 || d >= n || D.DeviceGet(&cd, d) != CUDA_SUCCESS) {
        s.vmm = s.multicast = false;
        s.why = "device " + std::to_string(d) + " not visible";
        return s;
      }
      if (D.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, cd) != CUDA_SUCCESS || !v) {
        s.vmm = s.multicast = false;
        s.why = "device " + std::to_string(d) + ": no virtual memory management";
        return s;
      }
      if (D.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, cd) != CUDA_SUCCESS || !v) {
        s.vmm = s.multicast = false;
        s.why = "device " + std::to_string(d) + ": POSIX fd handles unsupported";
        return s;
      }
      if (D.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cd) != CUDA_SUCCESS || !v) {
        if (s.multicast) s.why = "device " + std::to_string(d) + ": no multicast (NVLS) support";
        s.multicast = false;
      }
    }
    if (devices.size() < 2 && s.multicast) {
      s.multicast = false;
      s.why = "multicast needs at least two devices";
    }
    return s;
  }

  // ---- all ranks in this process ----
  SymmetricMemory(const std::vector<int>& devices, size_t bytes, bool want_multicast = true) : devices_(devices), my_rank_(-1) {
    const Support sup = probe(devices);
    if (!sup.vmm) throw std::runtime_error("SymmetricMemory: " + sup.why);
    const bool mc = want_multicast && sup.multicast;
    init_sizes(bytes, mc);
    const int W = (int)devices_.size();
    phys_.assign((size_t)W, 0);
    va_.assign((size_t)W, 0);
    try {
      for (int r = 0; r < W; ++r) {
        make_context(devices_[(size_t)r]);
        phys_[(size_t)r] = create_physical(devices_[(size_t)r]);
      }
      for (int r = 0; r < W; ++r) va_[(size_t)r] = map_handle(phys_[(size_t)r]);
      if (mc) {
        // the multicast window is an optimisation: a box that advertises it but refuses the object (fabric manager down,
        // NVLS disabled) still gets the peer mappings, and the kernels take their peer-load path (has_multicast() == false)
        try {
          CUmulticastObjectProp mp = multicast_prop();
          check(D().MulticastCreate(&mc_handle_, &mp), "cuMulticastCreate");
          for (int r = 0; r < W; ++r) check(D().MulticastAddDevice(mc_handle_, cu_device(devices_[(size_t)r])), "cuMulticastAddDevice");
          for (int r = 0; r < W; ++r) check(D().MulticastBindMem(mc_handle_, 0, phys_[(size_t)r], 0, bytes_, 0), "cuMulticastBindMem");
          mc_va_ = map_handle(mc_handle_);
        } catch (const std::exception& e) {
          multicast_error_ = e.what();
          unmap(mc_va_);
          if (mc_handle_) {
            D().MemRelease(mc_handle_);
            mc_handle_ = 0;
          }
        }
      }
    } catch (...) {
      destroy();
      throw;
    }
    lists_.resize((size_t)W);
    for (auto& l : lists_) l = std::make_unique<OffsetFreeList>(bytes_, 256);
  }

  // ---- one process per rank ----
  // Every rank calls this with the same `dir` (an existing directory private to the job), `bytes` and `world`.
  static std::unique_ptr<SymmetricMemory> rendezvous(int rank, int world, int device, size_t bytes, const std::string& dir, bool want_multicast = true,
                                                     int timeout_ms = 60000) {
    std::unique_ptr<SymmetricMemory> s(new SymmetricMemory());
    s->join(rank, world, device, bytes, dir, want_multicast, timeout_ms);
    return s;
  }

  ~SymmetricMemory() { destroy(); }
  SymmetricMemory(const SymmetricMemory&) = delete;
  SymmetricMemory& operator=(const SymmetricMemory&) = delete;

  int world() const { return (int)va_.size(); }
  int my_rank() const { return my_rank_; }                               // -1: every rank lives in this process
  size_t bytes() const { return bytes_; }                                // per rank, rounded up to the mapping granularity
  size_t granularity() const { return gran_; }
  bool has_multicast() const { return mc_va_ != 0; }
  const std::string& multicast_error() const { return multicast_error_; }   // why a requested window is missing ("" otherwise)
  void* base(int rank) const { return reinterpret_cast<void*>(va_.at((size_t)rank)); }
  void* multicast_base() const { return reinterpret_cast<void*>(mc_va_); }
  int device_of(int rank) const { return devices_.at((size_t)rank); }
  // the same offset in another rank's slab / in the multicast window
  uint64_t offset_of(int rank, const void* p) const {
    const uint64_t a = reinterpret_cast<uint64_t>(p), b = va_.at((size_t)rank);
    if (a < b || a >= b + bytes_) throw std::out_of_range("SymmetricMemory: pointer outside rank " + std::to_string(rank) + "'s slab");
    return a - b;
  }
  void* peer(int rank, const void* p, int other) const { return reinterpret_cast<char*>(base(other)) + offset_of(rank, p); }
  void* multicast(int rank, const void* p) const { return mc_va_ ? reinterpret_cast<char*>(multicast_base()) + offset_of(rank, p) : nullptr; }

  // symmetric blocks: rank r's k-th allocation has the same offset as every other rank's k-th allocation
  void* allocate(int rank, size_t n) {
    std::lock_guard<std::mutex> g(mu_);
    const uint64_t off = list(rank).allocate(n);
    if (off == OffsetFreeList::kNone)
      throw std::bad_alloc();
    return reinterpret_cast<char*>(base(rank)) + off;
  }
  void deallocate(int rank, void* p) {
    std::lock_guard<std::mutex> g(mu_);
    list(rank).release(offset_of(rank, p));
  }
  const OffsetFreeList& free_list(int rank) const { return *lists_.at(my_rank_ >= 0 ? 0 : (size_t)rank); }

  // shareable handle of rank's physical allocation (caller closes the descriptor)
  int export_fd(int rank) const {
    int fd = -1;
    check(D().MemExportToShareableHandle(&fd, phys_.at((size_t)rank), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle");
    return fd;
  }

 private:
  SymmetricMemory() = default;
  static const VmmDriver& D() { return VmmDriver::get(); }
  static void check(CUresult r, const char* what) {
    if (r != CUDA_SUCCESS) throw std::runtime_error(std::string("SymmetricMemory: ") + what + ": " + D().err(r));
  }
  static void make_context(int dev) {
    if (D().fake()) return;
    if (cudaSetDevice(dev) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) {
      const std::string e = cudaGetErrorString(cudaGetLastError());
      throw std::runtime_error("SymmetricMemory: cannot initialise device " + std::to_string(dev) + ": " + e);
    }
  }
  static CUdevice cu_device(int dev) {
    CUdevice d;
    check(D().DeviceGet(&d, dev), "cuDeviceGet");
    return d;
  }
  OffsetFreeList& list(int rank) {
    if (my_rank_ >= 0 && rank != my_rank_) throw std::invalid_argument("SymmetricMemory: only the local rank allocates in a multi-process group");
    return *lists_.at(my_rank_ >= 0 ? 0 : (size_t)rank);
  }
  CUmemAllocationProp alloc_prop(int dev) const {
    CUmemAllocationProp p{};
    p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    p.location.id = dev;
    p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return p;
  }
  CUmulticastObjectProp multicast_prop() const {
    CUmulticastObjectProp mp{};
    mp.numDevices = (unsigned)devices_.size();
    mp.size = bytes_;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return mp;
  }
  void init_sizes(size_t bytes, bool mc) {
    if (bytes == 0) throw std::invalid_argument("SymmetricMemory: zero-sized slab");
    size_t g = 0;
    for (int d : devices_) {
      CUmemAllocationProp p = alloc_prop(d);
      size_t gi = 0;
      check(D().MemGetAllocationGranularity(&gi, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
      g = std::max(g, gi);
    }
    if (mc) {
      bytes_ = bytes;                                  // the property's size field only matters for numDevices / handle types here
      CUmulticastObjectProp mp = multicast_prop();
      size_t gm = 0;
      check(D().MulticastGetGranularity(&gm, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED), "cuMulticastGetGranularity");
      g = std::max(g, gm);
    }
    gran_ = g ? g : (size_t)2 << 20;
    bytes_ = (bytes + gran_ - 1) / gran_ * gran_;
  }
  CUmemGenericAllocationHandle create_physical(int dev) const {
    CUmemAllocationProp p = alloc_prop(dev);
    CUmemGenericAllocationHandle h = 0;
    check(D().MemCreate(&h, bytes_, &p, 0), "cuMemCreate");
    return h;
  }
  // reserve a range, map the handle, open it read/write for every device of the group that lives in this process
  CUdeviceptr map_handle(CUmemGenericAllocationHandle h) const {
    CUdeviceptr va = 0;
    check(D().MemAddressReserve(&va, bytes_, gran_, 0, 0), "cuMemAddressReserve");
    CUresult r = D().MemMap(va, bytes_, 0, h, 0);
    if (r != CUDA_SUCCESS) {
      D().MemAddressFree(va, bytes_);
      check(r, "cuMemMap");
    }
    std::vector<CUmemAccessDesc> acc;
    for (int d : local_devices()) {
      CUmemAccessDesc a{};
      a.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
      a.location.id = d;
      a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
      acc.push_back(a);
    }
    r = D().MemSetAccess(va, bytes_, acc.data(), acc.size());
    if (r != CUDA_SUCCESS) {
      D().MemUnmap(va, bytes_);
      D().MemAddressFree(va, bytes_);
      check(r, "cuMemSetAccess");
    }
    return va;
  }
  std::vector<int> local_devices() const {
    if (my_rank_ < 0) return devices_;
    return {devices_.at((size_t)my_rank_)};
  }
  void unmap(CUdeviceptr& va) {
    if (!va) return;
    D().MemUnmap(va, bytes_);
    D().MemAddressFree(va, bytes_);
    va = 0;
  }
  void destroy() {
    if (!D().ok()) return;
    if (!D().fake())
      for (int d : local_devices_safe()) {
        cudaSetDevice(d);
        cudaDeviceSynchronize();
      }
    unmap(mc_va_);
    if (mc_handle_) {
      for (size_t r = 0; r < phys_.size(); ++r)
        if (phys_[r] && (my_rank_ < 0 || (int)r == my_rank_)) D().MulticastUnbind(mc_handle_, cu_device_noexcept(devices_[r]), 0, bytes_);
      D().MemRelease(mc_handle_);
      mc_handle_ = 0;
    }
    for (auto& v : va_) unmap(v);
    for (auto& h : phys_)
      if (h) {
        D().MemRelease(h);
        h = 0;
      }
    if (!D().fake()) cudaGetLastError();
  }
  std::vector<int> local_devices_safe() const {
    if (my_rank_ < 0) return devices_;
    if ((size_t)my_rank_ < devices_.size()) return {devices_[(size_t)my_rank_]};
    return {};
  }
  static CUdevice cu_device_noexcept(int dev) {
    CUdevice d = 0;
    D().DeviceGet(&d, dev);
    return d;
  }

  static constexpr uint64_t TAG_PHYS = 1ull << 32, TAG_MC = 2ull << 32, TAG_ADDED = 3ull << 32, TAG_BOUND = 4ull << 32, TAG_NO_MC = 5ull << 32;
  void join(int rank, int world, int device, size_t bytes, const std::string& dir, bool want_multicast, int timeout_ms) {
    if (world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("SymmetricMemory::rendezvous: bad rank / world");
    my_rank_ = rank;
    devices_.assign((size_t)world, -1);
    devices_[(size_t)rank] = device;
    Support sup = probe({device});
    if (!sup.vmm) throw std::runtime_error("SymmetricMemory: " + sup.why);
    int v = 0;
    const bool dev_mc = D().DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cu_device(device)) == CUDA_SUCCESS && v;
    bool mc = want_multicast && dev_mc && world > 1;
    // sizes: every rank computes the same numbers from its own (identical) device
    {
      std::vector<int> saved = devices_;
      devices_.assign((size_t)world, device);          // numDevices = world for the multicast granularity query
      init_sizes(bytes, mc);
      devices_ = saved;
    }
    phys_.assign((size_t)world, 0);
    va_.assign((size_t)world, 0);
    auto box_name = [](int r) { return "symm." + std::to_string(r) + ".sock"; };
    FdMailbox box(dir, box_name(rank));
    try {
      make_context(device);
      phys_[(size_t)rank] = create_physical(device);
      va_[(size_t)rank] = map_handle(phys_[(size_t)rank]);
      // 1. everybody hands its physical allocation to everybody (tag carries the sender's rank and device)
      for (int q = 0; q < world; ++q) {
        if (q == rank) continue;
        const int fd = export_fd(rank);
        try {
          box.send(dir + "/" + box_name(q), TAG_PHYS | ((uint64_t)device << 16) | (uint64_t)rank, fd, timeout_ms);
        } catch (...) {
          ::close(fd);
          throw;
        }
        ::close(fd);
      }
      // 2. rank 0 creates the multicast object and hands it out (or tells everybody there is none)
      std::vector<std::pair<uint64_t, int>> early;      // datagrams of a later phase that overtook this one
      auto kind_of = [](uint64_t t) { return t & ~0xffffffffull; };
      // next datagram of kind `a` (or `b`): held-back ones first, others are held back
      auto next = [&](uint64_t a, uint64_t b, int* fd) -> uint64_t {
        for (size_t i = 0; i < early.size(); ++i)
          if (kind_of(early[i].first) == a || kind_of(early[i].first) == b) {
            const auto e = early[i];
            early.erase(early.begin() + (long)i);
            *fd = e.second;
            return e.first;
          }
        for (;;) {
          int f = -1;
          const uint64_t t = box.recv(&f, timeout_ms);
          if (kind_of(t) == a || kind_of(t) == b) {
            *fd = f;
            return t;
          }
          early.emplace_back(t, f);
        }
      };
      for (int i = 0; i < world - 1; ++i) {
        int fd = -1;
        const uint64_t t = next(TAG_PHYS, TAG_PHYS, &fd);
        const int q = (int)(t & 0xffff), qdev = (int)((t >> 16) & 0xffff);
        if (fd < 0 || q < 0 || q >= world || q == rank || phys_[(size_t)q]) {
          if (fd >= 0) ::close(fd);
          throw std::runtime_error("SymmetricMemory::rendezvous: unexpected handle datagram");
        }
        devices_[(size_t)q] = qdev;
        CUresult r = D().MemImportFromShareableHandle(&phys_[(size_t)q], reinterpret_cast<void*>((uintptr_t)fd), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        ::close(fd);
        check(r, "cuMemImportFromShareableHandle");
        va_[(size_t)q] = map_handle(phys_[(size_t)q]);
      }
      if (world > 1) {
        if (rank == 0) {
          if (mc) {
            CUmulticastObjectProp mp = multicast_prop();
            if (D().MulticastCreate(&mc_handle_, &mp) != CUDA_SUCCESS) {
              mc = false;
              mc_handle_ = 0;
            }
          }
          for (int q = 1; q < world; ++q) {
            if (mc) {
              int fd = -1;
              check(D().MemExportToShareableHandle(&fd, mc_handle_, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle(multicast)");
              try {
                box.send(dir + "/" + box_name(q), TAG_MC, fd, timeout_ms);
              } catch (...) {
                ::close(fd);
                throw;
              }
              ::close(fd);
            } else {
              box.send(dir + "/" + box_name(q), TAG_NO_MC, -1, timeout_ms);
            }
          }
        } else {
          int fd = -1;
          const uint64_t t = next(TAG_MC, TAG_NO_MC, &fd);      // whichever rank 0 decided
          if (kind_of(t) == TAG_NO_MC) {
            if (fd >= 0) ::close(fd);
            fd = -1;
            mc = false;
          } else {
            mc = true;
          }
          if (mc && fd < 0) throw std::runtime_error("SymmetricMemory::rendezvous: multicast datagram without a descriptor");
          if (fd >= 0) {
            CUresult r = D().MemImportFromShareableHandle(&mc_handle_, reinterpret_cast<void*>((uintptr_t)fd), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
            ::close(fd);
            check(r, "cuMemImportFromShareableHandle(multicast)");
          }
        }
        if (mc) {
          // 3. every rank adds its device, then (after all have) binds its own allocation
          check(D().MulticastAddDevice(mc_handle_, cu_device(device)), "cuMulticastAddDevice");
          for (int q = 0; q < world; ++q)
            if (q != rank) box.send(dir + "/" + box_name(q), TAG_ADDED | (uint64_t)rank, -1, timeout_ms);
          for (int i = 0; i < world - 1; ++i) {
            int f = -1;
            next(TAG_ADDED, TAG_ADDED, &f);
            if (f >= 0) ::close(f);
          }
          check(D().MulticastBindMem(mc_handle_, 0, phys_[(size_t)rank], 0, bytes_, 0), "cuMulticastBindMem");
          mc_va_ = map_handle(mc_handle_);
        }
        // 4. nobody leaves (and unlinks its mailbox) before everybody is mapped
        for (int q = 0; q < world; ++q)
          if (q != rank) box.send(dir + "/" + box_name(q), TAG_BOUND | (uint64_t)rank, -1, timeout_ms);
        for (int i = 0; i < world - 1; ++i) {
          int f = -1;
          next(TAG_BOUND, TAG_BOUND, &f);
          if (f >= 0) ::close(f);
        }
        for (auto& e : early)
          if (e.second >= 0) ::close(e.second);
      }
    } catch (...) {
      destroy();
      throw;
    }
    lists_.clear();
    lists_.push_back(std::make_unique<OffsetFreeList>(bytes_, 256));
  }

  std::vector<int> devices_;
  int my_rank_ = -1;
  size_t bytes_ = 0, gran_ = 0;
  std::vector<CUmemGenericAllocationHandle> phys_;
  std::vector<CUdeviceptr> va_;
  CUmemGenericAllocationHandle mc_handle_ = 0;
  CUdeviceptr mc_va_ = 0;
  std::string multicast_error_;
  std::vector<std::unique_ptr<OffsetFreeList>> lists_;
  std::mutex mu_;
};

}  // namespace tnn
