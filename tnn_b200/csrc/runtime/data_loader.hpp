// Native data path: memory-mapped dataset files, threaded batch assembly (uint8 CHW records ->
// float NHWC with augmentation) straight into caller-provided (pinned) buffers.
// Reference: include/data_loading/cifar10_data_loader.hpp:41-170 (mmap + parallel_for
// uint8->float), mnist CSV loader, open_webtext_data_loader.hpp:11-87 (uint16 token mmap).
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace tnn {

class MappedFile {
 public:
  explicit MappedFile(const std::string& path) {
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) throw std::runtime_error("cannot open " + path);
    struct stat st;
    if (fstat(fd_, &st) != 0) throw std::runtime_error("cannot stat " + path);
    size_ = (size_t)st.st_size;
    if (size_ > 0) {
      ptr_ = (const uint8_t*)mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
      if (ptr_ == MAP_FAILED) throw std::runtime_error("mmap failed for " + path);
    }
  }
  ~MappedFile() {
    if (ptr_ && ptr_ != MAP_FAILED) munmap((void*)ptr_, size_);
    if (fd_ >= 0) ::close(fd_);
  }
  MappedFile(const MappedFile&) = delete;
  const uint8_t* data() const { return ptr_; }
  size_t size() const { return size_; }

 private:
  int fd_ = -1;
  const uint8_t* ptr_ = nullptr;
  size_t size_ = 0;
};

struct AugmentParams {
  int random_crop_pad = 0;     // 0 = off; pad then random crop back to HxW
  float hflip_p = 0.f;
  float vflip_p = 0.f;
  float brightness = 0.f;      // +- delta
  float contrast = 0.f;        // factor in [1-c, 1+c]
  float noise_std = 0.f;
  int cutout = 0;              // square size, 0 = off
  int normalize = 0;
  float mean[3] = {0, 0, 0};
  float stdv[3] = {1, 1, 1};
};

// Image dataset of fixed-size records: `label_bytes` label bytes (the LAST one is used) followed
// by C*H*W uint8 pixels in CHW order (CIFAR binary layout).  Several files may be appended.
class RecordImageDataset {
 public:
  RecordImageDataset(int C, int H, int W, int label_bytes) : C_(C), H_(H), W_(W), lb_(label_bytes) {}
  void add_file(const std::string& path) {
    files_.emplace_back(new MappedFile(path));
    const size_t rec = record_bytes();
    const size_t n = files_.back()->size() / rec;
    for (size_t i = 0; i < n; ++i) recs_.push_back(files_.back()->data() + i * rec);
  }
  // in-memory records (MNIST CSV decode, synthetic data)
  void add_records(const uint8_t* data, size_t n) {
    owned_.emplace_back(data, data + n * record_bytes());
    const uint8_t* base = owned_.back().data();
    for (size_t i = 0; i < n; ++i) recs_.push_back(base + i * record_bytes());
  }
  size_t size() const { return recs_.size(); }
  size_t record_bytes() const { return (size_t)lb_ + (size_t)C_ * H_ * W_; }
  int C() const { return C_; }
  int H() const { return H_; }
  int W() const { return W_; }

  // out: float [B, H, W, C] (NHWC, /255), labels: int32 [B]
  void assemble(const int64_t* indices, int B, float* out, int32_t* labels, const AugmentParams& aug,
                uint64_t seed, int nthreads) const {
    nthreads = std::max(1, std::min(nthreads, B));
    auto work = [&](int t) {
      for (int b = t; b < B; b += nthreads) one(indices[b], out + (size_t)b * H_ * W_ * C_, labels + b, aug,
                                                 seed * 0x9E3779B97F4A7C15ull + (uint64_t)indices[b]);
    };
    if (nthreads == 1) {
      work(0);
      return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }

 private:
  void one(int64_t idx, float* o, int32_t* label, const AugmentParams& a, uint64_t seed) const {
    const uint8_t* rec = recs_.at((size_t)idx);
    *label = rec[lb_ - 1];
    const uint8_t* px = rec + lb_;
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    int dy = 0, dx = 0;
    if (a.random_crop_pad > 0) {
      dy = (int)(rng() % (2 * a.random_crop_pad + 1)) - a.random_crop_pad;
      dx = (int)(rng() % (2 * a.random_crop_pad + 1)) - a.random_crop_pad;
    }
    const bool hf = a.hflip_p > 0.f && U(rng) < a.hflip_p;
    const bool vf = a.vflip_p > 0.f && U(rng) < a.vflip_p;
    const float br = a.brightness > 0.f ? (U(rng) * 2.f - 1.f) * a.brightness : 0.f;
    const float ct = a.contrast > 0.f ? 1.f + (U(rng) * 2.f - 1.f) * a.contrast : 1.f;
    int cy = -1, cx = -1;
    if (a.cutout > 0) {
      cy = (int)(rng() % H_);
      cx = (int)(rng() % W_);
    }
    std::normal_distribution<float> Nn(0.f, a.noise_std > 0.f ? a.noise_std : 1.f);
    for (int h = 0; h < H_; ++h) {
      for (int w = 0; w < W_; ++w) {
        int sh = (vf ? H_ - 1 - h : h) + dy, sw = (hf ? W_ - 1 - w : w) + dx;
        const bool inside = sh >= 0 && sh < H_ && sw >= 0 && sw < W_;
        const bool cut = a.cutout > 0 && std::abs(h - cy) <= a.cutout / 2 && std::abs(w - cx) <= a.cutout / 2;
        for (int c = 0; c < C_; ++c) {
          float v = 0.f;
          if (inside && !cut) {
            v = px[((size_t)c * H_ + sh) * W_ + sw] * (1.f / 255.f);
            if (ct != 1.f) v = (v - 0.5f) * ct + 0.5f;
            v += br;
            if (a.noise_std > 0.f) v += Nn(rng);
            if (a.brightness > 0.f || a.contrast > 0.f || a.noise_std > 0.f) v = std::min(1.f, std::max(0.f, v));
          }
          if (a.normalize) v = (v - a.mean[c % 3]) / a.stdv[c % 3];
          o[((size_t)h * W_ + w) * C_ + c] = v;
        }
      }
    }
  }
  int C_, H_, W_, lb_;
  std::vector<std::unique_ptr<MappedFile>> files_;
  std::vector<std::vector<uint8_t>> owned_;
  std::vector<const uint8_t*> recs_;
};

// MNIST CSV ("label,p0,...,p783" per line, optional header) -> records {label, 784 pixels}
inline std::vector<uint8_t> parse_mnist_csv(const std::string& path, size_t* n_out) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open " + path);
  std::vector<uint8_t> recs;
  std::string line;
  size_t n = 0;
  while (std::getline(f, line)) {
    if (line.empty()) continue;
    if (!(line[0] >= '0' && line[0] <= '9')) continue;  // header
    size_t start = recs.size();
    recs.resize(start + 785, 0);
    size_t i = 0, k = 0;
    int val = 0;
    bool have = false;
    for (; i <= line.size() && k < 785; ++i) {
      char ch = i < line.size() ? line[i] : ',';
      if (ch >= '0' && ch <= '9') {
        val = val * 10 + (ch - '0');
        have = true;
      } else if (ch == ',' || ch == '\r') {
        if (have) recs[start + k++] = (uint8_t)std::min(val, 255);
        val = 0;
        have = false;
      }
    }
    ++n;
  }
  *n_out = n;
  return recs;
}

// uint16 token stream (OpenWebText .bin): random windows x[B,L], y[B,L] = next tokens
class TokenDataset {
 public:
  explicit TokenDataset(const std::string& path) : f_(path) {}
  size_t num_tokens() const { return f_.size() / 2; }
  void sample(int B, int L, uint64_t seed, int32_t* x, int32_t* y) const {
    const uint16_t* t = reinterpret_cast<const uint16_t*>(f_.data());
    const size_t n = num_tokens();
    if (n < (size_t)L + 1) throw std::runtime_error("token file shorter than sequence length");
    std::mt19937_64 rng(seed);
    for (int b = 0; b < B; ++b) {
      size_t s = rng() % (n - L - 1);
      for (int i = 0; i < L; ++i) {
        x[(size_t)b * L + i] = t[s + i];
        y[(size_t)b * L + i] = t[s + i + 1];
      }
    }
  }
  void window(size_t start, int L, int32_t* x) const {
    const uint16_t* t = reinterpret_cast<const uint16_t*>(f_.data());
    for (int i = 0; i < L; ++i) x[i] = (start + i < num_tokens()) ? t[start + i] : 0;
  }

 private:
  MappedFile f_;
};

}  // namespace tnn
