// ImageFolderDataLoader: JPEG trees / decoded .npy caches (see tnn/data_loader.hpp).
#include <dirent.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <fstream>
#include <map>
#include <sstream>

#include "tnn/allocator.hpp"
#include "tnn/data_loader.hpp"
#include "tnn/jpeg.hpp"
#include "tnn/utils.hpp"

namespace tnn {

namespace {
bool is_dir(const std::string& p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
bool is_file(const std::string& p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}
std::vector<std::string> list_dir(const std::string& p) {
  std::vector<std::string> out;
  if (DIR* d = opendir(p.c_str())) {
    while (dirent* e = readdir(d)) {
      const std::string n = e->d_name;
      if (n != "." && n != "..") out.push_back(n);
    }
    closedir(d);
  }
  std::sort(out.begin(), out.end());
  return out;
}
bool is_jpeg_name(std::string n) {
  for (auto& c : n) c = (char)tolower((unsigned char)c);
  auto ends = [&](const char* s) { const size_t l = strlen(s); return n.size() >= l && n.compare(n.size() - l, l, s) == 0; };
  return ends(".jpeg") || ends(".jpg");
}
// NumPy .npy (format 1.0 / 2.0 / 3.0): maps the file, checks dtype / C order, returns the payload and the shape
struct Npy {
  void* map = nullptr;
  size_t bytes = 0;
  const uint8_t* payload = nullptr;
  std::vector<size_t> shape;
  std::string descr;
};
bool map_npy(const std::string& path, Npy* out, std::string* err) {
  const int fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) { *err = "cannot open " + path; return false; }
  struct stat st;
  fstat(fd, &st);
  out->bytes = (size_t)st.st_size;
  out->map = mmap(nullptr, out->bytes, PROT_READ, MAP_PRIVATE, fd, 0);
  ::close(fd);
  if (out->map == MAP_FAILED || out->bytes < 12) { *err = "cannot map " + path; out->map = nullptr; return false; }
  const uint8_t* b = static_cast<const uint8_t*>(out->map);
  if (std::memcmp(b, "\x93NUMPY", 6) != 0) { *err = path + ": not an .npy file"; return false; }
  const int major = b[6];
  size_t hlen, hoff;
  if (major == 1) { hlen = (size_t)b[8] | ((size_t)b[9] << 8); hoff = 10; }
  else { hlen = (size_t)b[8] | ((size_t)b[9] << 8) | ((size_t)b[10] << 16) | ((size_t)b[11] << 24); hoff = 12; }
  if (hoff + hlen > out->bytes) { *err = path + ": truncated header"; return false; }
  const std::string h(reinterpret_cast<const char*>(b + hoff), hlen);
  auto field = [&](const std::string& key) -> std::string {
    const size_t k = h.find("'" + key + "'");
    if (k == std::string::npos) return "";
    const size_t c = h.find(':', k);
    return c == std::string::npos ? "" : h.substr(c + 1);
  };
  std::string d = field("descr");
  const size_t q0 = d.find('\''), q1 = d.find('\'', q0 + 1);
  if (q0 == std::string::npos || q1 == std::string::npos) { *err = path + ": no descr"; return false; }
  out->descr = d.substr(q0 + 1, q1 - q0 - 1);
  if (field("fortran_order").find("False") == std::string::npos) { *err = path + ": Fortran order is not supported"; return false; }
  std::string sh = field("shape");
  const size_t p0 = sh.find('('), p1 = sh.find(')');
  if (p0 == std::string::npos || p1 == std::string::npos) { *err = path + ": no shape"; return false; }
  std::stringstream ss(sh.substr(p0 + 1, p1 - p0 - 1));
  std::string tok;
  while (std::getline(ss, tok, ',')) {
    bool digit = false;
    for (char c : tok) digit |= isdigit((unsigned char)c) != 0;
    if (digit) out->shape.push_back((size_t)strtoull(tok.c_str(), nullptr, 10));
  }
  out->payload = b + hoff + hlen;
  return true;
}
}  // namespace

ImageFolderDataLoader::~ImageFolderDataLoader() {
  if (map_x_) munmap(map_x_, map_x_bytes_);
  if (map_y_) munmap(map_y_, map_y_bytes_);
}

bool ImageFolderDataLoader::load_data(const std::string& path) {
  const std::string split = train_ ? "train" : "val";
  const size_t S = sample_shape_[0];
  // 1. decoded caches written by the Python host's ImageFolderDataLoader.cache() (or any tool): no decoding at all
  const std::string xf = path + "/" + split + "_x.npy", yf = path + "/" + split + "_y.npy";
  if (is_file(xf) && is_file(yf)) {
    Npy nx, ny;
    std::string err;
    if (!map_npy(xf, &nx, &err) || !map_npy(yf, &ny, &err)) throw std::runtime_error("ImageFolderDataLoader: " + err);
    if (nx.descr != "|u1" || nx.shape.size() != 4 || nx.shape[1] != S || nx.shape[2] != S || nx.shape[3] != 3)
      throw std::runtime_error("ImageFolderDataLoader: " + xf + " must be uint8 [N, " + std::to_string(S) + ", " + std::to_string(S) + ", 3]");
    if (ny.descr != "<i4" || ny.shape.size() != 1 || ny.shape[0] != nx.shape[0]) throw std::runtime_error("ImageFolderDataLoader: " + yf + " must be int32 [N]");
    map_x_ = nx.map;
    map_x_bytes_ = nx.bytes;
    map_y_ = ny.map;
    map_y_bytes_ = ny.bytes;
    cache_x_ = nx.payload;
    cache_y_ = reinterpret_cast<const int32_t*>(ny.payload);
    n_ = nx.shape[0];
    return true;
  }
  // 2. the image tree
  const std::string root = path + "/" + split;
  if (!is_dir(root)) return false;
  if (is_file(path + "/wnids.txt")) {
    std::ifstream f(path + "/wnids.txt");
    std::string l;
    while (std::getline(f, l)) {
      while (!l.empty() && isspace((unsigned char)l.back())) l.pop_back();
      if (!l.empty()) names_.push_back(l);
    }
  } else {
    const std::string base = is_dir(path + "/train") ? path + "/train" : root;
    for (auto& d : list_dir(base))
      if (is_dir(base + "/" + d)) names_.push_back(d);
  }
  std::map<std::string, int32_t> index;
  for (size_t i = 0; i < names_.size(); ++i) index[names_[i]] = (int32_t)i;
  const std::string ann = root + "/val_annotations.txt";
  if (!train_ && is_file(ann)) {                               // TinyImageNet validation: flat directory + annotation file
    std::ifstream f(ann);
    std::string l;
    while (std::getline(f, l)) {
      std::stringstream ss(l);
      std::string file, wnid;
      if ((ss >> file >> wnid) && index.count(wnid)) {
        files_.push_back(root + "/images/" + file);
        labels_.push_back(index[wnid]);
      }
    }
  } else {
    for (auto& n : names_) {
      std::string d = root + "/" + n;
      if (is_dir(d + "/images")) d += "/images";
      if (!is_dir(d)) continue;
      for (auto& f : list_dir(d))
        if (is_jpeg_name(f)) {
          files_.push_back(d + "/" + f);
          labels_.push_back(index[n]);
        }
    }
  }
  n_ = files_.size();
  if (names_.size() > 0) classes_ = std::max(classes_, names_.size());
  return n_ > 0;
}

void ImageFolderDataLoader::fill_sample(size_t index, float* dst, int32_t* label) {
  const size_t S = sample_shape_[0], per = S * S * 3;
  if (cache_x_) {
    const uint8_t* src = cache_x_ + index * per;
    for (size_t i = 0; i < per; ++i) dst[i] = (float)src[i] / 255.f;
    *label = cache_y_[index];
    return;
  }
  jpeg::Image im, rs;
  std::string err;
  if (!jpeg::decode_file(files_[index], &im, &err)) throw std::runtime_error("ImageFolderDataLoader: " + files_[index] + ": " + err);
  const jpeg::Image* use = &im;
  if ((size_t)im.w != S || (size_t)im.h != S) {
    jpeg::resize_bilinear(im, (int)S, (int)S, &rs);
    use = &rs;
  }
  for (size_t i = 0; i < per; ++i) dst[i] = (float)use->rgb[i] / 255.f;
  *label = labels_[index];
}

bool ImageFolderDataLoader::get_batch(size_t B, Tensor& x, Tensor& y) {
  if (cursor_ >= n_) return false;
  const size_t b = std::min(B, n_ - cursor_);
  Shape xs{b};
  xs.insert(xs.end(), sample_shape_.begin(), sample_shape_.end());
  IAllocator& a = pinned_ ? static_cast<IAllocator&>(DeviceAllocator::pinned()) : HostAllocator();
  if (!x || x->shape() != xs || x->is_gpu()) x = make_tensor(a, DType_t::FP32, xs);
  if (!y || y->size() != b || y->is_gpu()) y = make_tensor(a, DType_t::INT32_T, {b});
  const size_t per = shape_numel(sample_shape_);
  if (order_.size() != n_) {
    order_.resize(n_);
    for (size_t i = 0; i < n_; ++i) order_[i] = (uint32_t)i;
  }
  // decoding is the cost: one image per pool task (the reference decodes on the loader thread)
  std::string first_error;
  std::mutex mu;
  float* xp = x->data_as<float>();
  int32_t* yp = y->data_as<int32_t>();
  parallel_for(0, (int64_t)b, [&](int64_t i) {
    try {
      fill_sample(order_[cursor_ + (size_t)i], xp + (size_t)i * per, yp + i);
    } catch (const std::exception& e) {
      std::lock_guard<std::mutex> g(mu);
      if (first_error.empty()) first_error = e.what();
    }
  });
  if (!first_error.empty()) throw std::runtime_error(first_error);
  cursor_ += b;
  if (augmentation_) augmentation_->apply(*x, batches_served_);
  ++batches_served_;
  return true;
}

}  // namespace tnn
