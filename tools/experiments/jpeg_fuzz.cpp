// Mutation fuzzing of the C++ host's JPEG decoder (cpp/include/tnn/jpeg.hpp): truncations, byte flips and marker-length edits of
// valid files must produce an error or an image, never a crash / out-of-bounds access.  Build with the sanitizers:
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -Icpp/include tools/experiments/jpeg_fuzz.cpp -o /tmp/jpeg_fuzz
//   /tmp/jpeg_fuzz file.jpg [file.jpg ...] [--iters N]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <random>

#include "tnn/jpeg.hpp"

int main(int argc, char** argv) {
  int iters = 20000;
  std::vector<std::string> files;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--iters") && i + 1 < argc) iters = atoi(argv[++i]);
    else files.push_back(argv[i]);
  }
  std::mt19937_64 rng(1234);
  size_t ok = 0, rejected = 0;
  for (const auto& f : files) {
    std::ifstream in(f, std::ios::binary);
    std::vector<uint8_t> base((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    if (base.empty()) continue;
    for (int it = 0; it < iters; ++it) {
      std::vector<uint8_t> b = base;
      const int kind = (int)(rng() % 4);
      if (kind == 0) b.resize(rng() % b.size());                                   // truncate
      else if (kind == 1) for (int k = 0, n = 1 + (int)(rng() % 8); k < n; ++k) b[rng() % b.size()] = (uint8_t)rng();   // byte noise
      else if (kind == 2) {                                                       // damage the headers specifically
        const size_t lim = std::min<size_t>(b.size(), 700);
        for (int k = 0, n = 1 + (int)(rng() % 4); k < n; ++k) b[rng() % lim] = (uint8_t)rng();
      } else {                                                                     // splice two halves
        const size_t cut = rng() % b.size(), from = rng() % b.size();
        b.resize(cut);
        b.insert(b.end(), base.begin() + (long)from, base.end());
      }
      tnn::jpeg::Image im;
      std::string err;
      if (tnn::jpeg::decode(b.data(), b.size(), &im, &err)) ++ok;
      else ++rejected;
    }
  }
  std::printf("jpeg_fuzz: %zu decoded, %zu rejected, no crash\n", ok, rejected);
  return 0;
}
