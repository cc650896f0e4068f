// GPT-2 byte-level BPE tokenizer.
// The reference only decodes (include/tokenizer/tokenizer.hpp:11-68: vocab.bin = u32 n,
// {u32 len, bytes} x n exported from tiktoken by python/export_vocab.py).  Because tiktoken's
// ranks ARE the token ids, the same table is enough to ENCODE as well: split the text with the
// GPT-2 pre-tokenisation rule, then repeatedly merge the adjacent pair whose concatenation has
// the lowest rank.
#pragma once
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <limits>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace tnn {

class Tokenizer {
 public:
  void load(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open vocab file " + path);
    uint32_t n = 0;
    f.read(reinterpret_cast<char*>(&n), 4);
    std::vector<std::string> v(n);
    for (uint32_t i = 0; i < n; ++i) {
      uint32_t len = 0;
      f.read(reinterpret_cast<char*>(&len), 4);
      v[i].resize(len);
      if (len) f.read(&v[i][0], len);
      if (!f) throw std::runtime_error("truncated vocab file " + path);
    }
    set_vocab(std::move(v));
  }
  void set_vocab(std::vector<std::string> v) {
    vocab_ = std::move(v);
    rank_.clear();
    rank_.reserve(vocab_.size() * 2);
    for (size_t i = 0; i < vocab_.size(); ++i)
      if (!vocab_[i].empty()) rank_.emplace(vocab_[i], (int32_t)i);
  }
  void save(const std::string& path) const {
    std::ofstream f(path, std::ios::binary);
    uint32_t n = (uint32_t)vocab_.size();
    f.write(reinterpret_cast<const char*>(&n), 4);
    for (auto& s : vocab_) {
      uint32_t len = (uint32_t)s.size();
      f.write(reinterpret_cast<const char*>(&len), 4);
      f.write(s.data(), len);
    }
  }
  size_t vocab_size() const { return vocab_.size(); }
  std::string decode(const std::vector<int32_t>& ids) const {
    std::string out;
    for (int32_t id : ids)
      if (id >= 0 && (size_t)id < vocab_.size()) out += vocab_[id];
    return out;
  }
  std::string decode_one(int32_t id) const {
    return (id >= 0 && (size_t)id < vocab_.size()) ? vocab_[id] : std::string();
  }

  std::vector<int32_t> encode(const std::string& text) const {
    std::vector<int32_t> out;
    for (const std::string& piece : pretokenize(text)) bpe(piece, out);
    return out;
  }

 private:
  static bool is_letter(unsigned char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c >= 0x80; }
  static bool is_digit(unsigned char c) { return c >= '0' && c <= '9'; }
  static bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

  // 's|'t|'re|'ve|'m|'ll|'d| ?L+| ?N+| ?[^\sLN]+|\s+(?!\S)|\s+
  static std::vector<std::string> pretokenize(const std::string& t) {
    std::vector<std::string> out;
    size_t i = 0, n = t.size();
    auto cls = [&](size_t k) -> int {
      unsigned char c = (unsigned char)t[k];
      if (is_letter(c)) return 0;
      if (is_digit(c)) return 1;
      if (is_space(c)) return 3;
      return 2;
    };
    while (i < n) {
      if (t[i] == '\'' && i + 1 < n) {
        static const char* sfx[] = {"'s", "'t", "'re", "'ve", "'m", "'ll", "'d"};
        bool hit = false;
        for (const char* s : sfx) {
          size_t l = strlen(s);
          if (t.compare(i, l, s) == 0) {
            out.emplace_back(t.substr(i, l));
            i += l;
            hit = true;
            break;
          }
        }
        if (hit) continue;
      }
      size_t j = i;
      if (t[j] == ' ' && j + 1 < n && cls(j + 1) != 3) ++j;  // optional leading space
      if (j < n && cls(j) != 3) {
        int c = cls(j);
        size_t k = j;
        while (k < n && cls(k) == c && !(c == 2 && t[k] == '\'' && k > j && false)) ++k;
        out.emplace_back(t.substr(i, k - i));
        i = k;
        continue;
      }
      // whitespace run; leave the last space for the next word if one follows
      size_t k = i;
      while (k < n && cls(k) == 3) ++k;
      if (k < n && k - i > 1 && t[k - 1] == ' ') --k;
      if (k == i) k = i + 1;
      out.emplace_back(t.substr(i, k - i));
      i = k;
    }
    return out;
  }

  void bpe(const std::string& piece, std::vector<int32_t>& out) const {
    auto whole = rank_.find(piece);
    if (whole != rank_.end()) {
      out.push_back(whole->second);
      return;
    }
    std::vector<std::string> parts;
    for (char c : piece) parts.emplace_back(1, c);
    while (parts.size() > 1) {
      int32_t best = std::numeric_limits<int32_t>::max();
      size_t bi = 0;
      for (size_t i = 0; i + 1 < parts.size(); ++i) {
        auto it = rank_.find(parts[i] + parts[i + 1]);
        if (it != rank_.end() && it->second < best) {
          best = it->second;
          bi = i;
        }
      }
      if (best == std::numeric_limits<int32_t>::max()) break;
      parts[bi] += parts[bi + 1];
      parts.erase(parts.begin() + bi + 1);
    }
    for (auto& p : parts) {
      auto it = rank_.find(p);
      if (it != rank_.end()) out.push_back(it->second);
    }
  }

  std::vector<std::string> vocab_;
  std::unordered_map<std::string, int32_t> rank_;
};

}  // namespace tnn
