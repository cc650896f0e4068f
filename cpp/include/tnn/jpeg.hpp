// Baseline JPEG decoder for the image-tree data loaders (reference: data_loading/image_data_loader.hpp decodes with the vendored
// stb_image; nothing is vendored here).  Written against ITU-T T.81: sequential DCT (SOF0 / SOF1, 8-bit), Huffman coding, 1 or 3
// components, sampling factors 1 or 2 per axis (4:4:4, 4:2:2, 4:4:0, 4:2:0), restart intervals, JFIF YCbCr (Adobe transform 0 =
// RGB).  Progressive / arithmetic / 12-bit / CMYK files are rejected with a message (convert them once with the Python host's
// `ImageFolderDataLoader.cache()` and load the .npy cache instead).  Chroma is upsampled with the triangle ("fancy") filter.
// Output: interleaved 8-bit RGB.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

namespace tnn {
namespace jpeg {

struct Image {
  int w = 0, h = 0;
  std::vector<uint8_t> rgb;      // [h][w][3]
};

namespace detail {

static const uint8_t kZigZag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
  // T.81 F.2.2.3: per code length the smallest / largest code and the index of its first symbol
  int mincode[17] = {}, maxcode[18] = {}, valptr[17] = {};
  uint8_t vals[256] = {};
  bool present = false;
  void build(const uint8_t* bits, const uint8_t* v, int n) {
    std::memcpy(vals, v, (size_t)n);
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
      valptr[l] = k;
      mincode[l] = code;
      code += bits[l - 1];
      k += bits[l - 1];
      maxcode[l] = bits[l - 1] ? code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7fffffff;
    present = true;
  }
};

struct Component {
  int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
  int pred = 0;
  int pw = 0, ph = 0;             // padded plane size (multiple of 8 * blocks per MCU)
  std::vector<uint8_t> plane;
};

struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint32_t acc = 0;
  int nbits = 0;
  bool hit_marker = false;
  BitReader(const uint8_t* b, const uint8_t* e) : p(b), end(e) {}
  void fill() {
    while (nbits <= 24) {
      uint32_t byte = 0;
      if (!hit_marker && p < end) {
        if (*p == 0xFF) {
          if (p + 1 < end && p[1] == 0x00) {
            byte = 0xFF;
            p += 2;
          } else {
            hit_marker = true;        // RSTn / EOI / anything else: feed zeros until the caller handles it
          }
        } else {
          byte = *p++;
        }
      }
      acc |= byte << (24 - nbits);
      nbits += 8;
    }
  }
  int bit() {
    if (nbits == 0) fill();
    const int b = (int)(acc >> 31);
    acc <<= 1;
    --nbits;
    return b;
  }
  int bits(int n) {
    if (n == 0) return 0;
    if (nbits < n) fill();
    const int v = (int)(acc >> (32 - n));
    acc <<= n;
    nbits -= n;
    return v;
  }
  void reset() {
    acc = 0;
    nbits = 0;
    hit_marker = false;
  }
};

inline int decode_symbol(BitReader& br, const Huff& h) {
  int code = 0;
  for (int l = 1; l <= 16; ++l) {
    code = (code << 1) | br.bit();
    if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
  }
  return -1;
}
inline int extend(int v, int t) { return t == 0 ? 0 : (v < (1 << (t - 1)) ? v - (1 << t) + 1 : v); }

// separable 8x8 inverse DCT in float, output level-shifted and clamped to [0, 255]
inline void idct8x8(const float* in, uint8_t* out, int stride) {
  struct Table {
    float c[8][8];
    Table() {
      for (int x = 0; x < 8; ++x)
        for (int u = 0; u < 8; ++u) c[x][u] = (u == 0 ? 0.353553390593f : 0.5f) * std::cos((2 * x + 1) * u * 3.14159265358979323846f / 16.f);
    }
  };
  static const Table table;            // thread-safe one-time initialisation (decoding runs on a thread pool)
  const auto& c = table.c;
  float tmp[64];
  for (int y = 0; y < 8; ++y)          // rows: tmp[y][x] = sum_u c[x][u] * in[y][u]
    for (int x = 0; x < 8; ++x) {
      float s = 0.f;
      for (int u = 0; u < 8; ++u) s += c[x][u] * in[y * 8 + u];
      tmp[y * 8 + x] = s;
    }
  for (int x = 0; x < 8; ++x)          // columns
    for (int y = 0; y < 8; ++y) {
      float s = 0.f;
      for (int v = 0; v < 8; ++v) s += c[y][v] * tmp[v * 8 + x];
      const int q = (int)std::lround(s + 128.f);
      out[y * stride + x] = (uint8_t)(q < 0 ? 0 : q > 255 ? 255 : q);
    }
}

// triangle-filter upsampling of one plane by (fx, fy) in {1, 2} into a w x h destination (edge samples replicated)
inline void upsample(const Component& c, int fx, int fy, int w, int h, std::vector<uint8_t>* dst) {
  dst->assign((size_t)w * h, 0);
  const int sw = (w + fx - 1) / fx, sh = (h + fy - 1) / fy;     // meaningful source extent
  auto S = [&](int x, int y) -> int {
    x = x < 0 ? 0 : x >= sw ? sw - 1 : x;
    y = y < 0 ? 0 : y >= sh ? sh - 1 : y;
    return c.plane[(size_t)y * c.pw + x];
  };
  for (int y = 0; y < h; ++y) {
    const int sy = y / fy, ny = fy == 2 ? ((y & 1) ? sy + 1 : sy - 1) : sy;
    for (int x = 0; x < w; ++x) {
      const int sx = x / fx, nx = fx == 2 ? ((x & 1) ? sx + 1 : sx - 1) : sx;
      int v;
      if (fx == 1 && fy == 1) v = S(sx, sy);
      else if (fx == 2 && fy == 1) v = (3 * S(sx, sy) + S(nx, sy) + ((x & 1) ? 2 : 1)) >> 2;
      else if (fx == 1 && fy == 2) v = (3 * S(sx, sy) + S(sx, ny) + ((y & 1) ? 2 : 1)) >> 2;
      else {
        const int near_row = 3 * S(sx, sy) + S(sx, ny), far_row = 3 * S(nx, sy) + S(nx, ny);
        v = (3 * near_row + far_row + ((x & 1) ? 7 : 8)) >> 4;
      }
      (*dst)[(size_t)y * w + x] = (uint8_t)v;
    }
  }
}

}  // namespace detail

inline bool decode(const uint8_t* data, size_t n, Image* out, std::string* err) {
  using namespace detail;
  auto fail = [&](const std::string& m) {
    if (err) *err = m;
    return false;
  };
  if (n < 4 || data[0] != 0xFF || data[1] != 0xD8) return fail("not a JPEG (no SOI marker)");
  uint16_t qt[4][64] = {};
  bool qt_ok[4] = {};
  Huff dc[4], ac[4];
  std::vector<Component> comps;
  int W = 0, H = 0, restart = 0, adobe_transform = -1;
  size_t pos = 2;
  auto u16 = [&](size_t p) { return (int)((data[p] << 8) | data[p + 1]); };
  while (pos + 4 <= n) {
    if (data[pos] != 0xFF) return fail("corrupt marker stream");
    while (pos < n && data[pos] == 0xFF) ++pos;              // fill bytes
    if (pos >= n) break;
    const int m = data[pos++];
    if (m == 0xD9) break;                                     // EOI before any scan
    if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;     // no payload
    if (pos + 2 > n) return fail("truncated segment");
    const int len = u16(pos);
    if (len < 2 || pos + (size_t)len > n) return fail("truncated segment");
    const uint8_t* seg = data + pos + 2;
    const int sl = len - 2;
    if (m == 0xDB) {                                          // DQT
      int i = 0;
      while (i < sl) {
        const int pq = seg[i] >> 4, tq = seg[i] & 15;
        ++i;
        if (tq > 3 || i + (pq ? 128 : 64) > sl) return fail("bad quantisation table");
        for (int k = 0; k < 64; ++k) {
          qt[tq][kZigZag[k]] = pq ? (uint16_t)((seg[i] << 8) | seg[i + 1]) : seg[i];
          i += pq ? 2 : 1;
        }
        qt_ok[tq] = true;
      }
    } else if (m == 0xC4) {                                   // DHT
      int i = 0;
      while (i + 17 <= sl) {
        const int tc = seg[i] >> 4, th = seg[i] & 15;
        if (tc > 1 || th > 3) return fail("bad Huffman table id");
        int cnt = 0;
        for (int k = 0; k < 16; ++k) cnt += seg[i + 1 + k];
        if (cnt > 256 || i + 17 + cnt > sl) return fail("bad Huffman table");
        (tc ? ac[th] : dc[th]).build(seg + i + 1, seg + i + 17, cnt);
        i += 17 + cnt;
      }
    } else if (m == 0xC0 || m == 0xC1) {                      // SOF0 / SOF1: sequential DCT
      if (sl < 6 || seg[0] != 8) return fail("only 8-bit samples are supported");
      H = u16(pos + 3);
      W = u16(pos + 5);
      const int nc = seg[5];
      if (W <= 0 || H <= 0 || (nc != 1 && nc != 3) || sl < 6 + 3 * nc) return fail("unsupported frame header (components: " + std::to_string(nc) + ")");
      if ((long long)W * H > 100ll * 1000 * 1000) return fail("image larger than 100 megapixels");     // a damaged header must not allocate gigabytes
      comps.resize(nc);
      for (int k = 0; k < nc; ++k) {
        comps[k].id = seg[6 + 3 * k];
        comps[k].h = seg[7 + 3 * k] >> 4;
        comps[k].v = seg[7 + 3 * k] & 15;
        comps[k].tq = seg[8 + 3 * k];
        if (comps[k].h < 1 || comps[k].h > 2 || comps[k].v < 1 || comps[k].v > 2 || comps[k].tq > 3) return fail("unsupported sampling factors");
      }
    } else if (m == 0xC2) {
      return fail("progressive JPEG is not supported (decode it once with the Python host's cache())");
    } else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      return fail("unsupported JPEG process (SOF" + std::to_string(m - 0xC0) + ")");
    } else if (m == 0xDD) {                                   // DRI
      if (sl >= 2) restart = u16(pos + 2);
    } else if (m == 0xEE) {                                   // APP14 Adobe
      if (sl >= 12 && std::memcmp(seg, "Adobe", 5) == 0) adobe_transform = seg[11];
    } else if (m == 0xDA) {                                   // SOS: the (single, interleaved) baseline scan
      if (comps.empty()) return fail("scan before frame header");
      const int ns = seg[0];
      if (ns != (int)comps.size() || sl < 1 + 2 * ns + 3) return fail("non-interleaved scans are not supported");
      for (int k = 0; k < ns; ++k) {
        const int cid = seg[1 + 2 * k];
        Component* c = nullptr;
        for (auto& cc : comps)
          if (cc.id == cid) c = &cc;
        if (!c) return fail("scan refers to an unknown component");
        c->td = seg[2 + 2 * k] >> 4;
        c->ta = seg[2 + 2 * k] & 15;
        if (c->td > 3 || c->ta > 3 || !dc[c->td].present || !ac[c->ta].present || !qt_ok[c->tq]) return fail("scan uses a missing table");
      }
      int hmax = 1, vmax = 1;
      for (auto& c : comps) {
        hmax = std::max(hmax, c.h);
        vmax = std::max(vmax, c.v);
      }
      if (comps.size() == 1) comps[0].h = comps[0].v = hmax = vmax = 1;      // single-component scans are never subsampled
      const int mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
      for (auto& c : comps) {
        c.pw = mcux * c.h * 8;
        c.ph = mcuy * c.v * 8;
        c.plane.assign((size_t)c.pw * c.ph, 0);
        c.pred = 0;
      }
      BitReader br(data + pos + (size_t)len, data + n);
      float blk[64];
      int until_restart = restart;
      for (int my = 0; my < mcuy; ++my)
        for (int mx = 0; mx < mcux; ++mx) {
          if (restart && until_restart == 0) {
            // byte-align, expect RSTn, reset the DC predictors
            br.reset();
            while (br.p + 1 < br.end && !(br.p[0] == 0xFF && br.p[1] >= 0xD0 && br.p[1] <= 0xD7)) ++br.p;
            if (br.p + 1 < br.end) br.p += 2;
            for (auto& c : comps) c.pred = 0;
            until_restart = restart;
          }
          for (auto& c : comps)
            for (int by = 0; by < c.v; ++by)
              for (int bx = 0; bx < c.h; ++bx) {
                std::memset(blk, 0, sizeof(blk));
                const int t = decode_symbol(br, dc[c.td]);
                if (t < 0 || t > 11) return fail("corrupt DC coefficient");
                c.pred += extend(br.bits(t), t);
                blk[0] = (float)(c.pred * (int)qt[c.tq][0]);
                for (int k = 1; k < 64;) {
                  const int rs = decode_symbol(br, ac[c.ta]);
                  if (rs < 0) return fail("corrupt AC coefficient");
                  const int r = rs >> 4, s = rs & 15;
                  if (s == 0) {
                    if (r == 15) {
                      k += 16;
                      continue;
                    }
                    break;                                    // end of block
                  }
                  k += r;
                  if (k > 63) return fail("corrupt AC run");
                  const int z = kZigZag[k];
                  blk[z] = (float)(extend(br.bits(s), s) * (int)qt[c.tq][z]);
                  ++k;
                }
                idct8x8(blk, &c.plane[(size_t)((my * c.v + by) * 8) * c.pw + (mx * c.h + bx) * 8], c.pw);
              }
          if (restart) --until_restart;
        }
      // planes -> RGB
      out->w = W;
      out->h = H;
      out->rgb.assign((size_t)W * H * 3, 0);
      if (comps.size() == 1) {
        for (int y = 0; y < H; ++y)
          for (int x = 0; x < W; ++x) {
            const uint8_t g = comps[0].plane[(size_t)y * comps[0].pw + x];
            uint8_t* px = &out->rgb[((size_t)y * W + x) * 3];
            px[0] = px[1] = px[2] = g;
          }
        return true;
      }
      std::vector<uint8_t> full[3];
      for (int k = 0; k < 3; ++k) upsample(comps[k], hmax / comps[k].h, vmax / comps[k].v, W, H, &full[k]);
      const bool ycc = adobe_transform != 0;
      for (size_t i = 0; i < (size_t)W * H; ++i) {
        uint8_t* px = &out->rgb[i * 3];
        if (!ycc) {
          px[0] = full[0][i];
          px[1] = full[1][i];
          px[2] = full[2][i];
          continue;
        }
        const float Y = full[0][i], cb = (float)full[1][i] - 128.f, cr = (float)full[2][i] - 128.f;
        const int r = (int)std::lround(Y + 1.402f * cr), g = (int)std::lround(Y - 0.344136f * cb - 0.714136f * cr), b = (int)std::lround(Y + 1.772f * cb);
        px[0] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
        px[1] = (uint8_t)(g < 0 ? 0 : g > 255 ? 255 : g);
        px[2] = (uint8_t)(b < 0 ? 0 : b > 255 ? 255 : b);
      }
      return true;
    }
    pos += (size_t)len;
  }
  return fail("no image data (missing SOS)");
}

inline bool decode_file(const std::string& path, Image* out, std::string* err) {
  std::ifstream f(path, std::ios::binary);
  if (!f) {
    if (err) *err = "cannot open " + path;
    return false;
  }
  std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  return decode(buf.data(), buf.size(), out, err);
}

// bilinear resize (pixel centres aligned, no anti-aliasing): used when a tree's images are not already size x size
inline void resize_bilinear(const Image& src, int W, int H, Image* dst) {
  dst->w = W;
  dst->h = H;
  dst->rgb.assign((size_t)W * H * 3, 0);
  const float sx = (float)src.w / W, sy = (float)src.h / H;
  for (int y = 0; y < H; ++y) {
    float fy = (y + 0.5f) * sy - 0.5f;
    fy = fy < 0 ? 0 : fy;
    const int y0 = std::min((int)fy, src.h - 1), y1 = std::min(y0 + 1, src.h - 1);
    const float wy = fy - y0;
    for (int x = 0; x < W; ++x) {
      float fx = (x + 0.5f) * sx - 0.5f;
      fx = fx < 0 ? 0 : fx;
      const int x0 = std::min((int)fx, src.w - 1), x1 = std::min(x0 + 1, src.w - 1);
      const float wx = fx - x0;
      for (int c = 0; c < 3; ++c) {
        const float a = src.rgb[((size_t)y0 * src.w + x0) * 3 + c], b = src.rgb[((size_t)y0 * src.w + x1) * 3 + c];
        const float d = src.rgb[((size_t)y1 * src.w + x0) * 3 + c], e = src.rgb[((size_t)y1 * src.w + x1) * 3 + c];
        const float v = (a * (1 - wx) + b * wx) * (1 - wy) + (d * (1 - wx) + e * wx) * wy;
        dst->rgb[((size_t)y * W + x) * 3 + c] = (uint8_t)std::lround(v);
      }
    }
  }
}

}  // namespace jpeg
}  // namespace tnn
