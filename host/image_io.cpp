#include "image_io.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <thread>

namespace asuna_host {

namespace {

std::vector<uint8_t> read_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("failed to load " + path);
  return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
void write_file(const std::string& path, const std::vector<uint8_t>& data) {
  std::ofstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot write " + path);
  f.write((const char*)data.data(), (std::streamsize)data.size());
}
std::string lower_ext(const std::string& path) {
  size_t dot = path.find_last_of('.');
  std::string e = dot == std::string::npos ? "" : path.substr(dot + 1);
  for (auto& c : e) c = (char)tolower(c);
  return e;
}
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void put_be32(std::vector<uint8_t>& v, uint32_t x) {
  v.push_back(x >> 24), v.push_back(x >> 16), v.push_back(x >> 8), v.push_back(x);
}
template <class T>
void put_le(std::vector<uint8_t>& v, T x) {
  uint8_t b[sizeof(T)];
  std::memcpy(b, &x, sizeof(T));
  v.insert(v.end(), b, b + sizeof(T));
}

std::vector<uint8_t> zlib_inflate(const uint8_t* src, size_t n, size_t expected) {
  std::vector<uint8_t> out(expected);
  uLongf len = (uLongf)expected;
  int rc = uncompress(out.data(), &len, src, (uLong)n);
  if (rc != Z_OK) throw std::runtime_error("zlib: inflate failed (" + std::to_string(rc) + ")");
  out.resize(len);
  return out;
}
// One zlib stream deflated by several threads (the pigz construction): the input is cut into bands, every band becomes
// raw deflate data that ends on a byte boundary (Z_SYNC_FLUSH; the last band ends the stream with Z_FINISH), and the
// pieces are concatenated behind one zlib header with the Adler-32 of the whole input.  Any inflater reads it as an
// ordinary stream; a 1080p RGBA8 image takes ~20 ms instead of ~250 (it was a fifth of a 256-spp job's wall time).
std::vector<uint8_t> zlib_deflate(const std::vector<uint8_t>& src) {
  const size_t kMinBand = 256 * 1024;
  const size_t hw = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  const size_t bands = std::max<size_t>(1, std::min(hw, (src.size() + kMinBand - 1) / kMinBand));
  const size_t per = (src.size() + bands - 1) / bands;
  std::vector<std::vector<uint8_t>> piece(bands);
  std::vector<int> ok(bands, 0);
  auto work = [&](size_t b) {
    const size_t begin = std::min(src.size(), b * per), end = std::min(src.size(), begin + per);
    z_stream z{};
    if (deflateInit2(&z, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return;  // raw deflate
    piece[b].resize(deflateBound(&z, (uLong)(end - begin)) + 16);
    z.next_in = const_cast<Bytef*>(src.data() + begin), z.avail_in = (uInt)(end - begin);
    z.next_out = piece[b].data(), z.avail_out = (uInt)piece[b].size();
    const int rc = deflate(&z, b + 1 == bands ? Z_FINISH : Z_SYNC_FLUSH);
    ok[b] = (b + 1 == bands ? rc == Z_STREAM_END : rc == Z_OK) && z.avail_in == 0;
    piece[b].resize(z.total_out);
    deflateEnd(&z);
  };
  std::vector<std::thread> th;
  for (size_t b = 1; b < bands; b++) th.emplace_back(work, b);
  work(0);
  for (auto& t : th) t.join();
  std::vector<uint8_t> out = {0x78, 0x9C};
  for (size_t b = 0; b < bands; b++) {
    if (!ok[b]) throw std::runtime_error("zlib: deflate failed");
    out.insert(out.end(), piece[b].begin(), piece[b].end());
  }
  uLong adler = adler32(0L, Z_NULL, 0);
  for (size_t off = 0; off < src.size(); off += (size_t)1 << 30)
    adler = adler32(adler, src.data() + off, (uInt)std::min<size_t>((size_t)1 << 30, src.size() - off));
  put_be32(out, (uint32_t)adler);
  return out;
}

// ---------------------------------------------------------------------------------------------- PNG
struct Png {
  int w = 0, h = 0, channels = 0, depth = 0;
  std::vector<uint16_t> samples;  // w*h*channels, native bit depth
  int maxval = 255;
};

Png decode_png(const std::vector<uint8_t>& f, const std::string& name) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (f.size() < 8 || std::memcmp(f.data(), sig, 8) != 0) throw std::runtime_error(name + ": not a PNG file");
  Png img;
  int color = 0, interlace = 0;
  std::vector<uint8_t> idat, plte, trns;
  size_t p = 8;
  while (p + 12 <= f.size()) {
    uint32_t len = be32(&f[p]);
    std::string type((const char*)&f[p + 4], 4);
    const uint8_t* d = &f[p + 8];
    if (p + 12 + len > f.size()) throw std::runtime_error(name + ": truncated PNG");
    if (type == "IHDR") {
      if (len != 13) throw std::runtime_error(name + ": bad PNG IHDR length");
      img.w = (int)be32(d), img.h = (int)be32(d + 4);
      img.depth = d[8], color = d[9], interlace = d[12];
    } else if (type == "PLTE") plte.assign(d, d + len);
    else if (type == "tRNS") trns.assign(d, d + len);
    else if (type == "IDAT") idat.insert(idat.end(), d, d + len);
    else if (type == "IEND") break;
    p += 12 + len;
  }
  if (interlace) throw std::runtime_error(name + ": interlaced PNG is not supported");
  int spp = color == 0 ? 1 : color == 2 ? 3 : color == 3 ? 1 : color == 4 ? 2 : color == 6 ? 4 : 0;
  if (!spp || img.w <= 0 || img.h <= 0) throw std::runtime_error(name + ": bad PNG header");
  // bit depths the PNG specification allows per colour type; anything else would divide by a zero maxval or shift
  // by a negative amount below
  const int dep = img.depth;
  const bool depth_ok = color == 0 ? (dep == 1 || dep == 2 || dep == 4 || dep == 8 || dep == 16)
                        : color == 3 ? (dep == 1 || dep == 2 || dep == 4 || dep == 8)
                                     : (dep == 8 || dep == 16);
  if (!depth_ok) throw std::runtime_error(name + ": bad PNG bit depth");
  if ((uint64_t)img.w * (uint64_t)img.h > (1ull << 28)) throw std::runtime_error(name + ": PNG larger than 2^28 pixels");
  size_t bpp_bits = (size_t)spp * img.depth, stride = (img.w * bpp_bits + 7) / 8, bpp = std::max<size_t>(1, bpp_bits / 8);
  std::vector<uint8_t> raw = zlib_inflate(idat.data(), idat.size(), (stride + 1) * img.h);
  if (raw.size() != (stride + 1) * img.h) throw std::runtime_error(name + ": PNG data size mismatch");
  std::vector<uint8_t> prev(stride, 0), cur(stride);
  std::vector<uint16_t> s((size_t)img.w * img.h * spp);
  for (int y = 0; y < img.h; y++) {
    const uint8_t* row = &raw[(stride + 1) * y];
    int filter = row[0];
    for (size_t i = 0; i < stride; i++) {
      int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0, x = row[1 + i];
      int v;
      switch (filter) {
        case 0: v = x; break;
        case 1: v = x + a; break;
        case 2: v = x + b; break;
        case 3: v = x + ((a + b) >> 1); break;
        case 4: {
          int pa = std::abs(b - c), pb = std::abs(a - c), pc = std::abs(a + b - 2 * c);
          v = x + (pa <= pb && pa <= pc ? a : (pb <= pc ? b : c));
          break;
        }
        default: throw std::runtime_error(name + ": bad PNG filter");
      }
      cur[i] = (uint8_t)v;
    }
    for (int x = 0; x < img.w * spp; x++) {
      uint16_t v;
      if (img.depth == 8) v = cur[x];
      else if (img.depth == 16) v = (uint16_t)((cur[2 * x] << 8) | cur[2 * x + 1]);
      else {
        size_t bit = (size_t)x * img.depth;
        v = (cur[bit / 8] >> (8 - img.depth - bit % 8)) & ((1 << img.depth) - 1);
      }
      s[((size_t)y * img.w) * spp + x] = v;
    }
    prev = cur;
  }
  img.maxval = (1 << img.depth) - 1;
  if (color == 3) {  // palette -> RGBA8
    std::vector<uint16_t> o((size_t)img.w * img.h * 4);
    for (size_t i = 0; i < (size_t)img.w * img.h; i++) {
      size_t k = s[i];
      if (3 * k + 2 >= plte.size()) throw std::runtime_error(name + ": palette index out of range");
      o[4 * i] = plte[3 * k], o[4 * i + 1] = plte[3 * k + 1], o[4 * i + 2] = plte[3 * k + 2];
      o[4 * i + 3] = k < trns.size() ? trns[k] : 255;
    }
    img.samples = std::move(o), img.channels = 4, img.maxval = 255;
  } else {
    img.samples = std::move(s), img.channels = spp;
  }
  return img;
}

void png_chunk(std::vector<uint8_t>& out, const char* type, const std::vector<uint8_t>& data) {
  put_be32(out, (uint32_t)data.size());
  size_t start = out.size();
  out.insert(out.end(), type, type + 4);
  out.insert(out.end(), data.begin(), data.end());
  put_be32(out, (uint32_t)crc32(0L, out.data() + start, (uInt)(out.size() - start)));
}

void encode_png_rgba8(const std::string& path, int w, int h, const uint8_t* rgba) {
  std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  std::vector<uint8_t> ihdr;
  put_be32(ihdr, (uint32_t)w), put_be32(ihdr, (uint32_t)h);
  ihdr.insert(ihdr.end(), {8, 6, 0, 0, 0});
  png_chunk(out, "IHDR", ihdr);
  std::vector<uint8_t> raw((size_t)(w * 4 + 1) * h);
  for (int y = 0; y < h; y++) {
    raw[(size_t)(w * 4 + 1) * y] = 0;
    std::memcpy(&raw[(size_t)(w * 4 + 1) * y + 1], rgba + (size_t)w * 4 * y, (size_t)w * 4);
  }
  png_chunk(out, "IDAT", zlib_deflate(raw));
  png_chunk(out, "IEND", {});
  write_file(path, out);
}

// ---------------------------------------------------------------------------------------------- Radiance .hdr
ImageF decode_hdr(const std::vector<uint8_t>& f, const std::string& name) {
  size_t p = 0;
  auto line = [&]() {
    std::string s;
    while (p < f.size() && f[p] != '\n') s += (char)f[p++];
    p++;
    return s;
  };
  std::string first = line();
  if (first.rfind("#?", 0) != 0) throw std::runtime_error(name + ": not a Radiance HDR file");
  while (p < f.size()) {
    std::string s = line();
    if (s.empty()) break;
  }
  std::string res = line();
  int w = 0, h = 0;
  if (sscanf(res.c_str(), "-Y %d +X %d", &h, &w) != 2) throw std::runtime_error(name + ": unsupported HDR orientation");
  ImageF img;
  img.w = w, img.h = h, img.px.resize((size_t)w * h * 4);
  std::vector<uint8_t> scan((size_t)w * 4);
  for (int y = 0; y < h; y++) {
    if (p + 4 > f.size()) throw std::runtime_error(name + ": truncated HDR");
    bool rle = w >= 8 && w < 32768 && f[p] == 2 && f[p + 1] == 2 && ((f[p + 2] << 8) | f[p + 3]) == w;
    if (rle) {
      p += 4;
      for (int c = 0; c < 4; c++) {
        int x = 0;
        while (x < w) {
          if (p >= f.size()) throw std::runtime_error(name + ": truncated HDR");
          int n = f[p++];
          if (n > 128) {
            n -= 128;
            uint8_t v = f[p++];
            while (n-- && x < w) scan[4 * x++ + c] = v;
          } else {
            while (n-- && x < w) scan[4 * x++ + c] = f[p++];
          }
        }
      }
    } else {
      if (p + (size_t)w * 4 > f.size()) throw std::runtime_error(name + ": truncated HDR");
      std::memcpy(scan.data(), &f[p], (size_t)w * 4);
      p += (size_t)w * 4;
    }
    for (int x = 0; x < w; x++) {
      const uint8_t* e = &scan[4 * x];
      float* o = &img.px[((size_t)y * w + x) * 4];
      if (e[3]) {
        float s = std::ldexp(1.0f, (int)e[3] - (128 + 8));  // stbi__hdr_convert
        o[0] = e[0] * s, o[1] = e[1] * s, o[2] = e[2] * s;
      } else
        o[0] = o[1] = o[2] = 0.f;
      o[3] = 1.f;
    }
  }
  return img;
}

void encode_hdr(const std::string& path, int w, int h, const float* rgba) {
  std::string head = "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y " + std::to_string(h) + " +X " + std::to_string(w) + "\n";
  std::vector<uint8_t> out(head.begin(), head.end());
  for (size_t i = 0; i < (size_t)w * h; i++) {
    const float* c = rgba + 4 * i;
    float m = std::max(c[0], std::max(c[1], c[2]));
    uint8_t e[4] = {0, 0, 0, 0};
    if (m >= 1e-32f) {
      int ex;
      float n = std::frexp(m, &ex) * 256.0f / m;
      e[0] = (uint8_t)(c[0] * n), e[1] = (uint8_t)(c[1] * n), e[2] = (uint8_t)(c[2] * n), e[3] = (uint8_t)(ex + 128);
    }
    out.insert(out.end(), e, e + 4);
  }
  write_file(path, out);
}

// ---------------------------------------------------------------------------------------------- PFM / NPY
ImageF decode_pfm(const std::vector<uint8_t>& f, const std::string& name) {
  std::string s((const char*)f.data(), std::min<size_t>(f.size(), 128));
  std::istringstream is(s);
  std::string magic;
  int w, h;
  float scale;
  is >> magic >> w >> h >> scale;
  int nc = magic == "PF" ? 3 : (magic == "Pf" ? 1 : 0);
  if (!nc) throw std::runtime_error(name + ": not a PFM file");
  size_t off = (size_t)is.tellg() + 1;
  if (off + (size_t)w * h * nc * 4 > f.size()) throw std::runtime_error(name + ": truncated PFM");
  ImageF img;
  img.w = w, img.h = h, img.px.assign((size_t)w * h * 4, 1.f);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      for (int c = 0; c < 3; c++) {
        float v;
        std::memcpy(&v, &f[off + (((size_t)(h - 1 - y) * w + x) * nc + (nc == 3 ? c : 0)) * 4], 4);  // bottom-up, little endian
        img.px[((size_t)y * w + x) * 4 + c] = v;
      }
  if (scale > 0) throw std::runtime_error(name + ": big-endian PFM is not supported");
  return img;
}

ImageF decode_npy(const std::vector<uint8_t>& f, const std::string& name) {
  if (f.size() < 10 || std::memcmp(f.data(), "\x93NUMPY", 6) != 0) throw std::runtime_error(name + ": not an NPY file");
  size_t hlen = f[6] == 1 ? (size_t)(f[8] | (f[9] << 8)) : (size_t)(f[8] | (f[9] << 8) | (f[10] << 16) | ((size_t)f[11] << 24));
  size_t hoff = f[6] == 1 ? 10 : 12;
  std::string head((const char*)&f[hoff], hlen);
  if (head.find("'<f4'") == std::string::npos || head.find("'fortran_order': False") == std::string::npos)
    throw std::runtime_error(name + ": only C-order float32 NPY is supported");
  size_t a = head.find('(', head.find("shape")), b = head.find(')', a);
  std::vector<size_t> shape;
  std::istringstream is(head.substr(a + 1, b - a - 1));
  std::string tok;
  while (std::getline(is, tok, ','))
    if (tok.find_first_of("0123456789") != std::string::npos) shape.push_back((size_t)std::stoul(tok));
  if (shape.size() < 2) throw std::runtime_error(name + ": NPY image must be at least 2-D");
  int h = (int)shape[0], w = (int)shape[1], nc = shape.size() > 2 ? (int)shape[2] : 1;
  const float* d = (const float*)&f[hoff + hlen];
  ImageF img;
  img.w = w, img.h = h, img.px.assign((size_t)w * h * 4, 1.f);
  for (size_t i = 0; i < (size_t)w * h; i++)
    for (int c = 0; c < 4; c++) {
      if (nc == 1 && c < 3) img.px[4 * i + c] = d[i];
      else if (c < nc) img.px[4 * i + c] = d[i * nc + c];
    }
  return img;
}

// ---------------------------------------------------------------------------------------------- OpenEXR (scanline)
struct ExrChannel {
  std::string name;
  int type;  // 0 uint, 1 half, 2 float
};

ImageF decode_exr(const std::vector<uint8_t>& f, const std::string& name) {
  if (f.size() < 8 || f[0] != 0x76 || f[1] != 0x2f || f[2] != 0x31 || f[3] != 0x01) throw std::runtime_error(name + ": not an EXR file");
  if (f[5] & 0x1a) throw std::runtime_error(name + ": tiled / deep / multi-part EXR is not supported");
  size_t p = 8;
  std::vector<ExrChannel> chans;
  int compression = 0, dw[4] = {0, 0, 0, 0};
  auto cstr = [&]() {
    std::string s;
    while (p < f.size() && f[p]) s += (char)f[p++];
    p++;
    return s;
  };
  while (p < f.size() && f[p]) {
    std::string an = cstr(), at = cstr();
    int32_t sz;
    std::memcpy(&sz, &f[p], 4);
    p += 4;
    size_t end = p + sz;
    if (an == "channels") {
      while (p < end && f[p]) {
        ExrChannel c;
        c.name = cstr();
        int32_t t;
        std::memcpy(&t, &f[p], 4);
        c.type = t;
        p += 16;
        chans.push_back(c);
      }
    } else if (an == "compression") compression = f[p];
    else if (an == "dataWindow") std::memcpy(dw, &f[p], 16);
    p = end;
  }
  p++;
  int w = dw[2] - dw[0] + 1, h = dw[3] - dw[1] + 1;
  int lines_per_block = compression == 0 || compression == 2 ? 1 : (compression == 3 ? 16 : 0);
  if (!lines_per_block) throw std::runtime_error(name + ": only uncompressed / ZIP EXR is supported (compression " + std::to_string(compression) + ")");
  size_t row_bytes = 0;
  for (auto& c : chans) row_bytes += (size_t)w * (c.type == 1 ? 2 : 4);
  int n_blocks = (h + lines_per_block - 1) / lines_per_block;
  ImageF img;
  img.w = w, img.h = h, img.px.assign((size_t)w * h * 4, 0.f);
  for (size_t i = 0; i < (size_t)w * h; i++) img.px[4 * i + 3] = 1.f;
  for (int b = 0; b < n_blocks; b++) {
    uint64_t off;
    std::memcpy(&off, &f[p + 8 * (size_t)b], 8);
    int32_t y0, sz;
    std::memcpy(&y0, &f[off], 4);
    std::memcpy(&sz, &f[off + 4], 4);
    int lines = std::min(lines_per_block, dw[3] + 1 - y0);
    size_t raw_size = row_bytes * lines;
    std::vector<uint8_t> raw;
    if (compression == 0 || (size_t)sz == raw_size) raw.assign(&f[off + 8], &f[off + 8] + sz);
    else {
      std::vector<uint8_t> t = zlib_inflate(&f[off + 8], (size_t)sz, raw_size);
      for (size_t i = 1; i < t.size(); i++) t[i] = (uint8_t)(t[i - 1] + t[i] - 128);  // predictor
      raw.resize(t.size());
      size_t half = (t.size() + 1) / 2;  // de-interleave
      for (size_t i = 0; i < t.size(); i++) raw[i] = i % 2 == 0 ? t[i / 2] : t[half + i / 2];
    }
    for (int l = 0; l < lines; l++) {
      const uint8_t* q = &raw[row_bytes * l];
      int y = y0 - dw[1] + l;
      for (auto& c : chans) {
        int dst = c.name == "R" ? 0 : c.name == "G" ? 1 : c.name == "B" ? 2 : c.name == "A" ? 3 : (c.name == "Y" ? 4 : -1);
        for (int x = 0; x < w; x++) {
          float v;
          if (c.type == 1) {
            uint16_t hv;
            std::memcpy(&hv, q + 2 * x, 2);
            v = half_to_float(hv);
          } else if (c.type == 2) std::memcpy(&v, q + 4 * x, 4);
          else {
            uint32_t u;
            std::memcpy(&u, q + 4 * x, 4);
            v = (float)u;
          }
          float* o = &img.px[((size_t)y * w + x) * 4];
          if (dst == 4) o[0] = o[1] = o[2] = v;
          else if (dst >= 0) o[dst] = v;
        }
        q += (size_t)w * (c.type == 1 ? 2 : 4);
      }
    }
  }
  return img;
}

void exr_attr(std::vector<uint8_t>& out, const char* name, const char* type, const std::vector<uint8_t>& data) {
  out.insert(out.end(), name, name + strlen(name) + 1);
  out.insert(out.end(), type, type + strlen(type) + 1);
  put_le<int32_t>(out, (int32_t)data.size());
  out.insert(out.end(), data.begin(), data.end());
}

// Half-float RGB scanline file, one scanline per block, no compression: what RgbaOutputFile(name, display, data,
// WRITE_RGB) stores (reference src/core/texture.cpp:50-75) except for the compression method.
void encode_exr_half_rgb(const std::string& path, int w, int h, const float* rgba) {
  std::vector<uint8_t> out = {0x76, 0x2f, 0x31, 0x01, 2, 0, 0, 0};
  {
    std::vector<uint8_t> ch;
    for (const char* n : {"B", "G", "R"}) {  // alphabetical
      ch.push_back((uint8_t)n[0]), ch.push_back(0);
      put_le<int32_t>(ch, 1);  // HALF
      ch.insert(ch.end(), {0, 0, 0, 0});
      put_le<int32_t>(ch, 1), put_le<int32_t>(ch, 1);
    }
    ch.push_back(0);
    exr_attr(out, "channels", "chlist", ch);
  }
  exr_attr(out, "compression", "compression", {0});
  std::vector<uint8_t> box;
  put_le<int32_t>(box, 0), put_le<int32_t>(box, 0), put_le<int32_t>(box, w - 1), put_le<int32_t>(box, h - 1);
  exr_attr(out, "dataWindow", "box2i", box);
  exr_attr(out, "displayWindow", "box2i", box);
  exr_attr(out, "lineOrder", "lineOrder", {0});
  std::vector<uint8_t> f1, v2;
  put_le<float>(f1, 1.0f);
  exr_attr(out, "pixelAspectRatio", "float", f1);
  put_le<float>(v2, 0.f), put_le<float>(v2, 0.f);
  exr_attr(out, "screenWindowCenter", "v2f", v2);
  exr_attr(out, "screenWindowWidth", "float", f1);
  out.push_back(0);
  size_t table = out.size(), row = (size_t)w * 6;
  out.resize(table + 8 * (size_t)h);
  for (int y = 0; y < h; y++) {
    uint64_t off = out.size();
    std::memcpy(&out[table + 8 * (size_t)y], &off, 8);
    put_le<int32_t>(out, y), put_le<int32_t>(out, (int32_t)row);
    for (int c : {2, 1, 0})
      for (int x = 0; x < w; x++) put_le<uint16_t>(out, float_to_half(rgba[((size_t)y * w + x) * 4 + c]));
  }
  write_file(path, out);
}

}  // namespace

uint16_t float_to_half(float f) {  // round to nearest even, like Imath's half(float)
  uint32_t x;
  std::memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u, mant = x & 0x7FFFFFu;
  int exp = (int)((x >> 23) & 0xFF) - 127 + 15;
  if (((x >> 23) & 0xFF) == 0xFF) return (uint16_t)(sign | 0x7C00u | (mant ? 0x200u | (mant >> 13) : 0));
  if (exp >= 31) return (uint16_t)(sign | 0x7C00u);
  if (exp <= 0) {
    if (exp < -10) return (uint16_t)sign;
    mant |= 0x800000u;
    int shift = 14 - exp;
    uint32_t h = mant >> shift, rem = mant & ((1u << shift) - 1), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (h & 1))) h++;
    return (uint16_t)(sign | h);
  }
  uint32_t h = ((uint32_t)exp << 10) | (mant >> 13), rem = mant & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) h++;
  return (uint16_t)(sign | h);
}

float half_to_float(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1F, mant = h & 0x3FFu, x;
  if (exp == 0) {
    if (!mant) x = sign;
    else {
      int e = -1;
      do {
        e++;
        mant <<= 1;
      } while (!(mant & 0x400u));
      x = sign | ((uint32_t)(127 - 15 - e) << 23) | ((mant & 0x3FFu) << 13);
    }
  } else if (exp == 31) x = sign | 0x7F800000u | (mant << 13);
  else x = sign | ((exp - 15 + 127) << 23) | (mant << 13);
  float f;
  std::memcpy(&f, &x, 4);
  return f;
}


// ---------------------------------------------------------------------------------------------- JPEG (baseline)
// The reference reads jpg textures / env maps through stb_image (src/core/texture.cpp:307-336).  This is a baseline
// sequential decoder (SOF0 / SOF1, 8-bit, Huffman, 1 or 3 components, any sampling factors up to 4x4, restart
// intervals): float IDCT, stb-style triangle-filter chroma upsampling, JFIF YCbCr -> RGB.  Progressive files (SOF2) are rejected by name.
struct JpegDecoder {
  const uint8_t* d;
  size_t n, p = 0;
  std::string name;
  struct Huff {
    uint8_t bits[17] = {0};
    uint8_t vals[256] = {0};
    int mincode[17], maxcode[18], valptr[17];
    bool set = false;
    void build() {
      int code = 0, k = 0;
      for (int l = 1; l <= 16; l++) {
        valptr[l] = k, mincode[l] = code;
        code += bits[l], k += bits[l];
        maxcode[l] = bits[l] ? code - 1 : -1;
        code <<= 1;
      }
      maxcode[17] = 0x7FFFFFFF;
      set = true;
    }
  } dc[4], ac[4];
  struct Comp {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, pred = 0, bw = 0, bh = 0;
    std::vector<uint8_t> px;  // bw*8 x bh*8 samples
  } comp[3];
  uint16_t qt[4][64];
  bool qt_set[4] = {false, false, false, false};
  int w = 0, h = 0, nc = 0, hmax = 1, vmax = 1, restart = 0;
  uint32_t bitbuf = 0;
  int bitcnt = 0;
  bool hit_marker = false;

  [[noreturn]] void fail(const std::string& why) const { throw std::runtime_error(name + ": " + why); }
  uint8_t u8() {
    if (p >= n) fail("truncated JPEG");
    return d[p++];
  }
  int u16() {
    int a = u8();
    return (a << 8) | u8();
  }
  int bit() {
    if (bitcnt == 0) {
      uint8_t b = 0;
      if (!hit_marker && p < n) {
        b = d[p++];
        if (b == 0xFF) {
          uint8_t m = p < n ? d[p] : 0;
          if (m == 0) p++;  // stuffed zero
          else hit_marker = true, p--, b = 0;  // a marker ends the entropy-coded segment: feed zeros
        }
      }
      bitbuf = b, bitcnt = 8;
    }
    bitcnt--;
    return (bitbuf >> bitcnt) & 1;
  }
  int receive(int s) {
    int v = 0;
    for (int i = 0; i < s; i++) v = (v << 1) | bit();
    return v;
  }
  static int extend(int v, int s) { return s && v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }
  int decode(const Huff& t) {
    int code = 0;
    for (int l = 1; l <= 16; l++) {
      code = (code << 1) | bit();
      if (t.maxcode[l] >= 0 && code <= t.maxcode[l] && code >= t.mincode[l]) return t.vals[t.valptr[l] + code - t.mincode[l]];
    }
    fail("bad Huffman code in JPEG");
  }
  static void idct8x8(const float* in, uint8_t* out, int stride) {
    static float c[8][8];
    static bool init = false;
    if (!init) {
      for (int x = 0; x < 8; x++)
        for (int u = 0; u < 8; u++) c[x][u] = (u == 0 ? std::sqrt(0.125f) : 0.5f) * std::cos((2 * x + 1) * u * 3.14159265358979f / 16.0f);
      init = true;
    }
    float tmp[64];
    for (int y = 0; y < 8; y++)  // rows
      for (int x = 0; x < 8; x++) {
        float s = 0.f;
        for (int u = 0; u < 8; u++) s += c[x][u] * in[y * 8 + u];
        tmp[y * 8 + x] = s;
      }
    for (int x = 0; x < 8; x++)  // columns
      for (int y = 0; y < 8; y++) {
        float s = 0.f;
        for (int v = 0; v < 8; v++) s += c[y][v] * tmp[v * 8 + x];
        int q = (int)std::lround(s + 128.0f);
        out[y * stride + x] = (uint8_t)std::min(255, std::max(0, q));
      }
  }
  void decode_block(Comp& cp, int bx, int by) {
    static const uint8_t zz[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                   41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                   30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    float blk[64] = {0};
    int t = decode(dc[cp.td]);
    cp.pred += extend(receive(t), t);
    blk[0] = (float)(cp.pred * qt[cp.tq][0]);
    for (int k = 1; k < 64;) {
      int rs = decode(ac[cp.ta]), r = rs >> 4, sz = rs & 15;
      if (sz == 0) {
        if (r != 15) break;  // EOB
        k += 16;
        continue;
      }
      k += r;
      if (k > 63) fail("bad AC run in JPEG");
      blk[zz[k]] = (float)(extend(receive(sz), sz) * qt[cp.tq][k]);
      k++;
    }
    idct8x8(blk, &cp.px[((size_t)by * 8) * ((size_t)cp.bw * 8) + (size_t)bx * 8], cp.bw * 8);
  }
  void scan() {
    const int mcux = (w + 8 * hmax - 1) / (8 * hmax), mcuy = (h + 8 * vmax - 1) / (8 * vmax);
    for (int c = 0; c < nc; c++) {
      comp[c].bw = mcux * comp[c].h, comp[c].bh = mcuy * comp[c].v;
      comp[c].px.assign((size_t)comp[c].bw * 8 * comp[c].bh * 8, 0);
      comp[c].pred = 0;
      if (!qt_set[comp[c].tq] || !dc[comp[c].td].set || !ac[comp[c].ta].set) fail("JPEG scan refers to a missing table");
    }
    bitcnt = 0, hit_marker = false;
    int count = 0;
    for (int my = 0; my < mcuy; my++)
      for (int mx = 0; mx < mcux; mx++) {
        if (restart && count && count % restart == 0) {  // RSTn: byte-align, skip the marker, reset predictors
          bitcnt = 0, hit_marker = false;
          while (p + 1 < n && !(d[p] == 0xFF && d[p + 1] >= 0xD0 && d[p + 1] <= 0xD7)) p++;
          p += 2;
          for (int c = 0; c < nc; c++) comp[c].pred = 0;
        }
        for (int c = 0; c < nc; c++)
          for (int v = 0; v < comp[c].v; v++)
            for (int hh = 0; hh < comp[c].h; hh++) decode_block(comp[c], mx * comp[c].h + hh, my * comp[c].v + v);
        count++;
      }
  }
  // returns w*h*nc interleaved 8-bit samples (nc = 1 grey, 3 RGB)
  std::vector<uint8_t> run() {
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) fail("not a JPEG file");
    p = 2;
    bool have_frame = false, done = false;
    while (!done) {
      while (p < n && d[p] != 0xFF) p++;
      while (p < n && d[p] == 0xFF) p++;
      if (p >= n) fail("JPEG without a scan");
      int m = d[p++];
      if (m == 0xD9) break;
      if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
      int len = u16();
      if (len < 2 || p + (size_t)len - 2 > n) fail("truncated JPEG segment");
      size_t end = p + (size_t)len - 2;
      if (m == 0xC0 || m == 0xC1) {
        if (u8() != 8) fail("only 8-bit JPEG is supported");
        h = u16(), w = u16(), nc = u8();
        if (w <= 0 || h <= 0 || (nc != 1 && nc != 3) || (uint64_t)w * h > (1ull << 28)) fail("unsupported JPEG frame");
        for (int c = 0; c < nc; c++) {
          comp[c].id = u8();
          int hv = u8();
          comp[c].h = hv >> 4, comp[c].v = hv & 15, comp[c].tq = u8() & 3;
          if (comp[c].h < 1 || comp[c].h > 4 || comp[c].v < 1 || comp[c].v > 4) fail("bad JPEG sampling factors");
          hmax = std::max(hmax, comp[c].h), vmax = std::max(vmax, comp[c].v);
        }
        have_frame = true;
      } else if (m == 0xC2 || (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC)) {
        fail("progressive / lossless / arithmetic-coded JPEG is not supported (baseline only) -- re-save it as baseline or png");
      } else if (m == 0xC4) {
        while (p < end) {
          int tc_th = u8(), cls = tc_th >> 4, id = tc_th & 3;
          Huff& t = cls ? ac[id] : dc[id];
          int total = 0;
          for (int l = 1; l <= 16; l++) t.bits[l] = u8(), total += t.bits[l];
          if (total > 256) fail("bad JPEG Huffman table");
          for (int i = 0; i < total; i++) t.vals[i] = u8();
          t.build();
        }
      } else if (m == 0xDB) {
        while (p < end) {
          int pq_tq = u8(), id = pq_tq & 3;
          for (int i = 0; i < 64; i++) qt[id][i] = (pq_tq >> 4) ? (uint16_t)u16() : u8();
          qt_set[id] = true;
        }
      } else if (m == 0xDD) {
        restart = u16();
      } else if (m == 0xDA) {
        if (!have_frame) fail("JPEG scan before frame header");
        int ns = u8();
        if (ns != nc) fail("multi-scan baseline JPEG is not supported");
        for (int i = 0; i < ns; i++) {
          int id = u8(), tt = u8();
          for (int c = 0; c < nc; c++)
            if (comp[c].id == id) comp[c].td = tt >> 4, comp[c].ta = tt & 15;
        }
        p += 3;  // Ss, Se, Ah/Al
        scan();
        done = true;
        continue;
      }
      p = end;
    }
    if (!done) fail("JPEG without image data");
    // chroma upsampling as stb_image does it (the decoder the reference uses): 2x factors through the centred triangle
    // filter (3 near + 1 far per axis, stbi__resample_row_h_2 / _v_2 / _hv_2), anything else by replication
    std::vector<std::vector<uint8_t>> full((size_t)nc);
    for (int c = 0; c < nc; c++) {
      const Comp& cp = comp[c];
      const int fx = hmax / cp.h, fy = vmax / cp.v, stride = cp.bw * 8;
      const int cw = (w * cp.h + hmax - 1) / hmax, ch = (h * cp.v + vmax - 1) / vmax;
      full[(size_t)c].resize((size_t)w * h);
      auto at = [&](int x, int y) { return (int)cp.px[(size_t)std::min(std::max(y, 0), ch - 1) * stride + std::min(std::max(x, 0), cw - 1)]; };
      const bool exact = hmax % cp.h == 0 && vmax % cp.v == 0;
      for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
          int v;
          if (exact && fx == 2 && fy == 2) {
            const int cx = x >> 1, cy = y >> 1, nx = (x & 1) ? cx + 1 : cx - 1, ny = (y & 1) ? cy + 1 : cy - 1;
            const int t_near = 3 * at(cx, cy) + at(cx, ny), t_far = 3 * at(nx, cy) + at(nx, ny);
            v = (x == 0 || x == 2 * cw - 1) ? (t_near + 2) >> 2 : (3 * t_near + t_far + 8) >> 4;
          } else if (exact && fx == 2 && fy == 1) {
            const int cx = x >> 1, nx = (x & 1) ? cx + 1 : cx - 1;
            v = (x == 0 || x == 2 * cw - 1) ? at(cx, y) : (3 * at(cx, y) + at(nx, y) + 2) >> 2;
          } else if (exact && fx == 1 && fy == 2) {
            const int cy = y >> 1, ny = (y & 1) ? cy + 1 : cy - 1;
            v = (3 * at(x, cy) + at(x, ny) + 2) >> 2;
          } else {
            v = at(x * cp.h / hmax, y * cp.v / vmax);
          }
          full[(size_t)c][(size_t)y * w + x] = (uint8_t)v;
        }
    }
    std::vector<uint8_t> out((size_t)w * h * nc);
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) {
        float s[3] = {0, 0, 0};
        for (int c = 0; c < nc; c++) s[c] = full[(size_t)c][(size_t)y * w + x];
        uint8_t* o = &out[((size_t)y * w + x) * nc];
        if (nc == 1) {
          o[0] = (uint8_t)s[0];
        } else {  // JFIF: full-range BT.601
          float Y = s[0], cb = s[1] - 128.f, cr = s[2] - 128.f;
          float rgb[3] = {Y + 1.402f * cr, Y - 0.344136f * cb - 0.714136f * cr, Y + 1.772f * cb};
          for (int k = 0; k < 3; k++) o[k] = (uint8_t)std::min(255.f, std::max(0.f, std::round(rgb[k])));
        }
      }
    return out;
  }
};

ImageF read_image(const std::string& path, float gamma) {
  std::string ext = lower_ext(path);
  std::vector<uint8_t> f = read_file(path);
  if (ext == "hdr") return decode_hdr(f, path);
  if (ext == "exr") return decode_exr(f, path);
  if (ext == "pfm") return decode_pfm(f, path);
  if (ext == "npy") return decode_npy(f, path);
  if (ext == "png") {
    Png p = decode_png(f, path);
    ImageF img;
    img.w = p.w, img.h = p.h, img.px.resize((size_t)p.w * p.h * 4);
    const float inv = 1.0f / 255.0f;
    for (size_t i = 0; i < (size_t)p.w * p.h; i++) {
      // stb converts 16-bit and low-bit-depth files to 8 bits per channel first, then to float (stbi__ldr_to_hdr)
      uint8_t c8[4] = {0, 0, 0, 255};
      for (int c = 0; c < p.channels; c++) {
        uint16_t s = p.samples[i * p.channels + c];
        c8[c] = p.depth == 16 ? (uint8_t)(s >> 8) : (p.depth == 8 || p.maxval == 255 ? (uint8_t)s : (uint8_t)(s * 255 / p.maxval));
      }
      uint8_t r, g, b, a = 255;
      if (p.channels == 1) r = g = b = c8[0];
      else if (p.channels == 2) r = g = b = c8[0], a = c8[1];
      else r = c8[0], g = c8[1], b = c8[2], a = p.channels == 4 ? c8[3] : 255;
      float* o = &img.px[4 * i];
      o[0] = std::pow(r * inv, gamma), o[1] = std::pow(g * inv, gamma), o[2] = std::pow(b * inv, gamma);
      o[3] = a * inv;
    }
    return img;
  }
  if (ext == "jpg" || ext == "jpeg") {
    std::vector<uint8_t> f = read_file(path);
    JpegDecoder j;
    j.d = f.data(), j.n = f.size(), j.name = path;
    std::vector<uint8_t> px = j.run();
    ImageF img;
    img.w = j.w, img.h = j.h, img.px.resize((size_t)j.w * j.h * 4);
    const float inv = 1.0f / 255.0f;
    for (size_t i = 0; i < (size_t)j.w * j.h; i++) {
      uint8_t r = px[i * j.nc], g = j.nc == 3 ? px[i * 3 + 1] : r, b = j.nc == 3 ? px[i * 3 + 2] : r;
      float* o = &img.px[4 * i];
      o[0] = std::pow(r * inv, gamma), o[1] = std::pow(g * inv, gamma), o[2] = std::pow(b * inv, gamma), o[3] = 1.0f;
    }
    return img;
  }
  throw std::runtime_error("textures only support extensions (hdr exr png jpg pfm npy) while [" + ext + "] is passed in (" + path + ")");
}

void write_image(const std::string& path, int w, int h, const float* rgba) {
  std::string ext = lower_ext(path);
  if (ext == "exr") return encode_exr_half_rgb(path, w, h, rgba);
  if (ext == "hdr") return encode_hdr(path, w, h, rgba);
  if (ext == "pfm") return write_pfm(path, w, h, rgba);
  if (ext == "npy") return write_npy_f32(path, {(size_t)h, (size_t)w, 4}, rgba);
  if (ext == "png") {
    std::vector<uint8_t> ldr((size_t)w * h * 4);
    for (size_t i = 0; i < ldr.size(); i++) {  // stbi__hdr_to_ldr, gamma 1, scale 1
      float z = rgba[i] * 255.0f + 0.5f;
      if (!(z >= 0.f)) z = 0.f;
      if (z > 255.f) z = 255.f;
      ldr[i] = (uint8_t)(int)z;
    }
    return encode_png_rgba8(path, w, h, ldr.data());
  }
  throw std::runtime_error("output images support extensions (exr png hdr pfm npy) while [" + ext + "] is passed in");
}

void write_npy_f32(const std::string& path, const std::vector<size_t>& shape, const float* data) {
  std::string dict = "{'descr': '<f4', 'fortran_order': False, 'shape': (";
  size_t n = 1;
  for (size_t i = 0; i < shape.size(); i++) {
    dict += std::to_string(shape[i]) + (shape.size() == 1 || i + 1 < shape.size() ? ", " : "");
    n *= shape[i];
  }
  dict += "), }";
  while ((10 + dict.size() + 1) % 64) dict += ' ';
  dict += '\n';
  std::vector<uint8_t> out = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, (uint8_t)(dict.size() & 0xFF), (uint8_t)(dict.size() >> 8)};
  out.insert(out.end(), dict.begin(), dict.end());
  size_t off = out.size();
  out.resize(off + n * 4);
  if (n) std::memcpy(&out[off], data, n * 4);
  write_file(path, out);
}

void write_pfm(const std::string& path, int w, int h, const float* rgba) {
  std::string head = "PF\n" + std::to_string(w) + " " + std::to_string(h) + "\n-1.0\n";
  std::vector<uint8_t> out(head.begin(), head.end());
  for (int y = h - 1; y >= 0; y--)
    for (int x = 0; x < w; x++)
      for (int c = 0; c < 3; c++) put_le<float>(out, rgba[((size_t)y * w + x) * 4 + c]);
  write_file(path, out);
}

}  // namespace asuna_host
