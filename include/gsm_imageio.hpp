// gsm_imageio.hpp -- minimal host-side image file I/O for the GUI-free front-end (gsm_caller): PNG (8-bit, gray / RGB /
// gray+alpha / RGBA / palette, non-interlaced; zlib does the inflate) and binary PNM (P5 / P6) in, PGM / gray PNG /
// raw out.  Replaces cv::imread / cv::imshow of BlockMatching/Caller.cpp:12-13,23-24 -- image loading stays on the host
// (north_star) and no OpenCV is needed.  Header-only; link with -lz.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace gsm_io {

struct Image {
  int rows = 0, cols = 0, channels = 0;  // channels in FILE order (PNG / PPM: R, G, B [, A])
  std::vector<uint8_t> data;              // interleaved, tight
};

inline bool read_file(const char* path, std::vector<uint8_t>& out, std::string& err) {
  FILE* f = std::fopen(path, "rb");
  if (!f) { err = std::string("cannot open ") + path; return false; }
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  out.resize(n > 0 ? (size_t)n : 0);
  const bool ok = n >= 0 && std::fread(out.data(), 1, out.size(), f) == out.size();
  std::fclose(f);
  if (!ok) err = std::string("cannot read ") + path;
  return ok;
}

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline bool decode_png(const std::vector<uint8_t>& buf, Image& img, std::string& err) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (buf.size() < 33 || std::memcmp(buf.data(), sig, 8) != 0) { err = "not a PNG file"; return false; }
  size_t pos = 8;
  int width = 0, height = 0, depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> idat, plte;
  bool seen_ihdr = false;
  while (pos + 12 <= buf.size()) {
    const uint32_t len = be32(&buf[pos]);
    const char* type = reinterpret_cast<const char*>(&buf[pos + 4]);
    if (pos + 12 + (size_t)len > buf.size()) { err = "truncated PNG chunk"; return false; }
    const uint8_t* d = &buf[pos + 8];
    if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
      width = (int)be32(d); height = (int)be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12];
      seen_ihdr = true;
    } else if (!std::memcmp(type, "PLTE", 4)) {
      plte.assign(d, d + len);
    } else if (!std::memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), d, d + len);
    } else if (!std::memcmp(type, "IEND", 4)) {
      break;
    }
    pos += 12 + (size_t)len;
  }
  if (!seen_ihdr || width < 1 || height < 1) { err = "PNG without a valid IHDR"; return false; }
  if (depth != 8 || interlace != 0) { err = "only 8-bit non-interlaced PNGs are supported"; return false; }
  int spp;  // samples per pixel in the file
  switch (ctype) {
    case 0: spp = 1; break;
    case 2: spp = 3; break;
    case 3: spp = 1; break;
    case 4: spp = 2; break;
    case 6: spp = 4; break;
    default: err = "unsupported PNG colour type"; return false;
  }
  const size_t stride = (size_t)width * spp;
  std::vector<uint8_t> raw((stride + 1) * height);
  uLongf rawlen = (uLongf)raw.size();
  if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) {
    err = "PNG inflate failed";
    return false;
  }
  std::vector<uint8_t> pix(stride * height);
  for (int y = 0; y < height; ++y) {
    const uint8_t ft = raw[(stride + 1) * y];
    const uint8_t* in = &raw[(stride + 1) * y + 1];
    uint8_t* out = &pix[stride * y];
    const uint8_t* up = y ? out - stride : nullptr;
    for (size_t i = 0; i < stride; ++i) {
      const int a = i >= (size_t)spp ? out[i - spp] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)spp) ? up[i - spp] : 0;
      int v = in[i];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: {
          const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
          v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
          break;
        }
        default: err = "bad PNG filter type"; return false;
      }
      out[i] = (uint8_t)v;
    }
  }
  img.rows = height;
  img.cols = width;
  if (ctype == 3) {  // palette -> RGB
    img.channels = 3;
    img.data.resize((size_t)width * height * 3);
    for (size_t i = 0; i < (size_t)width * height; ++i) {
      const size_t k = 3 * (size_t)pix[i];
      for (int ch = 0; ch < 3; ++ch) img.data[3 * i + ch] = k + ch < plte.size() ? plte[k + ch] : 0;
    }
  } else {
    img.channels = spp;
    img.data.swap(pix);
  }
  return true;
}

inline bool decode_pnm(const std::vector<uint8_t>& buf, Image& img, std::string& err) {
  if (buf.size() < 7 || buf[0] != 'P' || (buf[1] != '5' && buf[1] != '6')) { err = "not a binary PGM/PPM file"; return false; }
  size_t pos = 2;
  int vals[3], got = 0;
  while (got < 3 && pos < buf.size()) {
    while (pos < buf.size() && (buf[pos] == ' ' || buf[pos] == '\n' || buf[pos] == '\r' || buf[pos] == '\t')) ++pos;
    if (pos < buf.size() && buf[pos] == '#') { while (pos < buf.size() && buf[pos] != '\n') ++pos; continue; }
    int v = 0, digits = 0;
    while (pos < buf.size() && buf[pos] >= '0' && buf[pos] <= '9') { v = v * 10 + (buf[pos++] - '0'); ++digits; }
    if (!digits) { err = "bad PNM header"; return false; }
    vals[got++] = v;
  }
  ++pos;  // the single whitespace byte after maxval
  const int ch = buf[1] == '5' ? 1 : 3;
  if (got < 3 || vals[2] != 255 || vals[0] < 1 || vals[1] < 1) { err = "only 8-bit PNM files are supported"; return false; }
  const size_t n = (size_t)vals[0] * vals[1] * ch;
  if (pos + n > buf.size()) { err = "truncated PNM file"; return false; }
  img.cols = vals[0];
  img.rows = vals[1];
  img.channels = ch;
  img.data.assign(buf.begin() + pos, buf.begin() + pos + n);
  return true;
}

inline bool read_image(const char* path, Image& img, std::string& err) {
  std::vector<uint8_t> buf;
  if (!read_file(path, buf, err)) return false;
  if (buf.size() >= 2 && buf[0] == 'P') return decode_pnm(buf, img, err);
  return decode_png(buf, img, err);
}

// 8-bit gray as cv::cvtColor(..., CV_BGR2GRAY) computes it (Caller.cpp:15-16), alpha dropped like cv::imread's default
// flags do.  OpenCV 3.x / 4.x fixed point: (R*9798 + G*19235 + B*3735 + 16384) >> 15 -- the version the fixtures of
// this repository were made with (cv2 4.13, tests/golden/make_fixtures.py) and pinned against them byte for byte by
// tests/test_abi.py.  (OpenCV 2.4, which the reference linked, used 14-bit coefficients 4899 / 9617 / 1868: the two
// differ by one grey level on ~0.06 % of the pixels; nothing in the reference pins either.)
inline std::vector<uint8_t> to_gray(const Image& img) {
  const size_t n = (size_t)img.rows * img.cols;
  std::vector<uint8_t> g(n);
  if (img.channels <= 2) {
    for (size_t i = 0; i < n; ++i) g[i] = img.data[i * img.channels];
  } else {
    for (size_t i = 0; i < n; ++i) {
      const uint8_t* p = &img.data[i * img.channels];
      g[i] = (uint8_t)((p[0] * 9798 + p[1] * 19235 + p[2] * 3735 + 16384) >> 15);
    }
  }
  return g;
}

inline bool ends_with(const char* s, const char* suffix) {
  const size_t a = std::strlen(s), b = std::strlen(suffix);
  return a >= b && std::strcmp(s + a - b, suffix) == 0;
}

inline bool write_pgm(const char* path, const uint8_t* g, int rows, int cols, std::string& err) {
  FILE* f = std::fopen(path, "wb");
  if (!f) { err = std::string("cannot write ") + path; return false; }
  std::fprintf(f, "P5\n%d %d\n255\n", cols, rows);
  const bool ok = std::fwrite(g, 1, (size_t)rows * cols, f) == (size_t)rows * cols;
  std::fclose(f);
  if (!ok) err = std::string("short write to ") + path;
  return ok;
}

inline bool write_png_gray(const char* path, const uint8_t* g, int rows, int cols, std::string& err) {
  std::vector<uint8_t> raw((size_t)(cols + 1) * rows);
  for (int y = 0; y < rows; ++y) {
    raw[(size_t)(cols + 1) * y] = 0;
    std::memcpy(&raw[(size_t)(cols + 1) * y + 1], g + (size_t)cols * y, cols);
  }
  uLongf clen = compressBound((uLong)raw.size());
  std::vector<uint8_t> comp(clen);
  if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) { err = "deflate failed"; return false; }
  FILE* f = std::fopen(path, "wb");
  if (!f) { err = std::string("cannot write ") + path; return false; }
  auto chunk = [&](const char* type, const uint8_t* d, uint32_t len) {
    uint8_t hdr[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len,
                      (uint8_t)type[0], (uint8_t)type[1], (uint8_t)type[2], (uint8_t)type[3]};
    std::fwrite(hdr, 1, 8, f);
    if (len) std::fwrite(d, 1, len, f);
    uLong c = crc32(0L, hdr + 4, 4);
    if (len) c = crc32(c, d, len);
    const uint8_t cb[4] = {(uint8_t)(c >> 24), (uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c};
    std::fwrite(cb, 1, 4, f);
  };
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  std::fwrite(sig, 1, 8, f);
  const uint8_t ihdr[13] = {(uint8_t)(cols >> 24), (uint8_t)(cols >> 16), (uint8_t)(cols >> 8), (uint8_t)cols,
                            (uint8_t)(rows >> 24), (uint8_t)(rows >> 16), (uint8_t)(rows >> 8), (uint8_t)rows, 8, 0, 0, 0, 0};
  chunk("IHDR", ihdr, 13);
  chunk("IDAT", comp.data(), (uint32_t)clen);
  chunk("IEND", nullptr, 0);
  const bool ok = std::fclose(f) == 0;
  if (!ok) err = std::string("short write to ") + path;
  return ok;
}

// by extension: .png -> gray PNG, .pgm -> P5, anything else -> raw bytes
inline bool write_gray(const char* path, const uint8_t* g, int rows, int cols, std::string& err) {
  if (ends_with(path, ".png")) return write_png_gray(path, g, rows, cols, err);
  if (ends_with(path, ".pgm")) return write_pgm(path, g, rows, cols, err);
  FILE* f = std::fopen(path, "wb");
  if (!f) { err = std::string("cannot write ") + path; return false; }
  const bool ok = std::fwrite(g, 1, (size_t)rows * cols, f) == (size_t)rows * cols;
  std::fclose(f);
  return ok;
}

}  // namespace gsm_io
