// Package loading: minimal JSON, .npy and zip (zlib inflate) readers.
#include "znicz_native.h"

#include <zlib.h>

#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <sys/stat.h>
#include <dirent.h>

namespace znicz {

// ------------------------------------------------------------------------------- JSON
namespace {
struct P {
  const std::string& s; size_t i = 0;
  explicit P(const std::string& t) : s(t) {}
  void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i; }
  [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("json: ") + m + " at " + std::to_string(i)); }
  Json value() {
    ws();
    if (i >= s.size()) fail("unexpected end");
    char c = s[i];
    if (c == '{') return object();
    if (c == '[') return array();
    if (c == '"') { Json j; j.type = Json::String; j.str = string(); return j; }
    if (!s.compare(i, 4, "true")) { i += 4; Json j; j.type = Json::Bool; j.b = true; return j; }
    if (!s.compare(i, 5, "false")) { i += 5; Json j; j.type = Json::Bool; j.b = false; return j; }
    if (!s.compare(i, 4, "null")) { i += 4; return Json(); }
    return number();
  }
  std::string string() {
    std::string o; ++i;
    while (i < s.size() && s[i] != '"') {
      if (s[i] == '\\' && i + 1 < s.size()) {
        char e = s[++i];
        switch (e) { case 'n': o += '\n'; break; case 't': o += '\t'; break; case 'r': o += '\r'; break;
          case 'u': i += 4; o += '?'; break; default: o += e; }
        ++i;
      } else o += s[i++];
    }
    if (i >= s.size()) fail("unterminated string");
    ++i; return o;
  }
  Json number() {
    size_t st = i;
    while (i < s.size() && (isdigit((unsigned char)s[i]) || s[i] == '-' || s[i] == '+' || s[i] == '.' || s[i] == 'e' || s[i] == 'E')) ++i;
    if (st == i) fail("bad value");
    Json j; j.type = Json::Number; j.num = std::stod(s.substr(st, i - st)); return j;
  }
  Json array() {
    Json j; j.type = Json::Array; ++i; ws();
    if (s[i] == ']') { ++i; return j; }
    for (;;) { j.arr.push_back(value()); ws(); if (s[i] == ',') { ++i; continue; } if (s[i] == ']') { ++i; break; } fail("expected , or ]"); }
    return j;
  }
  Json object() {
    Json j; j.type = Json::Object; ++i; ws();
    if (s[i] == '}') { ++i; return j; }
    for (;;) {
      ws(); if (s[i] != '"') fail("expected key");
      std::string k = string(); ws(); if (s[i] != ':') fail("expected :"); ++i;
      j.obj[k] = value(); ws();
      if (s[i] == ',') { ++i; continue; } if (s[i] == '}') { ++i; break; } fail("expected , or }");
    }
    return j;
  }
};
}  // namespace

const Json& Json::at(const std::string& k) const {
  auto it = obj.find(k);
  if (type != Object || it == obj.end()) throw std::runtime_error("json: missing key " + k);
  return it->second;
}
Json Json::parse(const std::string& text) { P p(text); Json j = p.value(); return j; }

// ------------------------------------------------------------------------------- npy
static float half_to_float(uint16_t h) {
  uint32_t sign = (h >> 15) & 1, exp = (h >> 10) & 0x1f, man = h & 0x3ff, f;
  if (exp == 0) {
    if (man == 0) f = sign << 31;
    else { exp = 127 - 15 + 1; while (!(man & 0x400)) { man <<= 1; --exp; } man &= 0x3ff; f = (sign << 31) | (exp << 23) | (man << 13); }
  } else if (exp == 31) f = (sign << 31) | 0x7f800000u | (man << 13);
  else f = (sign << 31) | ((exp + 127 - 15) << 23) | (man << 13);
  float o; std::memcpy(&o, &f, 4); return o;
}

int64_t NpyArray::size() const { int64_t s = 1; for (auto d : shape) s *= d; return s; }

NpyArray parse_npy(const std::string& b) {
  if (b.size() < 10 || std::memcmp(b.data(), "\x93NUMPY", 6)) throw std::runtime_error("npy: bad magic");
  int major = (unsigned char)b[6];
  size_t hlen, off;
  if (major == 1) { hlen = (unsigned char)b[8] | ((unsigned char)b[9] << 8); off = 10; }
  else { uint32_t l; std::memcpy(&l, b.data() + 8, 4); hlen = l; off = 12; }
  std::string hdr = b.substr(off, hlen);
  auto find = [&](const std::string& key) { size_t p = hdr.find(key); if (p == std::string::npos) throw std::runtime_error("npy: no " + key); return p; };
  size_t p = find("'descr'"); p = hdr.find('\'', p + 7); size_t q = hdr.find('\'', p + 1);
  std::string descr = hdr.substr(p + 1, q - p - 1);
  if (hdr.find("'fortran_order': True") != std::string::npos) throw std::runtime_error("npy: fortran order unsupported");
  p = find("'shape'"); p = hdr.find('(', p); q = hdr.find(')', p);
  NpyArray a;
  { std::string sh = hdr.substr(p + 1, q - p - 1); std::stringstream ss(sh); std::string tok;
    while (std::getline(ss, tok, ',')) { size_t t0 = tok.find_first_not_of(' '); if (t0 == std::string::npos) continue; a.shape.push_back(std::stoll(tok.substr(t0))); } }
  int64_t n = a.size();
  const char* d = b.data() + off + hlen;
  size_t avail = b.size() - off - hlen;
  a.data.resize(n);
  char kind = descr.size() >= 2 ? descr[1] : '?';
  int isz = descr.size() >= 3 ? std::stoi(descr.substr(2)) : 0;
  if (descr[0] == '>') throw std::runtime_error("npy: big endian unsupported");
  if ((size_t)n * isz > avail) throw std::runtime_error("npy: truncated");
  for (int64_t i = 0; i < n; ++i) {
    if (kind == 'f' && isz == 2) { uint16_t h; std::memcpy(&h, d + i * 2, 2); a.data[i] = half_to_float(h); }
    else if (kind == 'f' && isz == 4) { float v; std::memcpy(&v, d + i * 4, 4); a.data[i] = v; }
    else if (kind == 'f' && isz == 8) { double v; std::memcpy(&v, d + i * 8, 8); a.data[i] = (float)v; }
    else if (kind == 'i' && isz == 4) { int32_t v; std::memcpy(&v, d + i * 4, 4); a.data[i] = (float)v; }
    else if (kind == 'i' && isz == 8) { int64_t v; std::memcpy(&v, d + i * 8, 8); a.data[i] = (float)v; }
    else throw std::runtime_error("npy: unsupported dtype " + descr);
  }
  return a;
}

// ------------------------------------------------------------------------------- zip / dir
static std::string slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  std::stringstream ss; ss << f.rdbuf(); return ss.str();
}
static uint32_t rd32(const std::string& s, size_t o) { uint32_t v; std::memcpy(&v, s.data() + o, 4); return v; }
static uint16_t rd16(const std::string& s, size_t o) { uint16_t v; std::memcpy(&v, s.data() + o, 2); return v; }

std::map<std::string, std::string> read_package_files(const std::string& path) {
  std::map<std::string, std::string> out;
  struct stat st;
  if (stat(path.c_str(), &st) != 0) throw std::runtime_error("no such package: " + path);
  if (S_ISDIR(st.st_mode)) {
    DIR* d = opendir(path.c_str());
    if (!d) throw std::runtime_error("cannot list " + path);
    while (dirent* e = readdir(d)) { std::string n = e->d_name; if (n == "." || n == "..") continue; out[n] = slurp(path + "/" + n); }
    closedir(d);
    return out;
  }
  std::string z = slurp(path);
  // end of central directory
  if (z.size() < 22) throw std::runtime_error("zip: too small");
  size_t eocd = std::string::npos;
  for (size_t i = z.size() - 22;; --i) { if (rd32(z, i) == 0x06054b50u) { eocd = i; break; } if (i == 0) break; }
  if (eocd == std::string::npos) throw std::runtime_error("zip: no end-of-central-directory (only .zip or directories are supported)");
  int count = rd16(z, eocd + 10); size_t cd = rd32(z, eocd + 16);
  for (int k = 0; k < count; ++k) {
    if (rd32(z, cd) != 0x02014b50u) throw std::runtime_error("zip: bad central header");
    int method = rd16(z, cd + 10); uint32_t csize = rd32(z, cd + 20), usize = rd32(z, cd + 24);
    int nlen = rd16(z, cd + 28), elen = rd16(z, cd + 30), clen = rd16(z, cd + 32);
    size_t lho = rd32(z, cd + 42);
    std::string name = z.substr(cd + 46, nlen);
    int lnlen = rd16(z, lho + 26), lelen = rd16(z, lho + 28);
    size_t data = lho + 30 + lnlen + lelen;
    if (method == 0) out[name] = z.substr(data, usize);
    else if (method == 8) {
      std::string o(usize, '\0');
      z_stream zs{}; zs.next_in = (Bytef*)(z.data() + data); zs.avail_in = csize; zs.next_out = (Bytef*)&o[0]; zs.avail_out = usize;
      if (inflateInit2(&zs, -MAX_WBITS) != Z_OK) throw std::runtime_error("zip: inflateInit failed");
      int r = inflate(&zs, Z_FINISH); inflateEnd(&zs);
      if (r != Z_STREAM_END) throw std::runtime_error("zip: inflate failed for " + name);
      out[name] = std::move(o);
    } else throw std::runtime_error("zip: unsupported compression method");
    cd += 46 + nlen + elen + clen;
  }
  return out;
}

}  // namespace znicz
