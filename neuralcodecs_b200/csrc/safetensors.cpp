#include "safetensors.h"

#include <cmath>
#include <cstdio>
#include <cstring>

#include "runtime.h"

namespace nc {
namespace {

struct Parser {
  const char* s;
  size_t n, i = 0;
  void ws() {
    while (i < n && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i;
  }
  bool eat(char c) {
    ws();
    if (i < n && s[i] == c) { ++i; return true; }
    return false;
  }
  void expect(char c) {
    if (!eat(c)) throw Error(NC_BAD_WEIGHTS, std::string("safetensors header: expected '") + c + "'");
  }
  std::string str() {
    ws();
    if (i >= n || s[i] != '"') throw Error(NC_BAD_WEIGHTS, "safetensors header: expected string");
    ++i;
    std::string out;
    while (i < n && s[i] != '"') {
      if (s[i] == '\\' && i + 1 < n) {
        ++i;
        switch (s[i]) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'u': i += 4; out += '?'; break;
          default: out += s[i];
        }
        ++i;
      } else {
        out += s[i++];
      }
    }
    if (i >= n) throw Error(NC_BAD_WEIGHTS, "safetensors header: unterminated string");
    ++i;
    return out;
  }
  int64_t integer() {
    ws();
    bool neg = false;
    if (i < n && s[i] == '-') { neg = true; ++i; }
    if (i >= n || s[i] < '0' || s[i] > '9') throw Error(NC_BAD_WEIGHTS, "safetensors header: expected integer");
    int64_t v = 0;
    while (i < n && s[i] >= '0' && s[i] <= '9') v = v * 10 + (s[i++] - '0');
    return neg ? -v : v;
  }
  std::vector<int64_t> int_array() {
    std::vector<int64_t> v;
    expect('[');
    if (eat(']')) return v;
    do { v.push_back(integer()); } while (eat(','));
    expect(']');
    return v;
  }
  // skip any JSON value
  void skip() {
    ws();
    if (i >= n) return;
    if (s[i] == '"') { str(); return; }
    if (s[i] == '{') {
      ++i;
      if (eat('}')) return;
      do { str(); expect(':'); skip(); } while (eat(','));
      expect('}');
      return;
    }
    if (s[i] == '[') {
      ++i;
      if (eat(']')) return;
      do { skip(); } while (eat(','));
      expect(']');
      return;
    }
    while (i < n && s[i] != ',' && s[i] != '}' && s[i] != ']') ++i;
  }
};

float half_to_float(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000) << 16;
  uint32_t exp = (h >> 10) & 0x1F, man = h & 0x3FF, u;
  if (exp == 0) {
    if (man == 0) {
      u = sign;
    } else {
      exp = 127 - 15 + 1;
      while (!(man & 0x400)) { man <<= 1; --exp; }
      man &= 0x3FF;
      u = sign | (exp << 23) | (man << 13);
    }
  } else if (exp == 31) {
    u = sign | 0x7F800000u | (man << 13);
  } else {
    u = sign | ((exp + 127 - 15) << 23) | (man << 13);
  }
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

}  // namespace

void load_safetensors(const std::string& path, TensorMap* out) {
  FILE* fp = std::fopen(path.c_str(), "rb");
  if (!fp) throw Error(NC_FILE_NOT_FOUND, "weights not found at " + path);
  struct Closer {
    FILE* f;
    ~Closer() { std::fclose(f); }
  } closer{fp};
  uint64_t hlen = 0;
  if (std::fread(&hlen, 8, 1, fp) != 1) throw Error(NC_BAD_WEIGHTS, path + ": truncated safetensors file");
  std::fseek(fp, 0, SEEK_END);
  const long fsize = std::ftell(fp);
  if (hlen == 0 || (long)(hlen + 8) > fsize || hlen > (1ull << 30))
    throw Error(NC_BAD_WEIGHTS, path + ": not a safetensors file (bad header length)");
  std::string header(hlen, '\0');
  std::fseek(fp, 8, SEEK_SET);
  if (std::fread(&header[0], 1, hlen, fp) != hlen) throw Error(NC_BAD_WEIGHTS, path + ": truncated header");
  const size_t data_base = 8 + hlen;

  Parser p{header.data(), header.size()};
  p.expect('{');
  if (p.eat('}')) return;
  std::vector<unsigned char> raw;
  do {
    const std::string name = p.str();
    p.expect(':');
    if (name == "__metadata__") {
      p.skip();
      continue;
    }
    std::string dtype;
    std::vector<int64_t> shape, offs;
    p.expect('{');
    do {
      const std::string key = p.str();
      p.expect(':');
      if (key == "dtype") dtype = p.str();
      else if (key == "shape") shape = p.int_array();
      else if (key == "data_offsets") offs = p.int_array();
      else p.skip();
    } while (p.eat(','));
    p.expect('}');
    if (offs.size() != 2 || offs[1] < offs[0] || (long)(data_base + offs[1]) > fsize)
      throw Error(NC_BAD_WEIGHTS, path + ": bad data_offsets for " + name);
    HostTensor t;
    t.shape = shape;
    const size_t numel = t.numel(), nbytes = (size_t)(offs[1] - offs[0]);
    raw.resize(nbytes);
    std::fseek(fp, (long)(data_base + offs[0]), SEEK_SET);
    if (nbytes && std::fread(raw.data(), 1, nbytes, fp) != nbytes) throw Error(NC_BAD_WEIGHTS, path + ": truncated data for " + name);
    auto need = [&](size_t esz) {
      if (numel * esz != nbytes) throw Error(NC_BAD_WEIGHTS, path + ": size mismatch for " + name);
    };
    if (dtype == "F32") {
      need(4);
      t.f32.resize(numel);
      std::memcpy(t.f32.data(), raw.data(), nbytes);
    } else if (dtype == "F64") {
      need(8);
      t.f32.resize(numel);
      for (size_t i = 0; i < numel; ++i) { double d; std::memcpy(&d, raw.data() + 8 * i, 8); t.f32[i] = (float)d; }
    } else if (dtype == "F16") {
      need(2);
      t.f32.resize(numel);
      for (size_t i = 0; i < numel; ++i) { uint16_t h; std::memcpy(&h, raw.data() + 2 * i, 2); t.f32[i] = half_to_float(h); }
    } else if (dtype == "BF16") {
      need(2);
      t.f32.resize(numel);
      for (size_t i = 0; i < numel; ++i) {
        uint16_t h; std::memcpy(&h, raw.data() + 2 * i, 2);
        uint32_t u = (uint32_t)h << 16; std::memcpy(&t.f32[i], &u, 4);
      }
    } else if (dtype == "I64") {
      need(8);
      t.is_int = true;
      t.i64.resize(numel);
      std::memcpy(t.i64.data(), raw.data(), nbytes);
    } else if (dtype == "I32") {
      need(4);
      t.is_int = true;
      t.i64.resize(numel);
      for (size_t i = 0; i < numel; ++i) { int32_t v; std::memcpy(&v, raw.data() + 4 * i, 4); t.i64[i] = v; }
    } else if (dtype == "BOOL" || dtype == "U8" || dtype == "I8") {
      need(1);
      t.is_int = true;
      t.i64.resize(numel);
      for (size_t i = 0; i < numel; ++i) t.i64[i] = raw[i];
    } else {
      throw Error(NC_BAD_WEIGHTS, path + ": unsupported dtype " + dtype + " for " + name);
    }
    (*out)[name] = std::move(t);
  } while (p.eat(','));
  p.expect('}');
}

}  // namespace nc
