// Reader for PyTorch zip checkpoints (`torch.save`, the format of the official DAC `.pth` weights): a stored
// (uncompressed) zip archive holding `<root>/data.pkl` (pickle protocol 2) and one raw little-endian file per tensor
// storage `<root>/data/<key>`.  Replaces the reference's DACUnpickler.LoadFromStream
// (Config/DAC/DACUnpickler.cs:341-381: zip magic check, `data.pkl` lookup, unpickle, `state_dict` + `metadata`)
// and the Razorvine pickle machinery behind it.  Only the opcodes `torch.save` emits are implemented; anything else
// is rejected with NC_BAD_WEIGHTS.  Host code only.
#include "pth_reader.h"

#include <cstdio>
#include <cstring>
#include <memory>

#include "runtime.h"

namespace nc {
namespace {

[[noreturn]] void bad(const std::string& m) { throw Error(NC_BAD_WEIGHTS, "Failed to load weights: " + m); }

// ------------------------------------------------------------------------------------------------ zip (stored entries)
struct ZipEntry {
  std::string name;
  uint64_t offset = 0, size = 0;   // offset of the DATA inside the file
};

uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

std::vector<ZipEntry> zip_directory(const std::vector<uint8_t>& f) {
  const size_t n = f.size();
  if (n < 22 || rd32(f.data()) != 0x04034b50u)
    bad("Invalid .pth file format - must be a zip archive. File may be corrupted or saved in legacy format.");   // DACUnpickler.cs:352-357
  // end of central directory: scan backwards for PK\5\6
  size_t eocd = (size_t)-1;
  for (size_t i = n - 22;; --i) {
    if (rd32(&f[i]) == 0x06054b50u) { eocd = i; break; }
    if (i == 0 || n - i > 22 + 65535) break;
  }
  if (eocd == (size_t)-1) bad("zip end-of-central-directory record not found");
  uint64_t count = rd16(&f[eocd + 10]), cd_size = rd32(&f[eocd + 12]), cd_off = rd32(&f[eocd + 16]);
  if (count == 0xFFFF || cd_size == 0xFFFFFFFFu || cd_off == 0xFFFFFFFFu) {
    // zip64: locator PK\6\7 sits 20 bytes before the EOCD and points at the zip64 EOCD record PK\6\6
    if (eocd < 20 || rd32(&f[eocd - 20]) != 0x07064b50u) bad("zip64 locator missing");
    const uint64_t z = rd64(&f[eocd - 20 + 8]);
    if (z > n || n - z < 56 || rd32(&f[z]) != 0x06064b50u) bad("zip64 end-of-central-directory record missing");
    count = rd64(&f[z + 32]); cd_size = rd64(&f[z + 40]); cd_off = rd64(&f[z + 48]);
  }
  if (cd_off > n || cd_size > n - cd_off) bad("zip central directory out of range");   // written so crafted 64-bit values cannot wrap
  std::vector<ZipEntry> out;
  size_t p = (size_t)cd_off;
  for (uint64_t e = 0; e < count; ++e) {
    if (p + 46 > n || rd32(&f[p]) != 0x02014b50u) bad("zip central directory entry corrupted");
    const uint16_t method = rd16(&f[p + 10]);
    uint64_t csize = rd32(&f[p + 20]), usize = rd32(&f[p + 24]), lho = rd32(&f[p + 42]);
    const uint16_t nlen = rd16(&f[p + 28]), xlen = rd16(&f[p + 30]), clen = rd16(&f[p + 32]);
    if (p + 46 + nlen + xlen + clen > n) bad("zip central directory entry out of range");
    ZipEntry z;
    z.name.assign((const char*)&f[p + 46], nlen);
    // zip64 extra field (id 1): the 0xFFFFFFFF fields in order usize, csize, local header offset
    size_t x = p + 46 + nlen, xend = x + xlen;
    while (x + 4 <= xend) {
      const uint16_t id = rd16(&f[x]), len = rd16(&f[x + 2]);
      if (id == 1) {
        size_t q = x + 4;
        if (usize == 0xFFFFFFFFu && q + 8 <= xend) { usize = rd64(&f[q]); q += 8; }
        if (csize == 0xFFFFFFFFu && q + 8 <= xend) { csize = rd64(&f[q]); q += 8; }
        if (lho == 0xFFFFFFFFu && q + 8 <= xend) { lho = rd64(&f[q]); q += 8; }
      }
      x += 4 + (size_t)len;
    }
    if (method != 0 || csize != usize) bad("zip entry '" + z.name + "' is compressed; torch.save writes stored entries");
    if (lho > n || n - lho < 30 || rd32(&f[lho]) != 0x04034b50u) bad("zip local header corrupted");
    z.offset = lho + 30 + rd16(&f[lho + 26]) + rd16(&f[lho + 28]);
    z.size = usize;
    if (z.offset > n || z.size > n - z.offset) bad("zip entry '" + z.name + "' out of range");
    out.push_back(std::move(z));
    p += 46 + (size_t)nlen + xlen + clen;
  }
  return out;
}

// ------------------------------------------------------------------------------------------------ pickle values
struct PVal;
using P = std::shared_ptr<PVal>;
struct PVal {
  enum Kind { NONE, BOOL, INT, FLOAT, STR, TUPLE, LIST, DICT, GLOBAL, STORAGE, TENSOR, OBJECT, MARK } kind = NONE;
  int64_t i = 0;
  double f = 0;
  std::string s;                         // STR / GLOBAL ("module name") / STORAGE key
  std::vector<P> items;                  // TUPLE / LIST
  std::vector<std::pair<P, P>> dict;     // DICT (insertion order)
  // STORAGE
  std::string dtype;
  // TENSOR
  P storage;
  int64_t offset = 0;
  std::vector<int64_t> shape, stride;
};
P mk(PVal::Kind k) { auto p = std::make_shared<PVal>(); p->kind = k; return p; }

std::string storage_dtype(const std::string& global) {
  static const std::pair<const char*, const char*> map[] = {
      {"FloatStorage", "F32"}, {"HalfStorage", "F16"}, {"BFloat16Storage", "BF16"}, {"DoubleStorage", "F64"},
      {"LongStorage", "I64"}, {"IntStorage", "I32"}, {"ShortStorage", "I16"}, {"CharStorage", "I8"},
      {"ByteStorage", "U8"}, {"BoolStorage", "BOOL"}};
  for (auto& m : map)
    if (global.size() >= std::strlen(m.first) && global.compare(global.size() - std::strlen(m.first), std::string::npos, m.first) == 0)
      return m.second;
  bad("unsupported storage type '" + global + "'");
}

class Unpickler {
 public:
  Unpickler(const uint8_t* d, size_t n) : d_(d), n_(n) {}
  P load() {
    for (;;) {
      const uint8_t op = u8();
      switch (op) {
        case 0x80: u8(); break;                                               // PROTO
        case '}': push(mk(PVal::DICT)); break;                                // EMPTY_DICT
        case ']': push(mk(PVal::LIST)); break;                                // EMPTY_LIST
        case ')': push(mk(PVal::TUPLE)); break;                               // EMPTY_TUPLE
        case '(': push(mk(PVal::MARK)); break;                                // MARK
        case 'N': push(mk(PVal::NONE)); break;
        case 0x88: { auto v = mk(PVal::BOOL); v->i = 1; push(v); break; }     // NEWTRUE
        case 0x89: { auto v = mk(PVal::BOOL); v->i = 0; push(v); break; }     // NEWFALSE
        case 'J': { auto v = mk(PVal::INT); v->i = (int32_t)u32(); push(v); break; }   // BININT
        case 'K': { auto v = mk(PVal::INT); v->i = u8(); push(v); break; }             // BININT1
        case 'M': { auto v = mk(PVal::INT); v->i = u16(); push(v); break; }            // BININT2
        case 0x8a: {                                                                   // LONG1
          const int len = u8();
          if (len > 8) bad("pickle: integer too long");
          uint64_t x = 0;
          for (int k = 0; k < len; ++k) x |= (uint64_t)u8() << (8 * k);
          if (len > 0 && len < 8 && (x >> (8 * len - 1)) & 1) x |= ~0ull << (8 * len);   // sign-extend
          auto v = mk(PVal::INT); v->i = (int64_t)x; push(v); break;
        }
        case 'G': {                                                                    // BINFLOAT (big-endian double)
          uint64_t x = 0;
          for (int k = 0; k < 8; ++k) x = (x << 8) | u8();
          auto v = mk(PVal::FLOAT); std::memcpy(&v->f, &x, 8); push(v); break;
        }
        case 'X': { const uint32_t len = u32(); push(str(len)); break; }               // BINUNICODE
        case 0x8c: { const uint32_t len = u8(); push(str(len)); break; }               // SHORT_BINUNICODE (protocol 4)
        case 'U': { const uint32_t len = u8(); push(str(len)); break; }                // SHORT_BINSTRING
        case 'T': { const uint32_t len = u32(); push(str(len)); break; }               // BINSTRING
        case 'c': {                                                                    // GLOBAL "module\nname\n"
          auto v = mk(PVal::GLOBAL);
          const std::string module = line();   // two statements: operand evaluation order is unspecified
          const std::string name = line();
          v->s = module + " " + name;
          push(v); break;
        }
        case 'q': memo_[u8()] = top(); break;                                          // BINPUT
        case 'r': memo_[u32()] = top(); break;                                         // LONG_BINPUT
        case 0x94: memo_[(uint32_t)memo_.size()] = top(); break;                       // MEMOIZE (protocol 4)
        case 'h': push(get(u8())); break;                                              // BINGET
        case 'j': push(get(u32())); break;                                             // LONG_BINGET
        case 't': { auto v = mk(PVal::TUPLE); v->items = pop_mark(); push(v); break; } // TUPLE
        case 0x85: { auto v = mk(PVal::TUPLE); v->items = pop_n(1); push(v); break; }
        case 0x86: { auto v = mk(PVal::TUPLE); v->items = pop_n(2); push(v); break; }
        case 0x87: { auto v = mk(PVal::TUPLE); v->items = pop_n(3); push(v); break; }
        case 'l': { auto v = mk(PVal::LIST); v->items = pop_mark(); push(v); break; }  // LIST
        case 'a': { P x = pop(); need(top(), PVal::LIST)->items.push_back(x); break; } // APPEND
        case 'e': { auto xs = pop_mark(); auto l = need(top(), PVal::LIST); for (auto& x : xs) l->items.push_back(x); break; }   // APPENDS
        case 's': { P v = pop(), k = pop(); setitem(top(), k, v); break; }             // SETITEM
        case 'u': {                                                                    // SETITEMS
          auto xs = pop_mark();
          if (xs.size() % 2) bad("pickle: odd SETITEMS");
          for (size_t k = 0; k < xs.size(); k += 2) setitem(top(), xs[k], xs[k + 1]);
          break;
        }
        case 'Q': push(persistent(pop())); break;                                      // BINPERSID
        case 'R': { P args = pop(), fn = pop(); push(reduce(fn, args)); break; }       // REDUCE
        case 0x81: { P args = pop(), cls = pop(); push(reduce(cls, args)); break; }    // NEWOBJ
        case 'b': { pop(); break; }                                                    // BUILD: state (e.g. OrderedDict._metadata) ignored
        case '.': return pop();                                                        // STOP
        default: {
          char buf[64];
          std::snprintf(buf, sizeof buf, "pickle: unsupported opcode 0x%02x at byte %zu", op, pos_ - 1);
          bad(buf);
        }
      }
    }
  }

 private:
  uint8_t u8() { if (pos_ >= n_) bad("pickle: truncated"); return d_[pos_++]; }
  uint16_t u16() { const uint16_t a = u8(); return (uint16_t)(a | (u8() << 8)); }
  uint32_t u32() { uint32_t x = 0; for (int k = 0; k < 4; ++k) x |= (uint32_t)u8() << (8 * k); return x; }
  P str(uint32_t len) {
    if (pos_ + len > n_) bad("pickle: truncated string");
    auto v = mk(PVal::STR); v->s.assign((const char*)d_ + pos_, len); pos_ += len; return v;
  }
  std::string line() {
    std::string s;
    for (;;) { const char c = (char)u8(); if (c == '\n') break; s.push_back(c); }
    return s;
  }
  void push(P v) { stack_.push_back(std::move(v)); }
  P pop() { if (stack_.empty()) bad("pickle: stack underflow"); P v = stack_.back(); stack_.pop_back(); return v; }
  P top() { if (stack_.empty()) bad("pickle: stack underflow"); return stack_.back(); }
  P get(uint32_t k) { auto it = memo_.find(k); if (it == memo_.end()) bad("pickle: memo miss"); return it->second; }
  std::vector<P> pop_n(size_t n) {
    if (stack_.size() < n) bad("pickle: stack underflow");
    std::vector<P> v(stack_.end() - n, stack_.end());
    stack_.resize(stack_.size() - n);
    return v;
  }
  std::vector<P> pop_mark() {
    size_t m = stack_.size();
    while (m > 0 && stack_[m - 1]->kind != PVal::MARK) --m;
    if (m == 0) bad("pickle: MARK not found");
    std::vector<P> v(stack_.begin() + m, stack_.end());
    stack_.resize(m - 1);
    return v;
  }
  static PVal* need(const P& v, PVal::Kind k) { if (v->kind != k) bad("pickle: unexpected object on the stack"); return v.get(); }
  static void setitem(const P& d, const P& k, const P& v) {
    if (d->kind != PVal::DICT) bad("pickle: SETITEM on a non-dict");
    d->dict.emplace_back(k, v);
  }
  // ('storage', <storage type>, key, location, numel)  (torch/serialization.py persistent_id)
  static P persistent(const P& pid) {
    if (pid->kind != PVal::TUPLE || pid->items.size() < 5 || pid->items[0]->kind != PVal::STR || pid->items[0]->s != "storage")
      bad("pickle: unknown persistent id");
    if (pid->items[1]->kind != PVal::GLOBAL && pid->items[1]->kind != PVal::STR) bad("pickle: persistent id without a storage type");
    if (pid->items[2]->kind != PVal::STR) bad("pickle: persistent id without a storage key");
    auto v = mk(PVal::STORAGE);
    v->dtype = storage_dtype(pid->items[1]->s);
    v->s = pid->items[2]->s;
    v->i = pid->items[4]->i;
    return v;
  }
  static std::vector<int64_t> ints(const P& t) {
    std::vector<int64_t> v;
    if (t->kind != PVal::TUPLE && t->kind != PVal::LIST) bad("pickle: expected a tuple of integers");
    for (auto& x : t->items) v.push_back(x->i);
    return v;
  }
  static P reduce(const P& fn, const P& args) {
    if (fn->kind != PVal::GLOBAL || args->kind != PVal::TUPLE) bad("pickle: REDUCE on a non-callable");
    const std::string& g = fn->s;
    if (g == "collections OrderedDict") return mk(PVal::DICT);
    if (g == "torch._utils _rebuild_tensor_v2" || g == "torch._utils _rebuild_tensor") {
      if (args->items.size() < 4 || args->items[0]->kind != PVal::STORAGE) bad("pickle: malformed tensor record");
      auto t = mk(PVal::TENSOR);
      t->storage = args->items[0];
      t->offset = args->items[1]->i;
      t->shape = ints(args->items[2]);
      t->stride = ints(args->items[3]);
      if (t->stride.size() != t->shape.size()) bad("pickle: tensor record with mismatched shape / stride ranks");
      for (int64_t d : t->shape)
        if (d < 0 || d > ((int64_t)1 << 40)) bad("pickle: tensor record with a negative or absurd dimension");
      return t;
    }
    if (g == "torch._utils _rebuild_parameter" || g == "torch._utils _rebuild_parameter_with_state") {
      if (args->items.empty() || args->items[0]->kind != PVal::TENSOR) bad("pickle: malformed parameter record");
      return args->items[0];
    }
    auto o = mk(PVal::OBJECT);   // anything else (e.g. numpy scalars in metadata): opaque
    o->s = g;
    o->items = args->items;
    return o;
  }

  const uint8_t* d_;
  size_t n_, pos_ = 0;
  std::vector<P> stack_;
  std::map<uint32_t, P> memo_;
};

const P* dict_get(const P& d, const std::string& key) {
  if (d->kind != PVal::DICT) return nullptr;
  for (auto& kv : d->dict)
    if (kv.first->kind == PVal::STR && kv.first->s == key) return &kv.second;
  return nullptr;
}

void json_escape(const std::string& s, std::string* out) {
  out->push_back('"');
  for (char c : s) {
    if (c == '"' || c == '\\') { out->push_back('\\'); out->push_back(c); }
    else if ((unsigned char)c < 0x20) { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", c); *out += b; }
    else out->push_back(c);
  }
  out->push_back('"');
}

void to_json(const P& v, std::string* out) {
  char buf[64];
  switch (v->kind) {
    case PVal::NONE: *out += "null"; break;
    case PVal::BOOL: *out += v->i ? "true" : "false"; break;
    case PVal::INT: std::snprintf(buf, sizeof buf, "%lld", (long long)v->i); *out += buf; break;
    case PVal::FLOAT: std::snprintf(buf, sizeof buf, "%.17g", v->f); *out += buf; break;
    case PVal::STR: json_escape(v->s, out); break;
    case PVal::TUPLE: case PVal::LIST: {
      out->push_back('[');
      for (size_t k = 0; k < v->items.size(); ++k) { if (k) *out += ", "; to_json(v->items[k], out); }
      out->push_back(']');
      break;
    }
    case PVal::DICT: {
      out->push_back('{');
      bool first = true;
      for (auto& kv : v->dict) {
        if (kv.first->kind != PVal::STR) continue;
        if (!first) *out += ", ";
        first = false;
        json_escape(kv.first->s, out); *out += ": "; to_json(kv.second, out);
      }
      out->push_back('}');
      break;
    }
    default: *out += "null"; break;   // tensors / opaque objects inside metadata
  }
}

float half_to_float(uint16_t h) {
  const uint32_t s = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 0x1F, m = h & 0x3FF;
  uint32_t u;
  if (e == 0) {
    if (m == 0) u = s;
    else { int sh = 0; uint32_t mm = m; while (!(mm & 0x400)) { mm <<= 1; ++sh; } u = s | ((uint32_t)(113 - sh) << 23) | ((mm & 0x3FF) << 13); }
  } else if (e == 31) u = s | 0x7F800000u | (m << 13);
  else u = s | ((e + 112) << 23) | (m << 13);
  float f; std::memcpy(&f, &u, 4); return f;
}

size_t dtype_size(const std::string& d) {
  if (d == "F32" || d == "I32") return 4;
  if (d == "F16" || d == "BF16" || d == "I16") return 2;
  if (d == "F64" || d == "I64") return 8;
  return 1;
}

void materialise(const PVal& t, const std::vector<uint8_t>& file, const std::map<std::string, ZipEntry>& data, const std::string& name,
                 HostTensor* out) {
  const PVal& st = *t.storage;
  auto it = data.find(st.s);
  if (it == data.end()) bad("storage '" + st.s + "' of tensor '" + name + "' is missing from the archive");
  const size_t es = dtype_size(st.dtype);
  const uint8_t* base = file.data() + it->second.offset;
  const uint64_t avail = it->second.size / es;
  out->shape = t.shape;
  const size_t n = out->numel();
  out->is_int = st.dtype[0] == 'I' || st.dtype[0] == 'U' || st.dtype == "BOOL";
  if (out->is_int) out->i64.resize(n); else out->f32.resize(n);
  // general strided gather (state dicts are contiguous in practice)
  std::vector<int64_t> idx(t.shape.size(), 0);
  for (size_t k = 0; k < n; ++k) {
    int64_t e = t.offset;
    for (size_t d = 0; d < idx.size(); ++d) e += idx[d] * t.stride[d];
    if (e < 0 || (uint64_t)e >= avail) bad("tensor '" + name + "' reads outside its storage");
    const uint8_t* p = base + (size_t)e * es;
    if (st.dtype == "F32") { float v; std::memcpy(&v, p, 4); out->f32[k] = v; }
    else if (st.dtype == "F16") { uint16_t h; std::memcpy(&h, p, 2); out->f32[k] = half_to_float(h); }
    else if (st.dtype == "BF16") { uint16_t h; std::memcpy(&h, p, 2); const uint32_t u = (uint32_t)h << 16; float v; std::memcpy(&v, &u, 4); out->f32[k] = v; }
    else if (st.dtype == "F64") { double v; std::memcpy(&v, p, 8); out->f32[k] = (float)v; }
    else if (st.dtype == "I64") { int64_t v; std::memcpy(&v, p, 8); out->i64[k] = v; }
    else if (st.dtype == "I32") { int32_t v; std::memcpy(&v, p, 4); out->i64[k] = v; }
    else if (st.dtype == "I16") { int16_t v; std::memcpy(&v, p, 2); out->i64[k] = v; }
    else if (st.dtype == "I8") out->i64[k] = (int8_t)*p;
    else out->i64[k] = *p;
    for (int d = (int)idx.size() - 1; d >= 0; --d) { if (++idx[d] < t.shape[d]) break; idx[d] = 0; }
  }
}

}  // namespace

bool is_torch_zip(const std::string& path) {
  FILE* fp = std::fopen(path.c_str(), "rb");
  if (!fp) return false;
  uint8_t m[4] = {0, 0, 0, 0};
  const size_t got = std::fread(m, 1, 4, fp);
  std::fclose(fp);
  return got == 4 && m[0] == 0x50 && m[1] == 0x4B && m[2] == 0x03 && m[3] == 0x04;
}

void load_torch_zip(const std::string& path, TensorMap* out, std::string* metadata_json) {
  FILE* fp = std::fopen(path.c_str(), "rb");
  if (!fp) throw Error(NC_FILE_NOT_FOUND, "weights not found at " + path);
  std::fseek(fp, 0, SEEK_END);
  const long sz = std::ftell(fp);
  std::fseek(fp, 0, SEEK_SET);
  std::vector<uint8_t> file((size_t)std::max<long>(sz, 0));
  const size_t got = file.empty() ? 0 : std::fread(file.data(), 1, file.size(), fp);
  std::fclose(fp);
  if (got != file.size()) bad("short read of " + path);

  const auto entries = zip_directory(file);
  const ZipEntry* pkl = nullptr;
  for (auto& e : entries)
    if (e.name.size() >= 8 && e.name.compare(e.name.size() - 8, 8, "data.pkl") == 0) { pkl = &e; break; }
  if (!pkl) bad("Model archive missing data.pkl");                                  // DACUnpickler.cs:361-362
  const std::string root = pkl->name.substr(0, pkl->name.size() - 8);               // "<archive>/"
  std::map<std::string, ZipEntry> data;
  for (auto& e : entries)
    if (e.name.compare(0, root.size() + 5, root + "data/") == 0) data[e.name.substr(root.size() + 5)] = e;

  Unpickler up(file.data() + pkl->offset, (size_t)pkl->size);
  P top = up.load();
  if (top->kind != PVal::DICT) bad("Failed to unpickle model data");                 // :365-366
  // {"state_dict": OrderedDict, "metadata": {...}} (descript-audio-codec) or a bare state dict
  P sd = top;
  if (const P* s = dict_get(top, "state_dict")) sd = *s;
  else if (const P* s2 = dict_get(top, "model")) sd = *s2;
  if (sd->kind != PVal::DICT) bad("Missing or invalid state_dict");                  // :369-370
  out->clear();
  for (auto& kv : sd->dict) {
    if (kv.first->kind != PVal::STR || kv.second->kind != PVal::TENSOR) continue;
    materialise(*kv.second, file, data, kv.first->s, &(*out)[kv.first->s]);
  }
  if (out->empty()) bad("Missing or invalid state_dict");
  if (metadata_json) {
    metadata_json->clear();
    if (const P* m = dict_get(top, "metadata")) to_json(*m, metadata_json);
    else *metadata_json = "{}";
  }
}

}  // namespace nc
