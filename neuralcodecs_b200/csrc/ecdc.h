// .ecdc container header (host side): "ECDC" | version byte 0 | big-endian int32 JSON length | UTF-8 JSON metadata.
// Follows Modules/Encodec/BinaryIO.cs:11 (MAGIC), :44-100 (ReadHeaderAsync), :104-146 (ValidateMetadata),
// :152-190 (WriteHeaderAsync) and the metadata keys of EncodecCompressor.cs:98-111 / :253-275.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>

#include "runtime.h"

namespace nc {

struct EcdcMeta {
  std::string model;          // "m"
  int64_t audio_length = 0;   // "al"
  int n_codebooks = 0;        // "nc"
  bool use_lm = false;        // "lm"
  int channels = 1;           // "ch" (optional, default mono)
  int sample_rate = 0;        // "sr" (optional; 0 = absent)
  float bandwidth = 0.f;      // "bw" (optional; 0 = absent)
  bool has_bandwidth = false;
};

inline std::string ecdc_number(float v) {   // System.Text.Json writes the shortest round-trip form: 6 -> "6", 1.5 -> "1.5"
  char buf[32];
  std::snprintf(buf, sizeof buf, "%.9g", (double)v);
  return buf;
}

// Same key order and compact form as the reference's Dictionary serialisation.
inline std::string ecdc_header(const EcdcMeta& m) {
  std::string j = "{\"m\":\"" + m.model + "\",\"al\":" + std::to_string(m.audio_length) + ",\"nc\":" + std::to_string(m.n_codebooks) +
                  ",\"lm\":" + (m.use_lm ? "true" : "false") + ",\"ch\":" + std::to_string(m.channels) +
                  ",\"sr\":" + std::to_string(m.sample_rate);
  if (m.has_bandwidth) j += ",\"bw\":" + ecdc_number(m.bandwidth);
  j += "}";
  std::string h = "ECDC";
  h.push_back((char)0);
  const uint32_t n = (uint32_t)j.size();
  for (int s = 24; s >= 0; s -= 8) h.push_back((char)((n >> s) & 0xFF));
  return h + j;
}

// Flat JSON object -> key -> raw value text (strings unquoted).  Accepts whitespace and Python-style ": " separators so
// streams written by facebookresearch/encodec parse too.
inline std::map<std::string, std::string> ecdc_parse_flat_json(const std::string& s) {
  std::map<std::string, std::string> kv;
  size_t i = 0;
  auto ws = [&] { while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) ++i; };
  auto fail = [] { throw Error(NC_INVALID_ARGUMENT, "Failed to read Encodec header"); };
  auto str = [&]() {
    std::string out;
    if (i >= s.size() || s[i] != '"') fail();
    for (++i; i < s.size() && s[i] != '"'; ++i) {
      if (s[i] == '\\' && i + 1 < s.size()) ++i;
      out.push_back(s[i]);
    }
    if (i >= s.size()) fail();
    ++i;
    return out;
  };
  ws();
  if (i >= s.size() || s[i] != '{') fail();
  ++i;
  ws();
  if (i < s.size() && s[i] == '}') return kv;
  while (true) {
    ws();
    const std::string k = str();
    ws();
    if (i >= s.size() || s[i] != ':') fail();
    ++i;
    ws();
    std::string v;
    if (i < s.size() && s[i] == '"') {
      v = str();
    } else {
      while (i < s.size() && s[i] != ',' && s[i] != '}' && s[i] != ' ' && s[i] != '\n') v.push_back(s[i++]);
      if (v.empty()) fail();
    }
    kv[k] = v;
    ws();
    if (i < s.size() && s[i] == ',') { ++i; continue; }
    if (i < s.size() && s[i] == '}') break;
    fail();
  }
  return kv;
}

// Parses and validates a stream prefix; returns the payload offset (header size).
inline size_t ecdc_read_header(const uint8_t* p, size_t n, EcdcMeta* out) {
  if (n < 9) throw Error(NC_INVALID_ARGUMENT, "Stream ended too soon");                                  // BinaryIO.cs:57-62
  if (std::memcmp(p, "ECDC", 4) != 0) throw Error(NC_INVALID_ARGUMENT, "File is not in ECDC format");     // :66-70
  if (p[4] != 0) throw Error(NC_INVALID_ARGUMENT, "Version not supported: " + std::to_string((int)p[4])); // :73-76
  const uint32_t len = ((uint32_t)p[5] << 24) | ((uint32_t)p[6] << 16) | ((uint32_t)p[7] << 8) | (uint32_t)p[8];
  if (len == 0 || len > (1u << 20)) throw Error(NC_INVALID_ARGUMENT, "Invalid metadata length: " + std::to_string(len));
  if (n < 9 + (size_t)len) throw Error(NC_INVALID_ARGUMENT, "Stream ended too soon");
  const auto kv = ecdc_parse_flat_json(std::string((const char*)p + 9, len));
  for (const char* req : {"m", "al", "nc", "lm"})                                                        // ValidateMetadata :114-122
    if (!kv.count(req)) throw Error(NC_INVALID_ARGUMENT, std::string("Missing required metadata key: ") + req);
  EcdcMeta m;
  try {
    m.model = kv.at("m");
    m.audio_length = std::stoll(kv.at("al"));
    m.n_codebooks = std::stoi(kv.at("nc"));
    const std::string lm = kv.at("lm");
    m.use_lm = (lm == "true" || lm == "True");
    if (kv.count("ch")) m.channels = std::stoi(kv.at("ch"));
    if (kv.count("sr")) m.sample_rate = std::stoi(kv.at("sr"));
    if (kv.count("bw")) { m.bandwidth = std::stof(kv.at("bw")); m.has_bandwidth = true; }
  } catch (const Error&) {
    throw;
  } catch (const std::exception&) {
    throw Error(NC_INVALID_ARGUMENT, "Failed to read Encodec header");
  }
  if (m.audio_length <= 0 || m.n_codebooks <= 0) throw Error(NC_INVALID_ARGUMENT, "Invalid Encodec metadata");
  *out = m;
  return 9 + (size_t)len;
}

}  // namespace nc
