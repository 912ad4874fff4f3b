// Minimal safetensors reader: 8-byte little-endian header length, JSON header
// {"name": {"dtype": "F32", "shape": [...], "data_offsets": [begin, end]}, "__metadata__": {...}},
// then the raw tensor bytes.  Replaces TorchSharp's Safetensors.LoadStateDict / load_safetensors
// used by the reference's LoadWeights (Config/DAC/DACUnpickler.cs:328-342, Models/SNAC.cs:216-231,
// Models/Encodec.cs:367-385).  Everything is converted to fp32 (or kept as int64) on the host.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace nc {

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> f32;    // floating tensors, converted
  std::vector<int64_t> i64;  // integer tensors
  bool is_int = false;
  size_t numel() const {
    size_t n = 1;
    for (auto d : shape) n *= (size_t)d;
    return n;
  }
};

using TensorMap = std::map<std::string, HostTensor>;

// Throws nc::Error (NC_FILE_NOT_FOUND / NC_BAD_WEIGHTS).
void load_safetensors(const std::string& path, TensorMap* out);

}  // namespace nc
