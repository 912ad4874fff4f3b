// PyTorch zip checkpoint (`torch.save`) reader: see pth_reader.cpp.  Replaces DACUnpickler.LoadFromStream /
// LoadWithConfig (Config/DAC/DACUnpickler.cs:341-424) for the official DAC `.pth` weights.
#pragma once
#include <string>

#include "safetensors.h"

namespace nc {

// true when the file starts with the zip local-header magic PK\3\4 (DACUnpickler.cs:346-357)
bool is_torch_zip(const std::string& path);

// Loads every tensor of the checkpoint's state dict ({"state_dict": ..., "metadata": ...} or a bare state dict) as fp32 /
// int64 host tensors; metadata_json (nullable) receives the "metadata" entry as JSON ("{}" when absent).
// Throws nc::Error (NC_FILE_NOT_FOUND / NC_BAD_WEIGHTS).
void load_torch_zip(const std::string& path, TensorMap* out, std::string* metadata_json);

}  // namespace nc
