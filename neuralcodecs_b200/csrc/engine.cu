// Engine base + DAC engine.  The DAC graph follows (paths under /root/reference/NeuralCodecs.Torch/):
//   Models/DAC.cs:141-154 (Preprocess), :163-181 (Encode), :231-234 (Decode), :101-106 (FromCodes)
//   Modules/DAC/Encoder.cs:21-58, EncoderBlock.cs:20-43, ResidualUnit.cs:24-59, Decoder.cs:22-58,
//   DecoderBlock.cs:20-44, ResidualVectorQuantizer.cs:54-103,211-238
// with activations channels-last in HBM and every Snake fused into the consuming conv's prologue.
#include "engine.h"
#include "pth_reader.h"

#include <cstdlib>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

namespace nc {

// ------------------------------------------------------------------------------------ Engine
Engine::Engine(int device_index) : device_(device_index) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    throw Error(NC_CUDA_UNAVAILABLE, "CUDA requested but not available");
  }
  if (device_index < 0 || device_index >= count)
    throw Error(NC_INVALID_ARGUMENT, "device index " + std::to_string(device_index) + " out of range (" +
                                         std::to_string(count) + " devices)");
  cudaDeviceProp prop{};
  NC_CUDA(cudaGetDeviceProperties(&prop, device_index));
  if (prop.major != 10)
    throw Error(NC_CUDA_UNAVAILABLE, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                         std::to_string(prop.minor) + "; this library is built for sm_100a only");
  num_sms_ = prop.multiProcessorCount;
  NC_CUDA(cudaSetDevice(device_index));
  NC_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
}

Engine::~Engine() {
  cudaSetDevice(device_);
  if (copy_stream_) cudaStreamDestroy(copy_stream_);
  if (stream_) cudaStreamDestroy(stream_);
}

cudaStream_t Engine::copy_stream() {
  if (!copy_stream_) NC_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
  return copy_stream_;
}

void Engine::bind() const { NC_CUDA(cudaSetDevice(device_)); }

LaunchCtx Engine::ctx() {
  LaunchCtx c;
  c.stream = stream_;
  c.num_sms = num_sms_;
  c.prof = prof_.enabled ? &prof_ : nullptr;
  c.launches = &launches_;
  c.fast_sin = fast_sin_;
  c.fuse_ru = fuse_ru_;
  return c;
}

void Engine::sync() { NC_CUDA(cudaStreamSynchronize(stream_)); }

void Engine::set_option(const std::string& key, const std::string& value) {
  if (key == "profile") {
    prof_.enabled = value == "1" || value == "true" || value == "2";
    prof_.by_layer = value == "2";
  } else if (key == "fast_sin") {
    fast_sin_ = std::atoi(value.c_str());
  } else if (key == "fuse_ru") {
    fuse_ru_ = std::atoi(value.c_str());
  } else if (key == "max_workspace_mb") {
    const long long mb = std::atoll(value.c_str());
    if (mb < 64) throw Error(NC_INVALID_ARGUMENT, "max_workspace_mb must be >= 64");
    max_workspace_bytes_ = (size_t)mb << 20;
  } else {
    throw Error(NC_INVALID_ARGUMENT, "unknown option '" + key + "'");
  }
}

void Engine::load_weights(const std::string& path) {
  tensors_.clear();
  weights_metadata_ = "{}";
  if (is_torch_zip(path)) {
    load_torch_zip(path, &tensors_, &weights_metadata_);   // DACUnpickler path of the reference (DAC.cs:368-372)
  } else {
    load_safetensors(path, &tensors_);
  }
  normalize_names();
  finalize_weights();
}

const HostTensor& Engine::tensor(const std::string& name) const {
  auto it = tensors_.find(name);
  if (it == tensors_.end()) throw Error(NC_BAD_WEIGHTS, std::string("Failed to load ") + codec_name() +
                                                            " weights: missing tensor '" + name + "'");
  return it->second;
}

SnakeParams::~SnakeParams() {
  cudaFree(alpha);
  cudaFree(inv_alpha);
}
void SnakeParams::build(const std::vector<float>& a) {
  std::vector<float> inv(a.size());
  for (size_t i = 0; i < a.size(); ++i) inv[i] = a[i] == 0.f ? 0.f : 1.0f / a[i];
  cudaFree(alpha);
  cudaFree(inv_alpha);
  alpha = upload(a);
  inv_alpha = upload(inv);
}

// ------------------------------------------------------------------------------------ DAC
DacEngine::DacEngine(const nc_dac_config& c, int device_index) : Engine(device_index) {
  if (c.struct_size != sizeof(nc_dac_config)) throw Error(NC_INVALID_ARGUMENT, "nc_dac_config.struct_size mismatch");
  if (c.n_encoder_rates < 1 || c.n_encoder_rates > NC_MAX_RATES || c.n_decoder_rates < 1 ||
      c.n_decoder_rates > NC_MAX_RATES)
    throw Error(NC_INVALID_ARGUMENT, "DAC config: rate count out of range");
  cfg_.sample_rate = c.sample_rate;
  cfg_.encoder_dim = c.encoder_dim;
  cfg_.encoder_rates.assign(c.encoder_rates, c.encoder_rates + c.n_encoder_rates);
  cfg_.decoder_dim = c.decoder_dim;
  cfg_.decoder_rates.assign(c.decoder_rates, c.decoder_rates + c.n_decoder_rates);
  cfg_.n_codebooks = c.n_codebooks;
  cfg_.codebook_size = c.codebook_size;
  cfg_.codebook_dim = c.codebook_dim;
  // Models/DAC.cs:64 : LatentDim ?? encoderDim * 2^len(rates)
  cfg_.latent_dim = c.latent_dim > 0 ? c.latent_dim : c.encoder_dim * (1 << c.n_encoder_rates);
  if (cfg_.sample_rate <= 0 || cfg_.encoder_dim <= 0 || cfg_.decoder_dim <= 0 || cfg_.n_codebooks <= 0 ||
      cfg_.codebook_size <= 0 || cfg_.codebook_dim <= 0)
    throw Error(NC_INVALID_ARGUMENT, "DAC config: non-positive field");
  for (int r : cfg_.encoder_rates)
    if (r < 1) throw Error(NC_INVALID_ARGUMENT, "DAC config: bad encoder rate");
  for (int r : cfg_.decoder_rates)
    if (r < 1) throw Error(NC_INVALID_ARGUMENT, "DAC config: bad decoder rate");
  if (cfg_.decoder_dim % (1 << cfg_.decoder_rates.size()) != 0)
    throw Error(NC_INVALID_ARGUMENT, "DAC config: decoder_dim not divisible by 2^n_rates");
}

DacEngine::~DacEngine() {
  cudaSetDevice(device_);
  cudaFree(d_conv_in_w_);
  cudaFree(d_conv_in_b_);
  cudaFree(d_conv_out_w_);
  cudaFree(d_conv_out_b_);
  for (float* p : rvq_alloc_) cudaFree(p);
}

void DacEngine::set_option(const std::string& key, const std::string& value) {
  if (key == "encoder_precision") {
    enc_prec_ = parse_precision(value);
    if (ready_) throw Error(NC_INVALID_ARGUMENT, "precision options must be set before weights are loaded");
  } else if (key == "encoder_tail_fp32") {
    // experiment / tight-parity knob: the last N layer groups of the encoder run on the exact fp32 CUDA-core path
    // (1: encoder.conv2; 2: + last strided conv; 3: + last block's residual units; 4: + the block before it)
    enc_tail_fp32_ = std::atoi(value.c_str());
    if (ready_) throw Error(NC_INVALID_ARGUMENT, "precision options must be set before weights are loaded");
  } else if (key == "encoder_short_chains") {
    // accumulation-chain policy of the encoder (conv_plan.h: acc_split; DESIGN.md "accumulation chains"):
    // 0 off; 1 (default) folded partials on the frame-rate layers (last strided conv, encoder.conv2), separate lo-term
    // accumulator on the other layers whose chains exceed ~250 MMAs; 2: folded partials on every layer the plain conv
    // kernel runs; 3 / 4: also encoder block 1 / block 0 (they then run unfused)
    enc_short_chains_ = std::atoi(value.c_str());
    if (ready_) throw Error(NC_INVALID_ARGUMENT, "precision options must be set before weights are loaded");
  } else if (key == "decoder_h16") {
    // 1 (default): the decoder layers that take one fp16 product per term exchange fp16 activations and run on the
    // fp16-operand executor (conv_h16.cu); 0: fp32 activations + in-kernel operand transform (conv_umma.cu)
    dec_h16_opt_ = std::atoi(value.c_str());
    if (ready_) throw Error(NC_INVALID_ARGUMENT, "precision options must be set before weights are loaded");
  } else if (key == "decoder_boost") {
    dec_boost_ = value == "1" || value == "true";
    if (ready_) throw Error(NC_INVALID_ARGUMENT, "precision options must be set before weights are loaded");
  } else if (key == "decoder_wide_precision") {
    dec_wide_prec_ = parse_precision(value);
    if (ready_) throw Error(NC_INVALID_ARGUMENT, "precision options must be set before weights are loaded");
  } else if (key == "decoder_precision") {
    if (value == "mixed") {
      dec_prec_ = PREC_BF16X3;
      dec_wide_prec_ = PREC_F16;
    } else {
      dec_prec_ = dec_wide_prec_ = parse_precision(value);
    }
    if (ready_) throw Error(NC_INVALID_ARGUMENT, "precision options must be set before weights are loaded");
  } else {
    Engine::set_option(key, value);
  }
}

// The official `.pth` weights carry descript-audio-codec's nn.Sequential paths (the names the reference's own modules
// use, Config/DAC/StateDictNameConverter.cs:274-335 maps the HF names onto them); the engine looks tensors up by the HF
// DacModel names, so translate the other way.  Suffixes (weight_g / weight_v / bias / alpha) are kept.
static bool dac_native_to_hf(const std::string& key, int n_enc, int n_dec, std::string* out) {
  auto split = [](const std::string& s) {
    std::vector<std::string> v;
    size_t a = 0;
    for (;;) {
      const size_t b = s.find('.', a);
      v.push_back(s.substr(a, b == std::string::npos ? std::string::npos : b - a));
      if (b == std::string::npos) break;
      a = b + 1;
    }
    return v;
  };
  auto join = [](const std::vector<std::string>& v, size_t from) {
    std::string s;
    for (size_t i = from; i < v.size(); ++i) s += "." + v[i];
    return s;
  };
  auto num = [](const std::string& s, int* x) {
    if (s.empty() || s.find_first_not_of("0123456789") != std::string::npos) return false;
    *x = std::atoi(s.c_str());
    return true;
  };
  static const char* kUnit[4] = {"snake1", "conv1", "snake2", "conv2"};
  const auto t = split(key);
  int i = 0, u = 0, k = 0;
  if (t.size() >= 4 && t[0] == "encoder" && t[1] == "block" && num(t[2], &i)) {
    if (i == 0) { *out = "encoder.conv1" + join(t, 3); return true; }
    if (i == n_enc + 1) { *out = "encoder.snake1" + join(t, 3); return true; }
    if (i == n_enc + 2) { *out = "encoder.conv2" + join(t, 3); return true; }
    if (i >= 1 && i <= n_enc && t.size() >= 6 && t[3] == "block" && num(t[4], &u)) {
      const std::string blk = "encoder.block." + std::to_string(i - 1);
      if (u == 3) { *out = blk + ".snake1" + join(t, 5); return true; }
      if (u == 4) { *out = blk + ".conv1" + join(t, 5); return true; }
      if (u >= 0 && u <= 2 && t.size() >= 8 && t[5] == "block" && num(t[6], &k) && k >= 0 && k <= 3) {
        *out = blk + ".res_unit" + std::to_string(u + 1) + "." + kUnit[k] + join(t, 7);
        return true;
      }
    }
    return false;
  }
  if (t.size() >= 4 && t[0] == "decoder" && t[1] == "model" && num(t[2], &i)) {
    if (i == 0) { *out = "decoder.conv1" + join(t, 3); return true; }
    if (i == n_dec + 1) { *out = "decoder.snake1" + join(t, 3); return true; }
    if (i == n_dec + 2) { *out = "decoder.conv2" + join(t, 3); return true; }
    if (i >= 1 && i <= n_dec && t.size() >= 6 && t[3] == "block" && num(t[4], &u)) {
      const std::string blk = "decoder.block." + std::to_string(i - 1);
      if (u == 0) { *out = blk + ".snake1" + join(t, 5); return true; }
      if (u == 1) { *out = blk + ".conv_t1" + join(t, 5); return true; }
      if (u >= 2 && u <= 4 && t.size() >= 8 && t[5] == "block" && num(t[6], &k) && k >= 0 && k <= 3) {
        *out = blk + ".res_unit" + std::to_string(u - 1) + "." + kUnit[k] + join(t, 7);
        return true;
      }
    }
    return false;
  }
  return false;
}

void DacEngine::normalize_names() {
  TensorMap& tm = tensors_mut();
  if (!tm.count("decoder.model.0.weight_v") && !tm.count("encoder.block.0.weight_v") && !tm.count("decoder.model.0.weight")) return;
  TensorMap renamed;
  for (auto& kv : tm) {
    std::string hf;
    if (dac_native_to_hf(kv.first, (int)cfg_.encoder_rates.size(), (int)cfg_.decoder_rates.size(), &hf))
      renamed[hf] = std::move(kv.second);
    else
      renamed[kv.first] = std::move(kv.second);
  }
  tm.swap(renamed);
}

void DacEngine::require_ready() const {
  if (!ready_) throw Error(NC_BAD_WEIGHTS, "DAC weights have not been loaded");
}

// Fold weight normalisation once, in fp32, as the reference does on every forward
// (Modules/DAC/WNConv1d.cs:145-150, WNConvTranspose1d.cs:146-150): w = v / (||v||_(1,2) + 1e-7) * g.
// HF-layout files carry only the folded `weight`; the reference then sets weight_v = weight and
// weight_g = ||weight||_(1,2) (Config/DAC/StateDictNameConverter.cs:48-58).
std::vector<float> DacEngine::folded_conv(const std::string& prefix, int d0, int d1, int k, std::vector<float>* bias,
                                          int bias_n) {
  const bool hf = has_tensor(prefix + ".weight");
  const HostTensor& v = tensor(hf ? prefix + ".weight" : prefix + ".weight_v");
  if (v.is_int || v.shape.size() != 3 || v.shape[0] != d0 || v.shape[1] != d1 || v.shape[2] != k)
    throw Error(NC_SHAPE_MISMATCH, "Failed to load DAC weights: '" + prefix + "' has the wrong shape");
  const HostTensor* g = hf ? nullptr : &tensor(prefix + ".weight_g");
  if (g && (g->is_int || (int64_t)g->numel() != d0))
    throw Error(NC_SHAPE_MISMATCH, "Failed to load DAC weights: '" + prefix + ".weight_g' has the wrong shape");
  std::vector<float> w(v.f32.size());
  const size_t inner = (size_t)d1 * k;
  for (int i = 0; i < d0; ++i) {
    double ss = 0;
    for (size_t j = 0; j < inner; ++j) ss += (double)v.f32[i * inner + j] * v.f32[i * inner + j];
    const float norm = std::sqrt((float)ss);
    const float gi = g ? g->f32[i] : norm;
    const float denom = norm + 1e-7f;
    for (size_t j = 0; j < inner; ++j) w[i * inner + j] = (v.f32[i * inner + j] / denom) * gi;
  }
  if (bias) {
    bias->clear();
    if (has_tensor(prefix + ".bias")) {
      const HostTensor& b = tensor(prefix + ".bias");
      if (b.is_int || (int)b.numel() != bias_n)
        throw Error(NC_SHAPE_MISMATCH, "Failed to load DAC weights: '" + prefix + ".bias' has the wrong shape");
      *bias = b.f32;
    }
  }
  return w;
}

static std::vector<float> alpha_of(const HostTensor& t, int c, const std::string& name) {
  if (t.is_int || (int)t.numel() != c)
    throw Error(NC_SHAPE_MISMATCH, "Failed to load DAC weights: '" + name + "' has the wrong shape");
  return t.f32;
}

// Precision policy of the decoder's "tf32" mode (measured with a CPU emulation of tf32 operand
// rounding, DESIGN.md "Precision"): the final 96->1 conv and the 1x1 convs of the two narrow
// blocks contribute most of the output error and are HBM-bound anyway, so they run 3xTF32.
Precision DacEngine::boosted(Precision p, bool narrow) const {
  return (p == PREC_TF32 && dec_boost_ && narrow) ? PREC_3XTF32 : p;
}

void DacEngine::build_ru(ResUnit& ru, const std::string& p, int dim, int dil, Precision prec, int short_chains) {
  std::vector<float> b;
  ru.s1.build(alpha_of(tensor(p + ".snake1.alpha"), dim, p + ".snake1.alpha"));
  ru.s2.build(alpha_of(tensor(p + ".snake2.alpha"), dim, p + ".snake2.alpha"));
  ConvSpec s1;
  s1.cin = s1.cout = dim; s1.k = 7; s1.dilation = dil; s1.padding = (7 - 1) * dil / 2;  // ResidualUnit.cs:27
  auto w1 = folded_conv(p + ".conv1", dim, dim, 7, &b, dim);
  const bool wide = p.compare(0, 8, "decoder.") == 0 && dim > 128 && prec != PREC_FP32;
  const bool d16 = wide && dec_wide_prec_ == PREC_F16 && dec_h16_opt_ != 0;
  ru.c1.build(p + ".conv1", s1, w1, b, wide ? dec_wide_prec_ : prec, short_chains, d16);
  ConvSpec s2;
  s2.cin = s2.cout = dim; s2.k = 1;
  auto w2 = folded_conv(p + ".conv2", dim, dim, 1, &b, dim);
  const bool is_dec = p.compare(0, 8, "decoder.") == 0;
  // the 1x1 conv of a wide decoder unit follows its k7 conv's mode (one fp16 product by default): 68.6 dB instead
  // of 69.6 dB, +2.7 % sustained, and the C = 192 unit can then run as one fused launch with 64-byte weight tiles
  // the 1x1 conv's chain is dim/32 * 6 MMAs (<= 96 in the encoder): nothing to shorten unless forced
  ru.c2.build(p + ".conv2", s2, w2, b, wide ? dec_wide_prec_ : (is_dec ? boosted(prec, dim <= 192) : prec),
              short_chains == 1 ? 1 : 0, d16);
}

void DacEngine::finalize_weights() {
  bind();
  ready_ = false;
  std::vector<float> b;
  // ---- encoder (Modules/DAC/Encoder.cs:21-58)
  int d = cfg_.encoder_dim;
  {
    auto w = folded_conv("encoder.conv1", d, 1, 7, &b, d);
    if (b.empty()) b.assign(d, 0.f);
    cudaFree(d_conv_in_w_); cudaFree(d_conv_in_b_);
    d_conv_in_w_ = upload(w);
    d_conv_in_b_ = upload(b);
  }
  enc_blocks_.clear();
  for (size_t i = 0; i < cfg_.encoder_rates.size(); ++i) {
    const int s = cfg_.encoder_rates[i];
    auto blk = std::make_unique<EncBlock>();
    const std::string p = "encoder.block." + std::to_string(i);
    const int dils[3] = {1, 3, 9};
    const int from_end = (int)cfg_.encoder_rates.size() - 1 - (int)i;   // 0 = last block
    const Precision ru_prec = enc_tail_fp32_ >= 3 + 2 * from_end ? PREC_FP32 : enc_prec_;
    const Precision down_prec = enc_tail_fp32_ >= 2 + 2 * from_end ? PREC_FP32 : enc_prec_;
    // MMAs accumulated per output element: 3 products x 2 K steps per 32-channel chunk and tap
    const int ru_chain = 6 * (d * 7 / 32), down_chain = 6 * (d * 2 * s / 32);
    const bool last = i + 1 == cfg_.encoder_rates.size();
    int ru_short = 0, down_short = 0;
    if (enc_short_chains_ == 1) {
      ru_short = (ru_chain >= 250 && d > 128) ? 2 : 0;     // units the fused kernel runs (C <= 128) keep their chains
      down_short = down_chain >= 250 ? (last ? 1 : 2) : 0;
    } else if (enc_short_chains_ >= 2) {
      ru_short = (d > 128 || enc_short_chains_ >= (d > 64 ? 3 : 4)) ? 1 : 0;
      down_short = 1;
    }
    for (int u = 0; u < 3; ++u) build_ru(blk->ru[u], p + ".res_unit" + std::to_string(u + 1), d, dils[u], ru_prec, ru_short);
    blk->s.build(alpha_of(tensor(p + ".snake1.alpha"), d, p + ".snake1.alpha"));
    ConvSpec cs;
    cs.cin = d; cs.cout = 2 * d; cs.k = 2 * s; cs.stride = s; cs.padding = (s + 1) / 2;  // EncoderBlock.cs:27-33
    auto w = folded_conv(p + ".conv1", 2 * d, d, 2 * s, &b, 2 * d);
    blk->down.build(p + ".conv1", cs, w, b, down_prec, down_short);
    enc_blocks_.push_back(std::move(blk));
    d *= 2;
  }
  enc_snake_.build(alpha_of(tensor("encoder.snake1.alpha"), d, "encoder.snake1.alpha"));
  {
    ConvSpec cs;
    cs.cin = d; cs.cout = cfg_.latent_dim; cs.k = 3; cs.padding = 1;
    auto w = folded_conv("encoder.conv2", cfg_.latent_dim, d, 3, &b, cfg_.latent_dim);
    enc_out_.build("encoder.conv2", cs, w, b, enc_tail_fp32_ >= 1 ? PREC_FP32 : enc_prec_, enc_short_chains_ >= 1 ? 1 : 0);
  }
  // ---- quantiser (Modules/DAC/VectorQuantizer.cs:36-43)
  {
    for (float* p : rvq_alloc_) cudaFree(p);
    rvq_alloc_.clear();
    const int nq = cfg_.n_codebooks, D = cfg_.codebook_dim, K = cfg_.codebook_size, Dz = cfg_.latent_dim;
    std::vector<float> in_w((size_t)nq * D * Dz), in_b((size_t)nq * D), cb((size_t)nq * K * D), cb_sq((size_t)nq * K),
        out_w((size_t)nq * Dz * D), out_b((size_t)nq * Dz);
    for (int q = 0; q < nq; ++q) {
      const std::string p = "quantizer.quantizers." + std::to_string(q);
      auto wi = folded_conv(p + ".in_proj", D, Dz, 1, &b, D);
      if (b.empty()) b.assign(D, 0.f);
      std::copy(wi.begin(), wi.end(), in_w.begin() + (size_t)q * D * Dz);
      std::copy(b.begin(), b.end(), in_b.begin() + (size_t)q * D);
      auto wo = folded_conv(p + ".out_proj", Dz, D, 1, &b, Dz);
      if (b.empty()) b.assign(Dz, 0.f);
      std::copy(wo.begin(), wo.end(), out_w.begin() + (size_t)q * Dz * D);
      std::copy(b.begin(), b.end(), out_b.begin() + (size_t)q * Dz);
      const HostTensor& c = tensor(p + ".codebook.weight");
      if (c.is_int || c.shape.size() != 2 || c.shape[0] != K || c.shape[1] != D)
        throw Error(NC_SHAPE_MISMATCH, "Failed to load DAC weights: '" + p + ".codebook.weight' has the wrong shape");
      std::copy(c.f32.begin(), c.f32.end(), cb.begin() + (size_t)q * K * D);
      for (int k = 0; k < K; ++k) {
        float s = 0.f;
        for (int dd = 0; dd < D; ++dd) {
          const float x = c.f32[(size_t)k * D + dd];
          const float x2 = x * x;
          s += x2;
        }
        cb_sq[(size_t)q * K + k] = s;
      }
    }
    rvq_.n_stages = nq; rvq_.Dz = Dz; rvq_.D = D; rvq_.K = K;
    float* p;
    rvq_alloc_.push_back(p = upload(in_w)); rvq_.in_w = p;
    rvq_alloc_.push_back(p = upload(in_b)); rvq_.in_b = p;
    rvq_alloc_.push_back(p = upload(cb)); rvq_.cb = p;
    rvq_alloc_.push_back(p = upload(cb_sq)); rvq_.cb_sq = p;
    rvq_alloc_.push_back(p = upload(out_w)); rvq_.out_w = p;
    rvq_alloc_.push_back(p = upload(out_b)); rvq_.out_b = p;
    {   // shared-memory images of the stages for the block kernel (codec_kernels.cu)
      std::vector<float> blob;
      int per_stage = 0;
      for (int q = 0; q < nq; ++q) {
        auto v = rvq_stage_blob(Dz, D, K, in_w.data() + (size_t)q * D * Dz, in_b.data() + (size_t)q * D, cb.data() + (size_t)q * K * D,
                                cb_sq.data() + (size_t)q * K, out_w.data() + (size_t)q * Dz * D, out_b.data() + (size_t)q * Dz);
        if (v.empty()) { blob.clear(); per_stage = 0; break; }
        per_stage = (int)v.size();
        blob.insert(blob.end(), v.begin(), v.end());
      }
      rvq_.blob = nullptr; rvq_.blob_floats = 0;
      if (per_stage > 0 && per_stage % 4 == 0) {
        rvq_alloc_.push_back(p = upload(blob));
        rvq_.blob = p; rvq_.blob_floats = per_stage;
      }
    }
  }
  // ---- decoder (Modules/DAC/Decoder.cs:22-58)
  const int C = cfg_.decoder_dim;
  {
    ConvSpec cs;
    cs.cin = cfg_.latent_dim; cs.cout = C; cs.k = 7; cs.padding = 3;
    auto w = folded_conv("decoder.conv1", C, cfg_.latent_dim, 7, &b, C);
    dec_in_.build("decoder.conv1", cs, w, b, dec_prec_ != PREC_FP32 ? dec_wide_prec_ : dec_prec_, 0,
                  dec_prec_ != PREC_FP32 && dec_wide_prec_ == PREC_F16 && dec_h16_opt_ != 0);
  }
  dec_blocks_.clear();
  int cout = C;
  for (size_t i = 0; i < cfg_.decoder_rates.size(); ++i) {
    const int s = cfg_.decoder_rates[i];
    const int cin = C / (1 << i);
    cout = C / (1 << (i + 1));
    auto blk = std::make_unique<DecBlock>();
    const std::string p = "decoder.block." + std::to_string(i);
    blk->s.build(alpha_of(tensor(p + ".snake1.alpha"), cin, p + ".snake1.alpha"));
    ConvSpec cs;
    cs.transposed = true; cs.cin = cin; cs.cout = cout; cs.k = 2 * s; cs.stride = s; cs.padding = (s + 1) / 2;
    auto w = folded_conv(p + ".conv_t1", cin, cout, 2 * s, &b, cout);  // norm per in-channel (dims 1,2 of [Cin,Cout,k])
    blk->up.build(p + ".conv_t1", cs, w, b, (cin > 128 && dec_prec_ != PREC_FP32) ? dec_wide_prec_ : dec_prec_, 0,
                  cin > 128 && dec_prec_ != PREC_FP32 && dec_wide_prec_ == PREC_F16 && dec_h16_opt_ != 0);
    const int dils[3] = {1, 3, 9};
    for (int u = 0; u < 3; ++u) build_ru(blk->ru[u], p + ".res_unit" + std::to_string(u + 1), cout, dils[u], dec_prec_);
    dec_blocks_.push_back(std::move(blk));
  }
  dec_snake_.build(alpha_of(tensor("decoder.snake1.alpha"), cout, "decoder.snake1.alpha"));
  {
    ConvSpec cs;
    cs.cin = cout; cs.cout = 1; cs.k = 7; cs.padding = 3;
    auto w = folded_conv("decoder.conv2", 1, cout, 7, &b, 1);
    dec_out_.build("decoder.conv2", cs, w, b, boosted(dec_prec_, true));
    // Cout = 1: a GEMV, not a GEMM -- dedicated HBM-bound fp32 kernel when the channel count allows
    cudaFree(d_conv_out_w_); cudaFree(d_conv_out_b_);
    d_conv_out_w_ = d_conv_out_b_ = nullptr;
    conv_out_c_ = 0;
    if (cout % 32 == 0) {
      std::vector<float> wkc((size_t)7 * cout);
      for (int ci = 0; ci < cout; ++ci)
        for (int j = 0; j < 7; ++j) wkc[(size_t)j * cout + ci] = w[(size_t)ci * 7 + j];
      d_conv_out_w_ = upload(wkc);
      if (!b.empty()) d_conv_out_b_ = upload(b);
      conv_out_c_ = cout;
    }
  }
  // fp16-operand decoder path: decoder.conv1, every transposed conv with more than 128 input channels and every
  // residual unit with more than 128 channels must be on that executor (their activations are exchanged as fp16), and
  // the last block must be narrow (it hands fp32 to the fused units / the final conv)
  h16_ = dec_h16_opt_ != 0 && dec_in_.direct16() && !dec_blocks_.empty() && !wide_layer(cout);
  {
    int c_in = C;
    for (auto& blk : dec_blocks_) {
      const int c_out = c_in / 2;
      if (wide_layer(c_in) && !blk->up.direct16()) h16_ = false;
      if (wide_layer(c_out))
        for (auto& r : blk->ru) if (!r.c1.direct16() || !r.c2.direct16()) h16_ = false;
      c_in = c_out;
    }
    if (!wide_layer(C)) h16_ = false;
  }
  drop_tensors();
  ready_ = true;
}

static const char* prec_name(Precision p) { return precision_name(p); }

std::string DacEngine::describe() const {
  std::string s = "{\"codec\": \"DAC\", \"encoder_precision\": \"";
  s += prec_name(enc_prec_);
  s += "\", \"decoder_precision\": \"";
  s += prec_name(dec_prec_);
  s += "\", \"decoder_wide_precision\": \"";
  s += prec_name(dec_wide_prec_);
  s += dec_boost_ ? "\", \"decoder_boost\": true, \"layers\": {" : "\", \"decoder_boost\": false, \"layers\": {";
  bool first = true;
  auto add = [&](const ConvLayer& l) {
    if (l.name().empty()) return;
    s += first ? "\"" : ", \"";
    first = false;
    s += l.name() + "\": \"" + l.executor() + "\"";
  };
  for (auto& b : enc_blocks_) {
    for (auto& r : b->ru) { add(r.c1); add(r.c2); }
    add(b->down);
  }
  add(enc_out_);
  add(dec_in_);
  for (auto& b : dec_blocks_) {
    add(b->up);
    for (auto& r : b->ru) { add(r.c1); add(r.c2); }
  }
  add(dec_out_);
  s += "}}";
  return s;
}

int64_t DacEngine::padded_length(int64_t L) const {
  const int64_t hop = cfg_.hop();
  return (L + hop - 1) / hop * hop;  // Models/DAC.cs:151-153
}

int64_t DacEngine::decoded_length(int64_t T) const {
  int64_t t = dec_in_.out_len((int)T);
  for (auto& b : dec_blocks_) t = b->up.out_len((int)t);
  return dec_out_.out_len((int)t);
}

int DacEngine::micro_batch(int B, int64_t Lp) const {
  // per clip: 3 rotating buffers of max(Lp*enc_dim, decoder peak) floats
  const int64_t T = Lp / cfg_.hop();
  int64_t peak = Lp * cfg_.encoder_dim;
  int64_t t = dec_in_.out_len((int)T);
  peak = std::max<int64_t>(peak, t * cfg_.decoder_dim);
  int c = cfg_.decoder_dim;
  for (auto& b : dec_blocks_) {
    t = b->up.out_len((int)t);
    c /= 2;
    peak = std::max<int64_t>(peak, t * c);
  }
  // fp16 activations of the wide decoder layers (3 rotating buffers) + the fp16 latent
  int64_t peak16 = 0;
  if (h16_) {
    int64_t t16 = dec_in_.out_len((int)T);
    peak16 = t16 * cfg_.decoder_dim;
    int c16 = cfg_.decoder_dim;
    for (auto& b : dec_blocks_) {
      t16 = b->up.out_len((int)t16);
      c16 /= 2;
      if (wide_layer(c16)) peak16 = std::max<int64_t>(peak16, t16 * c16);
    }
  }
  const double per_clip = 3.0 * (double)peak * 4 + 3.0 * (double)peak16 * 2 + 2.5 * (double)T * cfg_.latent_dim * 4;
  int mb = (int)std::max(1.0, std::floor((double)max_workspace_bytes_ / per_clip));
  const_cast<DacEngine*>(this)->per_clip_elems_ = peak;
  const_cast<DacEngine*>(this)->per_clip_elems16_ = peak16;
  mb = std::min(mb, B);
  // even split: 512 clips at 21 per micro-batch would leave a short tail batch; 25 batches of 20-21 clips instead
  const int n_batches = (B + mb - 1) / mb;
  return (B + n_batches - 1) / n_batches;
}

void DacEngine::ensure_workspace(int mb, int64_t Lp) {
  micro_batch(mb, Lp);
  const int64_t T = Lp / cfg_.hop();
  for (auto& w : ws_) w.reserve((size_t)mb * per_clip_elems_ * sizeof(float));
  z_in_.reserve((size_t)mb * T * cfg_.latent_dim * sizeof(float));
  z_q_.reserve((size_t)mb * T * cfg_.latent_dim * sizeof(float));
  if (h16_) {
    for (auto& w : ws16_) w.reserve((size_t)mb * per_clip_elems16_ * 2);
    z16_.reserve((size_t)mb * T * cfg_.latent_dim * 2);
  }
}

// One ResidualUnit (Modules/DAC/ResidualUnit.cs:24-59): y = conv_k1(Snake2(conv_k7(Snake1(x)))) + x.
// Snake1 is the k7 conv's prologue (x stays raw for the residual); Snake2 is applied in the k7 conv's
// epilogue (its output has no other reader), so the 1x1 conv streams its input untouched; `post`, when
// given, is the Snake that follows this unit in the graph (block Snake before the strided / transposed
// conv, or the codec's final Snake) and is applied in the 1x1 conv's epilogue after the residual add.
int DacEngine::run_ru(const ResUnit& ru, int cur, int B, int T, const SnakeParams* post) {
  const LaunchCtx c = ctx();
  const int h = (cur + 1) % 3, y = (cur + 2) % 3;
  ConvRunArgs a;
  a.in = buf(cur); a.out = buf(h); a.batch = B; a.t_in = T;
  a.prologue = PRO_SNAKE; a.alpha = ru.s1.alpha; a.inv_alpha = ru.s1.inv_alpha;
  a.post = PRO_SNAKE; a.post_alpha = ru.s2.alpha; a.post_inv_alpha = ru.s2.inv_alpha;
  ConvRunArgs b;
  b.in = buf(h); b.out = buf(y); b.residual = buf(cur); b.batch = B; b.t_in = T;
  if (post) { b.post = PRO_SNAKE; b.post_alpha = post->alpha; b.post_inv_alpha = post->inv_alpha; }
  {   // whole unit in one launch (intermediate stays in TMEM / shared memory) when the shapes allow it
    ConvRunArgs fa = a;
    fa.out = buf(y); fa.residual = buf(cur);
    if (try_run_ru_fused(ru.c1, ru.c2, fa, b, c)) return y;
  }
  ru.c1.run(a, c);
  ru.c2.run(b, c);
  return y;
}

int DacEngine::run_encoder(const float* audio, long long audio_stride, int in_len, int B, int Lp, int* T_out) {
  const LaunchCtx c = ctx();
  launch_conv_cin1(audio, audio_stride, in_len, buf(0), Lp, cfg_.encoder_dim, d_conv_in_w_, d_conv_in_b_, 7, 1, 3, B, c);
  int cur = 0, T = Lp;
  for (size_t i = 0; i < enc_blocks_.size(); ++i) {
    auto& blk = enc_blocks_[i];
    for (int u = 0; u < 3; ++u) cur = run_ru(blk->ru[u], cur, B, T, u == 2 ? &blk->s : nullptr);
    ConvRunArgs a;   // input already carries the block's Snake (EncoderBlock.cs:26-33)
    a.in = buf(cur); a.out = buf((cur + 1) % 3); a.batch = B; a.t_in = T;
    if (i + 1 == enc_blocks_.size()) {   // Encoder.cs:44: Snake1d before the output conv
      a.post = PRO_SNAKE; a.post_alpha = enc_snake_.alpha; a.post_inv_alpha = enc_snake_.inv_alpha;
    }
    blk->down.run(a, c);
    T = blk->down.out_len(T);
    cur = (cur + 1) % 3;
  }
  ConvRunArgs a;
  a.in = buf(cur); a.out = z_in_.as<float>(); a.batch = B; a.t_in = T;
  enc_out_.run(a, c);
  *T_out = enc_out_.out_len(T);
  return cur;
}

// Wide ResidualUnit on the fp16-operand executor (Modules/DAC/ResidualUnit.cs:24-59): the k7 conv reads Snake1(x) as
// fp16 and writes Snake2(.) as fp16; the 1x1 conv reads that, adds the raw fp32 x and writes the raw fp32 result (for
// the next unit's residual) and / or the following Snake of it as fp16 (the next conv's operand).
void DacEngine::run_ru_h16(const ResUnit& ru, int* cur32, int* cur16, int B, int T, const SnakeParams* post, bool need32) {
  const LaunchCtx c = ctx();
  const int h16 = (*cur16 + 1) % 3, y16 = (*cur16 + 2) % 3, y32 = (*cur32 + 1) % 3;
  ConvRunArgs a;
  a.in16 = buf16(*cur16); a.out16 = buf16(h16); a.batch = B; a.t_in = T;
  a.post = PRO_SNAKE; a.post_alpha = ru.s2.alpha; a.post_inv_alpha = ru.s2.inv_alpha;
  ru.c1.run(a, c);
  ConvRunArgs b;
  b.in16 = buf16(h16); b.residual = buf(*cur32); b.batch = B; b.t_in = T;
  b.out = need32 ? buf(y32) : nullptr;
  b.out16 = buf16(y16);
  b.post = PRO_SNAKE; b.post_alpha = post->alpha; b.post_inv_alpha = post->inv_alpha;
  ru.c2.run(b, c);
  if (need32) *cur32 = y32;
  *cur16 = y16;
}

// input latent: z_q_ [B][T][latent] channels-last
int DacEngine::run_decoder(int, int B, int T, float* audio_out, long long) {
  const LaunchCtx c = ctx();
  if (h16_) {
    // ---- fp16-operand path for the wide layers (conv_h16.cu); narrow blocks continue on the fp32 / fused-unit path
    launch_f32_to_f16(z_q_.as<float>(), z16_.as<void>(), (long long)B * T * cfg_.latent_dim, c);
    ConvRunArgs a;
    a.in16 = z16_.as<void>(); a.out16 = buf16(0); a.batch = B; a.t_in = T;
    a.post = PRO_SNAKE; a.post_alpha = dec_blocks_[0]->s.alpha; a.post_inv_alpha = dec_blocks_[0]->s.inv_alpha;   // DecoderBlock.cs:26
    dec_in_.run(a, c);
    T = dec_in_.out_len(T);
    int cur32 = 0, cur16 = 0;
    bool on16 = true;    // the current activation is the fp16 buffer cur16 (already carrying the next conv's Snake)
    int ch = cfg_.decoder_dim;
    for (size_t i = 0; i < dec_blocks_.size(); ++i) {
      auto& blk = dec_blocks_[i];
      const int c_out = ch / 2;
      const SnakeParams* next = i + 1 < dec_blocks_.size() ? &dec_blocks_[i + 1]->s : &dec_snake_;
      if (on16) {
        // transposed conv: raw fp32 x (the first unit's residual) + fp16 Snake1(x) when the units are wide
        const bool wide_out = wide_layer(c_out);
        ConvRunArgs u;
        u.in16 = buf16(cur16); u.batch = B; u.t_in = T;
        u.out = buf(cur32);
        if (wide_out) {
          u.out16 = buf16((cur16 + 1) % 3);
          u.post = PRO_SNAKE; u.post_alpha = blk->ru[0].s1.alpha; u.post_inv_alpha = blk->ru[0].s1.inv_alpha;
        }
        blk->up.run(u, c);
        T = blk->up.out_len(T);
        if (wide_out) {
          cur16 = (cur16 + 1) % 3;
          for (int r = 0; r < 3; ++r)
            run_ru_h16(blk->ru[r], &cur32, &cur16, B, T, r < 2 ? &blk->ru[r + 1].s1 : next, r < 2);
        } else {
          on16 = false;
          for (int r = 0; r < 3; ++r) cur32 = run_ru(blk->ru[r], cur32, B, T, r == 2 ? next : nullptr);
        }
      } else {
        ConvRunArgs u;
        u.in = buf(cur32); u.out = buf((cur32 + 1) % 3); u.batch = B; u.t_in = T;
        blk->up.run(u, c);
        T = blk->up.out_len(T);
        cur32 = (cur32 + 1) % 3;
        for (int r = 0; r < 3; ++r) cur32 = run_ru(blk->ru[r], cur32, B, T, r == 2 ? next : nullptr);
      }
      ch = c_out;
    }
    ConvRunArgs o;   // Decoder.cs:44-46: (Snake already applied) -> conv -> tanh
    o.in = buf(cur32); o.out = audio_out; o.batch = B; o.t_in = T;
    o.act = ACT_TANH;
    if (conv_out_c_)
      launch_conv_cout1(buf(cur32), audio_out, T, conv_out_c_, d_conv_out_w_, d_conv_out_b_, 7, 3, 1, B, c);
    else
      dec_out_.run(o, c);
    return cur32;
  }
  ConvRunArgs a;
  a.in = z_q_.as<float>(); a.out = buf(0); a.batch = B; a.t_in = T;
  if (!dec_blocks_.empty()) {   // DecoderBlock.cs:26: Snake1d before the transposed conv
    a.post = PRO_SNAKE; a.post_alpha = dec_blocks_[0]->s.alpha; a.post_inv_alpha = dec_blocks_[0]->s.inv_alpha;
  }
  dec_in_.run(a, c);
  int cur = 0;
  T = dec_in_.out_len(T);
  for (size_t i = 0; i < dec_blocks_.size(); ++i) {
    auto& blk = dec_blocks_[i];
    ConvRunArgs u;
    u.in = buf(cur); u.out = buf((cur + 1) % 3); u.batch = B; u.t_in = T;
    blk->up.run(u, c);
    T = blk->up.out_len(T);
    cur = (cur + 1) % 3;
    const SnakeParams* next = i + 1 < dec_blocks_.size() ? &dec_blocks_[i + 1]->s : &dec_snake_;
    for (int r = 0; r < 3; ++r) cur = run_ru(blk->ru[r], cur, B, T, r == 2 ? next : nullptr);
  }
  ConvRunArgs o;   // Decoder.cs:44-46: (Snake already applied) -> conv -> tanh
  o.in = buf(cur); o.out = audio_out; o.batch = B; o.t_in = T;
  o.act = ACT_TANH;
  if (conv_out_c_)
    launch_conv_cout1(buf(cur), audio_out, T, conv_out_c_, d_conv_out_w_, d_conv_out_b_, 7, 3, 1, B, c);
  else
    dec_out_.run(o, c);
  return cur;
}

static int clamp_nq(int nq, int n_codebooks) { return (nq <= 0 || nq > n_codebooks) ? n_codebooks : nq; }

void DacEngine::encode_dev(const float* audio, int B, int64_t L, int nq_in, float* z, int64_t* codes, float* latents) {
  forward_impl(audio, B, L, nq_in, nullptr, codes, z, latents);
}

void DacEngine::forward_dev(const float* audio, int B, int64_t L, int nq_in, float* audio_out, int64_t* codes, float* z) {
  forward_impl(audio, B, L, nq_in, audio_out, codes, z, nullptr);
}

void DacEngine::forward_impl(const float* audio, int B, int64_t L, int nq_in, float* audio_out, int64_t* codes, float* z,
                             float* latents) {
  require_ready();
  bind();
  if (B <= 0 || L <= 0) throw Error(NC_INVALID_ARGUMENT, "batch and length must be positive");
  const int nq = clamp_nq(nq_in, cfg_.n_codebooks);
  const int64_t Lp = padded_length(L);
  if (Lp > (int64_t)1 << 30) throw Error(NC_INVALID_ARGUMENT, "clip too long");
  const int mb = micro_batch(B, Lp);
  ensure_workspace(mb, Lp);
  const int64_t T = Lp / cfg_.hop();
  const int64_t out_len = decoded_length(T);
  const LaunchCtx c = ctx();
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int nb = std::min(mb, B - b0);
    int Tm = 0;
    run_encoder(audio + (int64_t)b0 * L, L, (int)L, nb, (int)Lp, &Tm);
    launch_rvq_encode(rvq_, z_in_.as<float>(), z_q_.as<float>(), codes ? codes + (int64_t)b0 * nq * T : nullptr,
                      latents ? latents + (int64_t)b0 * nq * cfg_.codebook_dim * T : nullptr, nb, Tm, nq, c);
    if (z) launch_transpose_tc_to_ct(z_q_.as<float>(), z + (int64_t)b0 * cfg_.latent_dim * T, nb, cfg_.latent_dim, Tm, c);
    if (audio_out) run_decoder(0, nb, Tm, audio_out + (int64_t)b0 * out_len, out_len);
  }
  sync();
}

void DacEngine::decode_dev(const float* z, int B, int64_t T, float* audio_out) {
  require_ready();
  bind();
  if (B <= 0 || T <= 0) throw Error(NC_INVALID_ARGUMENT, "batch and frames must be positive");
  const int64_t Lp = T * cfg_.hop();
  const int mb = micro_batch(B, Lp);
  ensure_workspace(mb, Lp);
  const int64_t out_len = decoded_length(T);
  const LaunchCtx c = ctx();
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int nb = std::min(mb, B - b0);
    launch_transpose_ct_to_tc(z + (int64_t)b0 * cfg_.latent_dim * T, z_q_.as<float>(), nb, cfg_.latent_dim, (int)T, c);
    run_decoder(0, nb, (int)T, audio_out + (int64_t)b0 * out_len, out_len);
  }
  sync();
}

void DacEngine::from_codes_dev(const int64_t* codes, int B, int nq, int64_t T, float* z) {
  require_ready();
  bind();
  if (B <= 0 || T <= 0 || nq <= 0 || nq > cfg_.n_codebooks)
    throw Error(NC_INVALID_ARGUMENT, "from_codes: bad batch / frames / n_quantizers");
  const int64_t Lp = T * cfg_.hop();
  const int mb = micro_batch(B, Lp);
  ensure_workspace(mb, Lp);
  const LaunchCtx c = ctx();
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int nb = std::min(mb, B - b0);
    launch_rvq_from_codes(rvq_, codes + (int64_t)b0 * nq * T, z_q_.as<float>(), nb, (int)T, nq, c);
    launch_transpose_tc_to_ct(z_q_.as<float>(), z + (int64_t)b0 * cfg_.latent_dim * T, nb, cfg_.latent_dim, (int)T, c);
  }
  sync();
}

void DacEngine::decode_codes_dev(const int64_t* codes, int B, int nq, int64_t T, float* audio_out) {
  require_ready();
  bind();
  if (B <= 0 || T <= 0 || nq <= 0 || nq > cfg_.n_codebooks)
    throw Error(NC_INVALID_ARGUMENT, "decode_codes: bad batch / frames / n_quantizers");
  const int64_t Lp = T * cfg_.hop();
  const int mb = micro_batch(B, Lp);
  ensure_workspace(mb, Lp);
  const int64_t out_len = decoded_length(T);
  const LaunchCtx c = ctx();
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int nb = std::min(mb, B - b0);
    launch_rvq_from_codes(rvq_, codes + (int64_t)b0 * nq * T, z_q_.as<float>(), nb, (int)T, nq, c);
    run_decoder(0, nb, (int)T, audio_out + (int64_t)b0 * out_len, out_len);
  }
  sync();
}

void DacEngine::decode_dia_dev(const int64_t* generated, int B, int T, int C, const int* delay, const int64_t* lengths,
                               float* audio_out, int64_t audio_stride) {
  require_ready();
  bind();
  if (B <= 0 || T <= 0 || !generated || !delay || !lengths || !audio_out) throw Error(NC_INVALID_ARGUMENT, "decode_dia: bad arguments");
  if (C != cfg_.n_codebooks) throw Error(NC_INVALID_ARGUMENT, "decode_dia: channel count must equal the number of codebooks");
  int max_delay = 0;
  for (int c = 0; c < C; ++c) {
    if (delay[c] < 0) throw Error(NC_INVALID_ARGUMENT, "decode_dia: negative delay");
    max_delay = std::max(max_delay, delay[c]);
  }
  const int t_valid = T - max_delay;   // codebook[:, :-maxDelay, :]  (Dia.cs:1039)
  std::map<int64_t, std::vector<int>> groups;
  for (int b = 0; b < B; ++b) {
    if (lengths[b] < 0 || lengths[b] > t_valid) throw Error(NC_INVALID_ARGUMENT, "decode_dia: length exceeds T - max(delay)");
    if (lengths[b] > 0) groups[lengths[b]].push_back(b);
    if (lengths[b] * cfg_.hop() > audio_stride) throw Error(NC_INVALID_ARGUMENT, "decode_dia: audio_stride too small");
  }
  int* d_delay = static_cast<int*>(dia_idx_.reserve((size_t)(C + B) * sizeof(int)));
  int* d_items = d_delay + C;
  NC_CUDA(cudaMemcpyAsync(d_delay, delay, (size_t)C * sizeof(int), cudaMemcpyHostToDevice, stream_));
  const LaunchCtx c = ctx();
  for (auto& g : groups) {
    const int64_t len = g.first;
    const std::vector<int>& items = g.second;
    const int n = (int)items.size();
    NC_CUDA(cudaMemcpyAsync(d_items, items.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, stream_));
    int64_t* codes = static_cast<int64_t*>(dia_codes_.reserve((size_t)n * C * len * sizeof(int64_t)));
    launch_dia_revert(generated, d_items, d_delay, codes, n, T, C, (int)len, cfg_.codebook_size, c);
    const int64_t Lp = len * cfg_.hop();
    const int mb = micro_batch(n, Lp);
    ensure_workspace(mb, Lp);
    const int64_t out_len = decoded_length(len);
    float* tmp = static_cast<float*>(dia_audio_.reserve((size_t)mb * out_len * sizeof(float)));
    for (int b0 = 0; b0 < n; b0 += mb) {
      const int nb = std::min(mb, n - b0);
      launch_rvq_from_codes(rvq_, codes + (int64_t)b0 * C * len, z_q_.as<float>(), nb, (int)len, C, c);
      run_decoder(0, nb, (int)len, tmp, out_len);
      for (int i = 0; i < nb; ++i)   // scatter to each item's row of the caller's buffer
        NC_CUDA(cudaMemcpyAsync(audio_out + (int64_t)items[b0 + i] * audio_stride, tmp + (int64_t)i * out_len,
                                (size_t)std::min<int64_t>(out_len, audio_stride) * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
    }
    sync();   // d_items / codes are reused by the next group
  }
  sync();
}

Engine* create_engine(nc_codec_kind kind, const void* cfg, size_t cfg_size, int device_index) {
  if (!cfg) throw Error(NC_INVALID_ARGUMENT, "config is null");
  switch (kind) {
    case NC_CODEC_DAC: {
      if (cfg_size != sizeof(nc_dac_config)) throw Error(NC_INVALID_ARGUMENT, "cfg_size != sizeof(nc_dac_config)");
      return new DacEngine(*static_cast<const nc_dac_config*>(cfg), device_index);
    }
    case NC_CODEC_SNAC: {
      if (cfg_size != sizeof(nc_snac_config)) throw Error(NC_INVALID_ARGUMENT, "cfg_size != sizeof(nc_snac_config)");
      return new SnacEngine(*static_cast<const nc_snac_config*>(cfg), device_index);
    }
    case NC_CODEC_ENCODEC: {
      if (cfg_size != sizeof(nc_encodec_config)) throw Error(NC_INVALID_ARGUMENT, "cfg_size != sizeof(nc_encodec_config)");
      return new EncodecEngine(*static_cast<const nc_encodec_config*>(cfg), device_index);
    }
    default:
      throw Error(NC_UNSUPPORTED, "unknown codec kind");
  }
}

}  // namespace nc
