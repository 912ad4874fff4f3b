// Encodec engine: the 24 kHz preset (mono, causal, weight-norm, one frame per clip) and the 48 kHz preset (stereo, non-causal,
// GroupNorm(1, C) after every conv, 1 s segments with 1 % overlap, per-segment loudness scale, triangular overlap-add:
// Config/Encodec/EncodecConfig.cs:37-66, Modules/Encodec/NormConv1d.cs:52-100, AudioTools/AudioTensorDSP.cs:161-261).
// Graph (paths under /root/reference/NeuralCodecs.Torch/):
//   Models/Encodec.cs:213-296,436-489; Modules/Encodec/SEANetEncoder.cs:37-148, SEANetDecoder.cs:40-153,
//   SEANetResnetBlock.cs:30-86, SConv1d.cs:144-173,245-274, SConvTranspose1d.cs:116-139, SLSTM.cs:40-57,
//   ResidualVectorQuantizer.cs:107-157, EuclideanCodebook.cs:155-182.
// Activations are channels-last with 16 margin rows around every clip: the reflect padding of SConv1d (causal: k - s on the
// left; non-causal: split left / right; plus the stride-alignment extra on the right) is materialised in those rows by a
// tiny fix-up kernel, after which every conv
// is a plain valid convolution for the tcgen05 kernel.  ELU commutes with reflect padding, so it is applied in the
// producing layer's epilogue (or the consumer's prologue) like DAC's Snake.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "engine.h"

namespace nc {

EncodecEngine::EncodecEngine(const nc_encodec_config& c, int device_index) : Engine(device_index) {
  if (c.struct_size != sizeof(nc_encodec_config)) throw Error(NC_INVALID_ARGUMENT, "nc_encodec_config.struct_size mismatch");
  if (c.n_ratios < 1 || c.n_ratios > NC_MAX_RATES) throw Error(NC_INVALID_ARGUMENT, "Encodec config: ratio count out of range");
  cfg_.sample_rate = c.sample_rate;
  cfg_.channels = c.channels;
  cfg_.n_filters = c.n_filters;
  cfg_.dimension = c.dimension;
  cfg_.ratios.assign(c.ratios, c.ratios + c.n_ratios);
  cfg_.n_residual_layers = c.n_residual_layers;
  cfg_.lstm_layers = c.lstm_layers;
  cfg_.codebook_size = c.codebook_size;
  cfg_.n_quantizers = c.n_quantizers;
  cfg_.causal = c.causal != 0;
  cfg_.group_norm = c.norm_type == NC_ENCODEC_NORM_TIME_GROUP;
  cfg_.normalize = c.normalize != 0;
  cfg_.segment_s = c.segment_s;
  cfg_.overlap = c.overlap;
  if (c.norm_type != NC_ENCODEC_NORM_WEIGHT && c.norm_type != NC_ENCODEC_NORM_TIME_GROUP)
    throw Error(NC_UNSUPPORTED, "Encodec: norm must be weight_norm or time_group_norm");
  if (cfg_.channels < 1 || cfg_.channels > 2) throw Error(NC_INVALID_ARGUMENT, "Invalid number of channels: " + std::to_string(cfg_.channels));   // Encodec.cs:268-271
  if (cfg_.group_norm && cfg_.causal) throw Error(NC_INVALID_ARGUMENT, "GroupNorm doesn't support causal evaluation");   // NormConv1d.cs:143-147
  if (cfg_.segment_s < 0.f || !(cfg_.overlap >= 0.f && cfg_.overlap < 1.f)) throw Error(NC_INVALID_ARGUMENT, "Encodec config: segment / overlap out of range");
  if (cfg_.segment_s > 0.f && cfg_.segment_length() < 1) throw Error(NC_INVALID_ARGUMENT, "Encodec config: segment shorter than one sample");
  generic_io_ = cfg_.channels != 1 || !cfg_.causal || cfg_.group_norm || cfg_.normalize || cfg_.segment_s > 0.f;
  if (cfg_.n_residual_layers != 1) throw Error(NC_UNSUPPORTED, "Encodec: n_residual_layers must be 1");
  if (cfg_.lstm_layers != 0 && cfg_.lstm_layers != 2) throw Error(NC_UNSUPPORTED, "Encodec: lstm_layers must be 0 or 2");
  if (cfg_.dimension != 128) throw Error(NC_UNSUPPORTED, "Encodec: dimension (codebook dim) must be 128");
  if (cfg_.n_filters % 32 != 0) throw Error(NC_UNSUPPORTED, "Encodec: n_filters must be a multiple of 32");
  for (int r : cfg_.ratios)
    if (r < 1 || r > 8) throw Error(NC_UNSUPPORTED, "Encodec: ratios must be in 1..8");
  if (cfg_.n_quantizers < 1 || cfg_.codebook_size < 1) throw Error(NC_INVALID_ARGUMENT, "Encodec config: non-positive field");
}

EncodecEngine::~EncodecEngine() {
  cudaSetDevice(device_);
  cudaFree(d_conv_in_w_); cudaFree(d_conv_in_b_); cudaFree(d_conv_out_w_); cudaFree(d_conv_out_b_);
  for (float* p : embed_) cudaFree(p);
  for (float* p : embed_sq_) cudaFree(p);
  cudaFree(d_embed_ptrs_);
  for (auto& kv : gn_) { cudaFree(kv.second.gamma); cudaFree(kv.second.beta); }
  for (auto* l : {&enc_lstm_, &dec_lstm_})
    for (float*& p : l->whh) { cudaFree(p); p = nullptr; }
}

void EncodecEngine::set_option(const std::string& key, const std::string& value) {
  if (key == "encoder_short_chains") {
    // tensor-core encoder modes only: keep the accumulation chains short (conv_plan.h acc_split); 0 off, 1 folded partials
    enc_short_chains_ = std::atoi(value.c_str());
    if (ready_) throw Error(NC_INVALID_ARGUMENT, "precision options must be set before weights are loaded");
  } else if (key == "precision" || key == "encoder_precision" || key == "decoder_precision") {
    if (key != "decoder_precision") enc_prec_ = parse_precision(value);
    if (key != "encoder_precision") dec_prec_ = parse_precision(value);
    if (ready_) throw Error(NC_INVALID_ARGUMENT, "precision options must be set before weights are loaded");
  } else {
    Engine::set_option(key, value);
  }
}

// Accumulation-chain policy of the encoder's tensor-core layers: a layer accumulates cin*k/32 K chunks of 12 (3xTF32) or 6
// (16-bit splits) MMAs per output; chains of at most one fold interval (96 MMAs: the early, long-T layers) run the plain
// kernel with its wide N tiles, longer ones fold (conv_plan.h acc_split = 1).
int EncodecEngine::chain_mode(int cin, int k) const {
  if (enc_short_chains_ <= 0) return 0;
  if (enc_short_chains_ >= 2) return 1;
  const int per_chunk = enc_prec_ == PREC_3XTF32 ? 12 : 6;
  return (cin * k + 31) / 32 * per_chunk > 96 ? 1 : 0;
}

void EncodecEngine::require_ready() const {
  if (!ready_) throw Error(NC_BAD_WEIGHTS, "Encodec weights have not been loaded");
}

std::string EncodecEngine::describe() const {
  std::string s = "{\"codec\": \"Encodec\", \"encoder_precision\": \"";
  s += precision_name(enc_prec_);
  s += "\", \"decoder_precision\": \"";
  s += precision_name(dec_prec_);
  s += "\", \"layers\": {";
  bool first = true;
  auto add = [&](const ConvLayer& l) {
    if (l.name().empty()) return;
    s += first ? "\"" : ", \"";
    first = false;
    s += l.name() + "\": \"" + l.executor() + "\"";
  };
  add(conv_in_l_);
  for (size_t i = 0; i < enc_res_.size(); ++i) { add(enc_res_[i]->shortcut); add(enc_res_[i]->c3); add(enc_res_[i]->c1); add(*enc_down_[i]); }
  add(enc_lstm_.ih[0]); add(enc_lstm_.ih[1]); add(enc_out_); add(dec_in_); add(dec_lstm_.ih[0]); add(dec_lstm_.ih[1]);
  for (size_t i = 0; i < dec_res_.size(); ++i) { add(*dec_up_[i]); add(dec_res_[i]->shortcut); add(dec_res_[i]->c3); add(dec_res_[i]->c1); }
  add(conv_out_l_);
  s += "}";
  if (cfg_.group_norm) s += ", \"norm\": \"time_group_norm\"";
  if (segmented()) s += ", \"segment_length\": " + std::to_string(cfg_.segment_length()) + ", \"segment_stride\": " + std::to_string(cfg_.segment_stride());
  s += "}";
  return s;
}

// w = (v / ||v||_(1,2)) * (g - 1e-7)   (Modules/Encodec/WNConv1d.cs:113-127, WNConvTranspose1d.cs:124-156)
// time_group_norm models carry the plain conv weight (NormConv1d.cs:52-63; keys SConv1d.cs:119-128).
std::vector<float> EncodecEngine::folded(const std::string& p, int d0, int d1, int k, std::vector<float>* bias, int bias_n) {
  const HostTensor& v = tensor(p + (cfg_.group_norm ? ".conv.weight" : ".conv.weight_v"));
  if (v.is_int || v.shape.size() != 3 || v.shape[0] != d0 || v.shape[1] != d1 || v.shape[2] != k)
    throw Error(NC_SHAPE_MISMATCH, "Failed to load Encodec weights: '" + p + "' has the wrong shape");
  std::vector<float> w(v.f32.size());
  if (cfg_.group_norm) {
    w = v.f32;
  } else {
    const HostTensor& g = tensor(p + ".conv.weight_g");
    if (g.is_int || (int64_t)g.numel() != d0)
      throw Error(NC_SHAPE_MISMATCH, "Failed to load Encodec weights: '" + p + ".conv.weight_g' has the wrong shape");
    const size_t inner = (size_t)d1 * k;
    for (int i = 0; i < d0; ++i) {
      double ss = 0;
      for (size_t j = 0; j < inner; ++j) ss += (double)v.f32[i * inner + j] * v.f32[i * inner + j];
      const float norm = std::sqrt((float)ss), gi = g.f32[i] - 1e-7f;
      for (size_t j = 0; j < inner; ++j) w[i * inner + j] = (v.f32[i * inner + j] / norm) * gi;
    }
  }
  if (bias) {
    bias->clear();
    if (has_tensor(p + ".conv.bias")) {
      const HostTensor& b = tensor(p + ".conv.bias");
      if (b.is_int || (int)b.numel() != bias_n)
        throw Error(NC_SHAPE_MISMATCH, "Failed to load Encodec weights: '" + p + ".conv.bias' has the wrong shape");
      *bias = b.f32;
    }
  }
  return w;
}

static std::vector<float> pad3e(const std::vector<float>& w, int d0, int d1, int k, int p0, int p1) {
  if (d0 == p0 && d1 == p1) return w;
  std::vector<float> o((size_t)p0 * p1 * k, 0.f);
  for (int i = 0; i < d0; ++i)
    for (int j = 0; j < d1; ++j) std::memcpy(&o[((size_t)i * p1 + j) * k], &w[((size_t)i * d1 + j) * k], sizeof(float) * k);
  return o;
}
static std::vector<float> pad1e(const std::vector<float>& b, int n, int p) {
  std::vector<float> o((size_t)p, 0.f);
  for (int i = 0; i < n && i < (int)b.size(); ++i) o[i] = b[i];
  return o;
}
static int pad32e(int c) { return (c + 31) / 32 * 32; }

// GroupNorm(1, C, eps 1e-5, affine) of a conv (NormConv1d.cs:136-160): keys "<p>.norm.{weight,bias}"; padded channels get
// gamma = beta = 0 so they stay zero.
void EncodecEngine::load_gn(const std::string& p, int c_real, int c_pad) {
  if (!cfg_.group_norm) return;
  const HostTensor& g = tensor(p + ".norm.weight");
  const HostTensor& b = tensor(p + ".norm.bias");
  if (g.is_int || b.is_int || (int)g.numel() != c_real || (int)b.numel() != c_real)
    throw Error(NC_SHAPE_MISMATCH, "Failed to load Encodec weights: '" + p + ".norm' has the wrong shape");
  Gn& n = gn_[p];
  cudaFree(n.gamma); cudaFree(n.beta);
  n.gamma = upload(pad1e(g.f32, c_real, c_pad));
  n.beta = upload(pad1e(b.f32, c_real, c_pad));
  n.c_real = c_real;
}

void EncodecEngine::build_res(Res& r, const std::string& p, int dim) {
  std::vector<float> b;
  const Precision prec_ = p.compare(0, 8, "encoder.") == 0 ? enc_prec_ : dec_prec_;
  const int hid = dim / 2, hp = pad32e(hid);
  r.hidden_p = hp;
  ConvSpec s1;  // shortcut: SConv1d(dim, dim, 1)
  s1.cin = s1.cout = dim; s1.k = 1;
  auto ws = folded(p + ".shortcut", dim, dim, 1, &b, dim);
  const bool enc = p.compare(0, 8, "encoder.") == 0;
  r.shortcut.build(p + ".shortcut", s1, ws, b, prec_, enc ? chain_mode(dim, 1) : 0);
  load_gn(p + ".shortcut", dim, dim);
  load_gn(p + ".block.1", hid, hp);
  load_gn(p + ".block.3", dim, dim);
  ConvSpec s3;  // block.1: SConv1d(dim, dim/2, 3): valid conv over the left-padded input
  s3.cin = dim; s3.cout = hp; s3.k = 3;
  auto w3 = folded(p + ".block.1", hid, dim, 3, &b, hid);
  r.c3.build(p + ".block.1", s3, pad3e(w3, hid, dim, 3, hp, dim), pad1e(b, hid, hp), prec_, enc ? chain_mode(dim, 3) : 0);
  ConvSpec s2;  // block.3: SConv1d(dim/2, dim, 1)
  s2.cin = hp; s2.cout = dim; s2.k = 1;
  auto w1 = folded(p + ".block.3", dim, hid, 1, &b, dim);
  r.c1.build(p + ".block.3", s2, pad3e(w1, dim, hid, 1, dim, hp), b, prec_, enc ? chain_mode(hp, 1) : 0);
}

void EncodecEngine::build_lstm(Lstm& l, const std::string& p, int dim) {
  const Precision prec_ = p.compare(0, 8, "encoder.") == 0 ? enc_prec_ : dec_prec_;
  l.layers = cfg_.lstm_layers;
  for (int i = 0; i < l.layers; ++i) {
    const std::string sfx = "_l" + std::to_string(i);
    const HostTensor& wih = tensor(p + ".lstm.weight_ih" + sfx);
    const HostTensor& whh = tensor(p + ".lstm.weight_hh" + sfx);
    const HostTensor& bih = tensor(p + ".lstm.bias_ih" + sfx);
    const HostTensor& bhh = tensor(p + ".lstm.bias_hh" + sfx);
    if ((int64_t)wih.numel() != (int64_t)4 * dim * dim || (int64_t)whh.numel() != (int64_t)4 * dim * dim ||
        (int64_t)bih.numel() != 4 * dim || (int64_t)bhh.numel() != 4 * dim)
      throw Error(NC_SHAPE_MISMATCH, "Failed to load Encodec weights: '" + p + ".lstm' has the wrong shape");
    std::vector<float> bsum((size_t)4 * dim);
    for (int j = 0; j < 4 * dim; ++j) bsum[j] = bih.f32[j] + bhh.f32[j];
    ConvSpec s;   // hoisted input projection: one GEMM over all time steps
    s.cin = dim; s.cout = 4 * dim; s.k = 1;
    l.ih[i].build(p + ".lstm.weight_ih" + sfx, s, wih.f32, bsum, prec_ == PREC_FP32 ? PREC_FP32 : PREC_3XTF32,
                  p.compare(0, 8, "encoder.") == 0 ? chain_mode(dim, 1) : 0);
    cudaFree(l.whh[i]);
    l.whh[i] = upload(whh.f32);
  }
}

void EncodecEngine::finalize_weights() {
  bind();
  ready_ = false;
  std::vector<float> b;
  const int nf = cfg_.n_filters;
  // ---- encoder (SEANetEncoder.cs:60-125): Sequential indices
  if (!generic_io_) {
    auto w = folded("encoder.layers.0", nf, 1, 7, &b, nf);
    cudaFree(d_conv_in_w_); cudaFree(d_conv_in_b_);
    d_conv_in_w_ = upload(w);
    d_conv_in_b_ = upload(pad1e(b, nf, nf));
  } else {   // SConv1d(channels, n_filters, 7) over the channel-padded channels-last segment
    ConvSpec cs;
    cs.cin = cin_pad_; cs.cout = nf; cs.k = 7;
    auto w = folded("encoder.layers.0", nf, cfg_.channels, 7, &b, nf);
    conv_in_l_.build("encoder.layers.0", cs, pad3e(w, nf, cfg_.channels, 7, nf, cin_pad_), b, enc_prec_, chain_mode(cin_pad_, 7));
    load_gn("encoder.layers.0", nf, nf);
  }
  enc_res_.clear(); enc_down_.clear();
  int mult = 1, idx = 1;
  for (int i = (int)cfg_.ratios.size() - 1; i >= 0; --i) {
    const int r = cfg_.ratios[i], dim = mult * nf;
    auto res = std::make_unique<Res>();
    build_res(*res, "encoder.layers." + std::to_string(idx), dim);
    enc_res_.push_back(std::move(res));
    idx += 2;  // resnet, ELU
    auto down = std::make_unique<ConvLayer>();
    ConvSpec cs;
    cs.cin = dim; cs.cout = 2 * dim; cs.k = 2 * r; cs.stride = r;   // valid conv over the padded input
    const std::string p = "encoder.layers." + std::to_string(idx);
    auto w = folded(p, 2 * dim, dim, 2 * r, &b, 2 * dim);
    down->build(p, cs, w, b, enc_prec_, chain_mode(dim, 2 * r));
    load_gn(p, 2 * dim, 2 * dim);
    enc_down_.push_back(std::move(down));
    ++idx;
    mult *= 2;
  }
  const int dim_top = mult * nf;
  if (cfg_.lstm_layers > 0) { build_lstm(enc_lstm_, "encoder.layers." + std::to_string(idx), dim_top); ++idx; }
  ++idx;  // ELU
  {
    ConvSpec cs;
    cs.cin = dim_top; cs.cout = cfg_.dimension; cs.k = 7;
    const std::string p = "encoder.layers." + std::to_string(idx);
    auto w = folded(p, cfg_.dimension, dim_top, 7, &b, cfg_.dimension);
    enc_out_.build(p, cs, w, b, enc_prec_, chain_mode(dim_top, 7));
    load_gn(p, cfg_.dimension, cfg_.dimension);
  }
  // ---- quantiser codebooks (EuclideanCodebook.cs:22-25)
  {
    for (float* p : embed_) cudaFree(p);
    for (float* p : embed_sq_) cudaFree(p);
    embed_.clear(); embed_sq_.clear();
    std::vector<const float*> ptrs;
    for (int q = 0; q < cfg_.n_quantizers; ++q) {
      const std::string p = "quantizer.layers." + std::to_string(q) + ".codebook.embed";
      if (!has_tensor(p)) break;   // files may carry fewer layers than the constructor builds
      const HostTensor& e = tensor(p);
      if (e.is_int || e.shape.size() != 2 || e.shape[0] != cfg_.codebook_size || e.shape[1] != cfg_.dimension)
        throw Error(NC_SHAPE_MISMATCH, "Failed to load Encodec weights: '" + p + "' has the wrong shape");
      std::vector<float> sq((size_t)cfg_.codebook_size);
      for (int k = 0; k < cfg_.codebook_size; ++k) {
        float s = 0.f;
        for (int d = 0; d < cfg_.dimension; ++d) { const float x = e.f32[(size_t)k * cfg_.dimension + d]; s += x * x; }
        sq[k] = s;
      }
      embed_.push_back(upload(e.f32));
      embed_sq_.push_back(upload(sq));
      ptrs.push_back(embed_.back());
    }
    if (embed_.empty()) throw Error(NC_BAD_WEIGHTS, "Failed to load Encodec weights: no quantizer codebooks found");
    cudaFree(d_embed_ptrs_);
    NC_CUDA(cudaMalloc(&d_embed_ptrs_, ptrs.size() * sizeof(float*)));
    NC_CUDA(cudaMemcpy(d_embed_ptrs_, ptrs.data(), ptrs.size() * sizeof(float*), cudaMemcpyHostToDevice));
  }
  // ---- decoder (SEANetDecoder.cs:75-145)
  mult = 1 << cfg_.ratios.size();
  {
    ConvSpec cs;
    cs.cin = cfg_.dimension; cs.cout = mult * nf; cs.k = 7;
    auto w = folded("decoder.layers.0", mult * nf, cfg_.dimension, 7, &b, mult * nf);
    dec_in_.build("decoder.layers.0", cs, w, b, dec_prec_);
    load_gn("decoder.layers.0", mult * nf, mult * nf);
  }
  idx = 1;
  if (cfg_.lstm_layers > 0) { build_lstm(dec_lstm_, "decoder.layers." + std::to_string(idx), mult * nf); ++idx; }
  dec_res_.clear(); dec_up_.clear();
  for (size_t i = 0; i < cfg_.ratios.size(); ++i) {
    const int r = cfg_.ratios[i], cin = mult * nf, cout = cin / 2;
    ++idx;  // ELU
    auto up = std::make_unique<ConvLayer>();
    ConvSpec cs;
    cs.transposed = true; cs.cin = cin; cs.cout = cout; cs.k = 2 * r; cs.stride = r;   // padding 0; the tail is trimmed
    const std::string p = "decoder.layers." + std::to_string(idx);
    auto w = folded(p, cin, cout, 2 * r, &b, cout);
    up->build(p, cs, w, b, dec_prec_);
    load_gn(p, cout, cout);
    dec_up_.push_back(std::move(up));
    ++idx;
    auto res = std::make_unique<Res>();
    build_res(*res, "decoder.layers." + std::to_string(idx), cout);
    dec_res_.push_back(std::move(res));
    ++idx;
    mult /= 2;
  }
  ++idx;  // ELU
  {
    const std::string p = "decoder.layers." + std::to_string(idx);
    if (generic_io_) {   // SConv1d(n_filters, channels, 7) with the output channels padded to one N tile
      ConvSpec cs;
      cs.cin = nf; cs.cout = cout_pad_; cs.k = 7;
      auto w = folded(p, cfg_.channels, nf, 7, &b, cfg_.channels);
      conv_out_l_.build(p, cs, pad3e(w, cfg_.channels, nf, 7, cout_pad_, nf), pad1e(b, cfg_.channels, cout_pad_), dec_prec_);
      load_gn(p, cfg_.channels, cout_pad_);
      drop_tensors();
      ready_ = true;
      return;
    }
    auto w = folded(p, 1, nf, 7, &b, 1);
    std::vector<float> wkc((size_t)7 * nf);
    for (int ci = 0; ci < nf; ++ci)
      for (int j = 0; j < 7; ++j) wkc[(size_t)j * nf + ci] = w[(size_t)ci * 7 + j];
    cudaFree(d_conv_out_w_); cudaFree(d_conv_out_b_);
    d_conv_out_w_ = upload(wkc);
    d_conv_out_b_ = b.empty() ? nullptr : upload(b);
    conv_out_c_ = nf;
  }
  drop_tensors();
  ready_ = true;
}

// ------------------------------------------------------------------------------------ shapes
// SConv1d.forward's padding (SConv1d.cs:144-173): padding_total = k - s goes on the left (causal) or is split
// right = padding_total / 2, left = the rest (non-causal); the stride alignment extra goes on the right (computed with a
// float32 division, :245-250); and -- when the input is not longer than the larger pad -- Pad1d's short-input branch
// zero-extends on the right first (:258-272).  The zero extension is not trimmed afterwards, so such a layer lengthens
// the sequence.
EncodecEngine::SPad EncodecEngine::sconv_pad(int64_t T, int k, int s, bool causal) {
  const int pt = k - s;
  const float n_frames = ((float)(T - k + pt)) / (float)s + 1.0f;
  const int64_t ideal = ((int64_t)std::ceil(n_frames) - 1) * s + (k - pt);
  const int extra = (int)(ideal - T);
  SPad p;
  if (causal) {
    p.left = pt;
    p.right = extra;
  } else {
    const int right = pt / 2;
    p.left = pt - right;
    p.right = right + extra;
  }
  const int m = std::max(p.left, p.right);
  p.extra_zero = T <= m ? (int)(m - T + 1) : 0;
  p.t_out = (int)((T + p.extra_zero + p.left + p.right - k) / s + 1);
  return p;
}

int64_t EncodecEngine::frames(int64_t L) const {
  int64_t t = sconv_pad(L, 7, 1).t_out;                                    // SEANetEncoder: SConv1d(channels, n_filters, 7)
  for (int i = (int)cfg_.ratios.size() - 1; i >= 0; --i) {
    const int r = cfg_.ratios[i];
    const SPad k3 = sconv_pad(t, 3, 1);                                    // resnet block: the k3 branch must keep the length
    if (k3.t_out != t) throw Error(NC_INVALID_ARGUMENT, "Encodec: clip too short: a residual block's branches would differ in length (the reference's tensor add fails)");
    t = sconv_pad(t, 2 * r, r).t_out;
  }
  return sconv_pad(t, 7, 1).t_out;                                         // final SConv1d(.., dimension, 7)
}

int64_t EncodecEngine::decoded_length(int64_t T) const {
  int64_t t = sconv_pad(T, 7, 1).t_out;
  for (size_t i = 0; i < cfg_.ratios.size(); ++i) {
    t *= cfg_.ratios[i];
    const SPad k3 = sconv_pad(t, 3, 1);
    if (k3.t_out != t) throw Error(NC_INVALID_ARGUMENT, "Encodec: too few frames: a residual block's branches would differ in length (the reference's tensor add fails)");
  }
  return sconv_pad(t, 7, 1).t_out;
}

// Encodec.Encode's loop (Encodec.cs:273-282): offsets 0, stride, ... while offset < length; a frame ends at
// min(offset + segment, length).  SegmentLength / SegmentStride: Encodec.cs:190-196.
EncodecEngine::SegLayout EncodecEngine::seg_layout(int64_t L) const {
  SegLayout lay;
  lay.seg = segmented() ? cfg_.segment_length() : L;
  lay.stride = segmented() ? cfg_.segment_stride() : L;
  lay.n_seg = (int)((L + lay.stride - 1) / lay.stride);
  for (int s = 0; s < lay.n_seg; ++s) {
    const int64_t len = std::min<int64_t>(lay.seg, L - (int64_t)s * lay.stride);
    lay.len.push_back(len);
    if (len == lay.seg && lay.n_full == s) ++lay.n_full;
    const int64_t T = frames(len);
    lay.col.push_back(lay.t_total);
    lay.frames.push_back(T);
    lay.t_total += T;
    lay.ld.push_back(decoded_length(T));
    lay.ld_max = std::max(lay.ld_max, lay.ld.back());
  }
  lay.total_out = lay.stride * (lay.n_seg - 1) + lay.ld.back();            // AudioTensorDSP.cs:180
  return lay;
}

EncodecEngine::SegLayout EncodecEngine::seg_layout_from_frames(const int64_t* seg_frames, int n_seg) const {
  if (n_seg < 1 || !seg_frames) throw Error(NC_INVALID_ARGUMENT, "No frames provided to decode");               // Encodec.cs:215-218
  if (!segmented() && n_seg != 1) throw Error(NC_INVALID_ARGUMENT, "Expected single frame when no segmentation is used");   // :222-225
  SegLayout lay;
  lay.n_seg = n_seg;
  lay.seg = segmented() ? cfg_.segment_length() : 0;
  lay.stride = segmented() ? cfg_.segment_stride() : 0;
  for (int s = 0; s < n_seg; ++s) {
    const int64_t T = seg_frames[s];
    if (T <= 0 || T > ((int64_t)1 << 24)) throw Error(NC_INVALID_ARGUMENT, "Invalid frame codes in Encodec Decode");
    if (T == seg_frames[0] && lay.n_full == s) ++lay.n_full;     // equal-length leading frames decode as one batch
    lay.len.push_back(0);
    lay.col.push_back(lay.t_total);
    lay.frames.push_back(T);
    lay.t_total += T;
    lay.ld.push_back(decoded_length(T));
    lay.ld_max = std::max(lay.ld_max, lay.ld.back());
  }
  lay.total_out = lay.stride * (n_seg - 1) + lay.ld.back();
  return lay;
}

std::vector<EncodecEngine::Group> EncodecEngine::groups_of(const SegLayout& lay) const {
  std::vector<Group> g;
  if (lay.n_full > 0) g.push_back({lay.n_full, 0, lay.len[0], lay.frames[0], 0});
  for (int s = lay.n_full; s < lay.n_seg; ++s) g.push_back({1, s, lay.len[s], lay.frames[s], lay.col[s]});
  return g;
}

// LinearOverlapAdd narrows the output at [s*stride, s*stride + len_s) for every frame (AudioTensorDSP.cs:214): a frame that
// sticks out beyond stride*(n-1) + len_last makes the reference throw; weights come from the FIRST frame's length (:184-189).
void EncodecEngine::check_overlap_add(const SegLayout& lay) const {
  for (int s = 0; s < lay.n_seg; ++s)
    if ((int64_t)s * lay.stride + lay.ld[s] > lay.total_out || lay.ld[s] > lay.ld[0])
      throw Error(NC_INVALID_ARGUMENT, "Encodec Decode: frame " + std::to_string(s) + " does not fit the overlap-add output (the last frame is too short; the reference's narrow() fails)");
}

void EncodecEngine::conv_short(const ConvLayer& L, const Act& in, const SPad& pad, const Act& out, int B, int prologue, int post) {
  const LaunchCtx c = ctx();
  const int Tp = in.T + pad.extra_zero + pad.left + pad.right;
  float* tmp = static_cast<float*>(pad_tmp_.reserve((size_t)B * Tp * in.C * sizeof(float)));
  launch_pad1d_dense(in.base, in.stride, in.T, in.C, pad.extra_zero, pad.left, pad.right, tmp, B, c);
  ConvRunArgs a;
  a.in = tmp; a.in_clip_stride = (long long)Tp * in.C; a.t_in = Tp;
  a.out = out.base; a.out_clip_stride = out.stride; a.batch = B;
  a.prologue = prologue;    // ELU(0) = 0 and ELU commutes with reflection: applying it inside the conv is Pad1d(ELU(x))
  a.post = post;
  gn_fused_ = false;
  if (cfg_.group_norm && gn_.count(L.name())) {
    NC_CUDA(cudaMemsetAsync(gn_slot(), 0, (size_t)B * 2 * sizeof(double), stream_));
    a.gn_stats = gn_slot();
    a.gn_stats_done = &gn_fused_;
  }
  L.run(a, c);
}

int EncodecEngine::n_q_for_bandwidth(float kbps) const {
  const int frame_rate = (int)std::ceil((float)cfg_.sample_rate / (float)cfg_.hop());      // Encodec.cs:86
  const double bw_per_q = std::log2((double)cfg_.codebook_size) * frame_rate;
  int nq = cfg_.n_quantizers;
  if (kbps > 0) nq = (int)std::max(1.0, std::floor((double)kbps * 1000.0 / bw_per_q));
  return std::min<int>(nq, (int)embed_.size());
}

int EncodecEngine::micro_batch(int B, int64_t L) {
  const int64_t T = frames(L);
  const int64_t Lfull = std::max<int64_t>(L, decoded_length(T)) + 16;
  const int64_t per_buf = (Lfull + 2 * kMargin + 8) * cfg_.n_filters * 2;   // widest layers: L x 32 and L/2 x 64
  const int top = cfg_.n_filters << cfg_.ratios.size();
  double per_clip = 5.0 * per_buf * 4 + (double)T * (4.0 * top + cfg_.dimension) * 4 + (double)Lfull * 4;
  int mb = (int)std::max(1.0, std::floor((double)max_workspace_bytes_ / per_clip));
  mb = std::min(mb, B);
  mb = std::max(mb, 1);   // (the LSTM splits a micro-batch into slices of at most lstm_max_batch clips itself: run_lstm)
  for (auto& w : ws_) w.reserve((size_t)mb * per_buf * sizeof(float));
  xproj_.reserve((size_t)mb * T * 4 * top * sizeof(float));
  z_.reserve((size_t)mb * T * cfg_.dimension * sizeof(float));
  hbuf_.reserve((size_t)2 * ((std::min(mb, cfg_.lstm_layers > 0 ? lstm_max_batch(num_sms_, top) : mb) + 31) / 32 * 32) * top * sizeof(float));
  barriers_.reserve(64 * sizeof(unsigned int));
  audio_tmp_.reserve((size_t)mb * decoded_length(T) * sizeof(float));
  gn_stats_.reserve((size_t)2 * mb * 2 * sizeof(double));   // two slots: the conv being normalised, and a resnet block's shortcut
  gn_mb_ = mb;
  return mb;
}

EncodecEngine::Act EncodecEngine::act(int bufi, int B, int T, int C) {
  Act a;
  a.T = T; a.C = C;
  a.stride = (long long)(T + 2 * kMargin) * C;
  a.base = ws_[bufi].as<float>() + (long long)kMargin * C;
  if ((size_t)B * a.stride * sizeof(float) > ws_[bufi].capacity()) throw Error(NC_INTERNAL, "Encodec workspace too small");
  return a;
}

// A conv over rows [-left_pad, T + extra) of `in` (the margins must already hold the padding) -> rows [0, ...) of out
void EncodecEngine::conv(const ConvLayer& L, const Act& in, int left_pad, int extra, const Act& out, int B, int prologue,
                         int post, const Act* residual) {
  ConvRunArgs a;
  a.in = in.base - (long long)left_pad * in.C;
  a.in_clip_stride = in.stride;
  a.t_in = in.T + left_pad + extra;
  a.out = out.base;
  a.out_clip_stride = out.stride;
  a.batch = B;
  a.prologue = prologue;
  a.post = post;
  if (residual) a.residual = residual->base;
  gn_fused_ = false;
  if (cfg_.group_norm && gn_.count(L.name())) {   // the tcgen05 epilogue accumulates the GroupNorm statistics
    NC_CUDA(cudaMemsetAsync(gn_slot(), 0, (size_t)B * 2 * sizeof(double), stream_));
    a.gn_stats = gn_slot();
    a.gn_stats_done = &gn_fused_;
  }
  L.run(a, ctx());
}

// time_group_norm: y = GroupNorm(conv(x)), then whatever followed the conv in the weight-norm graph (residual add, ELU of
// the next layer).  Two HBM passes over the conv's output: fp64 sums, then the in-place affine.
void EncodecEngine::finish_norm(const ConvLayer& L, const Act& y, int B, int post, const Act* residual, int row0, int rows,
                                const ConvLayer* raw_residual_of) {
  if (!cfg_.group_norm) return;
  const LaunchCtx c = ctx();
  const auto it = gn_.find(L.name());
  if (it == gn_.end()) throw Error(NC_INTERNAL, "no GroupNorm parameters for " + L.name());
  const Gn& g = it->second;
  double* st = gn_slot();
  if (!gn_fused_) launch_gn_stats(y.base + (long long)row0 * y.C, y.stride, (long long)rows * y.C, st, B, c);   // CUDA-core fallback convs
  gn_fused_ = false;
  const Gn* g2 = raw_residual_of ? &gn_.at(raw_residual_of->name()) : nullptr;   // residual = that conv's raw output, statistics in slot 1
  launch_gn_apply(y.base, y.stride, y.T, y.C, st, (double)rows * g.c_real, 1e-5f, g.gamma, g.beta, residual ? residual->base : nullptr,
                  residual ? residual->stride : 0, post == PRO_ELU ? 1 : 0, B, c, g2 ? gn_stats_.as<double>() + (size_t)2 * gn_mb_ : nullptr,
                  g2 ? (double)residual->T * g2->c_real : 0.0, g2 ? g2->gamma : nullptr, g2 ? g2->beta : nullptr);
}

void EncodecEngine::conv_n(const ConvLayer& L, const Act& in, int left_pad, int extra, const Act& out, int B, int prologue, int post,
                           const Act* residual, int out_row0, int stat_rows) {
  const bool gn = cfg_.group_norm;
  Act o = out;
  o.base += (long long)out_row0 * out.C;      // transposed convs: the rows trimmed on the left land in the margin
  conv(L, in, left_pad, extra, o, B, prologue, gn ? PRO_NONE : post, gn ? nullptr : residual);
  finish_norm(L, out, B, post, residual, out_row0, stat_rows < 0 ? out.T : stat_rows);
}

void EncodecEngine::conv_short_n(const ConvLayer& L, const Act& in, const SPad& pad, const Act& out, int B, int prologue, int post) {
  conv_short(L, in, pad, out, B, prologue, cfg_.group_norm ? PRO_NONE : post);
  finish_norm(L, out, B, post, nullptr, 0, out.T);
}

int EncodecEngine::pick_free(int a, int b, int c, int d) const {
  for (int i = 0; i < 5; ++i)
    if (i != a && i != b && i != c && i != d) return i;
  throw Error(NC_INTERNAL, "no free workspace buffer");
}

// SEANetResnetBlock (SEANetResnetBlock.cs:70-86): shortcut(x) + conv_k1(ELU(conv_k3(ELU(x)))); x lives in buffer xb.
// Returns ELU(y) when post_elu (the block is always followed by an ELU in the SEANet stacks).
EncodecEngine::Act EncodecEngine::run_res(const Res& r, const Act& x, int B, int& xb, int& sb, int& hb, bool post_elu) {
  const LaunchCtx c = ctx();
  sb = pick_free(xb, -1);
  hb = pick_free(xb, sb);
  const int yb = pick_free(xb, sb, hb);
  Act S = act(sb, B, x.T, x.C), H = act(hb, B, x.T, r.hidden_p), Y = act(yb, B, x.T, x.C);
  const bool gn = cfg_.group_norm;
  if (gn) {   // the shortcut stays RAW (statistics in slot 1): its GroupNorm is applied inside the block output's apply pass
    gn_slot_ = 1;
    conv(r.shortcut, x, 0, 0, S, B, PRO_NONE, PRO_NONE, nullptr);
    if (!gn_fused_) launch_gn_stats(S.base, S.stride, (long long)S.T * S.C, gn_slot(), B, c);
    gn_fused_ = false;
    gn_slot_ = 0;
  } else {
    conv(r.shortcut, x, 0, 0, S, B, PRO_NONE, PRO_NONE, nullptr);
  }
  const SPad k3 = sconv_pad(x.T, 3, 1);                                   // padding_total = 2: causal (2, 0), else (1, 1)
  launch_reflect_pad(x.base, x.T, x.C, x.stride, k3.left, k3.right, B, c);
  conv_n(r.c3, x, k3.left, k3.right, H, B, PRO_ELU, PRO_ELU, nullptr);
  if (gn) {
    conv(r.c1, H, 0, 0, Y, B, PRO_NONE, PRO_NONE, nullptr);
    finish_norm(r.c1, Y, B, post_elu ? PRO_ELU : PRO_NONE, &S, 0, Y.T, &r.shortcut);
  } else {
    conv(r.c1, H, 0, 0, Y, B, PRO_NONE, post_elu ? PRO_ELU : PRO_NONE, &S);
  }
  xb = yb;
  return Y;
}

// SLSTM (SLSTM.cs:40-57): 2-layer LSTM + skip; the ELU that follows it in both SEANet stacks is applied on output.
EncodecEngine::Act EncodecEngine::run_lstm(const Lstm& l, const Act& x, int B, int out_buf) {
  const LaunchCtx c = ctx();
  const int H = x.C;
  Act cur = x;
  // dense projection buffer [B][T][4H]
  for (int i = 0; i < l.layers; ++i) {
    ConvRunArgs a;
    a.in = cur.base; a.in_clip_stride = cur.stride; a.t_in = cur.T; a.batch = B;
    a.out = xproj_.as<float>(); a.out_clip_stride = (long long)cur.T * 4 * H;
    l.ih[i].run(a, c);
    const bool last = i + 1 == l.layers;
    // layer outputs go to out_buf for the last layer, else to a scratch activation in z_-independent buffer 4
    Act o = act(last ? out_buf : 4, B, x.T, H);
    // one cooperative launch takes at most lstm_max_batch clips (all its CTAs must be co-resident); the convs around it run on
    // the whole micro-batch, so a batch of many short segments (48 kHz: 11 per 10 s clip) is not cut into small conv launches
    const int lmax = lstm_max_batch(num_sms_, H);
    const long long xstride = (long long)cur.T * 4 * H;
    for (int b0 = 0; b0 < B; b0 += lmax) {
      const int nb = std::min(lmax, B - b0);
      launch_lstm_layer(xproj_.as<float>() + (long long)b0 * xstride, xstride, l.whh[i], hbuf_.as<float>(), o.base + (long long)b0 * o.stride,
                        o.stride, last ? x.base + (long long)b0 * x.stride : nullptr, x.stride, last ? 1 : 0, barriers_.as<unsigned int>(),
                        nb, x.T, H, c);
    }
    cur = o;
  }
  return cur;
}

// audio: [B][L] mono samples (24 kHz-style path); the generic path expects the caller to have written the prepared,
// channel-padded segment into act(1, B, L, cin_pad_) (launch_encodec_segment_prep).
void EncodecEngine::run_encoder(const float* audio, int B, int64_t L, int64_t* T_out) {
  const LaunchCtx c = ctx();
  const int nf = cfg_.n_filters;
  int xb = 0, sb = -1, hb = -1;
  const SPad p_in = sconv_pad(L, 7, 1);
  Act x = act(xb, B, p_in.t_out, nf);
  if (generic_io_) {
    Act in = act(1, B, (int)L, cin_pad_);
    if (p_in.extra_zero == 0) {
      launch_reflect_pad(in.base, in.T, in.C, in.stride, p_in.left, p_in.right, B, c);
      conv_n(conv_in_l_, in, p_in.left, p_in.right, x, B, PRO_NONE, PRO_NONE, nullptr);
    } else {
      conv_short_n(conv_in_l_, in, p_in, x, B, PRO_NONE, PRO_NONE);
    }
  } else if (p_in.extra_zero == 0) {
    // SConv1d(1, 32, 7) causal: left reflect pad 6 handled by index reflection inside the Cin = 1 kernel
    launch_conv_cin1(audio, L, (int)L, x.base, (int)L, nf, d_conv_in_w_, d_conv_in_b_, 7, 1, 6, B, c, /*reflect=*/1, x.stride);
  } else {
    // short-input branch of Pad1d (L <= 6): zero-extend, reflect, then a valid convolution over the padded samples
    const int Tp = (int)L + p_in.extra_zero + p_in.left + p_in.right;
    float* tmp = static_cast<float*>(pad_tmp_.reserve((size_t)B * Tp * sizeof(float)));
    launch_pad1d_dense(audio, L, (int)L, 1, p_in.extra_zero, p_in.left, p_in.right, tmp, B, c);
    launch_conv_cin1(tmp, Tp, Tp, x.base, p_in.t_out, nf, d_conv_in_w_, d_conv_in_b_, 7, 1, 0, B, c, /*reflect=*/0, x.stride);
  }
  for (size_t i = 0; i < enc_res_.size(); ++i) {
    const int r = cfg_.ratios[cfg_.ratios.size() - 1 - i];
    Act y = run_res(*enc_res_[i], x, B, xb, sb, hb, /*post_elu=*/true);
    const SPad pd = sconv_pad(y.T, 2 * r, r);                            // padding_total = k - stride = r, + the alignment extra
    const int ob = pick_free(xb, -1);
    Act o = act(ob, B, pd.t_out, 2 * y.C);
    if (pd.extra_zero == 0) {
      launch_reflect_pad(y.base, y.T, y.C, y.stride, pd.left, pd.right, B, c);
      conv_n(*enc_down_[i], y, pd.left, pd.right, o, B, PRO_NONE, PRO_NONE, nullptr);
    } else {
      conv_short_n(*enc_down_[i], y, pd, o, B, PRO_NONE, PRO_NONE);      // y.T <= the pad: Pad1d's short-input branch
    }
    x = o;
    xb = ob;
  }
  Act top = x;
  int prologue = PRO_ELU;
  if (cfg_.lstm_layers > 0) {
    const int ob = pick_free(xb, 4);
    top = run_lstm(enc_lstm_, x, B, ob);   // = ELU(lstm(x) + x)
    xb = ob;
    prologue = PRO_NONE;
  }
  const SPad pf = sconv_pad(top.T, 7, 1);
  Act zo;
  zo.base = z_.as<float>(); zo.T = pf.t_out; zo.C = cfg_.dimension; zo.stride = (long long)pf.t_out * cfg_.dimension;
  if (pf.extra_zero == 0) {
    launch_reflect_pad(top.base, top.T, top.C, top.stride, pf.left, pf.right, B, c);
    conv_n(enc_out_, top, pf.left, pf.right, zo, B, prologue, PRO_NONE, nullptr);
  } else {                                                               // fewer frames than the pad: the final conv lengthens z
    conv_short_n(enc_out_, top, pf, zo, B, prologue, PRO_NONE);
  }
  *T_out = pf.t_out;
}

void EncodecEngine::run_decoder(int B, int T, float* audio_out, long long out_stride) {
  const LaunchCtx c = ctx();
  // zq activation was written into buffer 0 by the caller (with margins)
  int xb = 0, sb = -1, hb = -1;
  Act z = act(0, B, T, cfg_.dimension);
  const SPad pz = sconv_pad(T, 7, 1);
  const int top = cfg_.n_filters << cfg_.ratios.size();
  int ob = 1;
  Act x = act(ob, B, pz.t_out, top);
  const bool has_lstm = cfg_.lstm_layers > 0;
  if (pz.extra_zero == 0) {
    launch_reflect_pad(z.base, z.T, z.C, z.stride, pz.left, pz.right, B, c);
    conv_n(dec_in_, z, pz.left, pz.right, x, B, PRO_NONE, has_lstm ? PRO_NONE : PRO_ELU, nullptr);
  } else {
    conv_short_n(dec_in_, z, pz, x, B, PRO_NONE, has_lstm ? PRO_NONE : PRO_ELU);   // T <= the pad: short-input branch
  }
  xb = ob;
  if (has_lstm) {
    const int lb = pick_free(xb, 4);
    x = run_lstm(dec_lstm_, x, B, lb);     // = ELU(lstm(x) + x)
    xb = lb;
  }
  for (size_t i = 0; i < dec_up_.size(); ++i) {
    const int r = cfg_.ratios[i];
    const int ub = pick_free(xb, -1);
    // SConvTranspose1d (SConvTranspose1d.cs:116-139): conv_transpose1d yields (T+1)*r rows, GroupNorm (if any) sees all of
    // them, then padding_total = r rows are trimmed: all on the right (causal) or r - r/2 left, r/2 right.  The trimmed rows
    // land in the margins of u.
    const int trim_right = cfg_.causal ? r : r / 2, trim_left = r - trim_right;
    Act u = act(ub, B, x.T * r, x.C / 2);
    conv_n(*dec_up_[i], x, 0, 0, u, B, PRO_NONE, PRO_NONE, nullptr, -trim_left, (x.T + 1) * r);
    xb = ub;
    x = run_res(*dec_res_[i], u, B, xb, sb, hb, /*post_elu=*/true);
  }
  if (!generic_io_) {
    // SConv1d(32, 1, 7) causal: reflect handled by index reflection in the Cout = 1 kernel
    launch_conv_cout1(x.base, audio_out, x.T, conv_out_c_, d_conv_out_w_, d_conv_out_b_, 7, 6, 0, B, c, /*reflect=*/1, x.stride);
    (void)out_stride;
    return;
  }
  // SConv1d(n_filters, channels, 7) on the padded-N layer, then GroupNorm over the real channels, * scale (Encodec.cs:448-451)
  // and the planar frame for the overlap-add
  const SPad po = sconv_pad(x.T, 7, 1);
  const int rb = pick_free(xb, -1);
  Act raw = act(rb, B, po.t_out, cout_pad_);
  if (po.extra_zero == 0) {
    launch_reflect_pad(x.base, x.T, x.C, x.stride, po.left, po.right, B, c);
    conv(conv_out_l_, x, po.left, po.right, raw, B, PRO_NONE, PRO_NONE, nullptr);
  } else {
    conv_short(conv_out_l_, x, po, raw, B, PRO_NONE, PRO_NONE);
  }
  const double* st = nullptr;
  const float *gamma = nullptr, *beta = nullptr;
  if (cfg_.group_norm) {
    const Gn& g = gn_.at(conv_out_l_.name());
    if (!gn_fused_) launch_gn_stats(raw.base, raw.stride, (long long)raw.T * raw.C, gn_slot(), B, c);
    gn_fused_ = false;
    st = gn_slot(); gamma = g.gamma; beta = g.beta;
  }
  launch_encodec_frame_out(raw.base, raw.stride, raw.T, cout_pad_, cfg_.channels, st, 1e-5f, gamma, beta, map_.segs, map_.s0,
                           map_.item0, map_.n_seg, map_.scales, audio_out, out_stride, B, c);
}

// ------------------------------------------------------------------------------------ entry points
void EncodecEngine::encode_dev(const float* audio, int B, int64_t L, int nq, int64_t* codes) {
  forward_dev(audio, B, L, nq, nullptr, codes);
}

void EncodecEngine::forward_dev(const float* audio, int B, int64_t L, int nq, float* audio_out, int64_t* codes) {
  require_ready();
  bind();
  if (!simple()) { forward_frames_dev(audio, B, L, nq, audio_out, codes, nullptr); return; }
  if (B <= 0 || L <= 0) throw Error(NC_INVALID_ARGUMENT, "batch and length must be positive");
  if (nq <= 0 || nq > (int)embed_.size()) throw Error(NC_INVALID_ARGUMENT, "n_quantizers out of range");
  if (L > (int64_t)1 << 28) throw Error(NC_INVALID_ARGUMENT, "clip too long");
  const int mb = micro_batch(B, L);
  const int64_t T = frames(L);
  const LaunchCtx c = ctx();
  if (!codes) codes_tmp_.reserve((size_t)mb * nq * T * sizeof(int64_t));
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int nb = std::min(mb, B - b0);
    int64_t Tm = 0;
    run_encoder(audio + (int64_t)b0 * L, nb, L, &Tm);
    int64_t* cdst = codes ? codes + (int64_t)b0 * nq * T : codes_tmp_.as<int64_t>();
    for (int q = 0; q < nq; ++q)
      launch_encodec_vq_stage(z_.as<float>(), (long long)nb * T, embed_[q], embed_sq_[q], cfg_.codebook_size, cfg_.dimension,
                              cdst, (int)T, nq, q, c);
    if (audio_out) {
      Act z = act(0, nb, (int)T, cfg_.dimension);
      launch_encodec_decode_codes(cdst, d_embed_ptrs_, z.base, z.stride, nb, (int)T, nq, cfg_.codebook_size, cfg_.dimension, c);
      const int64_t Ld = decoded_length(T);
      run_decoder(nb, (int)T, audio_tmp_.as<float>(), Ld);
      launch_trim_rows(audio_tmp_.as<float>(), audio_out + (int64_t)b0 * L, nb, Ld, std::min<int64_t>(L, Ld), c);
    }
  }
  sync();
}

void EncodecEngine::decode_dev(const int64_t* codes, int B, int nq, int64_t T, float* audio_out) {
  require_ready();
  bind();
  if (!simple()) { decode_frames_dev(codes, nullptr, B, nq, &T, 1, audio_out); return; }
  if (B <= 0 || T <= 0 || !codes) throw Error(NC_INVALID_ARGUMENT, "Invalid frame codes in Encodec Decode");
  if (nq <= 0 || nq > (int)embed_.size()) throw Error(NC_INVALID_ARGUMENT, "n_quantizers out of range");
  const int64_t L = decoded_length(T);
  const int mb = micro_batch(B, L);
  const LaunchCtx c = ctx();
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int nb = std::min(mb, B - b0);
    Act z = act(0, nb, (int)T, cfg_.dimension);
    launch_encodec_decode_codes(codes + (int64_t)b0 * nq * T, d_embed_ptrs_, z.base, z.stride, nb, (int)T, nq,
                                cfg_.codebook_size, cfg_.dimension, c);
    run_decoder(nb, (int)T, audio_out + (int64_t)b0 * L, L);
  }
  sync();
}

// One group of equal-length segments, micro-batched over its B*segs items: prep (loudness scale, channels-last) -> encoder
// -> RVQ -> codes into the caller's layout; and / or codes -> decoder -> scaled planar frames for the overlap-add.
void EncodecEngine::run_group(const float* audio, int B, int64_t L, const SegLayout& lay, const Group& g, int nq, int64_t* codes_user,
                              float* scales, bool encode, bool decode) {
  const LaunchCtx c = ctx();
  const int items = B * g.segs;
  const int64_t T = g.frames;
  const int64_t Ld = decoded_length(T);
  const int mb = micro_batch(items, encode ? g.len : Ld);
  codes_tmp_.reserve((size_t)mb * nq * T * sizeof(int64_t));
  int64_t* dense = codes_tmp_.as<int64_t>();
  for (int i0 = 0; i0 < items; i0 += mb) {
    const int nb = std::min(mb, items - i0);
    if (encode) {
      Act in = act(1, nb, (int)g.len, cin_pad_);
      launch_encodec_segment_prep(audio, cfg_.channels, L, g.segs, g.s0, lay.stride, (int)g.len, i0, lay.n_seg,
                                  cfg_.normalize ? scales : nullptr, in.base, in.stride, cin_pad_, nb, c);
      int64_t Tm = 0;
      run_encoder(nullptr, nb, g.len, &Tm);
      if (Tm != T) throw Error(NC_INTERNAL, "Encodec: frame count mismatch");
      for (int q = 0; q < nq; ++q)
        launch_encodec_vq_stage(z_.as<float>(), (long long)nb * T, embed_[q], embed_sq_[q], cfg_.codebook_size, cfg_.dimension,
                                dense, (int)T, nq, q, c);
      if (codes_user) launch_encodec_codes_segment_copy(dense, codes_user, g.segs, i0, nq, (int)T, lay.t_total, g.col0, 1, nb, c);
    } else {
      launch_encodec_codes_segment_copy(dense, codes_user, g.segs, i0, nq, (int)T, lay.t_total, g.col0, 0, nb, c);
    }
    if (decode) {
      Act z = act(0, nb, (int)T, cfg_.dimension);
      launch_encodec_decode_codes(dense, d_embed_ptrs_, z.base, z.stride, nb, (int)T, nq, cfg_.codebook_size, cfg_.dimension, c);
      map_.segs = g.segs; map_.s0 = g.s0; map_.item0 = i0; map_.n_seg = lay.n_seg;
      map_.scales = scales;            // frames without a scale decode unscaled (Encodec.cs:448-451)
      run_decoder(nb, (int)T, frames_.as<float>(), lay.ld_max);
    }
  }
}

void EncodecEngine::overlap_add(const SegLayout& lay, int B, float* audio_out, int64_t out_len) {
  std::vector<int> lens(lay.ld.begin(), lay.ld.end());
  int* d = static_cast<int*>(seg_lens_.reserve(lens.size() * sizeof(int)));
  NC_CUDA(cudaMemcpyAsync(d, lens.data(), lens.size() * sizeof(int), cudaMemcpyHostToDevice, stream_));
  NC_CUDA(cudaStreamSynchronize(stream_));   // `lens` is a stack vector
  launch_encodec_overlap_add(frames_.as<float>(), B, lay.n_seg, cfg_.channels, lay.ld_max, d, (int)lay.ld_max, lay.stride,
                             segmented() ? 1 : 0, audio_out, out_len, ctx());
}

void EncodecEngine::forward_frames_dev(const float* audio, int B, int64_t L, int nq, float* audio_out, int64_t* codes, float* scales) {
  require_ready();
  bind();
  if (simple()) throw Error(NC_INTERNAL, "forward_frames_dev on the single-frame path");
  if (B <= 0 || L <= 0 || !audio) throw Error(NC_INVALID_ARGUMENT, "batch and length must be positive");
  if (nq <= 0 || nq > (int)embed_.size()) throw Error(NC_INVALID_ARGUMENT, "n_quantizers out of range");
  if (L > (int64_t)1 << 28) throw Error(NC_INVALID_ARGUMENT, "clip too long");
  const SegLayout lay = seg_layout(L);
  if (audio_out) check_overlap_add(lay);
  if (!cfg_.normalize) scales = nullptr;
  else if (!scales) scales = static_cast<float*>(scales_.reserve((size_t)B * lay.n_seg * sizeof(float)));
  if (audio_out) frames_.reserve((size_t)B * lay.n_seg * cfg_.channels * lay.ld_max * sizeof(float));
  for (const Group& g : groups_of(lay)) run_group(audio, B, L, lay, g, nq, codes, scales, /*encode=*/true, /*decode=*/audio_out != nullptr);
  if (audio_out) overlap_add(lay, B, audio_out, std::min<int64_t>(L, lay.total_out));   // Encodec.cs:295: sliced to the input length
  sync();
}

void EncodecEngine::decode_frames_dev(const int64_t* codes, const float* scales, int B, int nq, const int64_t* seg_frames, int n_seg,
                                      float* audio_out) {
  require_ready();
  bind();
  if (simple()) {
    if (n_seg != 1 || !seg_frames) throw Error(NC_INVALID_ARGUMENT, "Expected single frame when no segmentation is used");
    decode_dev(codes, B, nq, seg_frames[0], audio_out);
    return;
  }
  if (B <= 0 || !codes || !audio_out) throw Error(NC_INVALID_ARGUMENT, "Invalid frame codes in Encodec Decode");
  if (nq <= 0 || nq > (int)embed_.size()) throw Error(NC_INVALID_ARGUMENT, "n_quantizers out of range");
  const SegLayout lay = seg_layout_from_frames(seg_frames, n_seg);
  check_overlap_add(lay);
  frames_.reserve((size_t)B * lay.n_seg * cfg_.channels * lay.ld_max * sizeof(float));
  for (const Group& g : groups_of(lay))
    run_group(nullptr, B, 0, lay, g, nq, const_cast<int64_t*>(codes), const_cast<float*>(scales), /*encode=*/false, /*decode=*/true);
  overlap_add(lay, B, audio_out, lay.total_out);
  sync();
}

int EncodecEngine::bits_per_codebook() const {
  int bits = 0;
  while ((1 << bits) < cfg_.codebook_size) ++bits;
  if ((1 << bits) != cfg_.codebook_size) throw Error(NC_UNSUPPORTED, "Only codebooks with power-of-2 sizes are supported");  // Encodec.cs:130-133
  return bits;
}

void EncodecEngine::ecdc_pack_segment_dev(const int64_t* codes, int64_t t_total, int64_t col, int64_t T, int nq, uint8_t* bytes,
                                          int64_t stride, int B) {
  bind();
  launch_ecdc_pack(codes + col, bytes, stride, B, (int)T, nq, bits_per_codebook(), ctx(), t_total, (long long)nq * t_total);
  sync();
}

void EncodecEngine::ecdc_unpack_segment_dev(const uint8_t* bytes, int64_t stride, int64_t* codes, int64_t t_total, int64_t col,
                                            int64_t T, int nq, int B) {
  bind();
  launch_ecdc_unpack(bytes, stride, codes + col, B, (int)T, nq, bits_per_codebook(), ctx(), t_total, (long long)nq * t_total);
  sync();
}

void EncodecEngine::compress_dev(const float* audio, int B, int64_t L, int nq, uint8_t* payload, int64_t stride) {
  require_ready();
  bind();
  if (!simple()) throw Error(NC_INTERNAL, "compress_dev is the single-frame path");
  if (B <= 0 || L <= 0 || !audio) throw Error(NC_INVALID_ARGUMENT, "batch and length must be positive");
  if (nq <= 0 || nq > (int)embed_.size()) throw Error(NC_INVALID_ARGUMENT, "n_quantizers out of range");
  if (L > (int64_t)1 << 28) throw Error(NC_INVALID_ARGUMENT, "clip too long");
  const int64_t T = frames(L);
  if (!payload || stride < ecdc_payload_bytes(nq, T)) throw Error(NC_INVALID_ARGUMENT, "ecdc payload buffer too small");
  codes_tmp_.reserve((size_t)B * nq * T * sizeof(int64_t));
  int64_t* codes = codes_tmp_.as<int64_t>();
  // encode in micro-batches straight into the [B][nq][T] code tensor, then one pack launch for the whole batch
  const int mb = micro_batch(B, L);
  const LaunchCtx c = ctx();
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int nb = std::min(mb, B - b0);
    int64_t Tm = 0;
    run_encoder(audio + (int64_t)b0 * L, nb, L, &Tm);
    for (int q = 0; q < nq; ++q)
      launch_encodec_vq_stage(z_.as<float>(), (long long)nb * T, embed_[q], embed_sq_[q], cfg_.codebook_size, cfg_.dimension,
                              codes + (int64_t)b0 * nq * T, (int)T, nq, q, c);
  }
  launch_ecdc_pack(codes, payload, stride, B, (int)T, nq, bits_per_codebook(), c);
  sync();
}

void EncodecEngine::decompress_dev(const uint8_t* payload, int64_t stride, int B, int nq, int64_t L, float* audio_out) {
  require_ready();
  bind();
  if (!simple()) throw Error(NC_INTERNAL, "decompress_dev is the single-frame path");
  if (B <= 0 || L <= 0 || !payload || !audio_out) throw Error(NC_INVALID_ARGUMENT, "Invalid ecdc payload");
  if (nq <= 0 || nq > (int)embed_.size()) throw Error(NC_INVALID_ARGUMENT, "n_quantizers out of range");
  // frameLength = ceil(L * frame_rate / sample_rate) (EncodecCompressor.cs:296-297), frame_rate = ceil(sr / hop)
  const int64_t frame_rate = (cfg_.sample_rate + cfg_.hop() - 1) / cfg_.hop();
  const int64_t T = (L * frame_rate + cfg_.sample_rate - 1) / cfg_.sample_rate;
  if (stride < ecdc_payload_bytes(nq, T)) throw Error(NC_INVALID_ARGUMENT, "Stream ended too soon");   // EncodecCompressor.cs:390-393
  const int64_t Ld = decoded_length(T);
  const int mb = micro_batch(B, Ld);
  const LaunchCtx c = ctx();
  codes_tmp_.reserve((size_t)B * nq * T * sizeof(int64_t));
  int64_t* codes = codes_tmp_.as<int64_t>();
  launch_ecdc_unpack(payload, stride, codes, B, (int)T, nq, bits_per_codebook(), c);
  audio_tmp_.reserve((size_t)mb * Ld * sizeof(float));
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int nb = std::min(mb, B - b0);
    Act z = act(0, nb, (int)T, cfg_.dimension);
    launch_encodec_decode_codes(codes + (int64_t)b0 * nq * T, d_embed_ptrs_, z.base, z.stride, nb, (int)T, nq, cfg_.codebook_size,
                                cfg_.dimension, c);
    run_decoder(nb, (int)T, audio_tmp_.as<float>(), Ld);
    launch_trim_rows(audio_tmp_.as<float>(), audio_out + (int64_t)b0 * L, nb, Ld, std::min<int64_t>(L, Ld), c);   // :412-415
  }
  sync();
}

}  // namespace nc
