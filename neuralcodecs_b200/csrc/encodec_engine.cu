// Encodec engine (24 kHz mono, causal, weight-norm preset).  Graph (paths under /root/reference/NeuralCodecs.Torch/):
//   Models/Encodec.cs:213-296,436-489; Modules/Encodec/SEANetEncoder.cs:37-148, SEANetDecoder.cs:40-153,
//   SEANetResnetBlock.cs:30-86, SConv1d.cs:144-173,245-274, SConvTranspose1d.cs:116-139, SLSTM.cs:40-57,
//   ResidualVectorQuantizer.cs:107-157, EuclideanCodebook.cs:155-182.
// Activations are channels-last with 8 margin rows around every clip: the causal left reflect padding (and the
// right "extra" padding) of SConv1d is materialised in those rows by a tiny fix-up kernel, after which every conv
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
  if (cfg_.channels != 1) throw Error(NC_UNSUPPORTED, "Encodec: only the mono preset is built (channels = 1)");
  if (!cfg_.causal) throw Error(NC_UNSUPPORTED, "Encodec: only causal convolutions are built (24 kHz preset)");
  if (cfg_.n_residual_layers != 1) throw Error(NC_UNSUPPORTED, "Encodec: n_residual_layers must be 1");
  if (cfg_.lstm_layers != 0 && cfg_.lstm_layers != 2) throw Error(NC_UNSUPPORTED, "Encodec: lstm_layers must be 0 or 2");
  if (cfg_.dimension != 128) throw Error(NC_UNSUPPORTED, "Encodec: dimension (codebook dim) must be 128");
  if (cfg_.n_filters % 32 != 0) throw Error(NC_UNSUPPORTED, "Encodec: n_filters must be a multiple of 32");
  for (int r : cfg_.ratios)
    if (r < 1 || r > kMargin) throw Error(NC_UNSUPPORTED, "Encodec: ratios must be in 1..8");
  if (cfg_.n_quantizers < 1 || cfg_.codebook_size < 1) throw Error(NC_INVALID_ARGUMENT, "Encodec config: non-positive field");
}

EncodecEngine::~EncodecEngine() {
  cudaSetDevice(device_);
  cudaFree(d_conv_in_w_); cudaFree(d_conv_in_b_); cudaFree(d_conv_out_w_); cudaFree(d_conv_out_b_);
  for (float* p : embed_) cudaFree(p);
  for (float* p : embed_sq_) cudaFree(p);
  cudaFree(d_embed_ptrs_);
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
  for (size_t i = 0; i < enc_res_.size(); ++i) { add(enc_res_[i]->shortcut); add(enc_res_[i]->c3); add(enc_res_[i]->c1); add(*enc_down_[i]); }
  add(enc_lstm_.ih[0]); add(enc_lstm_.ih[1]); add(enc_out_); add(dec_in_); add(dec_lstm_.ih[0]); add(dec_lstm_.ih[1]);
  for (size_t i = 0; i < dec_res_.size(); ++i) { add(*dec_up_[i]); add(dec_res_[i]->shortcut); add(dec_res_[i]->c3); add(dec_res_[i]->c1); }
  s += "}}";
  return s;
}

// w = (v / ||v||_(1,2)) * (g - 1e-7)   (Modules/Encodec/WNConv1d.cs:113-127, WNConvTranspose1d.cs:124-156)
std::vector<float> EncodecEngine::folded(const std::string& p, int d0, int d1, int k, std::vector<float>* bias, int bias_n) {
  const HostTensor& v = tensor(p + ".conv.weight_v");
  const HostTensor& g = tensor(p + ".conv.weight_g");
  if (v.is_int || v.shape.size() != 3 || v.shape[0] != d0 || v.shape[1] != d1 || v.shape[2] != k)
    throw Error(NC_SHAPE_MISMATCH, "Failed to load Encodec weights: '" + p + "' has the wrong shape");
  if (g.is_int || (int64_t)g.numel() != d0)
    throw Error(NC_SHAPE_MISMATCH, "Failed to load Encodec weights: '" + p + ".conv.weight_g' has the wrong shape");
  std::vector<float> w(v.f32.size());
  const size_t inner = (size_t)d1 * k;
  for (int i = 0; i < d0; ++i) {
    double ss = 0;
    for (size_t j = 0; j < inner; ++j) ss += (double)v.f32[i * inner + j] * v.f32[i * inner + j];
    const float norm = std::sqrt((float)ss), gi = g.f32[i] - 1e-7f;
    for (size_t j = 0; j < inner; ++j) w[i * inner + j] = (v.f32[i * inner + j] / norm) * gi;
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
  {
    auto w = folded("encoder.layers.0", nf, 1, 7, &b, nf);
    cudaFree(d_conv_in_w_); cudaFree(d_conv_in_b_);
    d_conv_in_w_ = upload(w);
    d_conv_in_b_ = upload(pad1e(b, nf, nf));
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
// SConv1d.forward's padding for a causal conv (SConv1d.cs:144-173): padding_total = k - s on the left, the stride
// alignment extra on the right (computed with a float32 division, :245-250), and -- when the input is not longer than
// the larger pad -- Pad1d's short-input branch: zero-extend on the right first (:258-272).  The zero extension is not
// trimmed afterwards, so such a layer lengthens the sequence.
EncodecEngine::SPad EncodecEngine::sconv_pad(int64_t T, int k, int s) {
  const int pt = k - s;
  const float n_frames = ((float)(T - k + pt)) / (float)s + 1.0f;
  const int64_t ideal = ((int64_t)std::ceil(n_frames) - 1) * s + (k - pt);
  SPad p;
  p.left = pt;
  p.right = (int)(ideal - T);
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

int64_t EncodecEngine::decoded_length(int64_t T) const { return (int64_t)sconv_pad(T, 7, 1).t_out * cfg_.hop(); }

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
  if (cfg_.lstm_layers > 0) mb = std::min(mb, lstm_max_batch(num_sms_, top));
  mb = std::max(mb, 1);
  for (auto& w : ws_) w.reserve((size_t)mb * per_buf * sizeof(float));
  xproj_.reserve((size_t)mb * T * 4 * top * sizeof(float));
  z_.reserve((size_t)mb * T * cfg_.dimension * sizeof(float));
  hbuf_.reserve((size_t)2 * ((mb + 31) / 32 * 32) * top * sizeof(float));
  barriers_.reserve(64 * sizeof(unsigned int));
  audio_tmp_.reserve((size_t)mb * decoded_length(T) * sizeof(float));
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
  L.run(a, ctx());
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
  conv(r.shortcut, x, 0, 0, S, B, PRO_NONE, PRO_NONE, nullptr);
  launch_reflect_pad(x.base, x.T, x.C, x.stride, 2, 0, B, c);            // k3: padding_total = 2, causal -> left
  conv(r.c3, x, 2, 0, H, B, PRO_ELU, PRO_ELU, nullptr);
  conv(r.c1, H, 0, 0, Y, B, PRO_NONE, post_elu ? PRO_ELU : PRO_NONE, &S);
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
    launch_lstm_layer(xproj_.as<float>(), (long long)cur.T * 4 * H, l.whh[i], hbuf_.as<float>(), o.base, o.stride,
                      last ? x.base : nullptr, x.stride, last ? 1 : 0, barriers_.as<unsigned int>(), B, x.T, H, c);
    cur = o;
  }
  return cur;
}

void EncodecEngine::run_encoder(const float* audio, int B, int64_t L, int64_t* T_out) {
  const LaunchCtx c = ctx();
  const int nf = cfg_.n_filters;
  int xb = 0, sb = -1, hb = -1;
  const SPad p_in = sconv_pad(L, 7, 1);
  Act x = act(xb, B, p_in.t_out, nf);
  if (p_in.extra_zero == 0) {
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
    const SPad pd = sconv_pad(y.T, 2 * r, r);                            // padding_total = k - stride = r (left), extra (right)
    const int ob = pick_free(xb, -1);
    Act o = act(ob, B, pd.t_out, 2 * y.C);
    if (pd.extra_zero == 0) {
      launch_reflect_pad(y.base, y.T, y.C, y.stride, pd.left, pd.right, B, c);
      conv(*enc_down_[i], y, pd.left, pd.right, o, B, PRO_NONE, PRO_NONE, nullptr);
    } else {
      conv_short(*enc_down_[i], y, pd, o, B, PRO_NONE, PRO_NONE);        // y.T <= r: Pad1d's short-input branch
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
  if (pf.extra_zero == 0) {
    launch_reflect_pad(top.base, top.T, top.C, top.stride, 6, 0, B, c);
    ConvRunArgs a;
    a.in = top.base - (long long)6 * top.C; a.in_clip_stride = top.stride; a.t_in = top.T + 6; a.batch = B;
    a.out = z_.as<float>(); a.out_clip_stride = (long long)top.T * cfg_.dimension;
    a.prologue = prologue;
    enc_out_.run(a, c);
  } else {                                                               // fewer than 7 frames: the final conv lengthens z
    Act zo;
    zo.base = z_.as<float>(); zo.T = pf.t_out; zo.C = cfg_.dimension; zo.stride = (long long)pf.t_out * cfg_.dimension;
    conv_short(enc_out_, top, pf, zo, B, prologue, PRO_NONE);
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
    launch_reflect_pad(z.base, z.T, z.C, z.stride, 6, 0, B, c);
    conv(dec_in_, z, 6, 0, x, B, PRO_NONE, has_lstm ? PRO_NONE : PRO_ELU, nullptr);
  } else {
    conv_short(dec_in_, z, pz, x, B, PRO_NONE, has_lstm ? PRO_NONE : PRO_ELU);   // T <= 6 frames: short-input branch
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
    Act u = act(ub, B, x.T * r, x.C / 2);   // conv_transpose1d yields (T+1)*r rows; the last r land in the margin = trimmed
    conv(*dec_up_[i], x, 0, 0, u, B, PRO_NONE, PRO_NONE, nullptr);
    xb = ub;
    x = run_res(*dec_res_[i], u, B, xb, sb, hb, /*post_elu=*/true);
  }
  // SConv1d(32, 1, 7) causal: reflect handled by index reflection in the Cout = 1 kernel
  launch_conv_cout1(x.base, audio_out, x.T, conv_out_c_, d_conv_out_w_, d_conv_out_b_, 7, 6, 0, B, c, /*reflect=*/1, x.stride);
  (void)out_stride;
}

// ------------------------------------------------------------------------------------ entry points
void EncodecEngine::encode_dev(const float* audio, int B, int64_t L, int nq, int64_t* codes) {
  forward_dev(audio, B, L, nq, nullptr, codes);
}

void EncodecEngine::forward_dev(const float* audio, int B, int64_t L, int nq, float* audio_out, int64_t* codes) {
  require_ready();
  bind();
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

int EncodecEngine::bits_per_codebook() const {
  int bits = 0;
  while ((1 << bits) < cfg_.codebook_size) ++bits;
  if ((1 << bits) != cfg_.codebook_size) throw Error(NC_UNSUPPORTED, "Only codebooks with power-of-2 sizes are supported");  // Encodec.cs:130-133
  return bits;
}

void EncodecEngine::compress_dev(const float* audio, int B, int64_t L, int nq, uint8_t* payload, int64_t stride) {
  require_ready();
  bind();
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
