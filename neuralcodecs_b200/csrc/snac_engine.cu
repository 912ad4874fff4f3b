// SNAC engine.  Graph (paths under /root/reference/NeuralCodecs.Torch/):
//   Models/SNAC.cs:70-80 (Preprocess), :91-106 (forward), :129-150 (Encode), :157-192 (Decode)
//   Modules/SNAC/Encoder.cs:26-69, EncoderBlock.cs:27-55, ResidualUnit.cs:25-60, Decoder.cs:28-86,
//   DecoderBlock.cs:23-70, NoiseBlock.cs:23-45, VectorQuantizer.cs:82-157, ResidualVectorQuantizer.cs:69-131
// Channels-last activations; every channel count is padded to a multiple of 32 with zero weights so the
// dense layers run on the tcgen05 kernel (48 -> 64 for the 24 kHz preset's first block).  A ResidualUnit is
// two launches: depthwise k7 kernel (Snake1 prologue, Snake2 post) and the 1x1 GEMM (+ residual, + the
// Snake that follows the unit).  LocalMHA (32/44 kHz presets): LayerNorm kernel -> qkv GEMM -> rotary + windowed
// attention kernel (one warp per clip x window x head) -> out-projection GEMM with the residual in its epilogue.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "engine.h"

namespace nc {

DwConv::~DwConv() {
  cudaFree(w);
  cudaFree(b);
}

SnacEngine::LocalMha::~LocalMha() {
  cudaFree(ln_w);
  cudaFree(ln_b);
  cudaFree(inv_freq);
}

// LocalMHA weights: norm.{weight,bias}, to_qkv.weight [3C, C], to_out.weight [C, C] (no biases), rel_pos.inv_freq
void SnacEngine::build_mha(LocalMha& m, const std::string& p, int dim) {
  m.present = true;
  m.dim = dim;
  m.heads = dim / 64;
  const HostTensor& g = tensor(p + ".norm.weight");
  const HostTensor& b = tensor(p + ".norm.bias");
  const HostTensor& wq = tensor(p + ".to_qkv.weight");
  const HostTensor& wo = tensor(p + ".to_out.weight");
  if ((int)g.numel() != dim || (int)b.numel() != dim || (int64_t)wq.numel() != (int64_t)3 * dim * dim ||
      (int64_t)wo.numel() != (int64_t)dim * dim)
    throw Error(NC_SHAPE_MISMATCH, "Failed to load SNAC weights: '" + p + "' has the wrong shape");
  cudaFree(m.ln_w); cudaFree(m.ln_b); cudaFree(m.inv_freq);
  m.ln_w = upload(g.f32);
  m.ln_b = upload(b.f32);
  std::vector<float> inv(32);
  if (has_tensor(p + ".rel_pos.inv_freq") && tensor(p + ".rel_pos.inv_freq").numel() == 32) {
    inv = tensor(p + ".rel_pos.inv_freq").f32;
  } else {
    for (int i = 0; i < 32; ++i) inv[i] = 1.0f / std::pow(10000.0f, (float)(2 * i) / 64.0f);   // SinusoidalEmbedding.cs:37-40
  }
  m.inv_freq = upload(inv);
  ConvSpec sq;
  sq.cin = dim; sq.cout = 3 * dim; sq.k = 1;
  m.qkv.build(p + ".to_qkv", sq, wq.f32, std::vector<float>(), prec_);
  ConvSpec so;
  so.cin = so.cout = dim; so.k = 1;
  m.out.build(p + ".to_out", so, wo.f32, std::vector<float>(), prec_);
}

int SnacEngine::run_mha(const LocalMha& m, int cur, int B, int T, const SnakeParams* post) {
  const LaunchCtx c = ctx();
  const int a = (cur + 1) % 3, y = (cur + 2) % 3;
  launch_layernorm_rows(buf(cur), buf(a), m.ln_w, m.ln_b, (long long)B * T, m.dim, c);
  float* qkv = static_cast<float*>(qkv_buf_.reserve((size_t)B * T * 3 * m.dim * sizeof(float)));
  ConvRunArgs q;
  q.in = buf(a); q.out = qkv; q.batch = B; q.t_in = T;
  m.qkv.run(q, c);
  launch_local_attn(qkv, buf(a), m.inv_freq, B, T, m.dim, m.heads, cfg_.attn_window, c);
  ConvRunArgs o;
  o.in = buf(a); o.out = buf(y); o.residual = buf(cur); o.batch = B; o.t_in = T;
  if (post) { o.post = PRO_SNAKE; o.post_alpha = post->alpha; o.post_inv_alpha = post->inv_alpha; }
  m.out.run(o, c);
  return y;
}

SnacEngine::SnacEngine(const nc_snac_config& c, int device_index) : Engine(device_index) {
  if (c.struct_size != sizeof(nc_snac_config)) throw Error(NC_INVALID_ARGUMENT, "nc_snac_config.struct_size mismatch");
  if (c.n_encoder_rates < 1 || c.n_encoder_rates > NC_MAX_RATES || c.n_decoder_rates < 1 || c.n_decoder_rates > NC_MAX_RATES ||
      c.n_vq_strides < 1 || c.n_vq_strides > NC_MAX_RATES)
    throw Error(NC_INVALID_ARGUMENT, "SNAC config: rate / stride count out of range");
  cfg_.sample_rate = c.sample_rate;
  cfg_.encoder_dim = c.encoder_dim;
  cfg_.encoder_rates.assign(c.encoder_rates, c.encoder_rates + c.n_encoder_rates);
  cfg_.decoder_dim = c.decoder_dim;
  cfg_.decoder_rates.assign(c.decoder_rates, c.decoder_rates + c.n_decoder_rates);
  cfg_.vq_strides.assign(c.vq_strides, c.vq_strides + c.n_vq_strides);
  cfg_.latent_dim = c.latent_dim > 0 ? c.latent_dim : c.encoder_dim * (1 << c.n_encoder_rates);  // Models/SNAC.cs:37
  cfg_.attn_window = c.attn_window_size;
  cfg_.codebook_size = c.codebook_size;
  cfg_.codebook_dim = c.codebook_dim;
  cfg_.noise = c.noise != 0;
  cfg_.depthwise = c.depthwise != 0;
  if (cfg_.sample_rate <= 0 || cfg_.encoder_dim <= 0 || cfg_.decoder_dim <= 0 || cfg_.codebook_size <= 0)
    throw Error(NC_INVALID_ARGUMENT, "SNAC config: non-positive field");
  if (cfg_.codebook_dim != 8) throw Error(NC_UNSUPPORTED, "SNAC: codebook_dim must be 8");
  if (cfg_.attn_window < 0 || cfg_.attn_window > 256)
    throw Error(NC_UNSUPPORTED, "SNAC: attn_window_size must be in 1..256 (32 in the reference presets)");
  if (cfg_.attn_window != 0 && (cfg_.latent_dim % 64 != 0 || cfg_.decoder_dim % 64 != 0))
    throw Error(NC_UNSUPPORTED, "SNAC: LocalMHA needs latent_dim and decoder_dim to be multiples of the 64-wide heads");
  for (size_t i = 0; i < cfg_.vq_strides.size(); ++i)
    if (cfg_.vq_strides[i] < 1 || cfg_.vq_strides[0] % cfg_.vq_strides[i] != 0)
      throw Error(NC_INVALID_ARGUMENT, "SNAC config: every vq stride must divide the first");
  if (cfg_.decoder_dim % (1 << cfg_.decoder_rates.size()) != 0)
    throw Error(NC_INVALID_ARGUMENT, "SNAC config: decoder_dim not divisible by 2^n_rates");
}

SnacEngine::~SnacEngine() {
  cudaSetDevice(device_);
  cudaFree(d_conv_in_w_);
  cudaFree(d_conv_in_b_);
  cudaFree(d_conv_out_w_);
  cudaFree(d_conv_out_b_);
  for (float* p : vq_alloc_) cudaFree(p);
}

void SnacEngine::set_option(const std::string& key, const std::string& value) {
  if (key == "precision" || key == "encoder_precision" || key == "decoder_precision") {
    prec_ = parse_precision(value);
    if (ready_) throw Error(NC_INVALID_ARGUMENT, "precision options must be set before weights are loaded");
  } else if (key == "fuse_dw") {
    fuse_dw_ = value == "1" || value == "true";
  } else {
    Engine::set_option(key, value);
  }
}

void SnacEngine::require_ready() const {
  if (!ready_) throw Error(NC_BAD_WEIGHTS, "SNAC weights have not been loaded");
}

std::string SnacEngine::describe() const {
  std::string s = "{\"codec\": \"SNAC\", \"precision\": \"";
  s += precision_name(prec_);
  s += "\", \"layers\": {";
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
  add(enc_out_dense_);
  add(dec_in_);
  for (auto& b : dec_blocks_) {
    add(b->up); add(b->noise);
    for (auto& r : b->ru) { add(r.c1); add(r.c2); }
  }
  s += "}}";
  return s;
}

// w = (v / ||v||_(1,2)) * (g - 1e-7), folded once in fp32 (Modules/SNAC/WNConv1d.cs:122-135,
// WNConvTranspose1d.cs:126-133).  Also accepts a pre-folded "<name>.weight".
std::vector<float> SnacEngine::folded(const std::string& name, int d0, int d1, int k, std::vector<float>* bias, int bias_n) {
  std::vector<float> w;
  if (has_tensor(name + ".weight")) {
    const HostTensor& t = tensor(name + ".weight");
    if (t.is_int || (int64_t)t.numel() != (int64_t)d0 * d1 * k)
      throw Error(NC_SHAPE_MISMATCH, "Failed to load SNAC weights: '" + name + ".weight' has the wrong shape");
    w = t.f32;
  } else {
    const HostTensor& v = tensor(name + ".parametrizations.weight.original1");
    const HostTensor& g = tensor(name + ".parametrizations.weight.original0");
    if (v.is_int || v.shape.size() != 3 || v.shape[0] != d0 || v.shape[1] != d1 || v.shape[2] != k)
      throw Error(NC_SHAPE_MISMATCH, "Failed to load SNAC weights: '" + name + "' has the wrong shape");
    if (g.is_int || (int64_t)g.numel() != d0)
      throw Error(NC_SHAPE_MISMATCH, "Failed to load SNAC weights: '" + name + "' weight_g has the wrong shape");
    w.resize(v.f32.size());
    const size_t inner = (size_t)d1 * k;
    for (int i = 0; i < d0; ++i) {
      double ss = 0;
      for (size_t j = 0; j < inner; ++j) ss += (double)v.f32[i * inner + j] * v.f32[i * inner + j];
      const float norm = std::sqrt((float)ss);
      const float gi = g.f32[i] - 1e-7f;
      for (size_t j = 0; j < inner; ++j) w[i * inner + j] = (v.f32[i * inner + j] / norm) * gi;
    }
  }
  if (bias) {
    bias->clear();
    if (has_tensor(name + ".bias")) {
      const HostTensor& b = tensor(name + ".bias");
      if (b.is_int || (int)b.numel() != bias_n)
        throw Error(NC_SHAPE_MISMATCH, "Failed to load SNAC weights: '" + name + ".bias' has the wrong shape");
      *bias = b.f32;
    }
  }
  return w;
}

// zero-pad [d0][d1][k] -> [p0][p1][k]
static std::vector<float> pad3(const std::vector<float>& w, int d0, int d1, int k, int p0, int p1) {
  if (d0 == p0 && d1 == p1) return w;
  std::vector<float> o((size_t)p0 * p1 * k, 0.f);
  for (int i = 0; i < d0; ++i)
    for (int j = 0; j < d1; ++j)
      std::memcpy(&o[((size_t)i * p1 + j) * k], &w[((size_t)i * d1 + j) * k], sizeof(float) * k);
  return o;
}
static std::vector<float> pad1(const std::vector<float>& b, int n, int p, float fill) {
  std::vector<float> o((size_t)p, fill);
  for (int i = 0; i < n && i < (int)b.size(); ++i) o[i] = b[i];
  return o;
}

void SnacEngine::build_dw(DwConv& d, const std::string& name, int C, int dil) {
  std::vector<float> b;
  auto w = folded(name, C, 1, 7, &b, C);   // [C][1][7]
  const int Cp = pad32(C);
  std::vector<float> wkc((size_t)7 * Cp, 0.f);
  for (int c = 0; c < C; ++c)
    for (int j = 0; j < 7; ++j) wkc[(size_t)j * Cp + c] = w[(size_t)c * 7 + j];
  cudaFree(d.w); cudaFree(d.b);
  d.name = name; d.C = Cp; d.dil = dil;
  d.w = upload(wkc);
  d.b = upload(pad1(b, C, Cp, 0.f));
}

static std::vector<float> alpha_vec(const HostTensor& t, int c, int cp, const std::string& name) {
  if (t.is_int || (int)t.numel() != c)
    throw Error(NC_SHAPE_MISMATCH, "Failed to load SNAC weights: '" + name + "' has the wrong shape");
  return pad1(t.f32, c, cp, 1.0f);
}

void SnacEngine::build_ru(ResUnit& ru, const std::string& p, int dim, int dil) {
  const int dp = pad32(dim);
  std::vector<float> b;
  ru.s1.build(alpha_vec(tensor(p + ".block.0.alpha"), dim, dp, p + ".block.0.alpha"));
  ru.s2.build(alpha_vec(tensor(p + ".block.2.alpha"), dim, dp, p + ".block.2.alpha"));
  if (cfg_.depthwise) {
    build_dw(ru.dw, p + ".block.1", dim, dil);
  } else {
    ConvSpec s1;
    s1.cin = s1.cout = dp; s1.k = 7; s1.dilation = dil; s1.padding = 3 * dil;
    auto w1 = folded(p + ".block.1", dim, dim, 7, &b, dim);
    ru.c1.build(p + ".block.1", s1, pad3(w1, dim, dim, 7, dp, dp), pad1(b, dim, dp, 0.f), prec_);
  }
  ConvSpec s2;
  s2.cin = s2.cout = dp; s2.k = 1;
  auto w2 = folded(p + ".block.3", dim, dim, 1, &b, dim);
  ru.c2.build(p + ".block.3", s2, pad3(w2, dim, dim, 1, dp, dp), pad1(b, dim, dp, 0.f), prec_);
}

void SnacEngine::finalize_weights() {
  bind();
  ready_ = false;
  std::vector<float> b;
  // ---- encoder (Modules/SNAC/Encoder.cs:26-69)
  int d = cfg_.encoder_dim;
  c0p_ = pad32(d);
  {
    auto w = folded("encoder.block.0", d, 1, 7, &b, d);
    cudaFree(d_conv_in_w_); cudaFree(d_conv_in_b_);
    d_conv_in_w_ = upload(pad3(w, d, 1, 7, c0p_, 1));
    d_conv_in_b_ = upload(pad1(b, d, c0p_, 0.f));
  }
  enc_blocks_.clear();
  int idx = 1;
  for (size_t i = 0; i < cfg_.encoder_rates.size(); ++i, ++idx) {
    const int s = cfg_.encoder_rates[i];
    auto blk = std::make_unique<EncBlock>();
    const std::string p = "encoder.block." + std::to_string(idx);
    const int dils[3] = {1, 3, 9};
    for (int u = 0; u < 3; ++u) build_ru(blk->ru[u], p + ".block." + std::to_string(u), d, dils[u]);
    blk->s.build(alpha_vec(tensor(p + ".block.3.alpha"), d, pad32(d), p + ".block.3.alpha"));
    ConvSpec cs;
    cs.cin = pad32(d); cs.cout = pad32(2 * d); cs.k = 2 * s; cs.stride = s; cs.padding = (s + 1) / 2;  // EncoderBlock.cs:39-46
    auto w = folded(p + ".block.4", 2 * d, d, 2 * s, &b, 2 * d);
    blk->down.build(p + ".block.4", cs, pad3(w, 2 * d, d, 2 * s, cs.cout, cs.cin), pad1(b, 2 * d, cs.cout, 0.f), prec_);
    enc_blocks_.push_back(std::move(blk));
    d *= 2;
  }
  if (d != cfg_.latent_dim) throw Error(NC_INVALID_ARGUMENT, "SNAC: latent_dim must equal encoder_dim * 2^n_rates");
  dzp_ = pad32(d);
  enc_mha_.present = dec_mha_.present = false;
  if (cfg_.attn_window) {   // Encoder.cs:50-53
    build_mha(enc_mha_, "encoder.block." + std::to_string(idx), d);
    ++idx;
  }
  {
    const std::string p = "encoder.block." + std::to_string(idx);   // final conv (no Snake before it, Encoder.cs:55-62)
    if (cfg_.depthwise) {
      build_dw(enc_out_dw_, p, d, 1);
    } else {
      ConvSpec cs;
      cs.cin = cs.cout = dzp_; cs.k = 7; cs.padding = 3;
      auto w = folded(p, d, d, 7, &b, d);
      enc_out_dense_.build(p, cs, pad3(w, d, d, 7, dzp_, dzp_), pad1(b, d, dzp_, 0.f), prec_);
    }
  }
  // ---- quantiser (Modules/SNAC/VectorQuantizer.cs:44-56)
  {
    for (float* p : vq_alloc_) cudaFree(p);
    vq_alloc_.clear();
    stages_.clear();
    const int D = cfg_.codebook_dim, K = cfg_.codebook_size, Dz = cfg_.latent_dim;
    for (size_t q = 0; q < cfg_.vq_strides.size(); ++q) {
      const std::string p = "quantizer.quantizers." + std::to_string(q);
      auto wi = folded(p + ".in_proj", D, Dz, 1, &b, D);
      std::vector<float> in_w = pad3(wi, D, Dz, 1, D, dzp_), in_b = pad1(b, D, D, 0.f);
      auto wo = folded(p + ".out_proj", Dz, D, 1, &b, Dz);
      std::vector<float> out_w = pad3(wo, Dz, D, 1, dzp_, D), out_b = pad1(b, Dz, dzp_, 0.f);
      const HostTensor& c = tensor(p + ".codebook.weight");
      if (c.is_int || c.shape.size() != 2 || c.shape[0] != K || c.shape[1] != D)
        throw Error(NC_SHAPE_MISMATCH, "Failed to load SNAC weights: '" + p + ".codebook.weight' has the wrong shape");
      std::vector<float> sq((size_t)K);
      for (int k = 0; k < K; ++k) {
        float s = 0.f;
        for (int dd = 0; dd < D; ++dd) {
          const float x = c.f32[(size_t)k * D + dd];
          const float x2 = x * x;
          s += x2;
        }
        sq[k] = s;
      }
      SnacVqStage st;
      st.stride = cfg_.vq_strides[q];
      float* ptr;
      vq_alloc_.push_back(ptr = upload(in_w)); st.in_w = ptr;
      vq_alloc_.push_back(ptr = upload(in_b)); st.in_b = ptr;
      vq_alloc_.push_back(ptr = upload(c.f32)); st.cb = ptr;
      vq_alloc_.push_back(ptr = upload(sq)); st.cb_sq = ptr;
      vq_alloc_.push_back(ptr = upload(out_w)); st.out_w = ptr;
      vq_alloc_.push_back(ptr = upload(out_b)); st.out_b = ptr;
      stages_.push_back(st);
    }
  }
  // ---- decoder (Modules/SNAC/Decoder.cs:28-86)
  const int C = cfg_.decoder_dim;
  idx = 0;
  if (cfg_.depthwise) {
    build_dw(dec_in_dw_, "decoder.model.0", cfg_.latent_dim, 1);
    ConvSpec cs;
    cs.cin = dzp_; cs.cout = pad32(C); cs.k = 1;
    auto w = folded("decoder.model.1", C, cfg_.latent_dim, 1, &b, C);
    dec_in_.build("decoder.model.1", cs, pad3(w, C, cfg_.latent_dim, 1, cs.cout, cs.cin), pad1(b, C, cs.cout, 0.f), prec_);
    idx = 2;
  } else {
    ConvSpec cs;
    cs.cin = dzp_; cs.cout = pad32(C); cs.k = 7; cs.padding = 3;
    auto w = folded("decoder.model.0", C, cfg_.latent_dim, 7, &b, C);
    dec_in_.build("decoder.model.0", cs, pad3(w, C, cfg_.latent_dim, 7, cs.cout, cs.cin), pad1(b, C, cs.cout, 0.f), prec_);
    idx = 1;
  }
  if (cfg_.attn_window) {   // Decoder.cs:54-57
    build_mha(dec_mha_, "decoder.model." + std::to_string(idx), C);
    ++idx;
  }
  dec_blocks_.clear();
  int cout = C;
  for (size_t i = 0; i < cfg_.decoder_rates.size(); ++i, ++idx) {
    const int s = cfg_.decoder_rates[i];
    const int cin = C / (1 << i);
    cout = C / (1 << (i + 1));
    const int cinp = pad32(cin), coutp = pad32(cout);
    auto blk = std::make_unique<DecBlock>();
    const std::string p = "decoder.model." + std::to_string(idx);
    blk->s.build(alpha_vec(tensor(p + ".block.0.alpha"), cin, cinp, p + ".block.0.alpha"));
    ConvSpec cs;
    cs.transposed = true; cs.cin = cinp; cs.cout = coutp; cs.k = 2 * s; cs.stride = s; cs.padding = (s + 1) / 2;
    cs.output_padding = s % 2;                                           // DecoderBlock.cs:38-45
    auto w = folded(p + ".block.1", cin, cout, 2 * s, &b, cout);         // norm over dims (1,2) of [Cin,Cout,k]
    blk->up.build(p + ".block.1", cs, pad3(w, cin, cout, 2 * s, cinp, coutp), pad1(b, cout, coutp, 0.f), prec_);
    int bi = 2;
    if (cfg_.noise) {
      ConvSpec ns;
      ns.cin = ns.cout = coutp; ns.k = 1;
      auto wn = folded(p + ".block.2.linear", cout, cout, 1, nullptr, 0);   // NoiseBlock.cs:30: no bias
      blk->noise.build(p + ".block.2.linear", ns, pad3(wn, cout, cout, 1, coutp, coutp), std::vector<float>(), prec_);
      bi = 3;
    }
    const int dils[3] = {1, 3, 9};
    for (int u = 0; u < 3; ++u) build_ru(blk->ru[u], p + ".block." + std::to_string(bi + u), cout, dils[u]);
    dec_blocks_.push_back(std::move(blk));
  }
  {
    const std::string ps = "decoder.model." + std::to_string(idx), pc = "decoder.model." + std::to_string(idx + 1);
    const int cp = pad32(cout);
    dec_snake_.build(alpha_vec(tensor(ps + ".alpha"), cout, cp, ps + ".alpha"));
    auto w = folded(pc, 1, cout, 7, &b, 1);
    std::vector<float> wkc((size_t)7 * cp, 0.f);
    for (int ci = 0; ci < cout; ++ci)
      for (int j = 0; j < 7; ++j) wkc[(size_t)j * cp + ci] = w[(size_t)ci * 7 + j];
    cudaFree(d_conv_out_w_); cudaFree(d_conv_out_b_);
    d_conv_out_w_ = upload(wkc);
    d_conv_out_b_ = b.empty() ? nullptr : upload(b);
    conv_out_c_ = cp;
  }
  drop_tensors();
  ready_ = true;
}

// ------------------------------------------------------------------------------------ shapes
static int64_t lcm64(int64_t a, int64_t b) {
  int64_t x = a, y = b;
  while (y) { int64_t t = x % y; x = y; y = t; }
  return a / x * b;
}

int64_t SnacEngine::padded_length(int64_t L) const {
  const int64_t pad_to = (int64_t)cfg_.hop() * lcm64(cfg_.vq_strides[0], cfg_.attn_window > 0 ? cfg_.attn_window : 1);
  return (L + pad_to - 1) / pad_to * pad_to;   // Models/SNAC.cs:74-77
}

std::vector<int64_t> SnacEngine::noise_lengths(int64_t T) const {
  std::vector<int64_t> out;
  int64_t t = T;
  for (int s : cfg_.decoder_rates) {
    t = (t - 1) * s - 2 * ((s + 1) / 2) + 2 * s + (s % 2);
    out.push_back(t);
  }
  return out;
}

int64_t SnacEngine::decoded_length(int64_t T) const {
  auto v = noise_lengths(T);
  return v.empty() ? T : v.back();
}

int SnacEngine::micro_batch(int B, int64_t Lp) {
  int64_t peak = Lp * c0p_;
  int64_t t = Lp;
  int d = cfg_.encoder_dim;
  for (int s : cfg_.encoder_rates) { t /= s; d *= 2; peak = std::max<int64_t>(peak, t * pad32(d)); }
  const int64_t T = t;
  int64_t td = T;
  peak = std::max<int64_t>(peak, td * pad32(cfg_.decoder_dim));
  int c = cfg_.decoder_dim;
  for (int s : cfg_.decoder_rates) { td = (td - 1) * s - 2 * ((s + 1) / 2) + 2 * s + (s % 2); c /= 2; peak = std::max<int64_t>(peak, td * pad32(c)); }
  per_clip_elems_ = peak;
  const double per_clip = 3.0 * (double)peak * 4 + 2.0 * (double)T * dzp_ * 4 + 2.0 * (double)td * 4;
  int mb = (int)std::max(1.0, std::floor((double)max_workspace_bytes_ / per_clip));
  mb = std::min(mb, B);
  for (auto& w : ws_) w.reserve((size_t)mb * peak * sizeof(float));
  z_in_.reserve((size_t)mb * T * dzp_ * sizeof(float));
  z_q_.reserve((size_t)mb * T * dzp_ * sizeof(float));
  audio_tmp_.reserve((size_t)mb * td * sizeof(float));
  return mb;
}

// ------------------------------------------------------------------------------------ graph
int SnacEngine::run_ru(const ResUnit& ru, int cur, int B, int T, const SnakeParams* post) {
  const LaunchCtx c = ctx();
  const int h = (cur + 1) % 3, y = (cur + 2) % 3;
  if (cfg_.depthwise && fuse_dw_ && ru.dw.C == ru.c2.spec().cin) {
    // whole unit in one launch: Snake -> depthwise k7 -> Snake evaluated by the 1x1 conv's operand-transform warps,
    // x read once as operand (+ halo) and once as residual, y written once (5 -> 3 HBM passes per unit)
    ConvRunArgs f;
    f.in = buf(cur); f.out = buf(y); f.residual = buf(cur); f.batch = B; f.t_in = T;
    f.prologue = PRO_SNAKE; f.alpha = ru.s1.alpha; f.inv_alpha = ru.s1.inv_alpha;
    f.dw_w = ru.dw.w; f.dw_b = ru.dw.b; f.dw_dil = ru.dw.dil; f.dw_post_alpha = ru.s2.alpha; f.dw_post_inv_alpha = ru.s2.inv_alpha;
    if (post) { f.post = PRO_SNAKE; f.post_alpha = post->alpha; f.post_inv_alpha = post->inv_alpha; }
    ConvGemmParams probe;
    if (ru.c2.fill_umma(f, &probe)) {
      ru.c2.run(f, c);
      return y;
    }
  }
  if (cfg_.depthwise) {
    launch_dwconv7(buf(cur), buf(h), T, ru.dw.C, ru.dw.w, ru.dw.b, ru.dw.dil, ru.s1.alpha, ru.s2.alpha, B, c, ru.dw.name.c_str(),
                   prec_ != PREC_FP32 && prec_ != PREC_3XTF32);
  } else {
    ConvRunArgs a;
    a.in = buf(cur); a.out = buf(h); a.batch = B; a.t_in = T;
    a.prologue = PRO_SNAKE; a.alpha = ru.s1.alpha; a.inv_alpha = ru.s1.inv_alpha;
    a.post = PRO_SNAKE; a.post_alpha = ru.s2.alpha; a.post_inv_alpha = ru.s2.inv_alpha;
    ru.c1.run(a, c);
  }
  ConvRunArgs b;
  b.in = buf(h); b.out = buf(y); b.residual = buf(cur); b.batch = B; b.t_in = T;
  if (post) { b.post = PRO_SNAKE; b.post_alpha = post->alpha; b.post_inv_alpha = post->inv_alpha; }
  ru.c2.run(b, c);
  return y;
}

void SnacEngine::run_encoder(const float* audio, int in_len, int B, int Lp) {
  const LaunchCtx c = ctx();
  launch_conv_cin1(audio, in_len, in_len, buf(0), Lp, c0p_, d_conv_in_w_, d_conv_in_b_, 7, 1, 3, B, c);
  int cur = 0, T = Lp;
  for (auto& blk : enc_blocks_) {
    for (int u = 0; u < 3; ++u) cur = run_ru(blk->ru[u], cur, B, T, u == 2 ? &blk->s : nullptr);
    ConvRunArgs a;
    a.in = buf(cur); a.out = buf((cur + 1) % 3); a.batch = B; a.t_in = T;
    blk->down.run(a, c);
    T = blk->down.out_len(T);
    cur = (cur + 1) % 3;
  }
  if (enc_mha_.present) cur = run_mha(enc_mha_, cur, B, T, nullptr);
  if (cfg_.depthwise) {
    launch_dwconv7(buf(cur), z_in_.as<float>(), T, dzp_, enc_out_dw_.w, enc_out_dw_.b, 1, nullptr, nullptr, B, c,
                   enc_out_dw_.name.c_str());
  } else {
    ConvRunArgs a;
    a.in = buf(cur); a.out = z_in_.as<float>(); a.batch = B; a.t_in = T;
    enc_out_dense_.run(a, c);
  }
}

// z_in_ is consumed as the running residual; z_q_ accumulates the quantised latent
void SnacEngine::run_rvq(int B, int T, int64_t* const* codes, int b0) {
  const LaunchCtx c = ctx();
  NC_CUDA(cudaMemsetAsync(z_q_.as<float>(), 0, (size_t)B * T * dzp_ * sizeof(float), stream_));
  ++launches_;
  for (size_t q = 0; q < stages_.size(); ++q) {
    const int st = stages_[q].stride;
    int64_t* cq = (codes && codes[q]) ? codes[q] + (int64_t)b0 * (T / st) : nullptr;
    launch_snac_vq_stage(stages_[q], z_in_.as<float>(), z_q_.as<float>(), cq, B, T, dzp_, cfg_.codebook_size, c);
  }
}

void SnacEngine::run_decoder(int B, int T, const float* const* noise, uint64_t seed, int b0, float* audio_out) {
  const LaunchCtx c = ctx();
  const SnakeParams* first = dec_blocks_.empty() ? &dec_snake_ : &dec_blocks_[0]->s;
  ConvRunArgs a;
  a.batch = B; a.t_in = T; a.out = buf(0);
  if (!dec_mha_.present) { a.post = PRO_SNAKE; a.post_alpha = first->alpha; a.post_inv_alpha = first->inv_alpha; }
  if (cfg_.depthwise) {
    launch_dwconv7(z_q_.as<float>(), buf(1), T, dzp_, dec_in_dw_.w, dec_in_dw_.b, 1, nullptr, nullptr, B, c, dec_in_dw_.name.c_str());
    a.in = buf(1);
  } else {
    a.in = z_q_.as<float>();
  }
  dec_in_.run(a, c);
  int cur = 0;
  if (dec_mha_.present) cur = run_mha(dec_mha_, cur, B, T, first);
  const auto nlen = noise_lengths(T);
  for (size_t i = 0; i < dec_blocks_.size(); ++i) {
    auto& blk = dec_blocks_[i];
    ConvRunArgs u;
    u.in = buf(cur); u.out = buf((cur + 1) % 3); u.batch = B; u.t_in = T;
    blk->up.run(u, c);
    T = blk->up.out_len(T);
    cur = (cur + 1) % 3;
    if (cfg_.noise) {
      // x + randn[B,1,T] * linear(x)   (NoiseBlock.cs:38-45); explicit noise for parity, seeded generator otherwise
      const float* nz;
      if (noise && noise[i]) {
        nz = noise[i] + (int64_t)b0 * nlen[i];
      } else {
        float* gen = static_cast<float*>(noise_buf_.reserve((size_t)B * T * sizeof(float)));
        launch_randn(gen, B, T, b0, seed, (uint32_t)i, c);
        nz = gen;
      }
      ConvRunArgs n;
      n.in = buf(cur); n.out = buf((cur + 1) % 3); n.residual = buf(cur); n.noise = nz; n.batch = B; n.t_in = T;
      blk->noise.run(n, c);
      cur = (cur + 1) % 3;
    }
    const SnakeParams* next = i + 1 < dec_blocks_.size() ? &dec_blocks_[i + 1]->s : &dec_snake_;
    for (int r = 0; r < 3; ++r) cur = run_ru(blk->ru[r], cur, B, T, r == 2 ? next : nullptr);
  }
  launch_conv_cout1(buf(cur), audio_out, T, conv_out_c_, d_conv_out_w_, d_conv_out_b_, 7, 3, 1, B, c);
}

// ------------------------------------------------------------------------------------ entry points
void SnacEngine::encode_dev(const float* audio, int B, int64_t L, int64_t* const* codes) {
  forward_dev(audio, B, L, nullptr, 0, nullptr, codes);
}

void SnacEngine::forward_dev(const float* audio, int B, int64_t L, const float* const* noise, uint64_t seed, float* audio_out,
                             int64_t* const* codes) {
  require_ready();
  bind();
  if (B <= 0 || L <= 0) throw Error(NC_INVALID_ARGUMENT, "batch and length must be positive");
  const int64_t Lp = padded_length(L);
  if (Lp > (int64_t)1 << 30) throw Error(NC_INVALID_ARGUMENT, "clip too long");
  const int mb = micro_batch(B, Lp);
  const int64_t T = Lp / cfg_.hop();
  const int64_t out_len = decoded_length(T);
  const LaunchCtx c = ctx();
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int nb = std::min(mb, B - b0);
    run_encoder(audio + (int64_t)b0 * L, (int)L, nb, (int)Lp);
    run_rvq(nb, (int)T, codes, b0);
    if (audio_out) {
      run_decoder(nb, (int)T, noise, seed, b0, audio_tmp_.as<float>());
      launch_trim_rows(audio_tmp_.as<float>(), audio_out + (int64_t)b0 * L, nb, out_len, std::min<int64_t>(L, out_len), c);
    }
  }
  sync();
}

void SnacEngine::decode_dev(const int64_t* const* codes, int B, int64_t T, const float* const* noise, uint64_t seed,
                            float* audio_out) {
  require_ready();
  bind();
  if (B <= 0 || T <= 0 || !codes) throw Error(NC_INVALID_ARGUMENT, "decode: bad batch / frames / codes");
  if (T % cfg_.vq_strides[0] != 0) throw Error(NC_INVALID_ARGUMENT, "decode: frames must be a multiple of the first vq stride");
  const int64_t Lp = T * cfg_.hop();
  const int mb = micro_batch(B, Lp);
  const int64_t out_len = decoded_length(T);
  const LaunchCtx c = ctx();
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int nb = std::min(mb, B - b0);
    SnacFromCodes fc;
    fc.n_stages = (int)stages_.size();
    for (int q = 0; q < fc.n_stages; ++q) {
      if (!codes[q]) throw Error(NC_INVALID_ARGUMENT, "decode: null code array");
      fc.stride[q] = stages_[q].stride;
      fc.codes[q] = codes[q] + (int64_t)b0 * (T / stages_[q].stride);
      fc.cb[q] = stages_[q].cb; fc.out_w[q] = stages_[q].out_w; fc.out_b[q] = stages_[q].out_b;
    }
    launch_snac_from_codes(fc, z_q_.as<float>(), nb, (int)T, dzp_, cfg_.codebook_size, c);
    run_decoder(nb, (int)T, noise, seed, b0, audio_out + (int64_t)b0 * out_len);
  }
  sync();
}

}  // namespace nc
