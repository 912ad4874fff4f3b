// Engine objects behind an nc_handle.
#pragma once
#include <atomic>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "codec_kernels.h"
#include "conv_layer.h"
#include "runtime.h"
#include "safetensors.h"

namespace nc {

// Base of every codec engine: device binding, stream, profiler, named host tensors.
class Engine {
 public:
  Engine(int device_index);
  virtual ~Engine();
  virtual const char* codec_name() const = 0;
  virtual void finalize_weights() = 0;
  virtual void set_option(const std::string& key, const std::string& value);
  virtual std::string describe() const = 0;

  void load_weights(const std::string& path);   // .safetensors or a torch.save zip checkpoint (.pth / .bin)
  const std::string& weights_metadata() const { return weights_metadata_; }
  void set_tensor(const std::string& name, HostTensor&& t) { tensors_[name] = std::move(t); }
  void bind() const;  // cudaSetDevice
  LaunchCtx ctx();
  uint64_t launches() const { return launches_; }
  Profiler& profiler() { return prof_; }
  void sync();
  cudaStream_t stream() const { return stream_; }
  // second stream of the handle: host <-> device staging copies that overlap the kernels (host-pointer entry points)
  cudaStream_t copy_stream();

  // re-entrancy guard (a handle is one-thread-at-a-time)
  std::atomic<bool> busy{false};   // set for the duration of an entry point (BusyGuard): handles are not re-entrant

 protected:
  const HostTensor& tensor(const std::string& name) const;
  bool has_tensor(const std::string& name) const { return tensors_.count(name) != 0; }
  void drop_tensors() { tensors_.clear(); }
  // hook: translate a checkpoint's native parameter names into the names finalize_weights() looks up
  virtual void normalize_names() {}
  TensorMap& tensors_mut() { return tensors_; }

  int device_ = 0;
  int num_sms_ = 148;
  cudaStream_t stream_ = nullptr;
  cudaStream_t copy_stream_ = nullptr;
  Profiler prof_;
  uint64_t launches_ = 0;
  int fast_sin_ = -1, fuse_ru_ = 1;   // options "fast_sin" / "fuse_ru" (per handle)
  size_t max_workspace_bytes_ = (size_t)32 << 30;
  TensorMap tensors_;
  std::string weights_metadata_ = "{}";
  bool ready_ = false;
};

struct SnakeParams {
  float* alpha = nullptr;      // device [C]
  float* inv_alpha = nullptr;  // device [C]
  ~SnakeParams();
  void build(const std::vector<float>& a);
};

// ------------------------------------------------------------------------------------ DAC
struct DacConfig {
  int sample_rate = 44100;
  int encoder_dim = 64;
  std::vector<int> encoder_rates{2, 4, 8, 8};
  int decoder_dim = 1536;
  std::vector<int> decoder_rates{8, 8, 4, 2};
  int n_codebooks = 9, codebook_size = 1024, codebook_dim = 8;
  int latent_dim = 1024;
  int hop() const {
    int h = 1;
    for (int r : encoder_rates) h *= r;
    return h;
  }
};

class DacEngine : public Engine {
 public:
  DacEngine(const nc_dac_config& cfg, int device_index);
  ~DacEngine() override;
  const char* codec_name() const override { return "DAC"; }
  void finalize_weights() override;
  void normalize_names() override;   // descript-audio-codec module paths -> HF DacModel names
  void set_option(const std::string& key, const std::string& value) override;
  const DacConfig& config() const { return cfg_; }
  std::string describe() const override;

  int64_t padded_length(int64_t L) const;
  int64_t frames(int64_t L) const { return padded_length(L) / cfg_.hop(); }
  int64_t decoded_length(int64_t T) const;  // samples produced by Decode for T frames
  int micro_batch(int B, int64_t Lp) const; // clips the engine processes per internal pass for clips of Lp padded samples

  // All pointers are device memory on this engine's device; nullable outputs may be null.
  // audio [B][L]; audio_out [B][decoded_length(T)]; codes [B][nq][T]; z [B][latent][T]; latents [B][nq*D][T]
  void encode_dev(const float* audio, int B, int64_t L, int nq, float* z, int64_t* codes, float* latents);
  void decode_dev(const float* z, int B, int64_t T, float* audio_out);
  void from_codes_dev(const int64_t* codes, int B, int nq, int64_t T, float* z);
  void decode_codes_dev(const int64_t* codes, int B, int nq, int64_t T, float* audio_out);
  void forward_dev(const float* audio, int B, int64_t L, int nq, float* audio_out, int64_t* codes, float* z);
  // Batched Dia hand-off (Models/Dia.cs:1010-1060): generated [B][T][C] delayed codes (device), lengths [B] (host),
  // delay [C] (host) -> audio_out [B][audio_stride] (device), item b holding lengths[b]*hop samples.  Items of equal
  // length are decoded together (the reference decodes them one by one).
  void decode_dia_dev(const int64_t* generated, int B, int T, int C, const int* delay, const int64_t* lengths,
                      float* audio_out, int64_t audio_stride);

 private:
  struct ResUnit {
    SnakeParams s1, s2;
    ConvLayer c1, c2;
  };
  struct EncBlock {
    ResUnit ru[3];
    SnakeParams s;
    ConvLayer down;
  };
  struct DecBlock {
    SnakeParams s;
    ConvLayer up;
    ResUnit ru[3];
  };
  void require_ready() const;
  void forward_impl(const float* audio, int B, int64_t L, int nq, float* audio_out, int64_t* codes, float* z,
                    float* latents);
  std::vector<float> folded_conv(const std::string& prefix, int d0, int d1, int k, std::vector<float>* bias,
                                 int bias_n);
  Precision boosted(Precision p, bool narrow) const;
  void build_ru(ResUnit& ru, const std::string& prefix, int dim, int dil, Precision prec, int short_chains = 0);
  // returns the buffer index holding the result; T_io: in = input length, out = output length
  int run_ru(const ResUnit& ru, int cur, int B, int T, const SnakeParams* post);
  int run_encoder(const float* audio, long long audio_stride, int in_len, int B, int Lp, int* T_out);
  int run_decoder(int cur, int B, int T, float* audio_out, long long out_stride);
  void ensure_workspace(int mb, int64_t Lp);
  float* buf(int i) { return ws_[i].as<float>(); }
  void* buf16(int i) { return ws16_[i].as<void>(); }
  bool wide_layer(int channels) const { return channels > 128 && dec_prec_ != PREC_FP32; }
  // wide ResidualUnit on the fp16-operand executor: x32 (raw, buffer cur32) + x16 = Snake1(x) (buffer cur16) ->
  // out32 (raw, written only when `need32`) + out16 = post(out) ; returns through *cur32 / *cur16
  void run_ru_h16(const ResUnit& ru, int* cur32, int* cur16, int B, int T, const SnakeParams* post, bool need32);

  DacConfig cfg_;
  // Encoder: three-pass bf16 split everywhere (codes must match the fp32 reference).  Decoder: the wide layers
  // (convs, transposed convs and residual units with more than 128 channels -- blocks 0-2 and decoder.conv1) take one
  // fp16 product, everything else the three-pass split: 68.6 dB against the fp32 oracle (gate 60 dB), +26 % sustained
  // throughput (profiles/r01_decoder_precision_modes.txt).  decoder_precision=<mode> makes the decoder uniform again.
  Precision enc_prec_ = PREC_BF16X3, dec_prec_ = PREC_BF16X3, dec_wide_prec_ = PREC_F16;
  bool dec_boost_ = true;
  int enc_tail_fp32_ = 0;   // option encoder_tail_fp32
  int enc_short_chains_ = 1;  // option encoder_short_chains
  // encoder
  float* d_conv_in_w_ = nullptr;
  float* d_conv_in_b_ = nullptr;
  std::vector<std::unique_ptr<EncBlock>> enc_blocks_;
  SnakeParams enc_snake_;
  ConvLayer enc_out_;
  // quantiser
  RvqWeights rvq_{};
  std::vector<float*> rvq_alloc_;
  // decoder
  ConvLayer dec_in_;
  std::vector<std::unique_ptr<DecBlock>> dec_blocks_;
  SnakeParams dec_snake_;
  ConvLayer dec_out_;
  float* d_conv_out_w_ = nullptr;  // [7][C] for the Cout = 1 kernel (null: use dec_out_)
  float* d_conv_out_b_ = nullptr;
  int conv_out_c_ = 0;
  // workspaces: 3 rotating activation buffers + latent buffers
  DeviceBuffer ws_[3], z_in_, z_q_, dia_codes_, dia_idx_, dia_audio_;
  DeviceBuffer ws16_[3], z16_;  // fp16 activations of the fp16-operand decoder layers (conv_h16.cu)
  int64_t per_clip_elems_ = 0;  // per padded sample, see ensure_workspace
  int64_t per_clip_elems16_ = 0;
  bool h16_ = false;            // decoder's wide layers run on the fp16-operand executor (option decoder_h16, default on)
  int dec_h16_opt_ = 1;
};

Engine* create_engine(nc_codec_kind kind, const void* cfg, size_t cfg_size, int device_index);

}  // namespace nc

// ------------------------------------------------------------------------------------ SNAC
#include "snac_kernels.h"
namespace nc {

struct SnacConfig {
  int sample_rate = 24000, encoder_dim = 48, decoder_dim = 1024, latent_dim = 768;
  std::vector<int> encoder_rates{2, 4, 8, 8}, decoder_rates{8, 8, 4, 2}, vq_strides{4, 2, 1};
  int attn_window = 0, codebook_size = 4096, codebook_dim = 8;
  bool noise = true, depthwise = true;
  int hop() const {
    int h = 1;
    for (int r : encoder_rates) h *= r;
    return h;
  }
};

// Depthwise k=7 conv weights on the device ([7][C] + bias [C]).
struct DwConv {
  std::string name;
  int C = 0, dil = 1;
  float* w = nullptr;
  float* b = nullptr;
  ~DwConv();
};

class SnacEngine : public Engine {
 public:
  SnacEngine(const nc_snac_config& cfg, int device_index);
  ~SnacEngine() override;
  const char* codec_name() const override { return "SNAC"; }
  void finalize_weights() override;
  void set_option(const std::string& key, const std::string& value) override;
  std::string describe() const override;
  const SnacConfig& config() const { return cfg_; }

  int64_t padded_length(int64_t L) const;             // Models/SNAC.cs:70-80
  int64_t frames(int64_t L) const { return padded_length(L) / cfg_.hop(); }
  int64_t decoded_length(int64_t T) const;
  std::vector<int64_t> noise_lengths(int64_t T) const;  // per decoder block
  int n_stages() const { return (int)cfg_.vq_strides.size(); }

  // device pointers; codes[i]: [B][T / vq_strides[i]] int64; noise[i]: [B][noise_lengths(T)[i]] or null (seeded RNG)
  void encode_dev(const float* audio, int B, int64_t L, int64_t* const* codes);
  void decode_dev(const int64_t* const* codes, int B, int64_t T, const float* const* noise, uint64_t seed, float* audio_out);
  // audio_out [B][L] (trimmed to the input length, Models/SNAC.cs:103); codes nullable
  void forward_dev(const float* audio, int B, int64_t L, const float* const* noise, uint64_t seed, float* audio_out,
                   int64_t* const* codes);

 private:
  struct ResUnit {
    SnakeParams s1, s2;
    DwConv dw;        // depthwise variant
    ConvLayer c1;     // dense variant (depthwise = false)
    ConvLayer c2;     // 1x1
  };
  struct EncBlock { ResUnit ru[3]; SnakeParams s; ConvLayer down; };
  struct DecBlock { SnakeParams s; ConvLayer up; ConvLayer noise; ResUnit ru[3]; };
  struct LocalMha {          // Modules/SNAC/LocalMHA.cs:46-115
    bool present = false;
    int dim = 0, heads = 0;
    float* ln_w = nullptr;
    float* ln_b = nullptr;
    float* inv_freq = nullptr;
    ConvLayer qkv, out;
    ~LocalMha();
  };
  void build_mha(LocalMha& m, const std::string& p, int dim);
  // x (buffer cur) -> x + to_out(attn(to_qkv(LN(x)))) with `post` applied on the way out; returns the buffer index
  int run_mha(const LocalMha& m, int cur, int B, int T, const SnakeParams* post);
  void require_ready() const;
  std::vector<float> folded(const std::string& name, int d0, int d1, int k, std::vector<float>* bias, int bias_n);
  void build_ru(ResUnit& ru, const std::string& p, int dim, int dil);
  void build_dw(DwConv& d, const std::string& name, int C, int dil);
  int run_ru(const ResUnit& ru, int cur, int B, int T, const SnakeParams* post);
  void run_encoder(const float* audio, int in_len, int B, int Lp);   // -> z_in_
  void run_rvq(int B, int T, int64_t* const* codes, int b0);         // z_in_ -> z_q_
  void run_decoder(int B, int T, const float* const* noise, uint64_t seed, int b0, float* audio_out);  // z_q_ -> audio [B][T*hop]
  int micro_batch(int B, int64_t Lp);
  float* buf(int i) { return ws_[i].as<float>(); }
  static int pad32(int c) { return (c + 31) / 32 * 32; }

  SnacConfig cfg_;
  Precision prec_ = PREC_BF16X3;
  bool fuse_dw_ = true;   // depthwise conv folded into the 1x1 conv's operand prologue (tcgen05 path)
  int dzp_ = 0;  // padded latent dim
  float* d_conv_in_w_ = nullptr;
  float* d_conv_in_b_ = nullptr;
  int c0p_ = 0;
  std::vector<std::unique_ptr<EncBlock>> enc_blocks_;
  LocalMha enc_mha_, dec_mha_;
  DwConv enc_out_dw_;
  ConvLayer enc_out_dense_;
  std::vector<SnacVqStage> stages_;
  std::vector<float*> vq_alloc_;
  DwConv dec_in_dw_;
  ConvLayer dec_in_;
  std::vector<std::unique_ptr<DecBlock>> dec_blocks_;
  SnakeParams dec_snake_;
  float* d_conv_out_w_ = nullptr;
  float* d_conv_out_b_ = nullptr;
  int conv_out_c_ = 0;
  DeviceBuffer ws_[3], z_in_, z_q_, noise_buf_, audio_tmp_, qkv_buf_;
  int64_t per_clip_elems_ = 0;
};

}  // namespace nc

// ------------------------------------------------------------------------------------ Encodec
#include "encodec_kernels.h"
namespace nc {

struct EncodecConfig {
  int sample_rate = 24000, channels = 1, n_filters = 32, dimension = 128;
  std::vector<int> ratios{8, 5, 4, 2};   // decoder order
  int n_residual_layers = 1, lstm_layers = 2, codebook_size = 1024, n_quantizers = 32;
  bool causal = true;
  // 48 kHz preset (EncodecConfig.cs:37-66): GroupNorm(1, C) after every conv, loudness scale per frame, segments
  bool group_norm = false, normalize = false;
  float segment_s = 0.f, overlap = 0.01f;   // segment_s <= 0: one frame per clip
  int hop() const {
    int h = 1;
    for (int r : ratios) h *= r;
    return h;
  }
  // Models/Encodec.cs:190-196 (float32 products)
  int64_t segment_length() const { return segment_s > 0.f ? (int64_t)(int)(segment_s * (float)sample_rate) : 0; }
  int64_t segment_stride() const {
    if (segment_s <= 0.f) return 0;
    const int v = (int)((1.0f - overlap) * (float)(int)segment_length());
    return v > 1 ? v : 1;
  }
};

class EncodecEngine : public Engine {
 public:
  EncodecEngine(const nc_encodec_config& cfg, int device_index);
  ~EncodecEngine() override;
  const char* codec_name() const override { return "Encodec"; }
  void finalize_weights() override;
  void set_option(const std::string& key, const std::string& value) override;
  std::string describe() const override;
  const EncodecConfig& config() const { return cfg_; }

  int64_t frames(int64_t L) const;                     // encoder frames for L samples (per-layer SConv1d length chain)
  // samples Decode produces for T frames: T' * hop, T' = T except for T <= 6 where the first decoder conv takes the
  // short-input branch of Pad1d and lengthens the sequence (SConv1d.cs:258-272)
  int64_t decoded_length(int64_t T) const;
  // padding one SConv1d applies to an input of T samples (kernel k, stride s): SConv1d.cs:144-173,245-272
  struct SPad { int extra_zero, left, right, t_out; };
  static SPad sconv_pad(int64_t T, int k, int s, bool causal);
  SPad sconv_pad(int64_t T, int k, int s) const { return sconv_pad(T, k, s, cfg_.causal); }
  // Segment loop of Encodec.Encode / Decode (Encodec.cs:213-235,273-282): segment s covers samples
  // [s*stride, min(s*stride + seg, L)); its codes sit at columns [col[s], col[s] + frames[s]) of the clip's [n_q][t_total] code
  // matrix and its decoded frame (ld[s] samples) is overlap-added at s*stride.  The first n_full segments have the full length.
  struct SegLayout {
    int n_seg = 1, n_full = 0;
    int64_t seg = 0, stride = 0, t_total = 0, total_out = 0, ld_max = 0;
    std::vector<int64_t> len, frames, ld, col;
  };
  SegLayout seg_layout(int64_t L) const;
  SegLayout seg_layout_from_frames(const int64_t* seg_frames, int n_seg) const;
  bool segmented() const { return cfg_.segment_s > 0.f; }
  // true: the 24 kHz-style engine path (mono, causal, weight-norm, one un-normalised frame per clip)
  bool simple() const { return !generic_io_; }
  // Device pointers.  audio [B][C][L]; codes [B][n_q][t_total] int64 (nullable); scales [B][n_seg] float (nullable; written only
  // when the model normalises); audio_out [B][C][L] (forward: Decode(...) sliced to the input length, nullable = encode only).
  void forward_frames_dev(const float* audio, int B, int64_t L, int nq, float* audio_out, int64_t* codes, float* scales);
  // codes [B][n_q][sum seg_frames], scales [B][n_seg] (nullable) -> audio_out [B][C][total_out]
  void decode_frames_dev(const int64_t* codes, const float* scales, int B, int nq, const int64_t* seg_frames, int n_seg,
                         float* audio_out);
  int n_q_for_bandwidth(float kbps) const;             // ResidualVectorQuantizer.cs:133-144

  // device pointers.  codes [B][nq][T] int64; audio_out [B][T*hop] for decode, [B][L] (trimmed) for forward.
  void encode_dev(const float* audio, int B, int64_t L, int nq, int64_t* codes);
  void decode_dev(const int64_t* codes, int B, int nq, int64_t T, float* audio_out);
  void forward_dev(const float* audio, int B, int64_t L, int nq, float* audio_out, int64_t* codes);

  // .ecdc payload (no language model): bits per code = log2(codebook size) (Encodec.cs:87,128-133)
  int bits_per_codebook() const;
  int64_t ecdc_payload_bytes(int nq, int64_t T) const { return ((int64_t)nq * T * bits_per_codebook() + 7) / 8; }
  // audio [B][L] -> payload [B][stride] bytes (EncodecCompressor.cs:93-190 with useLm=false)
  void compress_dev(const float* audio, int B, int64_t L, int nq, uint8_t* payload, int64_t stride);
  // Segmented streams (EncodecCompressor.cs:116-190,303-400): per segment an optional scale block (int32 BE count, float32 BE
  // values: written by the host) and the segment's codes packed on their own, zero-padded to a byte.  These pack / unpack
  // ONE segment (columns [col, col+T) of codes [B][nq][t_total]) to / from bytes at `bytes` + b*stride, then synchronise.
  void ecdc_pack_segment_dev(const int64_t* codes, int64_t t_total, int64_t col, int64_t T, int nq, uint8_t* bytes, int64_t stride, int B);
  void ecdc_unpack_segment_dev(const uint8_t* bytes, int64_t stride, int64_t* codes, int64_t t_total, int64_t col, int64_t T, int nq, int B);
  // payload -> audio [B][L] trimmed to L = metadata "al" (EncodecCompressor.cs:288-420)
  void decompress_dev(const uint8_t* payload, int64_t stride, int B, int nq, int64_t L, float* audio_out);

 private:
  struct Act {            // channels-last activation with 8 margin rows on each side of every clip
    float* base = nullptr;  // row 0 of clip 0
    int T = 0, C = 0;
    long long stride = 0;   // floats between clips
  };
  struct Res { ConvLayer shortcut, c3, c1; int hidden_p = 0; };
  struct Lstm { ConvLayer ih[2]; float* whh[2] = {nullptr, nullptr}; int layers = 0; };
  static constexpr int kMargin = 16;   // >= the largest reflect pad: causal k - s <= 8 left, non-causal s/2 + (s - 1) right
  void require_ready() const;
  // GroupNorm affine of one conv (time_group_norm): gamma / beta padded like the conv's output channels
  struct Gn { float* gamma = nullptr; float* beta = nullptr; int c_real = 0; };
  std::unordered_map<std::string, Gn> gn_;
  void load_gn(const std::string& p, int c_real, int c_pad);
  // GroupNorm + follow-up of a conv's raw output `y` (no-op for weight-norm models, whose epilogue already did it):
  // statistics over rows [row0, row0 + rows) (the transposed convs normalise BEFORE their trim), result rows [0, y.T)
  void finish_norm(const ConvLayer& L, const Act& y, int B, int post, const Act* residual, int row0, int rows,
                   const ConvLayer* raw_residual_of = nullptr);
  int gn_slot_ = 0, gn_mb_ = 0;        // statistics slot the next conv accumulates into ([slot][micro-batch][2] fp64)
  double* gn_slot() { return gn_stats_.as<double>() + (size_t)gn_slot_ * 2 * gn_mb_; }
  // one group of equal-length segment items (clip-major: item = b*segs + j -> segment s0 + j) through encoder and / or decoder
  struct Group { int segs, s0; int64_t len, frames, col0; };
  void run_group(const float* audio, int B, int64_t L, const SegLayout& lay, const Group& g, int nq, int64_t* codes_user,
                 float* scales, bool encode, bool decode);
  std::vector<Group> groups_of(const SegLayout& lay) const;
  void check_overlap_add(const SegLayout& lay) const;
  void overlap_add(const SegLayout& lay, int B, float* audio_out, int64_t out_len);
  struct ItemMap { int segs = 1, s0 = 0, item0 = 0, n_seg = 1; const float* scales = nullptr; };
  ItemMap map_;                        // where run_decoder's generic output stage puts its frames
  bool gn_fused_ = false;              // the last conv's epilogue already accumulated its GroupNorm statistics
  bool generic_io_ = false;            // conv_in / conv_out as channel-padded ConvLayers (stereo, non-causal or GroupNorm)
  ConvLayer conv_in_l_, conv_out_l_;
  int cin_pad_ = 32, cout_pad_ = 32;
  DeviceBuffer gn_stats_, scales_, frames_, seg_lens_;
  std::vector<float> folded(const std::string& p, int d0, int d1, int k, std::vector<float>* bias, int bias_n);
  void build_res(Res& r, const std::string& p, int dim);
  void build_lstm(Lstm& l, const std::string& p, int dim);
  Act act(int buf, int B, int T, int C);
  void conv(const ConvLayer& L, const Act& in, int left_pad, int t_in_extra, const Act& out, int B, int prologue, int post,
            const Act* residual);
  Act run_res(const Res& r, const Act& x, int B, int& free_a, int& free_b, int& free_c, bool post_elu);
  // conv + finish_norm: in GroupNorm models the post activation / residual move behind the normalisation
  void conv_n(const ConvLayer& L, const Act& in, int left_pad, int t_in_extra, const Act& out, int B, int prologue, int post,
              const Act* residual, int out_row0 = 0, int stat_rows = -1);
  void conv_short_n(const ConvLayer& L, const Act& in, const SPad& pad, const Act& out, int B, int prologue, int post);
  Act run_lstm(const Lstm& l, const Act& x, int B, int out_buf);
  void run_encoder(const float* audio, int B, int64_t L, int64_t* T_out);   // -> z_ (dense [B][T][128])
  void run_decoder(int B, int T, float* audio_out, long long out_stride);   // zq_ act -> audio
  int micro_batch(int B, int64_t L);
  int pick_free(int a, int b, int c = -1, int d = -1) const;

  EncodecConfig cfg_;
  // The encoder feeds the argmin, and Encodec's later RVQ stages quantise residuals 10-50x smaller than the embedding, so
  // its error budget is ~10x tighter than DAC's.  Round 1 ran it in true fp32 (CUDA cores) because every tensor-core mode
  // flipped codes with margins up to 5e-5; the cause was the tensor core's truncating accumulation over long chains, not
  // the operands.  With short chains (conv_plan.h acc_split) 3xTF32 (22-bit operands at any magnitude) matches the fp32
  // path: 4 near-tie flips in 6000 frames, max margin 2.5e-7 (fp32: 5, 2.6e-7; profiles/r02_encodec_chains.txt).
  // bf16x3 (16-bit operands) still leaves one 3e-6 flip and stays opt-in.
  Precision enc_prec_ = PREC_3XTF32, dec_prec_ = PREC_BF16X3;
  int enc_short_chains_ = 1;   // option encoder_short_chains: 0 off, 1 layers whose chains exceed 96 MMAs, 2 every layer
  int chain_mode(int cin, int k) const;
  float* d_conv_in_w_ = nullptr;
  float* d_conv_in_b_ = nullptr;
  std::vector<std::unique_ptr<Res>> enc_res_, dec_res_;
  std::vector<std::unique_ptr<ConvLayer>> enc_down_, dec_up_;
  Lstm enc_lstm_, dec_lstm_;
  ConvLayer enc_out_, dec_in_;
  float* d_conv_out_w_ = nullptr;
  float* d_conv_out_b_ = nullptr;
  int conv_out_c_ = 0;
  std::vector<float*> embed_, embed_sq_;
  const float** d_embed_ptrs_ = nullptr;
  DeviceBuffer ws_[5], xproj_, z_, hbuf_, barriers_, audio_tmp_, codes_tmp_, pad_tmp_;
  // conv through the short-input branch: materialise Pad1d densely, then a plain valid convolution
  void conv_short(const ConvLayer& L, const Act& in, const SPad& pad, const Act& out, int B, int prologue, int post);
};

}  // namespace nc
