// A weight-normalised Conv1d / ConvTranspose1d of the reference, lowered once at load time to a
// multi-tap row-shifted GEMM plan (conv_plan.h) with packed device weights.
#pragma once
#include <string>
#include <vector>

#include "conv_plan.h"
#include "runtime.h"

namespace nc {

// *X3 = three tensor-core products per term (hi*hi + lo*hi + hi*lo, ~16-22 operand bits); F16X2 = activations split in
// two halves, weights rounded once to fp16 (11 bits); F16 = one fp16 product (11-bit operands, like TF32 at twice the rate).
enum Precision : int { PREC_FP32 = 0, PREC_TF32 = 1, PREC_3XTF32 = 3, PREC_BF16X3 = 4, PREC_F16X3 = 5, PREC_F16X2 = 6, PREC_F16 = 7 };
inline bool prec_is_h16(Precision p) { return p == PREC_BF16X3 || p == PREC_F16X3 || p == PREC_F16X2 || p == PREC_F16; }
const char* precision_name(Precision p);

Precision parse_precision(const std::string& s);

struct ConvSpec {
  bool transposed = false;
  int cin = 0, cout = 0, k = 1, stride = 1, dilation = 1, padding = 0, output_padding = 0;
};

struct ConvRunArgs {
  const float* in = nullptr;        // [B][T_in][Cin] channels-last
  float* out = nullptr;             // [B][T_out][Cout]
  // fp16-operand path (layers built with direct16): in16 = fp16 [B][T_in][Cin] already carrying this conv's input
  // activation; out (raw fp32, bias + residual) and / or out16 (fp16, `post` applied) are written
  const void* in16 = nullptr;
  void* out16 = nullptr;
  const float* residual = nullptr;  // same shape as out
  const float* noise = nullptr;     // [B][T_out] (SNAC NoiseBlock: out = residual + noise * conv)
  // Encodec time_group_norm: [B][2] fp64 {sum, sum of squares} of conv + bias over each clip's outputs, ADDED to by the
  // tcgen05 epilogue; *gn_stats_done tells the caller whether the executor did it (the CUDA-core fallback does not)
  double* gn_stats = nullptr;
  bool* gn_stats_done = nullptr;
  int batch = 0;
  int t_in = 0;
  int prologue = PRO_NONE;
  const float* alpha = nullptr;      // device [Cin]
  const float* inv_alpha = nullptr;  // device [Cin]
  int act = ACT_NONE;
  // activation of the NEXT layer applied in this layer's epilogue (after bias + residual)
  int post = PRO_NONE;
  const float* post_alpha = nullptr;      // device [Cout]
  const float* post_inv_alpha = nullptr;  // device [Cout]
  // floats between consecutive clips of the input / output (+ residual) tensors; 0 = dense (T*C).
  // Lets callers keep margin rows around each clip (Encodec's reflect padding is materialised there).
  long long in_clip_stride = 0;
  long long out_clip_stride = 0;
  // depthwise k7 conv (+ Snake after it) evaluated inside this 1x1 conv's operand prologue (tcgen05 path only):
  // in -> prologue Snake -> depthwise(dw_w [7][Cin], dw_b, dilation dw_dil) -> Snake(dw_post_*) -> this conv
  const float* dw_w = nullptr;
  const float* dw_b = nullptr;
  const float* dw_post_alpha = nullptr;
  const float* dw_post_inv_alpha = nullptr;
  int dw_dil = 1;
};

class ConvLayer {
 public:
  ConvLayer() = default;
  ConvLayer(const ConvLayer&) = delete;
  ConvLayer& operator=(const ConvLayer&) = delete;
  ~ConvLayer();

  // w: folded weights, conv [Cout][Cin][k] / transposed [Cin][Cout][k]; bias [Cout] or empty.
  // short_chains: keep the tensor core's accumulation chains short (conv_plan.h: acc_split = 1 | 2) -- for layers whose
  // output feeds an RVQ argmin; three-pass 16-bit modes only; 1 caps the N tile at 128 columns
  void build(const std::string& name, const ConvSpec& spec, const std::vector<float>& w,
             const std::vector<float>& bias, Precision requested, int short_chains = 0, bool direct16 = false);
  bool direct16() const { return direct16_; }
  bool fill_h16(const ConvRunArgs& a, ConvGemmParams* p, int fast_sin = -1) const;
  int out_len(int t_in) const;
  void run(const ConvRunArgs& a, const LaunchCtx& ctx) const;
  // fast_sin: LaunchCtx::fast_sin of the calling handle
  bool fill_umma(const ConvRunArgs& a, ConvGemmParams* p, int fast_sin = -1) const;

  const ConvSpec& spec() const { return spec_; }
  Precision precision() const { return mode_; }
  const std::string& name() const { return name_; }
  double flops(int batch, int t_in) const;
  // "tcgen05 tf32 BN=256" / "tcgen05 3xtf32 BN=128" / "simt fp32"
  std::string executor() const {
    if (mode_ == PREC_FP32) return "simt fp32";
    if (direct16_) return "tcgen05 f16 fp16-operands BN=" + std::to_string(bn16_);
    if (short_chains_) return std::string("tcgen05 ") + precision_name(mode_) + " BN=" + std::to_string(bn_) + (short_chains_ == 1 ? " folded" : " lo-acc");
    return std::string("tcgen05 ") + precision_name(mode_) + " BN=" + std::to_string(bn_);
  }

 private:
  struct Tap {
    int shift, koff, klen;
    std::vector<float> w;  // [n_logical][klen]
  };
  std::string name_;
  ConvSpec spec_;
  Precision mode_ = PREC_FP32;
  int n_logical_ = 0;  // GEMM N: Cout, or stride*Cout for transposed
  int n_pad_ = 0;      // rounded up to 16
  int k_view_ = 0;     // floats per A-view row
  std::vector<Tap> taps_;
  // device
  float* d_bias_ = nullptr;
  float* d_w_plain_ = nullptr;
  float* d_w_tiles_ = nullptr;
  int w_tile_floats_ = 0;
  bool w_hi_only_ = false;
  int short_chains_ = 0;   // 0 off, 1 = folded partials + lo accumulator (N tile <= 128), 2 = lo accumulator only (N tile <= 256)
  int fold_kc_ = 0;
  // fp16-operand executor: own tiling (64-channel K chunks, plain row-major fp16 weight tiles)
  bool direct16_ = false;
  float* d_w16_ = nullptr;
  int bn16_ = 0, n_tiles16_ = 0, tiles_per_ntile16_ = 0, kc_begin16_ = 0, n_kc16_ = 0;
  ConvTap utaps16_[kMaxTaps];
  // UMMA tiling
  int bn_ = 0, n_tiles_ = 0, tiles_per_ntile_ = 0;
  ConvTap utaps_[kMaxTaps];
  unsigned char tap_mask_[kMaxNTiles];
  int kc_begin_ = 0, n_kc_ = 0, smin_ = 0, span_ = 0, dense_step_ = -1;
  long long simt_w_off_[kMaxTaps];
  bool umma_ok_ = false;
};

// Whole ResidualUnit (k7 conv -> Snake -> 1x1 conv -> + x [-> following Snake]) in one launch when supported;
// a1: args of the k7 conv with out = the UNIT's output and residual = the unit's input; a2: the 1x1 conv's post.
bool try_run_ru_fused(const ConvLayer& c1, const ConvLayer& c2, const ConvRunArgs& a1, const ConvRunArgs& a2,
                      const LaunchCtx& ctx);

}  // namespace nc
