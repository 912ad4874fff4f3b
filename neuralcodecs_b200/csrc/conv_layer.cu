// Lowering of the reference's conv layers to multi-tap row-shifted GEMM plans + weight packing.
//   Conv1d          : Modules/DAC/WNConv1d.cs:152 (functional.conv1d), SNAC/WNConv1d.cs:137,
//                     Encodec/WNConv1d.cs:124
//   ConvTranspose1d : Modules/DAC/WNConvTranspose1d.cs:152-160 (functional.conv_transpose1d),
//                     SNAC/WNConvTranspose1d.cs:134-142, Encodec/WNConvTranspose1d.cs:140-148
#include "conv_layer.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>

#include "umma.cuh"

namespace nc {

Precision parse_precision(const std::string& s) {
  if (s == "fp32") return PREC_FP32;
  if (s == "tf32") return PREC_TF32;
  if (s == "3xtf32") return PREC_3XTF32;
  if (s == "bf16x3") return PREC_BF16X3;
  if (s == "f16x3") return PREC_F16X3;
  if (s == "f16x2") return PREC_F16X2;
  if (s == "f16") return PREC_F16;
  throw Error(NC_INVALID_ARGUMENT, "unknown precision '" + s + "' (fp32|tf32|3xtf32|bf16x3|f16x3)");
}

const char* precision_name(Precision p) {
  switch (p) {
    case PREC_FP32: return "fp32";
    case PREC_TF32: return "tf32";
    case PREC_3XTF32: return "3xtf32";
    case PREC_BF16X3: return "bf16x3";
    case PREC_F16X2: return "f16x2";
    case PREC_F16: return "f16";
    default: return "f16x3";
  }
}

static int mma_mode(Precision p) {
  switch (p) {
    case PREC_TF32: return MODE_TF32;
    case PREC_3XTF32: return MODE_TF32X3;
    case PREC_BF16X3: return MODE_BF16X3;
    default: return MODE_F16X3;
  }
}

// round-to-nearest-even fp32 -> bf16 / fp16 bit patterns on the host (finite inputs)
static inline uint16_t host_bf16(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float host_bf16_to_f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
static inline uint16_t host_f16(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  const uint32_t sign = (u >> 16) & 0x8000u;
  const int32_t e = (int32_t)((u >> 23) & 0xFF) - 127 + 15;
  uint32_t m = u & 0x7FFFFFu;
  if (e >= 31) return (uint16_t)(sign | 0x7C00u);           // overflow -> inf
  if (e <= 0) {                                             // subnormal / zero
    if (e < -10) return (uint16_t)sign;
    m |= 0x800000u;
    const int shift = 14 - e;                               // 14..24
    uint32_t r = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (r & 1u))) ++r;
    return (uint16_t)(sign | r);
  }
  uint32_t r = ((uint32_t)e << 10) | (m >> 13);
  const uint32_t rem = m & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) ++r;   // may carry into the exponent: still correct
  return (uint16_t)(sign | r);
}
static inline float host_f16_to_f(uint16_t h) {
  const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
  const uint32_t e = (h >> 10) & 0x1Fu, m = h & 0x3FFu;
  float f;
  if (e == 0) {
    f = std::ldexp((float)m, -24);
  } else if (e == 31) {
    f = m ? NAN : INFINITY;
  } else {
    f = std::ldexp((float)(m | 0x400u), (int)e - 25);
  }
  uint32_t u;
  std::memcpy(&u, &f, 4);
  u |= sign;
  std::memcpy(&f, &u, 4);
  return f;
}


static inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// cvt.rna.tf32.f32 on the host: round the magnitude to 10 mantissa bits, ties away from zero
static inline float host_rna_tf32(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return x;
  u += 0x1000u;
  u &= 0xFFFFE000u;
  float r;
  std::memcpy(&r, &u, 4);
  return r;
}

ConvLayer::~ConvLayer() {
  cudaFree(d_bias_);
  cudaFree(d_w_plain_);
  cudaFree(d_w_tiles_);
  cudaFree(d_w16_);
}

int ConvLayer::out_len(int t_in) const {
  const ConvSpec& s = spec_;
  if (s.transposed) return (t_in - 1) * s.stride - 2 * s.padding + s.dilation * (s.k - 1) + s.output_padding + 1;
  const int num = t_in + 2 * s.padding - s.dilation * (s.k - 1) - 1;
  return num < 0 ? 0 : num / s.stride + 1;
}

double ConvLayer::flops(int batch, int t_in) const {
  const ConvSpec& s = spec_;
  if (s.transposed) return 2.0 * s.cin * s.cout * s.k * (double)t_in * batch;
  return 2.0 * s.cin * s.cout * s.k * (double)out_len(t_in) * batch;
}

void ConvLayer::build(const std::string& name, const ConvSpec& spec, const std::vector<float>& w,
                      const std::vector<float>& bias, Precision requested, int short_chains, bool direct16) {
  name_ = name;
  short_chains_ = (requested == PREC_BF16X3 || requested == PREC_F16X3 || requested == PREC_3XTF32) ? short_chains : 0;
  spec_ = spec;
  const ConvSpec& s = spec_;
  if ((size_t)s.cin * s.cout * s.k != w.size())
    throw Error(NC_SHAPE_MISMATCH, name + ": weight has " + std::to_string(w.size()) + " elements, expected " +
                                       std::to_string((size_t)s.cin * s.cout * s.k));
  if (!bias.empty() && (int)bias.size() != s.cout) throw Error(NC_SHAPE_MISMATCH, name + ": bias size");
  if (s.transposed && s.dilation != 1) throw Error(NC_UNSUPPORTED, name + ": dilated transposed conv");

  // ---------------------------------------------------------------- logical taps
  std::map<int, std::map<int, int>> groups;  // conv: q -> (r -> kernel index j)
  taps_.clear();
  if (!s.transposed) {
    n_logical_ = s.cout;
    k_view_ = s.stride * s.cin;
    for (int j = 0; j < s.k; ++j) {
      const int u = j * s.dilation - s.padding;
      const int q = floordiv(u, s.stride);
      groups[q][u - q * s.stride] = j;
    }
    for (auto& g : groups) {
      const int rmin = g.second.begin()->first, rmax = g.second.rbegin()->first;
      Tap t;
      t.shift = g.first;
      t.koff = rmin * s.cin;
      t.klen = (rmax - rmin + 1) * s.cin;
      t.w.assign((size_t)n_logical_ * t.klen, 0.f);
      for (auto& rj : g.second) {
        const int r = rj.first, j = rj.second;
        for (int n = 0; n < s.cout; ++n)
          for (int ci = 0; ci < s.cin; ++ci)
            t.w[(size_t)n * t.klen + (size_t)(r - rmin) * s.cin + ci] = w[((size_t)n * s.cin + ci) * s.k + j];
      }
      taps_.push_back(std::move(t));
    }
  } else {
    // out[t*s + phi] += x[t - q] * W[ci][co][phi + p + q*s]  (valid kernel indices only)
    n_logical_ = s.stride * s.cout;
    k_view_ = s.cin;
    for (int q = -8; q <= 8; ++q) {
      bool any = false;
      for (int phi = 0; phi < s.stride; ++phi) {
        const int j = phi + s.padding + q * s.stride;
        if (j >= 0 && j < s.k) any = true;
      }
      if (!any) continue;
      Tap t;
      t.shift = -q;
      t.koff = 0;
      t.klen = s.cin;
      t.w.assign((size_t)n_logical_ * t.klen, 0.f);
      for (int phi = 0; phi < s.stride; ++phi) {
        const int j = phi + s.padding + q * s.stride;
        if (j < 0 || j >= s.k) continue;
        for (int co = 0; co < s.cout; ++co)
          for (int ci = 0; ci < s.cin; ++ci)
            t.w[(size_t)(phi * s.cout + co) * t.klen + ci] = w[((size_t)ci * s.cout + co) * s.k + j];
      }
      taps_.push_back(std::move(t));
    }
    std::sort(taps_.begin(), taps_.end(), [](const Tap& a, const Tap& b) { return a.shift < b.shift; });
  }
  if ((int)taps_.size() > kMaxTaps) throw Error(NC_UNSUPPORTED, name + ": more than 8 GEMM taps");
  n_pad_ = (n_logical_ + 15) / 16 * 16;
  smin_ = taps_.front().shift;
  int smax = smin_;
  for (auto& t : taps_) {
    smin_ = std::min(smin_, t.shift);
    smax = std::max(smax, t.shift);
  }
  span_ = smax - smin_;

  if (!bias.empty()) d_bias_ = upload(bias);

  // ---------------------------------------------------------------- tcgen05 eligibility + tiling
  umma_ok_ = requested != PREC_FP32 && span_ <= 64 && (k_view_ % 4) == 0;
  for (auto& t : taps_) umma_ok_ = umma_ok_ && (t.koff % 32 == 0) && (t.klen % 32 == 0);
  if (umma_ok_) {
    ConvGemmParams probe{};
    probe.span = span_;
    probe.mode = mma_mode(requested);
    probe.w_hi_only = 0;   // sized for full tiles; hi-only tiles only add weight stages
    int bn = 0;
    const int bn_max = short_chains_ == 1 ? 128 : 256;   // short chains: 2 main partials + lo accumulator + running sum in 512 TMEM columns
    if (n_pad_ <= bn_max) {
      bn = n_pad_;
    } else {
      for (int c = bn_max; c >= 16; c -= 16)
        if (n_pad_ % c == 0) { bn = c; break; }
    }
    // shrink until at least 2 weight stages fit
    UmmaLaunch L;
    while (bn >= 16) {
      probe.BN = bn;
      if (n_pad_ % bn == 0 && umma_smem_bytes(probe, &L) != 0 && L.w_stages >= 3) break;
      bn -= 16;
    }
    if (bn < 16 || n_pad_ / bn > kMaxNTiles) {
      umma_ok_ = false;
    } else {
      bn_ = bn;
      n_tiles_ = n_pad_ / bn;
    }
  }
  mode_ = umma_ok_ ? requested : PREC_FP32;

  std::memset(tap_mask_, 0, sizeof tap_mask_);
  if (umma_ok_) {
    kc_begin_ = 1 << 30;
    int kc_end = 0, tile_base = 0;
    for (size_t j = 0; j < taps_.size(); ++j) {
      utaps_[j].shift = taps_[j].shift;
      utaps_[j].kc_lo = taps_[j].koff / 32;
      utaps_[j].kc_hi = (taps_[j].koff + taps_[j].klen) / 32;
      utaps_[j].tile_base = tile_base;
      tile_base += utaps_[j].kc_hi - utaps_[j].kc_lo;
      kc_begin_ = std::min(kc_begin_, utaps_[j].kc_lo);
      kc_end = std::max(kc_end, utaps_[j].kc_hi);
    }
    tiles_per_ntile_ = tile_base;
    n_kc_ = kc_end - kc_begin_;
    // one smem image per (N tile, tap, K chunk): [BN rows][128 B], 16-byte chunks XOR-swizzled by row & 7.
    //   TF32   : 32 tf32-rounded floats per row
    //   TF32X3 : hi image followed by lo image (fetched with one bulk copy)
    //   H16X3  : 32 hi halves (64 B) then 32 lo halves (64 B) per row
    const size_t img = (size_t)bn_ * 32;
    // one- and two-pass fp16 never read the weights' lo halves: ship 64-byte rows (half the L2->smem weight stream,
    // which is what bounds those layers).  Layers that may run inside the fused ResidualUnit kernel (BN <= 128)
    // keep the full layout that kernel expects.
    w_hi_only_ = (mode_ == PREC_F16 || mode_ == PREC_F16X2) && bn_ > 128;
    w_tile_floats_ = (int)(img * (mode_ == PREC_3XTF32 ? 2 : 1)) / (w_hi_only_ ? 2 : 1);
    std::vector<float> tiles((size_t)n_tiles_ * tiles_per_ntile_ * w_tile_floats_, 0.f);
    for (int nt = 0; nt < n_tiles_; ++nt) {
      for (size_t j = 0; j < taps_.size(); ++j) {
        const Tap& t = taps_[j];
        bool any = false;
        for (int kcl = 0; kcl < utaps_[j].kc_hi - utaps_[j].kc_lo; ++kcl) {
          float* th = tiles.data() + ((size_t)nt * tiles_per_ntile_ + utaps_[j].tile_base + kcl) * w_tile_floats_;
          uint16_t* t16 = reinterpret_cast<uint16_t*>(th);
          for (int nl = 0; nl < bn_; ++nl) {
            const int n = nt * bn_ + nl;
            if (n >= n_logical_) continue;
            for (int kk = 0; kk < 32; ++kk) {
              const float v = t.w[(size_t)n * t.klen + (size_t)kcl * 32 + kk];
              if (v != 0.f) any = true;
              if (prec_is_h16(mode_)) {
                const bool bf = mode_ == PREC_BF16X3;
                const uint16_t h = bf ? host_bf16(v) : host_f16(v);
                const float hf = bf ? host_bf16_to_f(h) : host_f16_to_f(h);
                const uint16_t l = bf ? host_bf16(v - hf) : host_f16(v - hf);
                if (w_hi_only_) {
                  t16[ptx::sw64_offset((uint32_t)nl, (uint32_t)(kk / 8)) / 2 + (kk % 8)] = h;
                } else {
                  const size_t hi_off = ptx::sw128_offset((uint32_t)nl, (uint32_t)(kk / 8)) / 2 + (kk % 8);
                  const size_t lo_off = ptx::sw128_offset((uint32_t)nl, 4u + (uint32_t)(kk / 8)) / 2 + (kk % 8);
                  t16[hi_off] = h;
                  t16[lo_off] = l;
                }
              } else {
                const size_t off = ptx::sw128_offset((uint32_t)nl, (uint32_t)(kk / 4)) / 4 + (kk % 4);
                const float h = host_rna_tf32(v);
                th[off] = h;
                if (mode_ == PREC_3XTF32) th[img + off] = host_rna_tf32(v - h);
              }
            }
          }
        }
        if (any) tap_mask_[nt] |= (unsigned char)(1u << j);
      }
    }
    d_w_tiles_ = upload(tiles);
    // dense plan: every tap reads every K chunk, contributes to every N tile, and shifts are equally spaced
    {   // main-chain length per partial: 2 hi*hi MMAs per tap and K chunk; ~96 MMAs (32 main-chain steps) keep the truncation error of one
        // chain near 2e-6 relative (DESIGN.md "accumulation chains")
      static const int steps = getenv("NC_FOLD_STEPS") ? std::max(6, atoi(getenv("NC_FOLD_STEPS"))) : 96;   // MMAs (all three products) per partial
      // main-chain MMAs per tap and 32-channel chunk: 2 (K = 16 halves) or 4 (K = 8 tf32)
      fold_kc_ = std::max(1, steps / ((requested == PREC_3XTF32 ? 12 : 6) * (int)taps_.size()));
    }
    dense_step_ = taps_.size() == 1 ? 0 : taps_[1].shift - taps_[0].shift;
    for (size_t j = 0; j < taps_.size(); ++j) {
      if (utaps_[j].kc_lo != kc_begin_ || utaps_[j].kc_hi != kc_begin_ + n_kc_ ||
          taps_[j].shift != smin_ + (int)j * dense_step_)
        dense_step_ = -1;
    }
    for (int nt = 0; nt < n_tiles_ && dense_step_ >= 0; ++nt)
      if (tap_mask_[nt] != (unsigned char)((1u << taps_.size()) - 1)) dense_step_ = -1;
  }
  // plain weights for the CUDA-core executor: fp32 layers, and strided convs whose input length may
  // not be a whole number of rows (the TMA-fed tensor-core kernel needs whole rows per clip)
  if (!umma_ok_ || (!s.transposed && s.stride > 1)) {
    // plain [tap][n_pad][klen]
    size_t total = 0;
    for (size_t j = 0; j < taps_.size(); ++j) {
      simt_w_off_[j] = (long long)total;
      total += (size_t)n_pad_ * taps_[j].klen;
    }
    std::vector<float> plain(total, 0.f);
    for (size_t j = 0; j < taps_.size(); ++j)
      std::memcpy(plain.data() + simt_w_off_[j], taps_[j].w.data(), taps_[j].w.size() * sizeof(float));
    d_w_plain_ = upload(plain);
  }
  // ---------------------------------------------------------------- fp16-operand executor (conv_h16.cu)
  direct16_ = false;
  if (direct16 && umma_ok_ && mode_ == PREC_F16 && n_pad_ == n_logical_) {
    bool ok = span_ <= 64 && k_view_ % 64 == 0 && (s.transposed || s.stride == 1);
    // transposed convs: whole output rows only (out_len(t) == t * stride), i.e. not the one-sample-short odd strides
    if (s.transposed) ok = ok && s.k - 2 * s.padding + s.output_padding == s.stride;
    for (auto& t : taps_) ok = ok && t.koff % 64 == 0 && t.klen % 64 == 0;
    int bn = 0;
    for (int c = 256; c >= 64; c -= 64)    // epilogue stages are 64 columns wide
      if (n_logical_ % c == 0) { bn = c; break; }
    ok = ok && bn > 0 && n_logical_ / bn <= kMaxNTiles && s.cout % 32 == 0;
    if (ok) {
      bn16_ = bn;
      n_tiles16_ = n_logical_ / bn;
      kc_begin16_ = 1 << 30;
      int kc_end = 0, tile_base = 0;
      for (size_t j = 0; j < taps_.size(); ++j) {
        utaps16_[j].shift = taps_[j].shift;
        utaps16_[j].kc_lo = taps_[j].koff / 64;
        utaps16_[j].kc_hi = (taps_[j].koff + taps_[j].klen) / 64;
        utaps16_[j].tile_base = tile_base;
        tile_base += utaps16_[j].kc_hi - utaps16_[j].kc_lo;
        kc_begin16_ = std::min(kc_begin16_, utaps16_[j].kc_lo);
        kc_end = std::max(kc_end, utaps16_[j].kc_hi);
      }
      tiles_per_ntile16_ = tile_base;
      n_kc16_ = kc_end - kc_begin16_;
      // [N tile][tap][64-channel chunk] tiles of [BN rows][64 halves], plain row-major (the TMA applies the swizzle)
      std::vector<uint16_t> w16((size_t)n_tiles16_ * tiles_per_ntile16_ * bn16_ * 64, 0);
      for (int nt = 0; nt < n_tiles16_; ++nt)
        for (size_t j = 0; j < taps_.size(); ++j)
          for (int kcl = 0; kcl < utaps16_[j].kc_hi - utaps16_[j].kc_lo; ++kcl) {
            uint16_t* tile = w16.data() + ((size_t)nt * tiles_per_ntile16_ + utaps16_[j].tile_base + kcl) * bn16_ * 64;
            for (int nl = 0; nl < bn16_; ++nl)
              for (int kk = 0; kk < 64; ++kk)
                tile[(size_t)nl * 64 + kk] = host_f16(taps_[j].w[(size_t)(nt * bn16_ + nl) * taps_[j].klen + (size_t)kcl * 64 + kk]);
          }
      std::vector<float> packed(w16.size() / 2);
      std::memcpy(packed.data(), w16.data(), w16.size() * 2);
      d_w16_ = upload(packed);
      direct16_ = true;
    }
  }
  for (auto& t : taps_) std::vector<float>().swap(t.w);
}

// Plan for the fp16-operand executor; false when this layer / call cannot use it.
bool ConvLayer::fill_h16(const ConvRunArgs& a, ConvGemmParams* pp, int fast_sin) const {
  const ConvSpec& s = spec_;
  const int t_out = out_len(a.t_in);
  if (!direct16_ || !a.in16 || a.batch <= 0 || t_out <= 0 || a.prologue != PRO_NONE || a.noise || a.dw_w) return false;
  if (!s.transposed && s.stride != 1) return false;
  const int a_rows = a.t_in;
  const int m_rows = s.transposed ? (t_out + s.stride - 1) / s.stride : t_out;
  const int n_total = s.transposed ? s.stride * s.cout : s.cout;
  const long long a_valid = (long long)a.t_in * s.cin, d_valid = (long long)t_out * s.cout;
  if (d_valid != (long long)m_rows * n_total) return false;   // transposed convs one sample short (odd strides)
  ConvGemmParams& p = *pp;
  p = ConvGemmParams{};
  p.a16 = 1;
  p.A = static_cast<const float*>(a.in16); p.a_clip_stride = a.in_clip_stride ? a.in_clip_stride : a_valid;
  p.a_rows = a_rows; p.a_pitch = k_view_; p.a_valid = a_valid;
  p.D = a.out; p.D16 = a.out16; p.R = a.residual; p.d_clip_stride = a.out_clip_stride ? a.out_clip_stride : d_valid;
  p.m_rows = m_rows; p.n_total = n_total; p.n_valid = n_logical_; p.d_valid = d_valid;
  p.bias = d_bias_; p.bias_period = s.cout;
  p.post = a.post; p.post_alpha = a.post_alpha; p.post_inv_alpha = a.post_inv_alpha; p.post_period = s.cout; p.act = a.act;
  p.W = d_w16_; p.w_tile_floats = bn16_ * 32; p.BN = bn16_; p.n_tiles = n_tiles16_; p.tiles_per_ntile = tiles_per_ntile16_;
  p.mode = MODE_F16X3; p.passes = 1;
  p.n_taps = (int)taps_.size();
  for (int j = 0; j < p.n_taps; ++j) p.taps[j] = utaps16_[j];
  // per-N-tile tap masks: a transposed conv's tap j (input shift -q) feeds output phase phi iff kernel index
  // phi + padding + q*stride is valid (ConvLayer::build); a tile of BN columns spans BN / cout phases (or part of one)
  for (int nt = 0; nt < n_tiles16_; ++nt) {
    unsigned m = 0;
    for (int j = 0; j < p.n_taps; ++j) {
      bool any = !s.transposed;
      if (s.transposed) {
        const int q = -utaps16_[j].shift;
        const int phi_lo = (nt * bn16_) / s.cout, phi_hi = ((nt + 1) * bn16_ - 1) / s.cout;
        for (int phi = phi_lo; phi <= phi_hi && !any; ++phi) {
          const int jj = phi + s.padding + q * s.stride;
          if (jj >= 0 && jj < s.k) any = true;
        }
      }
      if (any) m |= 1u << j;
    }
    p.tap_mask[nt] = (unsigned char)m;
  }
  p.n_kc = n_kc16_; p.kc_begin = kc_begin16_; p.smin = smin_; p.span = span_;
  p.dense_step = taps_.size() == 1 ? 0 : taps_[1].shift - taps_[0].shift;
  for (size_t j = 0; j < taps_.size(); ++j)
    if (utaps16_[j].kc_lo != kc_begin16_ || utaps16_[j].kc_hi != kc_begin16_ + n_kc16_ || taps_[j].shift != smin_ + (int)j * p.dense_step)
      p.dense_step = -1;
  for (int nt = 0; nt < n_tiles16_ && p.dense_step >= 0; ++nt)
    if (p.tap_mask[nt] != (unsigned char)((1u << taps_.size()) - 1)) p.dense_step = -1;
  p.batch = a.batch; p.m_tiles_per_clip = (m_rows + 127) / 128;
  const int fast = fast_sin >= 0 ? fast_sin : 1;
  p.precise_sin = fast ? 0 : 1;
  return h16_supported(p);
}

// Fill the tcgen05 plan for one call; returns false when this layer / call cannot use the tensor-core kernel.
bool ConvLayer::fill_umma(const ConvRunArgs& a, ConvGemmParams* pp, int fast_sin) const {
  const ConvSpec& s = spec_;
  const int t_out = out_len(a.t_in);
  if (mode_ == PREC_FP32 || a.batch <= 0 || t_out <= 0) return false;
  int a_rows, m_rows, n_total;
  const long long a_valid = (long long)a.t_in * s.cin;
  long long d_valid;
  if (!s.transposed) {
    a_rows = (a.t_in + s.stride - 1) / s.stride; m_rows = t_out; n_total = s.cout; d_valid = (long long)t_out * s.cout;
  } else {
    a_rows = a.t_in; m_rows = (t_out + s.stride - 1) / s.stride; n_total = s.stride * s.cout; d_valid = (long long)t_out * s.cout;
  }
  const long long a_stride = a.in_clip_stride ? a.in_clip_stride : a_valid;
  const long long d_stride = a.out_clip_stride ? a.out_clip_stride : d_valid;
  const bool whole_rows = a_valid == (long long)a_rows * k_view_ && (a_stride % 4) == 0 &&
                          (reinterpret_cast<uintptr_t>(a.in) & 15) == 0;
  if (!whole_rows && d_w_plain_) return false;
  ConvGemmParams& p = *pp;
  p = ConvGemmParams{};
  p.A = a.in; p.a_clip_stride = a_stride; p.a_rows = a_rows; p.a_pitch = k_view_; p.a_valid = a_valid;
  p.D = a.out; p.R = a.residual; p.d_clip_stride = d_stride; p.m_rows = m_rows; p.n_total = n_total;
  p.n_valid = n_logical_; p.d_valid = d_valid;
  p.bias = d_bias_; p.bias_period = s.cout; p.noise = a.noise; p.gn_stats = a.gn_stats;
  p.alpha = a.alpha; p.inv_alpha = a.inv_alpha; p.alpha_period = s.cin; p.prologue = a.prologue; p.act = a.act;
  p.post = a.post; p.post_alpha = a.post_alpha; p.post_inv_alpha = a.post_inv_alpha; p.post_period = s.cout;
  p.W = d_w_tiles_; p.w_tile_floats = w_tile_floats_; p.BN = bn_; p.n_tiles = n_tiles_; p.tiles_per_ntile = tiles_per_ntile_;
  p.mode = mma_mode(mode_);
  p.passes = mode_ == PREC_F16 ? 1 : (mode_ == PREC_F16X2 ? 2 : 3);
  p.w_hi_only = w_hi_only_ ? 1 : 0;
  p.acc_split = (short_chains_ && p.passes == 3) ? ((short_chains_ == 2 && bn_ > 128) ? 2 : 1) : 0;
  p.fold_kc = (p.acc_split == 1 && short_chains_ == 1 && fold_kc_ < n_kc_) ? fold_kc_ : 0;
  p.n_taps = (int)taps_.size();
  for (int j = 0; j < p.n_taps; ++j) p.taps[j] = utaps_[j];
  std::memcpy(p.tap_mask, tap_mask_, sizeof tap_mask_);
  p.n_kc = n_kc_; p.kc_begin = kc_begin_; p.smin = smin_; p.span = span_;
  p.dense_step = dense_step_;
  if (a.dw_w) {
    // fused depthwise prologue: plain 1x1 stride-1 conv only; the A tile carries a halo of 3*dil rows on each side and
    // the transform warps leave the GEMM operand in stage rows [0, 128) (tap 0 of the dense path reads from row 0)
    if (s.k != 1 || s.stride != 1 || s.transposed || p.n_taps != 1 || a.prologue != PRO_SNAKE || 6 * a.dw_dil > 64) return false;
    p.dw_w = a.dw_w; p.dw_b = a.dw_b; p.dw_post_alpha = a.dw_post_alpha; p.dw_post_inv_alpha = a.dw_post_inv_alpha;
    p.dw_dil = a.dw_dil;
    p.smin = -3 * a.dw_dil; p.span = 6 * a.dw_dil; p.dense_step = 0;
  }
  p.batch = a.batch; p.m_tiles_per_clip = (m_rows + 127) / 128;
  const int fast = fast_sin >= 0 ? fast_sin : ((mode_ == PREC_TF32 || mode_ == PREC_BF16X3 || mode_ == PREC_F16X2 || mode_ == PREC_F16) ? 1 : 0);
  p.precise_sin = fast ? 0 : 1;
  return true;
}


// y = c2(post1(c1(pro(x)))) + x [-> post2] in one launch when the shapes allow it (C <= 256, bf16x3 / f16x3).
bool try_run_ru_fused(const ConvLayer& c1, const ConvLayer& c2, const ConvRunArgs& a1, const ConvRunArgs& a2,
                      const LaunchCtx& ctx) {
  if (!ctx.fuse_ru) return false;
  ConvGemmParams p, p2;
  if (!c1.fill_umma(a1, &p, ctx.fast_sin) || !c2.fill_umma(a2, &p2, ctx.fast_sin)) return false;
  if (!ru_fused_supported(p, p2)) return false;
  const int ev = ctx.begin();
  const int rc = launch_ru_fused(p, p2, ctx.num_sms, ctx.stream);
  if (rc < 0) return false;
  check_launch(rc, c1.name().c_str());
  const double fl = c1.flops(a1.batch, a1.t_in) + c2.flops(a1.batch, a1.t_in);
  const double bytes = 4.0 * a1.batch * (double)a1.t_in * c1.spec().cin * 3;   // read x (operand + residual), write y
  ctx.end(ev, std::string("ru_fused_") + precision_name(c1.precision()), fl, bytes, c1.name());
  return true;
}

void ConvLayer::run(const ConvRunArgs& a, const LaunchCtx& ctx) const {
  const ConvSpec& s = spec_;
  const int t_out = out_len(a.t_in);
  if (a.batch <= 0 || t_out <= 0) return;
  int a_rows, m_rows, n_total;
  long long a_valid = (long long)a.t_in * s.cin, d_valid;
  if (!s.transposed) {
    a_rows = (a.t_in + s.stride - 1) / s.stride;
    m_rows = t_out;
    n_total = s.cout;
    d_valid = (long long)t_out * s.cout;
  } else {
    a_rows = a.t_in;
    m_rows = (t_out + s.stride - 1) / s.stride;
    n_total = s.stride * s.cout;
    d_valid = (long long)t_out * s.cout;
  }
  const int m_tiles = (m_rows + 127) / 128;
  const double fl = flops(a.batch, a.t_in);
  const double bytes = 4.0 * a.batch * ((double)a_valid + (double)d_valid * (a.residual ? 2 : 1));
  const int ev = ctx.begin();
  const long long a_stride = a.in_clip_stride ? a.in_clip_stride : a_valid;
  const long long d_stride = a.out_clip_stride ? a.out_clip_stride : d_valid;
  ConvGemmParams up;
  if (a.in16) {   // fp16-operand executor: no fallback (the fp32 executors cannot read this input)
    if (!fill_h16(a, &up, ctx.fast_sin))
      throw Error(NC_INTERNAL, name_ + ": fp16-operand call on a layer / shape the fp16 executor does not support");
    check_launch(launch_conv_h16(up, ctx.num_sms, ctx.stream), name_.c_str());
    const double b16 = (double)a.batch * (2.0 * a_valid + ((a.out ? 4.0 : 0.0) + (a.out16 ? 2.0 : 0.0) + (a.residual ? 4.0 : 0.0)) * d_valid);
    ctx.end(ev, "conv_h16", fl, b16, name_);
    return;
  }
  if (a.gn_stats_done) *a.gn_stats_done = false;
  if (fill_umma(a, &up, ctx.fast_sin)) {
    check_launch(launch_conv_umma(up, ctx.num_sms, ctx.stream), name_.c_str());
    if (a.gn_stats_done) *a.gn_stats_done = a.gn_stats != nullptr;
    ctx.end(ev, std::string(a.dw_w ? "conv_umma_dw_" : "conv_umma_") + precision_name(mode_),
            fl + (a.dw_w ? 2.0 * 7 * s.cin * (double)a.t_in * a.batch : 0.0), bytes, name_);
  } else {
    if (a.dw_w) throw Error(NC_INTERNAL, "depthwise-fused conv is only available on the tcgen05 path");
    ConvSimtParams p{};
    p.A = a.in; p.a_clip_stride = a_stride; p.a_rows = a_rows; p.a_pitch = k_view_; p.a_valid = a_valid;
    p.D = a.out; p.R = a.residual; p.d_clip_stride = d_stride; p.m_rows = m_rows; p.n_total = n_total;
    p.n_valid = n_logical_; p.d_valid = d_valid;
    p.bias = d_bias_; p.bias_period = s.cout; p.noise = a.noise;
    p.alpha = a.alpha; p.alpha_period = s.cin; p.prologue = a.prologue; p.act = a.act;
    p.post = a.post; p.post_alpha = a.post_alpha; p.post_period = s.cout;
    p.W = d_w_plain_; p.n_pad = n_pad_;
    p.n_taps = (int)taps_.size();
    for (int j = 0; j < p.n_taps; ++j) {
      p.taps[j].shift = taps_[j].shift; p.taps[j].koff = taps_[j].koff; p.taps[j].klen = taps_[j].klen;
      p.taps[j].w_off = simt_w_off_[j];
    }
    p.mask_bn = 0;
    p.batch = a.batch; p.m_tiles_per_clip = m_tiles;
    check_launch(launch_conv_simt(p, ctx.stream), name_.c_str());
    ctx.end(ev, "conv_simt_fp32", fl, bytes, name_);
  }
}

}  // namespace nc
