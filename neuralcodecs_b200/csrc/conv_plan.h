// Multi-tap row-shifted GEMM: the one formulation every conv on the codec hot path is
// lowered to.  Activations live in HBM channels-last, [clip][time][channel] fp32.
//
//   D[b, m, n] = act( post( bias[n % bias_period]
//                           + sum_taps sum_{k in [kc_lo*32, kc_hi*32)} f(A[b, m + shift_tap, k]) * W_tap[n, k]
//                           (+ R[b, m, n]) ) )
//
// where the "A view" re-reads the channels-last input as rows of `a_pitch` floats:
//   stride-1 conv (k taps, dilation d, padding p) : row = time step,  shift_j = j*d - p, K = Cin
//   stride-s conv (kernel 2s)                      : row = s time steps (s*Cin floats); the kernel
//                                                    taps fold into <=3 super-taps with partial K
//   transposed conv (kernel 2s, stride s)          : row = input time step, N = s*Cout (phase-major),
//                                                    <=3 super-taps, some masked per N tile
// f    = Snake / ELU prologue of THIS conv (alpha indexed k % alpha_period), applied once per
//        staged input element.  Out-of-range rows read as zero (f(0) = 0 for both prologues).
// post = Snake / ELU of the NEXT layer (alpha indexed n % post_period) applied in the epilogue, so
//        the consumer needs no prologue: used whenever the raw value has no other reader.
#pragma once
#include <cstdint>

namespace nc {

constexpr int kMaxTaps = 8;
constexpr int kMaxNTiles = 64;

enum Prologue : int { PRO_NONE = 0, PRO_SNAKE = 1, PRO_ELU = 2 };
enum Activation : int { ACT_NONE = 0, ACT_TANH = 1 };
// tensor-core operand modes of the tcgen05 executor
//   TF32   : one kind::tf32 MMA per K step                                  (10-bit mantissa operands)
//   TF32X3 : hi/lo TF32 split of both operands, 3 MMAs (lo*hi, hi*lo, hi*hi)  (~21 bits)
//   BF16X3 : hi/lo BF16 split packed in one 128-byte row (32 hi | 32 lo), 3 kind::f16 MMAs of K=16
//            per 16 channels -> 1.5x the MMA time of TF32, ~16-bit operands
//   F16X3  : same with FP16 halves (~22 bits while values stay in fp16's normal range)
enum MmaMode : int { MODE_TF32 = 0, MODE_TF32X3 = 1, MODE_BF16X3 = 2, MODE_F16X3 = 3 };

struct ConvTap {
  int shift;      // row shift in the A view
  int kc_lo;      // first 32-float K chunk of the A view this tap reads
  int kc_hi;      // one past the last chunk
  int tile_base;  // index of this tap's first weight tile inside an N tile's tile list
};

struct ConvGemmParams {
  // ---- A view (input)
  const float* A;
  long long a_clip_stride;  // floats between clips
  int a_rows;               // rows per clip
  int a_pitch;              // floats per row
  long long a_valid;        // valid floats per clip (guards a partial last row)
  // ---- output / residual (same layout)
  float* D;
  const float* R;           // nullable
  long long d_clip_stride;
  int m_rows;               // output rows per clip
  int n_total;              // floats per output row
  int n_valid;              // logical N (<= n_tiles*BN)
  long long d_valid;        // valid floats per clip
  const float* bias;        // nullable
  int bias_period;
  const float* noise;       // nullable; [clip][m_rows]: D = R + noise*acc (SNAC NoiseBlock)
  double* gn_stats;         // nullable; [clip][2]: the epilogue adds sum / sum of squares of (acc + bias) over the clip's valid
                            // outputs (Encodec time_group_norm: the statistics of the GroupNorm that follows the conv)
  // ---- prologue / post-activation / activation
  const float* alpha;       // [alpha_period]
  const float* inv_alpha;   // [alpha_period] (1/alpha, 0 where alpha == 0)
  int alpha_period;
  int prologue;
  const float* post_alpha;      // [post_period] or null
  const float* post_inv_alpha;
  int post_period;
  int post;                 // Prologue enum applied after bias (+ residual)
  int act;
  // ---- weights: pre-swizzled smem images, [n_tile][tiles_per_ntile][w_tile_floats]
  const float* W;
  int w_tile_floats;        // BN*32 (TF32, BF16X3, F16X3) or 2*BN*32 (TF32X3: hi image then lo image)
  int BN;                   // N tile (multiple of 16, <= 256)
  int n_tiles;
  int tiles_per_ntile;
  int mode;                 // MmaMode
  // Optional depthwise k7 conv folded into the operand prologue of a 1x1 conv (SNAC ResidualUnit):
  //   operand[t, c] = snake2(dw_b[c] + sum_j dw_w[j][c] * snake1(x[t + (j-3)*dw_dil, c]))      (zero padding)
  // snake1 = alpha / inv_alpha above (prologue must be PRO_SNAKE), snake2 = dw_post_*.  smin / span describe the halo.
  const float* dw_w;        // [7][Cin] (null = no depthwise stage)
  const float* dw_b;        // [Cin] or null
  const float* dw_post_alpha; const float* dw_post_inv_alpha;   // [Cin] or null
  int dw_dil;
  int w_hi_only;            // 16-bit modes with passes < 3: weight tiles are [BN][32 hi halves] = 64-byte rows (SWIZZLE_64B)
  int passes;               // 16-bit modes: 3 = hi*hi + lo*hi + hi*lo, 2 = hi*hi + lo(A)*hi, 1 = hi*hi only
  // Short accumulation chains (three-pass 16-bit modes, BN <= 128, TMA epilogue).  The tensor core adds every MMA's
  // K=16 partial dot product to the fp32 accumulator with truncation (measured: relative error ~4e-8 per accumulation
  // step, always toward zero, i.e. 4e-5 over the ~700-1500 MMAs of the encoder's deep layers -- enough to flip RVQ
  // codes whose margin is far above the 1e-6 near-tie gate).  acc_split = 1 keeps the chains short: hi*hi products go
  // to a "main" partial accumulator that is folded into a running fp32 sum (round-to-nearest adds by the epilogue
  // warps, kept in TMEM) every fold_kc K chunks; the small lo*hi and hi*lo products (2^-8 of the main terms, so
  // their truncation error is negligible) go to a separate accumulator added once at the end.
  // acc_split = 2: N tiles up to 256 columns -- main accumulator at TMEM columns [0,BN), lo-term accumulator at
  // [256,256+BN), no folding and one tile in flight (the epilogue no longer overlaps the next tile's MMAs).
  int acc_split;
  int fold_kc;              // K chunks per main partial (0 = the whole tile in one partial)
  // fp16-operand executor (conv_h16.cu): A holds fp16 activations (a_pitch / a_clip_stride / a_valid count halves; taps'
  // kc_*, n_kc and kc_begin count 64-channel chunks; W = plain row-major fp16 tiles [BN][64]); D = raw fp32 output
  // (bias + residual, nullable), D16 = post-activated fp16 output in the consumer's operand format (nullable).
  int a16;
  void* D16;
  int n_taps;
  ConvTap taps[kMaxTaps];
  unsigned char tap_mask[kMaxNTiles];  // bit j set => tap j contributes to this N tile
  int n_kc;                 // K chunks of the A view touched by any tap: [kc_begin, kc_begin + n_kc)
  int kc_begin;
  int dense_step;           // >= 0: every tap covers every K chunk of every N tile and shift_j = smin + j*dense_step
                            // (stride-1 convs: dense_step = dilation); -1: general plan (use taps[] / tap_mask[])
  int smin;                 // min shift over taps
  int span;                 // max shift - min shift
  // ---- grid
  int batch;
  int m_tiles_per_clip;
  int precise_sin;          // 1: Cody-Waite + polynomial sin in Snake; 0: MUFU sin (abs err ~4e-7)
};

// ---- fp32 CUDA-core executor (conv_simt.cu): general tap extents, plain weights
struct SimtTap {
  int shift;
  int koff;         // floats into the A-view row
  int klen;         // floats
  long long w_off;  // floats into W: this tap's [N_pad][klen] block
};

struct ConvSimtParams {
  const float* A; long long a_clip_stride; int a_rows; int a_pitch; long long a_valid;
  float* D; const float* R; long long d_clip_stride; int m_rows; int n_total; int n_valid; long long d_valid;
  const float* bias; int bias_period;
  const float* noise;
  const float* alpha; int alpha_period; int prologue;
  const float* post_alpha; int post_period; int post;
  int act;
  const float* W; int n_pad;
  int n_taps; SimtTap taps[kMaxTaps];
  unsigned char tap_mask[kMaxNTiles]; int mask_bn;  // mask granularity in columns (0 = no masks)
  int batch; int m_tiles_per_clip;
};

struct UmmaLaunch {
  int a_stages;
  int w_stages;
  int a_rows_alloc;  // multiple of 8
  int tma_epilogue;  // 1: output / residual tiles move by TMA through a swizzled smem ring
  int epi_stages;    // conv_h16 / fused unit: depth of the epilogue's smem ring (residual prefetch / stores in flight)
  int epi_teams;     // conv_umma_kernel: 1 = the 8 epilogue warps work as two teams of 4 on alternate 32-column groups
  int dual_issue;    // conv_umma_kernel: 1 = a second MMA-issuing warp takes the odd tiles (see mma_role)
  int h_stages;      // fused unit: depth of the ring of 1x1-conv operand tiles written by the acc1 drain
  unsigned long long* trace;   // measurement only (env NC_TRACE_RU): CTA 0 writes clock64() at role events of its first tiles
  int knock;         // measurement only (env NC_KNOCK, results become WRONG): 1 = weight copies skipped after the ring
                     // filled once, 2 = operand transform skipped, 4 = A loads skipped, 8 = MMAs skipped, 16 = epilogue math/stores skipped
};

}  // namespace nc

#ifdef __CUDACC__
#include <cuda_runtime.h>
namespace nc {
// return 0 on success, a cudaError_t (>0) on launch failure, -1 when the shape does not fit
int launch_conv_simt(const ConvSimtParams& p, cudaStream_t stream);
int launch_conv_umma(const ConvGemmParams& p, int num_sms, cudaStream_t stream);
// dynamic smem the tcgen05 kernel needs for this plan (0 = does not fit); fills L
size_t umma_smem_bytes(const ConvGemmParams& p, UmmaLaunch* L);
// whether the A view of p can be fetched by TMA (whole rows per clip, 16-byte aligned strides)
bool umma_view_ok(const ConvGemmParams& p);
// Whole DAC-style ResidualUnit in one launch: p = the k7 conv's plan (p.D = unit output, p.R = unit input,
// p.post = Snake2), p2 = the 1x1 conv's plan (weights, bias, post = the Snake that follows the unit).
bool ru_fused_supported(const ConvGemmParams& p, const ConvGemmParams& p2);
// fp16-operand executor (conv_h16.cu)
bool h16_supported(const ConvGemmParams& p);
int launch_conv_h16(const ConvGemmParams& p, int num_sms, cudaStream_t stream);
int launch_ru_fused(const ConvGemmParams& p, const ConvGemmParams& p2, int num_sms, cudaStream_t stream);
}  // namespace nc
#endif
