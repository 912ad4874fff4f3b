// Encodec-specific kernels (encodec_kernels.cu).
#pragma once
#include <cstdint>

#include "runtime.h"

namespace nc {

// Materialise F.pad(mode="reflect") in the margin rows around each clip: x -> row 0 of clip 0, clips
// `clip_stride` floats apart; rows [-left, 0) and [T, T+right) are written.  (SConv1d.cs:252-274)
void launch_reflect_pad(float* x, int T, int C, long long clip_stride, int left, int right, int batch, const LaunchCtx& ctx);

// Pad1d of a SHORT input (SConv1d.cs:258-272: length <= max(left, right)): zero-extend by `extra_zero` samples on the right,
// then reflect-pad (left, right) around the extended signal.  in: rows [0, T) of each clip (clip_stride floats apart),
// out: dense [batch][T + extra_zero + left + right][C].  Also serves the ordinary case (extra_zero = 0).
void launch_pad1d_dense(const float* in, long long in_clip_stride, int T, int C, int extra_zero, int left, int right, float* out,
                        int batch, const LaunchCtx& ctx);

// One residual-VQ stage with a 128-d Euclidean codebook (EuclideanCodebook.cs:155-182,
// ResidualVectorQuantizer.cs:146-154): codes[b, stage, t] = argmin_k (|x|^2 + |e_k|^2) + (-2 x.e_k); residual -= e.
// residual: [frames][128] dense rows (frames = B*T); codes: [B][nq][T] int64.
void launch_encodec_vq_stage(float* residual, long long frames, const float* embed, const float* embed_sq, int K, int D,
                             int64_t* codes, int T, int nq, int stage, const LaunchCtx& ctx);
// out[b, t, :] = sum_i embed_i[codes[b, i, t]]  (ResidualVectorQuantizer.cs:107-124); embeds_dev: device array of nq pointers
void launch_encodec_decode_codes(const int64_t* codes, const float* const* embeds_dev, float* out, long long out_clip_stride,
                                 int batch, int T, int nq, int K, int D, const LaunchCtx& ctx);

// One nn.LSTM layer over T steps as a persistent cooperative kernel (zero initial state, SLSTM.cs:47-48).
// xproj: hoisted input projection W_ih x_t + b_ih + b_hh, [B][T][4H] (gate order i,f,g,o); out: [B][T][H] = h_t
// (+ skip[b,t,:] if given; ELU applied when post_elu).  hbuf: 2*ceil(B/16)*16*H floats; barriers: ceil(B/16) uints.
void launch_lstm_layer(const float* xproj, long long xproj_clip_stride, const float* w_hh, float* hbuf, float* out,
                       long long out_clip_stride, const float* skip, long long skip_clip_stride, int post_elu,
                       unsigned int* barriers, int batch, int T, int H, const LaunchCtx& ctx);
// largest batch one launch can take (all CTAs must be co-resident)
int lstm_max_batch(int num_sms, int H);

// .ecdc payload without the language model (EncodecCompressor.cs:170-190, BitPacker.cs:60-110): codes [B][nq][T] ->
// per clip ceil(T*nq*bits/8) bytes, values in (t outer, k inner) order, LSB first, last byte zero-padded.
// row_stride / clip_stride: int64 elements between codebook rows / clips of `codes` (0 = dense T / nq*T), so one segment
// of a [B][nq][t_total] code matrix packs in place.
void launch_ecdc_pack(const int64_t* codes, uint8_t* out, long long out_stride, int batch, int T, int nq, int bits,
                      const LaunchCtx& ctx, long long row_stride = 0, long long clip_stride = 0);
// inverse (EncodecCompressor.cs:383-398, BitUnpacker.cs:60-95)
void launch_ecdc_unpack(const uint8_t* in, long long in_stride, int64_t* codes, int batch, int T, int nq, int bits,
                        const LaunchCtx& ctx, long long row_stride = 0, long long clip_stride = 0);

// ---- 48 kHz preset (stereo, non-causal, time_group_norm, segments; EncodecConfig.cs:37-66)
// GroupNorm(1, C) of NormConv1d / NormConvTranspose1d (NormConv1d.cs:87-100,136-160): stats[b] = {sum, sum of squares} (fp64) over
// n_per_clip contiguous floats of each clip starting at x; apply in place over rows [0, T): (x - mean) * rstd * gamma + beta,
// then + residual and / or ELU.  count = the elements the reference normalises over (padded channels hold zeros and do not count).
void launch_gn_stats(const float* x, long long clip_stride, long long n_per_clip, double* stats, int batch, const LaunchCtx& ctx);
// stats2 != nullptr: `residual` holds another conv's RAW output, normalised on the fly with (stats2, count2, gamma2, beta2)
void launch_gn_apply(float* y, long long clip_stride, int T, int C, const double* stats, double count, float eps, const float* gamma,
                     const float* beta, const float* residual, long long res_clip_stride, int elu, int batch, const LaunchCtx& ctx,
                     const double* stats2 = nullptr, double count2 = 0.0, const float* gamma2 = nullptr, const float* beta2 = nullptr);
// Segment items: launch item i = global item item0 + i -> clip b = item / segs, segment s = s0 + item % segs, samples
// [s*seg_stride, +seg_len) of audio [B][C][L].  Writes the loudness scale (Encodec.cs:469-480) to scales[b*n_seg_total + s]
// (scales == nullptr: Normalize = false) and the normalised segment channels-last [seg_len][Cpad] into out.
void launch_encodec_segment_prep(const float* audio, int C, long long L, int segs, int s0, long long seg_stride, int seg_len, int item0,
                                 int n_seg_total, float* scales, float* out, long long out_clip_stride, int Cpad, int items,
                                 const LaunchCtx& ctx);
// raw rows [T][Cpad] of the last decoder conv -> GroupNorm (stats != nullptr) -> * scale (Encodec.cs:448-451) -> planar
// frames[b][s][c][frame_ld]
void launch_encodec_frame_out(const float* raw, long long clip_stride, int T, int Cpad, int C, const double* stats, float eps,
                              const float* gamma, const float* beta, int segs, int s0, int item0, int n_seg_total, const float* scales,
                              float* frames, long long frame_ld, int items, const LaunchCtx& ctx);
// DSP.LinearOverlapAdd (AudioTensorDSP.cs:161-261) over frames [B][n_seg][C][frame_ld] -> out [B][C][out_len]
// (frame s has lens_dev[s] samples, at most len_max)
void launch_encodec_overlap_add(const float* frames, int batch, int n_seg, int C, long long frame_ld, const int* lens_dev, int len_max,
                                long long stride, int segmented, float* out, long long out_len, const LaunchCtx& ctx);
// dense [items][nq][T] (item = b*segs + j) <-> caller layout [B][nq][T_total], the group's j-th segment at column col0 + j*T
void launch_encodec_codes_segment_copy(int64_t* dense, int64_t* user, int segs, int item0, int nq, int T, long long T_total,
                                       long long col0, int to_user, int items, const LaunchCtx& ctx);

}  // namespace nc
