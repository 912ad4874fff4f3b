/* neuralcodecs_cuda.h -- C ABI of libneuralcodecs_cuda.so
 *
 * Blackwell (sm_100a) backend for the codec encode/decode hot path of
 * DillionLowry/NeuralCodecs.  A `NeuralCodecs.Cuda` C# project binds these entry
 * points through P/Invoke ([LibraryImport]); see INTEGRATION.md.  Every function cites
 * the reference interface it replaces (paths relative to
 * /root/reference/NeuralCodecs.Torch/ unless stated otherwise).
 *
 * Conventions
 *  - plain C, no exceptions cross the boundary: every call returns nc_status and sets
 *    a thread-local message readable through nc_last_error().
 *  - host arrays are row-major fp32 / int64 exactly as the reference's tensors:
 *    audio [B,1,L], latents z [B,C,T], codes [B,nq,T] (int64).
 *  - the CALLER allocates every output (sizes from the *_query_shapes calls); the
 *    library owns device weights and workspaces per handle.
 *  - a handle is bound to one CUDA device and is not re-entrant (one thread at a time,
 *    matching the reference's de-facto contract: Models/DAC.cs:354,381 mutate global
 *    state in LoadWeights; nothing is synchronised).  Different handles are independent.
 *  - there is NO CPU fallback: without a usable CUDA device nc_create fails with
 *    NC_CUDA_UNAVAILABLE (ref: Utils/TorchUtils.cs:97-99 "CUDA requested but not
 *    available").
 */
#ifndef NEURALCODECS_CUDA_H_
#define NEURALCODECS_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define NC_API __declspec(dllexport)
#else
#define NC_API __attribute__((visibility("default")))
#endif

typedef struct nc_handle_s* nc_handle;

/* Status codes; the C# layer maps them onto the reference's exception conventions
 * (NeuralCodecs.Core/Exceptions/ directory; Models/DAC.cs:53,146,207,347-388):
 *   INVALID_ARGUMENT -> ArgumentException        FILE_NOT_FOUND -> FileNotFoundException
 *   BAD_WEIGHTS / SHAPE_MISMATCH -> InvalidOperationException("Failed to load ... weights")
 *   CUDA_UNAVAILABLE -> InvalidOperationException("CUDA requested but not available")
 *   CUDA_ERROR / OUT_OF_MEMORY / INTERNAL -> CodecException(codec, CodecOperation, msg) */
typedef enum nc_status {
  NC_OK = 0,
  NC_INVALID_ARGUMENT = 1,
  NC_FILE_NOT_FOUND = 2,
  NC_BAD_WEIGHTS = 3,
  NC_SHAPE_MISMATCH = 4,
  NC_CUDA_UNAVAILABLE = 5,
  NC_CUDA_ERROR = 6,
  NC_OUT_OF_MEMORY = 7,
  NC_INTERNAL = 8,
  NC_UNSUPPORTED = 9
} nc_status;

typedef enum nc_codec_kind { NC_CODEC_DAC = 1, NC_CODEC_SNAC = 2, NC_CODEC_ENCODEC = 3 } nc_codec_kind;

#define NC_MAX_RATES 8

/* Flat mirror of Config/DAC/DACConfig.cs:8-100 (fields the model constructor reads,
 * Models/DAC.cs:51-93).  latent_dim = 0 means "encoder_dim * 2^n_encoder_rates". */
typedef struct nc_dac_config {
  uint32_t struct_size; /* = sizeof(nc_dac_config) */
  int32_t sample_rate;
  int32_t encoder_dim;
  int32_t n_encoder_rates;
  int32_t encoder_rates[NC_MAX_RATES];
  int32_t decoder_dim;
  int32_t n_decoder_rates;
  int32_t decoder_rates[NC_MAX_RATES];
  int32_t n_codebooks;
  int32_t codebook_size;
  int32_t codebook_dim;
  int32_t latent_dim;
} nc_dac_config;

/* Flat mirror of Config/SNAC/SNACConfig.cs:11-153 (Models/SNAC.cs:34-63). */
typedef struct nc_snac_config {
  uint32_t struct_size;
  int32_t sample_rate;
  int32_t encoder_dim;
  int32_t n_encoder_rates;
  int32_t encoder_rates[NC_MAX_RATES];
  int32_t decoder_dim;
  int32_t n_decoder_rates;
  int32_t decoder_rates[NC_MAX_RATES];
  int32_t latent_dim; /* 0 = encoder_dim * 2^n_encoder_rates */
  int32_t attn_window_size; /* 0 = no LocalMHA (24 kHz preset) */
  int32_t codebook_size;
  int32_t codebook_dim;
  int32_t n_vq_strides;
  int32_t vq_strides[NC_MAX_RATES];
  int32_t noise;     /* bool */
  int32_t depthwise; /* bool */
} nc_snac_config;

/* Flat mirror of Config/Encodec/EncodecConfig.cs:6-153: the 24 kHz preset (:9-34: mono, causal, weight_norm, one
 * frame per clip) and the 48 kHz preset (:37-66: stereo, non-causal, time_group_norm, normalize, 1 s segments with
 * 1 % overlap); constructor Models/Encodec.cs:46-90. */
#define NC_ENCODEC_NORM_WEIGHT 0     /* "weight_norm" */
#define NC_ENCODEC_NORM_TIME_GROUP 1 /* "time_group_norm": GroupNorm(1, C) after every conv, NormConv1d.cs:136-160 */
typedef struct nc_encodec_config {
  uint32_t struct_size;
  int32_t sample_rate;
  int32_t channels;
  int32_t n_filters;
  int32_t dimension;
  int32_t n_ratios;
  int32_t ratios[NC_MAX_RATES]; /* decoder order, e.g. 8,5,4,2 */
  int32_t n_residual_layers;
  int32_t lstm_layers;
  int32_t codebook_size;
  int32_t n_quantizers; /* layers constructed (32 for 24 kHz) */
  int32_t causal;       /* bool */
  int32_t norm_type;    /* NC_ENCODEC_NORM_* */
  int32_t normalize;    /* bool: per-frame loudness scale, Models/Encodec.cs:469-480 */
  float segment_s;      /* `Segment` in seconds (chunk_length_s); <= 0 = one frame per clip */
  float overlap;        /* fraction of a segment shared with the next one (0.01) */
} nc_encodec_config;

/* -- library ---------------------------------------------------------------------- */
NC_API const char* nc_version(void);
/* thread-local description of the last failure on this thread ("" if none) */
NC_API const char* nc_last_error(void);
/* number of usable sm_100 devices (0 when CUDA is unavailable) */
NC_API int nc_device_count(void);

/* -- lifecycle -------------------------------------------------------------------- */
/* replaces: model constructors reached through ModelRegistry.CreateModel<TModel,TConfig>
 * (NeuralCodecs.Core/Loading/ModelRegistry.cs:35-74): new DAC(cfg) Models/DAC.cs:51,
 * new SNAC(cfg) Models/SNAC.cs:34, new Encodec(cfg) Models/Encodec.cs:46.
 * cfg points at the nc_*_config matching `kind`; device_index = DeviceConfiguration.Index
 * (NeuralCodecs.Core/Configuration/DeviceConfiguration.cs:7-17). */
NC_API nc_status nc_create(nc_codec_kind kind, const void* cfg, size_t cfg_size, int device_index,
                           nc_handle* out);
/* replaces: IDisposable.Dispose on the model (Models/DAC.cs:328-337). */
NC_API nc_status nc_destroy(nc_handle h);
/* replaces: INeuralCodec.LoadWeights(path) (NeuralCodecs.Core/INeuralCodec.cs:8-20;
 * Models/DAC.cs:345-389, Models/SNAC.cs:200-246, Models/Encodec.cs:348-402).  Reads a
 * .safetensors file in the reference's key layout for the codec (DAC: HF DacModel layout
 * translated by Config/DAC/StateDictNameConverter.cs:40-65,274-376) and folds
 * weight-norm once, in fp32, with the reference's epsilon placement. */
NC_API nc_status nc_load_weights(nc_handle h, const char* utf8_path);
/* test hook: supply one named tensor from host memory instead of a file.  dtype: 0 = f32,
 * 1 = i64.  Call nc_finalize_weights when all tensors are set. */
NC_API nc_status nc_set_tensor(nc_handle h, const char* name, int dtype, int rank, const int64_t* shape,
                               const void* data);
NC_API nc_status nc_finalize_weights(nc_handle h);
/* engine options (strings), e.g. "encoder_precision" / "decoder_precision" =
 * "fp32" | "tf32" | "3xtf32"; "max_workspace_mb" = "<n>"; "profile" = "0|1". */
NC_API nc_status nc_set_option(nc_handle h, const char* key, const char* value);

/* -- DAC -------------------------------------------------------------------------- */
/* DAC.Preprocess length algebra (Models/DAC.cs:141-154): padded length and frame count. */
NC_API nc_status nc_dac_query_shapes(nc_handle h, int64_t length, int64_t* padded_length, int64_t* frames,
                                     int32_t* latent_dim, int32_t* n_codebooks, int32_t* codebook_dim);
/* replaces: DAC.Encode(Tensor audio, int? nQuantizers, int? sampleRate) Models/DAC.cs:163-181
 * (and EncodeAudio :188-198, Encode(float[]) :205-224 which return z only).
 * audio [B,1,L]; sample_rate 0 = model rate, otherwise must equal it (ArgumentException in
 * the ref, Models/DAC.cs:144-149); n_quantizers 0 = all.  Outputs (each nullable):
 * z [B,latent,T] fp32, codes [B,nq,T] int64, latents [B,nq*codebook_dim,T] fp32. */
NC_API nc_status nc_dac_encode(nc_handle h, const float* audio, int32_t batch, int64_t length,
                               int32_t sample_rate, int32_t n_quantizers, float* z, int64_t* codes,
                               float* latents, int64_t* frames_out);
/* replaces: DAC.Decode(Tensor z) Models/DAC.cs:231-234 and Decode(float[]) :241-253.
 * z [B,latent,T] -> audio [B,1,T*hop] (NOT trimmed, as in the reference). */
NC_API nc_status nc_dac_decode(nc_handle h, const float* z, int32_t batch, int64_t frames, float* audio);
/* replaces: DAC.FromCodes(Tensor codes) Models/DAC.cs:101-106 ->
 * ResidualVectorQuantizer.FromCodes Modules/DAC/ResidualVectorQuantizer.cs:211-238.
 * codes [B,nq,T] int64 -> z [B,latent,T]. */
NC_API nc_status nc_dac_from_codes(nc_handle h, const int64_t* codes, int32_t batch, int32_t n_quantizers,
                                   int64_t frames, float* z);
/* replaces: Dia.Decode(codes) Models/Dia.cs:973-981 = FromCodes + Decode, batched (the
 * reference loops serially per item, Models/Dia.cs:1057-1060).  codes [B,nq,T] int64 ->
 * audio [B,1,T*hop]. */
NC_API nc_status nc_dac_decode_codes(nc_handle h, const int64_t* codes, int32_t batch, int32_t n_quantizers,
                                     int64_t frames, float* audio);
/* replaces: Dia.GenerateOutput's codec stage Models/Dia.cs:1010-1060: generated [B,T,C] int64 are
 * the delayed codes as Dia emits them; the delay pattern is reverted (Modules/Dia/AudioUtils.cs:108-176), the last
 * max(delay) steps dropped, values outside [0, codebook_size) set to 0, and item b's first lengths[b] frames decoded.
 * audio [B, audio_stride]: item b holds lengths[b]*hop samples (the rest of its row is left untouched).  Items are
 * decoded in batches of equal length instead of the reference's serial loop. */
NC_API nc_status nc_dac_decode_dia(nc_handle h, const int64_t* generated, int32_t batch, int32_t steps,
                                   int32_t channels, const int32_t* delay_pattern, const int64_t* lengths,
                                   float* audio, int64_t audio_stride);
/* replaces: DAC.forward(Tensor) Models/DAC.cs:262-322 (Encode then Decode).  Outputs
 * nullable: audio_out [B,1,padded L], codes [B,nq,T], z [B,latent,T]. */
NC_API nc_status nc_dac_forward(nc_handle h, const float* audio, int32_t batch, int64_t length,
                                int32_t n_quantizers, float* audio_out, int64_t* codes, float* z,
                                int64_t* frames_out);
/* Device-pointer variant of nc_dac_forward for zero-copy callers and device-only timing:
 * all pointers are device memory on the handle's device; work is enqueued on the handle's
 * stream and the call returns after that stream is synchronised. */
NC_API nc_status nc_dac_forward_dev(nc_handle h, const float* audio_dev, int32_t batch, int64_t length,
                                    int32_t n_quantizers, float* audio_out_dev, int64_t* codes_dev,
                                    float* z_dev, int64_t* frames_out);
NC_API nc_status nc_dac_decode_codes_dev(nc_handle h, const int64_t* codes_dev, int32_t batch,
                                         int32_t n_quantizers, int64_t frames, float* audio_dev);

/* -- SNAC ------------------------------------------------------------------------- */
/* SNAC.Preprocess length algebra (Models/SNAC.cs:70-80): audio is right-padded to a multiple of
 * hop * lcm(vq_strides[0], attn_window or 1).  frames = padded_length / hop (finest code rate);
 * code_lengths[i] = frames / vq_strides[i]; noise_lengths[i] = time steps after decoder block i
 * (NoiseBlock.cs:38-45 draws randn[B,1,T_i]).  Arrays must hold NC_MAX_RATES entries. */
NC_API nc_status nc_snac_query_shapes(nc_handle h, int64_t length, int64_t* padded_length, int64_t* frames,
                                      int32_t* n_stages, int64_t* code_lengths, int32_t* n_noise,
                                      int64_t* noise_lengths);
/* replaces: SNAC.Encode(float[]) Models/SNAC.cs:129-150 (codes of the PADDED audio; the Tensor overload's
 * unpadded-input bug, SNAC.cs:117-118, is not reproduced).  codes[i]: [B, frames / vq_strides[i]] int64
 * (the reference's List<float[]> API casts them to float32, SNAC.cs:147). */
NC_API nc_status nc_snac_encode(nc_handle h, const float* audio, int32_t batch, int64_t length,
                                int64_t* const* codes);
/* replaces: SNAC.Decode(List<...>) Models/SNAC.cs:157-192 = RVQ.FromCodes + Decoder; audio [B,1,frames*hop]
 * (not trimmed).  noise[i]: [B, noise_lengths[i]] explicit N(0,1) tensors, or noise == NULL / noise[i] ==
 * NULL to draw them on the device from `seed` (the reference draws randn per call and is not reproducible). */
NC_API nc_status nc_snac_decode(nc_handle h, const int64_t* const* codes, int32_t batch, int64_t frames,
                                const float* const* noise, uint64_t seed, float* audio);
/* replaces: SNAC.forward Models/SNAC.cs:91-106 and ProcessAudio :255-282 (without the CPU resampler):
 * audio_out [B,1,length] trimmed to the input length; codes nullable (array of n_stages pointers). */
NC_API nc_status nc_snac_forward(nc_handle h, const float* audio, int32_t batch, int64_t length,
                                 const float* const* noise, uint64_t seed, float* audio_out,
                                 int64_t* const* codes);
/* device-pointer variant (arrays of pointers live on the host, the buffers they point to on the device) */
NC_API nc_status nc_snac_forward_dev(nc_handle h, const float* audio_dev, int32_t batch, int64_t length,
                                     const float* const* noise_dev, uint64_t seed, float* audio_out_dev,
                                     int64_t* const* codes_dev);

/* replaces: SNAC.ProcessAudio(float[], sampleRate) Models/SNAC.cs:255-282: linear resample to the model rate when the
 * rates differ (ResampleAudio :284-308, on the device, same double arithmetic), forward, output of the (resampled)
 * input length.  audio [B][length] -> audio_out [B][out_capacity], *out_length samples per row written.
 * audio_out == NULL only reports *out_length.  noise as in nc_snac_forward (lengths from nc_snac_query_shapes of
 * *out_length). */
NC_API nc_status nc_snac_process_audio(nc_handle h, const float* audio, int32_t batch, int64_t length,
                                       int32_t sample_rate, const float* const* noise, uint64_t seed,
                                       float* audio_out, int64_t out_capacity, int64_t* out_length);

/* replaces: DACUnpickler.LoadWithConfig / CreateConfigFromMetadata Config/DAC/DACUnpickler.cs:383-424 (reading the
 * checkpoint's metadata to build the config BEFORE the model exists).  Host only, no handle: JSON
 * {"format": "torch_zip"|"safetensors", "metadata": {...}, "tensors": {name: {"dtype", "shape"}}} into buf.
 * nc_load_weights itself accepts both formats: .safetensors (HF DacModel or native names) and torch.save zip
 * checkpoints such as the official DAC .pth ({"state_dict", "metadata"}; DACUnpickler.LoadFromStream :341-381). */
NC_API nc_status nc_inspect_weights(const char* path, char* buf, size_t buf_size);

/* -- input conditioning (any handle; runs on that handle's device and stream) ------- */
/* replaces: AudioUtils.ResampleLinear NeuralCodecs.Core/Utils/AudioUtils.cs:329-352 (= SNAC.ResampleAudio
 * Models/SNAC.cs:284-308).  *out_length = (int64)(length * dst/src); out == NULL only reports it. */
NC_API nc_status nc_resample_linear(nc_handle h, const float* audio, int32_t batch, int64_t length, int32_t src_rate,
                                    int32_t dst_rate, float* out, int64_t out_capacity, int64_t* out_length);
/* replaces: AudioUtils.ConvertToMono(float[], channels) NeuralCodecs.Core/Utils/AudioUtils.cs:45-62:
 * interleaved [frames][channels] -> out [frames], float sum in channel order divided by the channel count. */
NC_API nc_status nc_convert_to_mono(nc_handle h, const float* interleaved, int64_t frames, int32_t channels, float* out);

/* -- Encodec ---------------------------------------------------------------------- */
/* Single-frame models (24 kHz preset).  frames = ceil-chain of the strided SConv1d layers (SConv1d.cs:245-250);
 * n_q = max(1, floor(bandwidth*1000 / (log2(bins) * frame_rate))) (ResidualVectorQuantizer.cs:133-144);
 * decoded_length = frames * hop.  bandwidth_kbps <= 0 selects every codebook in the file. */
NC_API nc_status nc_encodec_query_shapes(nc_handle h, int64_t length, float bandwidth_kbps, int64_t* frames,
                                         int32_t* n_q, int64_t* decoded_length);
/* replaces: Encodec.Encode(float[]) / Encode(Tensor) Models/Encodec.cs:243-285 -> EncodeFrame :457-489 (one frame =
 * the whole clip for the 24 kHz preset, Normalize = false -> scale = null).  codes [B, n_q, frames] int64. */
NC_API nc_status nc_encodec_encode(nc_handle h, const float* audio, int32_t batch, int64_t length,
                                   float bandwidth_kbps, int64_t* codes);
/* replaces: Encodec.Decode(List<EncodedFrame>) Models/Encodec.cs:213-235 -> DecodeFrame :436-455.
 * codes [B, n_q, frames] -> audio [B, 1, frames*hop] (not trimmed). */
NC_API nc_status nc_encodec_decode(nc_handle h, const int64_t* codes, int32_t batch, int32_t n_q, int64_t frames,
                                   float* audio);
/* replaces: Encodec.forward Models/Encodec.cs:292-296: decode(encode(x)) sliced to the input length, for every preset
 * (segment loop, scales and overlap-add inside).  audio / audio_out [B, channels, length]; codes nullable,
 * [B, n_q, total_frames]. */
NC_API nc_status nc_encodec_forward(nc_handle h, const float* audio, int32_t batch, int64_t length,
                                    float bandwidth_kbps, float* audio_out, int64_t* codes);
NC_API nc_status nc_encodec_forward_dev(nc_handle h, const float* audio_dev, int32_t batch, int64_t length,
                                        float bandwidth_kbps, float* audio_out_dev, int64_t* codes_dev);

/* Segmented / normalised models (48 kHz preset) and, equally, single-frame ones.
 * replaces: the loop of Encodec.Encode Models/Encodec.cs:273-282 with SegmentLength / SegmentStride :190-196: segment s
 * covers samples [s*stride, min(s*stride + segment, length)).  Reports the number of segments, the code frames of each
 * (seg_frames[0 .. min(n_segments, capacity)), nullable), their sum, n_q for the bandwidth and the length Decode returns
 * (stride*(n_segments-1) + decoded length of the last frame, AudioTools/AudioTensorDSP.cs:180). */
NC_API nc_status nc_encodec_query_frames(nc_handle h, int64_t length, float bandwidth_kbps, int32_t* n_segments,
                                         int64_t* seg_frames, int32_t seg_frames_capacity, int64_t* total_frames,
                                         int32_t* n_q, int64_t* decoded_length);
/* replaces: Encodec.Encode(Tensor) -> List<EncodedFrame> Models/Encodec.cs:259-285 (EncodeFrame :457-489), batched.
 * audio [B, channels, length] planar; codes [B, n_q, total_frames] int64 with the segments' codes concatenated in time;
 * scales [B, n_segments] float = EncodedFrame.Scale (written when the model normalises; nullable). */
NC_API nc_status nc_encodec_encode_frames(nc_handle h, const float* audio, int32_t batch, int64_t length,
                                          float bandwidth_kbps, int64_t* codes, float* scales);
/* replaces: Encodec.Decode(List<EncodedFrame>) Models/Encodec.cs:213-235: DecodeFrame (* scale, :436-455) of every frame,
 * then DSP.LinearOverlapAdd(frames, SegmentStride) AudioTools/AudioTensorDSP.cs:161-261.  seg_frames[n_segments] = code
 * frames of each segment; scales nullable (frames without a scale); audio [B, channels, decoded_length] with
 * decoded_length = stride*(n_segments-1) + decoded length of the last frame.  A last frame too short to cover the end
 * of an earlier one returns NC_INVALID_ARGUMENT (the reference's narrow() throws). */
NC_API nc_status nc_encodec_decode_frames(nc_handle h, const int64_t* codes, const float* scales, int32_t batch, int32_t n_q,
                                          const int64_t* seg_frames, int32_t n_segments, float* audio);
/* length nc_encodec_decode_frames writes per channel for these frames: stride*(n_segments-1) + decoded length of the
 * last frame (frames*hop, except that fewer frames than the first decoder conv's padding lengthen, SConv1d.cs:258-272) */
NC_API nc_status nc_encodec_query_decoded(nc_handle h, const int64_t* seg_frames, int32_t n_segments, int64_t* decoded_length);
/* device-pointer forms of the two calls above, asynchronous on the handle's stream up to the final synchronise */
NC_API nc_status nc_encodec_forward_frames_dev(nc_handle h, const float* audio_dev, int32_t batch, int64_t length,
                                               float bandwidth_kbps, float* audio_out_dev, int64_t* codes_dev,
                                               float* scales_dev);

/* -- Encodec .ecdc container, language-model entropy coder off --------------------
 * Stream = "ECDC" | version byte 0 | int32 big-endian JSON length | JSON {m,al,nc,lm,ch,sr,bw}
 * (Modules/Encodec/BinaryIO.cs:11,152-190) | codes bit-packed LSB-first at log2(bins) bits each, time-major then
 * codebook, last byte zero-padded (EncodecCompressor.cs:170-190, Modules/Encodec/BitPacker.cs:60-110).  The 24 kHz
 * preset has one frame per clip and Normalize = false, so no scale block is written. */
/* sizes of the header and of the whole stream nc_encodec_compress writes for one clip of `length` samples */
NC_API nc_status nc_encodec_ecdc_size(nc_handle h, int64_t length, float bandwidth_kbps, int64_t* header_bytes,
                                      int64_t* stream_bytes);
/* replaces: EncodecCompressor.CompressToStreamAsync(model, wav, stream, useLm: false)
 * Modules/Encodec/EncodecCompressor.cs:60-200 (and Compress :26-39), batched: audio [B,channels,length] -> B streams of
 * *stream_bytes bytes each, clip b at out + b*out_stride.  Encode and bit-pack run on the device. */
NC_API nc_status nc_encodec_compress(nc_handle h, const float* audio, int32_t batch, int64_t length,
                                     float bandwidth_kbps, uint8_t* out, int64_t out_stride, int64_t* stream_bytes);
/* replaces: BinaryIO.ReadHeaderAsync + ValidateMetadata Modules/Encodec/BinaryIO.cs:44-146 and the metadata defaults
 * of EncodecCompressor.cs:253-275.  Host only; every output nullable. */
NC_API nc_status nc_encodec_ecdc_info(const uint8_t* stream, int64_t stream_bytes, int64_t* audio_length, int32_t* n_q,
                                      int32_t* channels, int32_t* sample_rate, float* bandwidth_kbps, int32_t* use_lm,
                                      int64_t* payload_offset);
/* replaces: EncodecCompressor.DecompressFromStreamAsync Modules/Encodec/EncodecCompressor.cs:236-420 (and Decompress
 * :46-52), batched over streams with identical metadata: B streams of stream_bytes bytes, stream b at
 * streams + b*stream_stride -> audio [B][channels][audio_capacity] with the first *audio_length samples of each row
 * written (output trimmed to the stored length, :412-415).  audio == NULL only reports *audio_length / *sample_rate.
 * Segmented / normalised models (48 kHz): per segment a scale block [int32 BE count = 1][float32 BE scale] and that
 * segment's codes (frames per segment = ceil(segment samples * frame_rate / sample_rate), :303-309), then
 * Encodec.Decode with the overlap-add.
 * lm = true streams return NC_UNSUPPORTED; a truncated payload returns NC_INVALID_ARGUMENT "Stream ended too soon". */
NC_API nc_status nc_encodec_decompress(nc_handle h, const uint8_t* streams, int32_t batch, int64_t stream_stride,
                                       int64_t stream_bytes, float* audio, int64_t audio_capacity, int64_t* audio_length,
                                       int32_t* sample_rate);

/* The CUDA stream (cudaStream_t) every *_dev call of this handle enqueues on, so a caller
 * can bracket calls with its own events or order its own work against them. */
NC_API nc_status nc_get_stream(nc_handle h, void** stream_out);

/* JSON description of the loaded engine (codec, precision policy, per-layer executor) into
 * buf; NC_INVALID_ARGUMENT when buf is too small. */
NC_API nc_status nc_describe(nc_handle h, char* buf, size_t buf_size);

/* -- instrumentation ---------------------------------------------------------------- */
/* kernels launched by this handle since creation (bench.py's gpu_launches). */
NC_API uint64_t nc_launch_count(nc_handle h);
/* With option profile=1 every kernel launch is bracketed by CUDA events on the handle's
 * stream; this writes a JSON object {kernel: {launches, ms, flops, bytes}} into buf and
 * resets the counters.  Returns NC_INVALID_ARGUMENT when buf is too small. */
NC_API nc_status nc_profile_report(nc_handle h, char* buf, size_t buf_size);

#ifdef __cplusplus
}
#endif
#endif /* NEURALCODECS_CUDA_H_ */
