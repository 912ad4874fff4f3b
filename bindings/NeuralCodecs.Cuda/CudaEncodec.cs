// Drop-in counterpart of NeuralCodecs.Torch/Models/Encodec.cs (24 kHz and 48 kHz presets) and of the static
// EncodecCompressor (Modules/Encodec/EncodecCompressor.cs, useLm: false) over the C ABI.  NOT compiled in this
// repository (no dotnet); its executable twin is neuralcodecs_b200/encodec.py.
using System;
using System.Collections.Generic;
using System.Linq;
using NeuralCodecs.Core;
using NeuralCodecs.Core.Configuration;
using NeuralCodecs.Core.Exceptions;
using NeuralCodecs.Torch.Config.Encodec;

namespace NeuralCodecs.Cuda;

/// (codes [nq * frames] row-major [nq, frames], scale or null) -- Modules/Encodec/EncodedFrame.cs:8 with managed arrays
public sealed record CudaEncodedFrame(long[] Codes, int NumCodebooks, float[]? Scale);

public sealed unsafe class CudaEncodec : INeuralCodec
{
    private readonly EncodecConfig _config;
    private readonly NcHandle _h;
    private float _bandwidth;

    public IModelConfig Config => _config;
    public int SampleRate => _config.SampleRate;
    public int Channels => _config.Channels;
    public float? CurrentBandwidth => _bandwidth;

    public CudaEncodec(EncodecConfig config)                                   // Models/Encodec.cs:46-90
    {
        _config = config ?? throw new ArgumentNullException(nameof(config));
        if (config.Bandwidth is null || !config.TargetBandwidths.Contains(config.Bandwidth.Value))
            throw new ArgumentException($"Invalid bandwidth {config.Bandwidth}. Select one of {string.Join(", ", config.TargetBandwidths)}"); // Encodec.cs:47-53
        _bandwidth = config.Bandwidth.Value;
        int hop = config.UpsamplingRatios.Aggregate(1, (a, b) => a * b);
        int frameRate = (int)Math.Ceiling(config.SampleRate / (float)hop);       // Encodec.cs:86
        var c = new NcEncodecConfig
        {
            StructSize = (uint)sizeof(NcEncodecConfig), SampleRate = config.SampleRate, Channels = config.Channels,
            NFilters = config.NumFilters, Dimension = config.HiddenSize, NRatios = config.UpsamplingRatios.Length,
            NResidualLayers = config.NumResidualLayers, LstmLayers = config.NumLstmLayers, CodebookSize = config.CodebookSize,
            NQuantizers = (int)(1000 * config.TargetBandwidths.Max() / (frameRate * 10)),   // Encodec.cs:70-71
            Causal = config.UseCausalConv ? 1 : 0,
            NormType = config.NormType switch
            {
                "weight_norm" => 0, "time_group_norm" => 1,
                _ => throw new ArgumentException($"Unsupported normalization: {config.NormType}")   // NormConv1d.cs:156-157
            },
            Normalize = config.Normalize ? 1 : 0,
            SegmentS = config.ChunkLengthSeconds ?? 0f,                          // Encodec.cs:79
            Overlap = config.Overlap ?? 0f,                                      // Encodec.cs:84
        };
        for (int i = 0; i < config.UpsamplingRatios.Length; i++) c.Ratios[i] = config.UpsamplingRatios[i];
        Native.Check(Native.nc_create(NcCodecKind.Encodec, &c, (nuint)sizeof(NcEncodecConfig), config.Device?.Index ?? 0, out _h),
                     "Encodec", CodecOperation.Initialization);
    }

    public void LoadWeights(string path) =>                                      // Models/Encodec.cs:348-402
        Native.Check(Native.nc_load_weights(_h, path), "Encodec", CodecOperation.Initialization);

    public void SetTargetBandwidth(float bandwidth)                              // Encodec.cs:409-420
    {
        if (!_config.TargetBandwidths.Contains(bandwidth))
            throw new ArgumentException($"This model doesn't support the bandwidth {bandwidth}. Select one of {string.Join(", ", _config.TargetBandwidths)}");
        _bandwidth = bandwidth;
    }

    /// Encodec.Encode(float[]) (Encodec.cs:243-285): audioData is [Channels, L] planar (the reference reshapes it to
    /// (1, Channels, -1)); one frame per segment -- one for the whole clip in the 24 kHz preset, scale = null there.
    public List<CudaEncodedFrame> Encode(float[] audioData)
    {
        if (audioData is null) throw new ArgumentNullException(nameof(audioData));
        long L = audioData.Length / _config.Channels;
        Native.Check(Native.nc_encodec_query_frames(_h, L, _bandwidth, out int nSeg, null, 0, out long total, out int nq, out _), "Encodec", CodecOperation.Encoding);
        var seg = new long[nSeg];
        fixed (long* ps = seg)
            Native.Check(Native.nc_encodec_query_frames(_h, L, _bandwidth, out nSeg, ps, nSeg, out total, out nq, out _), "Encodec", CodecOperation.Encoding);
        var codes = new long[nq * total];
        var scales = _config.Normalize ? new float[nSeg] : null;
        fixed (float* a = audioData) fixed (long* pc = codes) fixed (float* psc = scales)
            Native.Check(Native.nc_encodec_encode_frames(_h, a, 1, L, _bandwidth, pc, psc), "Encodec", CodecOperation.Encoding);
        var frames = new List<CudaEncodedFrame>(nSeg);
        long col = 0;
        for (int s = 0; s < nSeg; col += seg[s], s++)
        {
            var fc = new long[nq * seg[s]];
            for (int q = 0; q < nq; q++) Array.Copy(codes, q * total + col, fc, q * seg[s], seg[s]);
            frames.Add(new CudaEncodedFrame(fc, nq, scales is null ? null : new[] { scales[s] }));
        }
        return frames;
    }

    /// Encodec.Decode(List<EncodedFrame>) (Encodec.cs:213-235): DecodeFrame (* scale) of every frame and, for segmented
    /// models, DSP.LinearOverlapAdd; audio [Channels, stride*(n-1) + len(last)] planar, not trimmed.
    public float[] Decode(List<CudaEncodedFrame> frames)
    {
        if (frames is null || frames.Count == 0) throw new ArgumentException("No frames provided to decode");
        if (_config.ChunkLengthSeconds is null && frames.Count != 1) throw new ArgumentException("Expected single frame when no segmentation is used");
        if (frames.Any(f => f.Codes is null)) throw new ArgumentException("Invalid frame codes in Encodec Decode");     // Encodec.cs:438-442
        int nq = frames[0].NumCodebooks;
        var seg = frames.Select(f => (long)(f.Codes.Length / nq)).ToArray();
        long total = seg.Sum();
        var codes = new long[nq * total];
        long col = 0;
        for (int s = 0; s < frames.Count; col += seg[s], s++)
            for (int q = 0; q < nq; q++) Array.Copy(frames[s].Codes, q * seg[s], codes, q * total + col, seg[s]);
        bool scaled = frames.All(f => f.Scale is { Length: > 0 });
        var scales = scaled ? frames.Select(f => f.Scale![0]).ToArray() : null;
        fixed (long* ps = seg) fixed (long* pc = codes) fixed (float* psc = scales)
        {
            Native.Check(Native.nc_encodec_query_decoded(_h, ps, seg.Length, out long n), "Encodec", CodecOperation.Decoding);
            var audio = new float[_config.Channels * n];
            fixed (float* pa = audio)
                Native.Check(Native.nc_encodec_decode_frames(_h, pc, psc, 1, nq, ps, seg.Length, pa), "Encodec", CodecOperation.Decoding);
            return audio;
        }
    }

    /// EncodecCompressor.Compress(model, wav, useLm: false) (EncodecCompressor.cs:26-39,60-200): the .ecdc byte stream.
    public byte[] Compress(float[] wav)
    {
        if (wav is null) throw new ArgumentNullException(nameof(wav));
        long L = wav.Length / _config.Channels;                                  // wav is [Channels, L] planar (:67-77)
        Native.Check(Native.nc_encodec_ecdc_size(_h, L, _bandwidth, out _, out long total), "Encodec", CodecOperation.Encoding);
        var outp = new byte[total];
        fixed (float* a = wav) fixed (byte* po = outp)
            Native.Check(Native.nc_encodec_compress(_h, a, 1, L, _bandwidth, po, total, out _), "Encodec", CodecOperation.Encoding);
        return outp;
    }

    /// EncodecCompressor.Decompress (EncodecCompressor.cs:46-52,236-420) for streams written without the language model.
    public (float[] wav, int sampleRate) Decompress(byte[] compressed)
    {
        if (compressed is null) throw new ArgumentNullException(nameof(compressed));
        fixed (byte* ps = compressed)
        {
            Native.Check(Native.nc_encodec_decompress(_h, ps, 1, compressed.Length, compressed.Length, null, 0, out long al, out int sr), "Encodec", CodecOperation.Decoding);
            var wav = new float[_config.Channels * al];                          // [Channels, al] planar
            fixed (float* pw = wav)
                Native.Check(Native.nc_encodec_decompress(_h, ps, 1, compressed.Length, compressed.Length, pw, al, out al, out sr), "Encodec", CodecOperation.Decoding);
            return (wav, sr);
        }
    }

    public void Dispose() => _h.Dispose();
}
