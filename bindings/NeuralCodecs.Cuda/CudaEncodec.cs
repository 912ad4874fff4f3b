// Drop-in counterpart of NeuralCodecs.Torch/Models/Encodec.cs (24 kHz mono causal preset) and of the static
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

    /// Encodec.Encode(float[]) (Encodec.cs:243-250): one frame for the whole clip (no segmenting in the 24 kHz preset), scale = null.
    public List<CudaEncodedFrame> Encode(float[] audioData)
    {
        if (audioData is null) throw new ArgumentNullException(nameof(audioData));
        Native.Check(Native.nc_encodec_query_shapes(_h, audioData.Length, _bandwidth, out long T, out int nq, out _), "Encodec", CodecOperation.Encoding);
        var codes = new long[nq * T];
        fixed (float* a = audioData) fixed (long* pc = codes)
            Native.Check(Native.nc_encodec_encode(_h, a, 1, audioData.Length, _bandwidth, pc), "Encodec", CodecOperation.Encoding);
        return new List<CudaEncodedFrame> { new(codes, nq, null) };
    }

    /// Encodec.Decode(List<EncodedFrame>) (Encodec.cs:213-235): audio [frames * hop], not trimmed.
    public float[] Decode(List<CudaEncodedFrame> frames)
    {
        if (frames is null || frames.Count == 0) throw new ArgumentException("No frames provided to decode");
        if (frames.Count != 1) throw new ArgumentException("Expected single frame when no segmentation is used");
        var f = frames[0];
        if (f.Codes is null) throw new ArgumentException("Invalid frame codes in Encodec Decode");     // Encodec.cs:438-442
        long T = f.Codes.Length / f.NumCodebooks;
        int hop = _config.UpsamplingRatios.Aggregate(1, (a, b) => a * b);
        var audio = new float[T * hop];
        fixed (long* pc = f.Codes) fixed (float* pa = audio)
            Native.Check(Native.nc_encodec_decode(_h, pc, 1, f.NumCodebooks, T, pa), "Encodec", CodecOperation.Decoding);
        if (f.Scale is { Length: > 0 }) for (int i = 0; i < audio.Length; i++) audio[i] *= f.Scale[0];  // Encodec.cs:449-452
        return audio;
    }

    /// EncodecCompressor.Compress(model, wav, useLm: false) (EncodecCompressor.cs:26-39,60-200): the .ecdc byte stream.
    public byte[] Compress(float[] wav)
    {
        if (wav is null) throw new ArgumentNullException(nameof(wav));
        Native.Check(Native.nc_encodec_ecdc_size(_h, wav.Length, _bandwidth, out _, out long total), "Encodec", CodecOperation.Encoding);
        var outp = new byte[total];
        fixed (float* a = wav) fixed (byte* po = outp)
            Native.Check(Native.nc_encodec_compress(_h, a, 1, wav.Length, _bandwidth, po, total, out _), "Encodec", CodecOperation.Encoding);
        return outp;
    }

    /// EncodecCompressor.Decompress (EncodecCompressor.cs:46-52,236-420) for streams written without the language model.
    public (float[] wav, int sampleRate) Decompress(byte[] compressed)
    {
        if (compressed is null) throw new ArgumentNullException(nameof(compressed));
        fixed (byte* ps = compressed)
        {
            Native.Check(Native.nc_encodec_decompress(_h, ps, 1, compressed.Length, compressed.Length, null, 0, out long al, out int sr), "Encodec", CodecOperation.Decoding);
            var wav = new float[al];
            fixed (float* pw = wav)
                Native.Check(Native.nc_encodec_decompress(_h, ps, 1, compressed.Length, compressed.Length, pw, al, out al, out sr), "Encodec", CodecOperation.Decoding);
            return (wav, sr);
        }
    }

    public void Dispose() => _h.Dispose();
}
