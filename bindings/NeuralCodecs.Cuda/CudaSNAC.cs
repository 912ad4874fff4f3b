// Drop-in counterpart of NeuralCodecs.Torch/Models/SNAC.cs over the C ABI: same config class (SNACConfig), same method
// names, managed arrays instead of TorchSharp tensors.  NOT compiled in this repository (no dotnet); its executable
// twin is neuralcodecs_b200/snac.py.
using System;
using System.Collections.Generic;
using System.Linq;
using NeuralCodecs.Core;
using NeuralCodecs.Core.Configuration;
using NeuralCodecs.Core.Exceptions;
using NeuralCodecs.Torch.Config.SNAC;

namespace NeuralCodecs.Cuda;

public sealed unsafe class CudaSNAC : INeuralCodec
{
    private readonly SNACConfig _config;
    private readonly NcHandle _h;

    public IModelConfig Config => _config;

    public CudaSNAC(SNACConfig config)                                         // Models/SNAC.cs:34-63
    {
        _config = config ?? throw new ArgumentNullException(nameof(config));
        var c = new NcSnacConfig
        {
            StructSize = (uint)sizeof(NcSnacConfig), SampleRate = config.SampleRate, EncoderDim = config.EncoderDim,
            NEncoderRates = config.EncoderRates.Length, DecoderDim = config.DecoderDim, NDecoderRates = config.DecoderRates.Length,
            LatentDim = config.LatentDim ?? 0, AttnWindowSize = config.AttnWindowSize ?? 0, CodebookSize = config.CodebookSize,
            CodebookDim = config.CodebookDim, NVqStrides = config.VQStrides.Length, Noise = config.Noise ? 1 : 0,
            Depthwise = config.Depthwise ? 1 : 0,
        };
        for (int i = 0; i < config.EncoderRates.Length; i++) c.EncoderRates[i] = config.EncoderRates[i];
        for (int i = 0; i < config.DecoderRates.Length; i++) c.DecoderRates[i] = config.DecoderRates[i];
        for (int i = 0; i < config.VQStrides.Length; i++) c.VqStrides[i] = config.VQStrides[i];
        Native.Check(Native.nc_create(NcCodecKind.Snac, &c, (nuint)sizeof(NcSnacConfig), config.Device?.Index ?? 0, out _h),
                     "SNAC", CodecOperation.Initialization);
    }

    public void LoadWeights(string path) =>                                     // Models/SNAC.cs:200-246 (.safetensors or pytorch_model.bin)
        Native.Check(Native.nc_load_weights(_h, path), "SNAC", CodecOperation.Initialization);

    private (long padded, long frames, long[] codeLengths, long[] noiseLengths) Shapes(long length)
    {
        var cl = new long[8]; var nl = new long[8];
        long padded, frames; int ns, nn;
        fixed (long* pc = cl, pn = nl)
            Native.Check(Native.nc_snac_query_shapes(_h, length, out padded, out frames, out ns, pc, out nn, pn), "SNAC", CodecOperation.Encoding);
        Array.Resize(ref cl, ns); Array.Resize(ref nl, nn);
        return (padded, frames, cl, nl);
    }

    /// SNAC.Encode(float[]) (Models/SNAC.cs:129-150): one array per VQ stage holding the codes of the padded audio,
    /// batch 1.  The reference returns the int64 indices cast to float32 (`code.to(torch.float32)`, :147), so this does
    /// too: a codebook index (< 2^24) is exact in float32.
    public List<float[]> Encode(float[] audioData) =>
        EncodeCodes(audioData).ConvertAll(stage => Array.ConvertAll(stage, c => (float)c));

    /// Same codes as integers (what the device produces); no counterpart in the reference's float[] API.
    public List<long[]> EncodeCodes(float[] audioData)
    {
        ArgumentNullException.ThrowIfNull(audioData);                                   // SNAC.cs:131
        var (_, _, cl, _) = Shapes(audioData.Length);
        var codes = new List<long[]>();
        foreach (var n in cl) codes.Add(new long[n]);
        WithPinned(codes, ptrs =>
        {
            fixed (float* a = audioData)
                Native.Check(Native.nc_snac_encode(_h, a, 1, audioData.Length, ptrs), "SNAC", CodecOperation.Encoding);
        });
        return codes;
    }

    /// SNAC.Decode(List<float[]>) (Models/SNAC.cs:173-192): codes as float32 arrays (converted to int64 exactly like
    /// `torch.tensor(code, dtype: torch.int64)`, :185: truncation toward zero), audio [frames * hop], not trimmed.
    /// NoiseBlock noise is a fresh draw per call, as in the reference (NoiseBlock.cs:38-45).
    public float[] Decode(List<float[]> codes)
    {
        ArgumentNullException.ThrowIfNull(codes);                                       // SNAC.cs:175
        if (codes.Count == 0 || codes.Any(c => c == null))
            throw new ArgumentException("Codes list cannot be empty or contain null arrays", nameof(codes));   // SNAC.cs:177-180
        return Decode(codes.ConvertAll(stage => Array.ConvertAll(stage, c => (long)c)), null);
    }

    /// Integer codes; `seed` = null draws a fresh 64-bit seed (a new noise realisation per call), a value makes the
    /// output reproducible and independent of how a batch is split.
    public float[] Decode(List<long[]> codes, ulong? seed = null)
    {
        ArgumentNullException.ThrowIfNull(codes);
        if (codes.Count == 0 || codes.Any(c => c == null))
            throw new ArgumentException("Codes list cannot be empty or contain null arrays", nameof(codes));
        long frames = (long)codes[^1].Length * _config.VQStrides[^1];
        long hop = 1; foreach (var r in _config.EncoderRates) hop *= r;
        var audio = new float[frames * hop];
        ulong s = seed ?? (ulong)Random.Shared.NextInt64();
        WithPinned(codes, ptrs =>
        {
            fixed (float* pa = audio)
                Native.Check(Native.nc_snac_decode(_h, ptrs, 1, frames, null, s, pa), "SNAC", CodecOperation.Decoding);
        });
        return audio;
    }

    private delegate void PinnedCall(long** ptrs);
    private static void WithPinned(List<long[]> arrays, PinnedCall call)
    {
        var ptrs = stackalloc long*[arrays.Count];
        var pins = new System.Runtime.InteropServices.GCHandle[arrays.Count];
        try
        {
            for (int i = 0; i < arrays.Count; i++)
            {
                pins[i] = System.Runtime.InteropServices.GCHandle.Alloc(arrays[i], System.Runtime.InteropServices.GCHandleType.Pinned);
                ptrs[i] = (long*)pins[i].AddrOfPinnedObject();
            }
            call(ptrs);
        }
        finally { foreach (var p in pins) if (p.IsAllocated) p.Free(); }
    }

    /// SNAC.ProcessAudio(float[], sampleRate) (SNAC.cs:255-282): resample on the device when needed, forward, trimmed output.
    public float[] ProcessAudio(float[] audioData, int sampleRate, ulong? noiseSeed = null)
    {
        ulong seed = noiseSeed ?? (ulong)Random.Shared.NextInt64();
        if (audioData is null || audioData.Length == 0) throw new ArgumentException("Audio data cannot be empty", nameof(audioData)); // SNAC.cs:257-258
        long n;
        fixed (float* a = audioData)
        {
            Native.Check(Native.nc_snac_process_audio(_h, a, 1, audioData.Length, sampleRate, null, seed, null, 0, out n), "SNAC", CodecOperation.Encoding);
            var outp = new float[n];
            fixed (float* po = outp)
                Native.Check(Native.nc_snac_process_audio(_h, a, 1, audioData.Length, sampleRate, null, seed, po, n, out n), "SNAC", CodecOperation.Encoding);
            return outp;
        }
    }

    public void Dispose() => _h.Dispose();
}
