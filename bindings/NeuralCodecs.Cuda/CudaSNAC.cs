// Drop-in counterpart of NeuralCodecs.Torch/Models/SNAC.cs over the C ABI: same config class (SNACConfig), same method
// names, managed arrays instead of TorchSharp tensors.  NOT compiled in this repository (no dotnet); its executable
// twin is neuralcodecs_b200/snac.py.
using System;
using System.Collections.Generic;
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

    /// SNAC.Encode(float[]) (SNAC.cs:129-150): one code array per VQ stage (codes of the padded audio), batch 1.
    public List<long[]> Encode(float[] audioData)
    {
        if (audioData is null) throw new ArgumentNullException(nameof(audioData));
        var (_, _, cl, _) = Shapes(audioData.Length);
        var codes = new List<long[]>();
        foreach (var n in cl) codes.Add(new long[n]);
        var ptrs = stackalloc long*[cl.Length];
        var pins = new System.Runtime.InteropServices.GCHandle[cl.Length];
        try
        {
            for (int i = 0; i < cl.Length; i++)
            {
                pins[i] = System.Runtime.InteropServices.GCHandle.Alloc(codes[i], System.Runtime.InteropServices.GCHandleType.Pinned);
                ptrs[i] = (long*)pins[i].AddrOfPinnedObject();
            }
            fixed (float* a = audioData)
                Native.Check(Native.nc_snac_encode(_h, a, 1, audioData.Length, ptrs), "SNAC", CodecOperation.Encoding);
        }
        finally { foreach (var p in pins) if (p.IsAllocated) p.Free(); }
        return codes;
    }

    /// SNAC.Decode(List<...>) (SNAC.cs:157-192): audio [frames * hop], not trimmed; NoiseBlock noise drawn on the device from `seed`.
    public float[] Decode(List<long[]> codes, ulong seed = 0)
    {
        if (codes is null || codes.Count == 0) throw new ArgumentException("Codes list cannot be empty or contain null arrays"); // SNAC.cs:177-180
        long frames = codes[^1].Length * _config.VQStrides[^1];
        long hop = 1; foreach (var r in _config.EncoderRates) hop *= r;
        var audio = new float[frames * hop];
        var ptrs = stackalloc long*[codes.Count];
        var pins = new System.Runtime.InteropServices.GCHandle[codes.Count];
        try
        {
            for (int i = 0; i < codes.Count; i++)
            {
                pins[i] = System.Runtime.InteropServices.GCHandle.Alloc(codes[i], System.Runtime.InteropServices.GCHandleType.Pinned);
                ptrs[i] = (long*)pins[i].AddrOfPinnedObject();
            }
            fixed (float* pa = audio)
                Native.Check(Native.nc_snac_decode(_h, ptrs, 1, frames, null, seed, pa), "SNAC", CodecOperation.Decoding);
        }
        finally { foreach (var p in pins) if (p.IsAllocated) p.Free(); }
        return audio;
    }

    /// SNAC.ProcessAudio(float[], sampleRate) (SNAC.cs:255-282): resample on the device when needed, forward, trimmed output.
    public float[] ProcessAudio(float[] audioData, int sampleRate, ulong seed = 0)
    {
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
