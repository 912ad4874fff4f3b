// Drop-in counterpart of NeuralCodecs.Torch/Models/DAC.cs over the C ABI: same config class
// (DACConfig), same method names, managed arrays instead of TorchSharp tensors.
// NOT compiled in this repository (no dotnet); its executable twin is neuralcodecs_b200/dac.py.
using System;
using NeuralCodecs.Core;
using NeuralCodecs.Core.Configuration;
using NeuralCodecs.Core.Exceptions;
using NeuralCodecs.Torch.Config.DAC;   // the reference's own DACConfig (JSON names, presets) is reused unchanged

namespace NeuralCodecs.Cuda;

public sealed unsafe class CudaDAC : INeuralCodec
{
    private readonly DACConfig _config;
    private readonly NcHandle _h;
    private readonly int _latentDim, _hop;

    public IModelConfig Config => _config;

    public CudaDAC(DACConfig config)                                        // Models/DAC.cs:51-93
    {
        _config = config ?? throw new ArgumentNullException(nameof(config)); // DAC.cs:53
        var er = config.EncoderRates ?? new[] { 2, 4, 8, 8 };                // DAC.cs:57
        var dr = config.DecoderRates ?? new[] { 8, 8, 4, 2 };                // DAC.cs:59
        var c = new NcDacConfig
        {
            StructSize = (uint)sizeof(NcDacConfig), SampleRate = config.SampleRate, EncoderDim = config.EncoderDim,
            NEncoderRates = er.Length, DecoderDim = config.DecoderDim, NDecoderRates = dr.Length,
            NCodebooks = config.NumCodebooks, CodebookSize = config.CodebookSize, CodebookDim = config.CodebookDim,
            LatentDim = config.LatentDim ?? 0,
        };
        _hop = 1;
        for (int i = 0; i < er.Length; i++) { c.EncoderRates[i] = er[i]; _hop *= er[i]; }
        for (int i = 0; i < dr.Length; i++) c.DecoderRates[i] = dr[i];
        _latentDim = config.LatentDim ?? config.EncoderDim * (1 << er.Length); // DAC.cs:64
        Native.Check(Native.nc_create(NcCodecKind.Dac, &c, (nuint)sizeof(NcDacConfig), config.Device?.Index ?? 0, out _h),
                     "DAC", CodecOperation.Initialization);
    }

    public void LoadWeights(string path) =>                                   // Models/DAC.cs:345-389
        Native.Check(Native.nc_load_weights(_h, path), "DAC", CodecOperation.Initialization);

    /// DAC.Encode(Tensor, int?, int?) (DAC.cs:163-181): audio [B,1,L] row-major.
    public (float[] z, long[] codes, float[] latents, int frames) Encode(float[] audio, int batch, int? nQuantizers = null, int? sampleRate = null)
    {
        if (audio is null) throw new ArgumentNullException(nameof(audio));
        long L = audio.Length / batch;
        Native.Check(Native.nc_dac_query_shapes(_h, L, out _, out long T, out int D, out int nCb, out int cbDim), "DAC", CodecOperation.Encoding);
        int nq = Math.Min(nQuantizers ?? nCb, nCb);
        var z = new float[batch * D * T]; var codes = new long[batch * nq * T]; var lat = new float[batch * nq * cbDim * T];
        fixed (float* a = audio, pz = z, pl = lat) fixed (long* pc = codes)
            Native.Check(Native.nc_dac_encode(_h, a, batch, L, sampleRate ?? 0, nq, pz, pc, pl, out _), "DAC", CodecOperation.Encoding);
        return (z, codes, lat, (int)T);
    }

    /// DAC.Encode(float[]) (DAC.cs:205-224): batch 1, returns the quantised latent z (NOT codes).
    public float[] Encode(float[] audioData)
    {
        if (audioData is null) throw new ArgumentNullException(nameof(audioData)); // DAC.cs:207
        Native.Check(Native.nc_dac_query_shapes(_h, audioData.Length, out _, out long T, out int D, out _, out _), "DAC", CodecOperation.Encoding);
        var z = new float[D * T];
        fixed (float* a = audioData, pz = z)
            Native.Check(Native.nc_dac_encode(_h, a, 1, audioData.Length, 0, 0, pz, null, null, out _), "DAC", CodecOperation.Encoding);
        return z;
    }

    /// DAC.Decode(float[]) (DAC.cs:241-253): z reshaped to [1, latentDim, -1]; output NOT trimmed.
    public float[] Decode(float[] qAudio)
    {
        if (qAudio is null) throw new ArgumentNullException(nameof(qAudio));
        long T = qAudio.Length / _latentDim;
        var audio = new float[T * _hop];
        fixed (float* pz = qAudio, pa = audio)
            Native.Check(Native.nc_dac_decode(_h, pz, 1, T, pa), "DAC", CodecOperation.Decoding);
        return audio;
    }

    /// DAC.FromCodes (DAC.cs:101-106): codes [B,nq,T] -> z [B,latent,T].
    public float[] FromCodes(long[] codes, int batch, int nQuantizers)
    {
        long T = codes.Length / ((long)batch * nQuantizers);
        var z = new float[batch * _latentDim * T];
        fixed (long* pc = codes) fixed (float* pz = z)
            Native.Check(Native.nc_dac_from_codes(_h, pc, batch, nQuantizers, T, pz), "DAC", CodecOperation.Decoding);
        return z;
    }

    /// Batched Dia stage (Models/Dia.cs:973-981,1057-1060): codes [B,nq,T] -> audio [B,1,T*hop] in one call.
    public float[] DecodeCodes(long[] codes, int batch, int nQuantizers)
    {
        long T = codes.Length / ((long)batch * nQuantizers);
        var audio = new float[batch * T * _hop];
        fixed (long* pc = codes) fixed (float* pa = audio)
            Native.Check(Native.nc_dac_decode_codes(_h, pc, batch, nQuantizers, T, pa), "DAC", CodecOperation.Decoding);
        return audio;
    }

    /// DAC.forward(float[]) (DAC.cs:310-322).
    public float[] forward(float[] audioData)
    {
        Native.Check(Native.nc_dac_query_shapes(_h, audioData.Length, out long Lp, out _, out _, out _, out _), "DAC", CodecOperation.Encoding);
        var outp = new float[Lp];
        fixed (float* a = audioData, po = outp)
            Native.Check(Native.nc_dac_forward(_h, a, 1, audioData.Length, 0, po, null, null, out _), "DAC", CodecOperation.Encoding);
        return outp;
    }

    public void Dispose() => _h.Dispose();                                    // Models/DAC.cs:328-337
}
