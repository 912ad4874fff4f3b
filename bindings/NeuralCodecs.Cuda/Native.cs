// P/Invoke surface of libneuralcodecs_cuda.so (include/neuralcodecs_cuda.h).
// Written against .NET 8 [LibraryImport]; NOT compiled in this repository (no dotnet toolchain in
// the build image) -- the executable twin of this file is neuralcodecs_b200/_lib.py (ctypes).
using System;
using System.Runtime.InteropServices;

namespace NeuralCodecs.Cuda;

internal enum NcStatus
{
    Ok = 0, InvalidArgument = 1, FileNotFound = 2, BadWeights = 3, ShapeMismatch = 4,
    CudaUnavailable = 5, CudaError = 6, OutOfMemory = 7, Internal = 8, Unsupported = 9,
}

internal enum NcCodecKind { Dac = 1, Snac = 2, Encodec = 3 }

[StructLayout(LayoutKind.Sequential)]
internal unsafe struct NcDacConfig
{
    public uint StructSize;
    public int SampleRate, EncoderDim, NEncoderRates;
    public fixed int EncoderRates[8];
    public int DecoderDim, NDecoderRates;
    public fixed int DecoderRates[8];
    public int NCodebooks, CodebookSize, CodebookDim, LatentDim;
}

[StructLayout(LayoutKind.Sequential)]
internal unsafe struct NcSnacConfig      // nc_snac_config (include/neuralcodecs_cuda.h)
{
    public uint StructSize;
    public int SampleRate, EncoderDim, NEncoderRates;
    public fixed int EncoderRates[8];
    public int DecoderDim, NDecoderRates;
    public fixed int DecoderRates[8];
    public int LatentDim, AttnWindowSize, CodebookSize, CodebookDim, NVqStrides;
    public fixed int VqStrides[8];
    public int Noise, Depthwise;
}

[StructLayout(LayoutKind.Sequential)]
internal unsafe struct NcEncodecConfig   // nc_encodec_config
{
    public uint StructSize;
    public int SampleRate, Channels, NFilters, Dimension, NRatios;
    public fixed int Ratios[8];
    public int NResidualLayers, LstmLayers, CodebookSize, NQuantizers, Causal;
    public int NormType;      // 0 = "weight_norm", 1 = "time_group_norm"
    public int Normalize;
    public float SegmentS;    // ChunkLengthSeconds, 0 = one frame per clip
    public float Overlap;
}

internal sealed class NcHandle : SafeHandle
{
    public NcHandle() : base(IntPtr.Zero, true) { }
    public override bool IsInvalid => handle == IntPtr.Zero;
    protected override bool ReleaseHandle() => Native.nc_destroy(handle) == NcStatus.Ok;
}

internal static unsafe partial class Native
{
    private const string Lib = "neuralcodecs_cuda";

    [LibraryImport(Lib)] internal static partial IntPtr nc_version();
    [LibraryImport(Lib)] internal static partial IntPtr nc_last_error();
    [LibraryImport(Lib)] internal static partial int nc_device_count();
    [LibraryImport(Lib)] internal static partial NcStatus nc_create(NcCodecKind kind, void* cfg, nuint cfgSize, int deviceIndex, out NcHandle handle);
    [LibraryImport(Lib)] internal static partial NcStatus nc_destroy(IntPtr handle);
    [LibraryImport(Lib, StringMarshalling = StringMarshalling.Utf8)] internal static partial NcStatus nc_load_weights(NcHandle h, string path);
    [LibraryImport(Lib, StringMarshalling = StringMarshalling.Utf8)] internal static partial NcStatus nc_set_option(NcHandle h, string key, string value);
    [LibraryImport(Lib)] internal static partial NcStatus nc_dac_query_shapes(NcHandle h, long length, out long paddedLength, out long frames, out int latentDim, out int nCodebooks, out int codebookDim);
    [LibraryImport(Lib)] internal static partial NcStatus nc_dac_encode(NcHandle h, float* audio, int batch, long length, int sampleRate, int nQuantizers, float* z, long* codes, float* latents, out long frames);
    [LibraryImport(Lib)] internal static partial NcStatus nc_dac_decode(NcHandle h, float* z, int batch, long frames, float* audio);
    [LibraryImport(Lib)] internal static partial NcStatus nc_dac_from_codes(NcHandle h, long* codes, int batch, int nQuantizers, long frames, float* z);
    [LibraryImport(Lib)] internal static partial NcStatus nc_dac_decode_codes(NcHandle h, long* codes, int batch, int nQuantizers, long frames, float* audio);
    [LibraryImport(Lib)] internal static partial NcStatus nc_dac_forward(NcHandle h, float* audio, int batch, long length, int nQuantizers, float* audioOut, long* codes, float* z, out long frames);
    [LibraryImport(Lib)] internal static partial NcStatus nc_dac_decode_dia(NcHandle h, long* generated, int batch, int steps, int channels, int* delayPattern, long* lengths, float* audio, long audioStride);
    // SNAC (Models/SNAC.cs): codes / noise are arrays of per-stage / per-block pointers
    [LibraryImport(Lib)] internal static partial NcStatus nc_snac_query_shapes(NcHandle h, long length, out long paddedLength, out long frames, out int nStages, long* codeLengths, out int nNoise, long* noiseLengths);
    [LibraryImport(Lib)] internal static partial NcStatus nc_snac_encode(NcHandle h, float* audio, int batch, long length, long** codes);
    [LibraryImport(Lib)] internal static partial NcStatus nc_snac_decode(NcHandle h, long** codes, int batch, long frames, float** noise, ulong seed, float* audio);
    [LibraryImport(Lib)] internal static partial NcStatus nc_snac_forward(NcHandle h, float* audio, int batch, long length, float** noise, ulong seed, float* audioOut, long** codes);
    [LibraryImport(Lib)] internal static partial NcStatus nc_snac_process_audio(NcHandle h, float* audio, int batch, long length, int sampleRate, float** noise, ulong seed, float* audioOut, long outCapacity, out long outLength);
    // weight files and input conditioning (Config/DAC/DACUnpickler.cs, NeuralCodecs.Core/Utils/AudioUtils.cs)
    [LibraryImport(Lib, StringMarshalling = StringMarshalling.Utf8)] internal static partial NcStatus nc_inspect_weights(string path, byte* buf, nuint bufSize);
    [LibraryImport(Lib)] internal static partial NcStatus nc_resample_linear(NcHandle h, float* audio, int batch, long length, int srcRate, int dstRate, float* output, long outCapacity, out long outLength);
    [LibraryImport(Lib)] internal static partial NcStatus nc_convert_to_mono(NcHandle h, float* interleaved, long frames, int channels, float* output);
    // Encodec (Models/Encodec.cs, Modules/Encodec/EncodecCompressor.cs)
    [LibraryImport(Lib)] internal static partial NcStatus nc_encodec_query_shapes(NcHandle h, long length, float bandwidthKbps, out long frames, out int nQ, out long decodedLength);
    [LibraryImport(Lib)] internal static partial NcStatus nc_encodec_encode(NcHandle h, float* audio, int batch, long length, float bandwidthKbps, long* codes);
    [LibraryImport(Lib)] internal static partial NcStatus nc_encodec_decode(NcHandle h, long* codes, int batch, int nQ, long frames, float* audio);
    [LibraryImport(Lib)] internal static partial NcStatus nc_encodec_forward(NcHandle h, float* audio, int batch, long length, float bandwidthKbps, float* audioOut, long* codes);
    [LibraryImport(Lib)] internal static partial NcStatus nc_encodec_query_frames(NcHandle h, long length, float bandwidthKbps, out int nSegments, long* segFrames, int segFramesCapacity, out long totalFrames, out int nQ, out long decodedLength);
    [LibraryImport(Lib)] internal static partial NcStatus nc_encodec_encode_frames(NcHandle h, float* audio, int batch, long length, float bandwidthKbps, long* codes, float* scales);
    [LibraryImport(Lib)] internal static partial NcStatus nc_encodec_decode_frames(NcHandle h, long* codes, float* scales, int batch, int nQ, long* segFrames, int nSegments, float* audio);
    [LibraryImport(Lib)] internal static partial NcStatus nc_encodec_query_decoded(NcHandle h, long* segFrames, int nSegments, out long decodedLength);
    [LibraryImport(Lib)] internal static partial NcStatus nc_encodec_ecdc_size(NcHandle h, long length, float bandwidthKbps, out long headerBytes, out long streamBytes);
    [LibraryImport(Lib)] internal static partial NcStatus nc_encodec_compress(NcHandle h, float* audio, int batch, long length, float bandwidthKbps, byte* output, long outStride, out long streamBytes);
    [LibraryImport(Lib)] internal static partial NcStatus nc_encodec_ecdc_info(byte* stream, long streamBytes, out long audioLength, out int nQ, out int channels, out int sampleRate, out float bandwidthKbps, out int useLm, out long payloadOffset);
    [LibraryImport(Lib)] internal static partial NcStatus nc_encodec_decompress(NcHandle h, byte* streams, int batch, long streamStride, long streamBytes, float* audio, long audioCapacity, out long audioLength, out int sampleRate);

    /// Status -> the reference's exception conventions (SURVEY 8b "Error conventions").
    internal static void Check(NcStatus s, string codec, Core.Exceptions.CodecOperation op)
    {
        if (s == NcStatus.Ok) return;
        string msg = Marshal.PtrToStringUTF8(nc_last_error()) ?? s.ToString();
        throw s switch
        {
            NcStatus.InvalidArgument => new ArgumentException(msg),                       // Models/DAC.cs:146,207
            NcStatus.FileNotFound => new System.IO.FileNotFoundException(msg),            // Models/DAC.cs:347-351
            NcStatus.BadWeights or NcStatus.ShapeMismatch or NcStatus.Unsupported
                => new InvalidOperationException(msg),                                    // Models/DAC.cs:386-388
            NcStatus.CudaUnavailable => new InvalidOperationException("CUDA requested but not available"), // Utils/TorchUtils.cs:97-99
            NcStatus.OutOfMemory => new OutOfMemoryException(msg),
            _ => new Core.Exceptions.CodecException(codec, op, msg),                      // Core/Exceptions/CodecException.cs:8-40
        };
    }
}
