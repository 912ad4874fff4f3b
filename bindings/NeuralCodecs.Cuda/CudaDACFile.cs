// `.dac` code container for the Cuda backend: the counterpart of NeuralCodecs.Torch/AudioTools/DACFile.cs:10-105 over
// managed arrays instead of TorchSharp tensors.  Same byte layout (BinaryWriter little-endian: Int32 JSON length,
// length-prefixed JSON string, Int32 tensor count, per tensor Int32 rank, Int64 dims, Int32 count, Int32 values), same
// System.Text.Json serialisation of the same DACConfig, so files written by either class load in the other.
// NOT compiled in this repository (no dotnet); its executable twin is neuralcodecs_b200/dac_file.py.
using System.Text.Json;
using NeuralCodecs.Torch.Config.DAC;

namespace NeuralCodecs.Cuda;

/// One code tensor of a `.dac` file: row-major values with their shape (DAC.Encode's codes are [batch, n_codebooks, frames]).
public sealed record DacCodes(long[] Shape, long[] Values);

public sealed class CudaDACFile
{
    public List<DacCodes> Codes { get; }
    public DACConfig Config { get; }

    public CudaDACFile(List<DacCodes> codes, DACConfig config)                 // DACFile.cs:14-18
    {
        Codes = codes ?? throw new ArgumentNullException(nameof(codes));
        Config = config ?? throw new ArgumentNullException(nameof(config));
    }

    /// DACFile.LoadAsync (DACFile.cs:27-62).
    public static async Task<CudaDACFile> LoadAsync(string path)
    {
        var bytes = await File.ReadAllBytesAsync(path);
        using var reader = new BinaryReader(new MemoryStream(bytes, writable: false));
        _ = reader.ReadInt32();                                                // configLength: stored, never used
        var config = JsonSerializer.Deserialize<DACConfig>(reader.ReadString())
                     ?? throw new InvalidDataException("missing DAC config in .dac file");
        var codes = new List<DacCodes>();
        int n = reader.ReadInt32();
        for (int i = 0; i < n; i++)
        {
            var shape = new long[reader.ReadInt32()];
            for (int j = 0; j < shape.Length; j++) shape[j] = reader.ReadInt64();
            var values = new long[reader.ReadInt32()];
            for (int j = 0; j < values.Length; j++) values[j] = reader.ReadInt32();
            long expect = 1; foreach (var d in shape) expect *= d;
            if (expect != values.Length)                                       // tensor(data).reshape(shape) throws in the reference
                throw new InvalidDataException($"code tensor {i}: {values.Length} values for shape [{string.Join(",", shape)}]");
            codes.Add(new DacCodes(shape, values));
        }
        return new CudaDACFile(codes, config);
    }

    /// DACFile.SaveAsync (DACFile.cs:72-103).
    public async Task SaveAsync(string path)
    {
        using var mem = new MemoryStream();
        using (var writer = new BinaryWriter(mem, System.Text.Encoding.UTF8, leaveOpen: true))
        {
            var json = JsonSerializer.Serialize(Config);
            writer.Write(json.Length);
            writer.Write(json);
            writer.Write(Codes.Count);
            foreach (var code in Codes)
            {
                writer.Write(code.Shape.Length);
                foreach (var dim in code.Shape) writer.Write(dim);
                writer.Write(code.Values.Length);
                foreach (var v in code.Values) writer.Write((int)v);          // code.to(int32)
            }
        }
        await File.WriteAllBytesAsync(path, mem.ToArray());
    }
}
