// Facade of the CUDA backend: the counterpart of NeuralCodecs.Torch/NeuralCodecs.cs:20-98
// (CreateTorchLoader / CreateSNACAsync / CreateDACAsync / CreateEncodecAsync).  A caller switches backend by
// replacing `NeuralCodecs.CreateDACAsync(path, config)` with `NeuralCodecsCuda.CreateDACCudaAsync(path, config)`.
// The reference calls model.eval() after loading; the CUDA engines have no training mode (inference only), so there
// is nothing to switch.  NOT compiled in this repository (no dotnet); see INTEGRATION.md.
using NeuralCodecs.Core.Loading;
using NeuralCodecs.Torch.Config.DAC;
using NeuralCodecs.Torch.Config.Encodec;
using NeuralCodecs.Torch.Config.SNAC;

namespace NeuralCodecs.Cuda;

public static class NeuralCodecsCuda
{
    /// Counterpart of NeuralCodecs.CreateTorchLoader (NeuralCodecs.cs:20-23).
    public static CudaModelLoader CreateCudaLoader() => new();

    /// Counterpart of NeuralCodecs.CreateSNACAsync (NeuralCodecs.cs:38-44).
    public static Task<CudaSNAC> CreateSNACCudaAsync(string path, SNACConfig? config = null, ModelLoadOptions? options = null) =>
        new CudaModelLoader().LoadModelAsync<CudaSNAC, SNACConfig>(path, config, options);

    /// Counterpart of NeuralCodecs.CreateDACAsync (NeuralCodecs.cs:57-64): same default options (no config file needed,
    /// no validation) -- a DAC `.pth` carries its own config in the checkpoint metadata (CudaDAC.LoadWeights).
    public static Task<CudaDAC> CreateDACCudaAsync(string path, DACConfig? config = null, ModelLoadOptions? options = null)
    {
        options ??= new ModelLoadOptions { HasConfigFile = false, ValidateModel = false };
        return new CudaModelLoader().LoadModelAsync<CudaDAC, DACConfig>(path, config ?? new DACConfig(), options);
    }

    /// Counterpart of NeuralCodecs.CreateEncodecAsync (NeuralCodecs.cs:75-81).
    public static Task<CudaEncodec> CreateEncodecCudaAsync(string path, EncodecConfig? config = null, ModelLoadOptions? options = null) =>
        new CudaModelLoader().LoadModelAsync<CudaEncodec, EncodecConfig>(path, config, options);

    /// Number of sm_100 devices the native library can use (0: the Cuda backend is unavailable on this machine).
    public static int DeviceCount => Native.nc_device_count();
}
