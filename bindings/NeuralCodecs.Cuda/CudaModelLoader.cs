// IModelLoader of the CUDA backend: the counterpart of NeuralCodecs.Torch/TorchModelLoader.cs:22-560.
// It reuses the Core package's cache and repositories unchanged (DefaultModelCache, HuggingFaceRepository,
// GitHubRepository, DirectUrlRepository); only the model factory differs: the registry creates CudaDAC / CudaSNAC /
// CudaEncodec from the SAME config classes, and LoadWeights hands the file path to the native library
// (nc_load_weights reads .safetensors and torch.save zip checkpoints itself).  Error wrapping follows the reference:
// LoadException / CacheException / ConfigurationException and the OnError event (TorchModelLoader.cs:198-224,371-383).
// NOT compiled in this repository (no dotnet); see INTEGRATION.md.
using System.Text.Json;
using NeuralCodecs.Core;
using NeuralCodecs.Core.Configuration;
using NeuralCodecs.Core.Events;
using NeuralCodecs.Core.Exceptions;
using NeuralCodecs.Core.Loading;
using NeuralCodecs.Core.Loading.Cache;
using NeuralCodecs.Core.Loading.Repository;
using NeuralCodecs.Core.Validation;
using NeuralCodecs.Torch.Config.DAC;
using NeuralCodecs.Torch.Config.Encodec;
using NeuralCodecs.Torch.Config.SNAC;

namespace NeuralCodecs.Cuda;

public sealed class CudaModelLoader : IModelLoader
{
    private static readonly TimeSpan WeightLoadTimeout = TimeSpan.FromSeconds(90);   // TorchModelLoader.cs:486
    private static readonly JsonSerializerOptions Json = new()
    {
        PropertyNameCaseInsensitive = true,
        ReadCommentHandling = JsonCommentHandling.Skip,
        Converters = { new ModelConfigJsonConverter<IModelConfig>() },
    };

    private readonly IModelCache _cache;
    private readonly ModelRegistry _models = new();
    private readonly Dictionary<Type, object> _validators = new();

    public event EventHandler<LoadErrorEventArgs>? OnError;
    public event EventHandler<LoadProgressEventArgs>? OnProgress;

    public CudaModelLoader(IModelCache? cache = null, IModelValidator<IModelConfig>? validator = null)
    {
        _cache = cache ?? new DefaultModelCache();
        if (validator is not null) RegisterValidator(validator);
        _models.RegisterModel<CudaDAC, DACConfig>(c => new CudaDAC(c));
        _models.RegisterModel<CudaSNAC, SNACConfig>(c => new CudaSNAC(c));
        _models.RegisterModel<CudaEncodec, EncodecConfig>(c => new CudaEncodec(c));
    }

    public void RegisterValidator<TConfig>(IModelValidator<TConfig> validator) where TConfig : IModelConfig =>
        _validators[typeof(TConfig)] = validator;

    public string GetDefaultCacheDirectory() => _cache.GetDefaultCacheDirectory();

    public void ClearCache(string? modelId = null)
    {
        try { _cache.ClearCache(modelId!); }
        catch (Exception ex) { Report(modelId ?? "all models", new LoadException("Failed to clear cache", ex)); }
    }

    /// Same classification as TorchModelLoader.IsLocalPath (:125-145): "owner/repo" and http(s) URLs are remote.
    public bool IsLocalPath(string source)
    {
        bool hubId = source.Count(ch => ch == '/') == 1 && !source.Contains(':') && !source.StartsWith('/') && !source.StartsWith('\\');
        if (hubId) return false;
        if (Uri.TryCreate(source, UriKind.Absolute, out var u) && (u.Scheme == Uri.UriSchemeHttp || u.Scheme == Uri.UriSchemeHttps)) return false;
        return Path.IsPathRooted(source) || File.Exists(source);
    }

    public async Task<ModelMetadata?> GetModelInfo(string source)
    {
        try
        {
            if (IsLocalPath(source))
            {
                if (!File.Exists(source)) return null;
                var fi = new FileInfo(source);
                return new ModelMetadata { Source = source, IsCached = false, LastModified = fi.LastWriteTimeUtc, Size = fi.Length, Backend = "Cuda" };
            }
            var info = await RepositoryFor(source).GetModelInfo(source, "main");
            return new ModelMetadata
            {
                Source = source, IsCached = await _cache.GetCachedPath(source, "main") != null, LastModified = info.LastModified,
                Author = info.Author, Tags = info.Tags, Size = info.Size, Backend = "Cuda",
            };
        }
        catch (Exception ex) { Report(source, ex); return null; }
    }

    public async Task<TModel> LoadModelAsync<TModel, TConfig>(string path, TConfig? config = default, ModelLoadOptions? options = null)
        where TModel : class, INeuralCodec where TConfig : class, IModelConfig
    {
        options ??= new ModelLoadOptions { Device = config?.Device, ValidateModel = config is null };
        config ??= await ReadConfig<TConfig>(ConfigPathFor(path));
        string local = IsLocalPath(path) ? path : await FetchAsync(path, config, options);
        return await CreateAndLoad(local, () => _models.CreateModel<TModel, TConfig>(config), options, isFactory: false);
    }

    public Task<TModel> LoadModelAsync<TModel, TConfig>(string path, Func<IModelConfig, TModel> modelFactory, TConfig config, ModelLoadOptions? options = null)
        where TModel : class, INeuralCodec where TConfig : class, IModelConfig =>
        CreateAndLoad(path, () => modelFactory(config), options ?? new ModelLoadOptions { ValidateModel = false }, isFactory: true);

    /// Config-less load (TorchModelLoader.cs:506-560): the model type's built-in defaults.
    public async Task<TModel> LoadModelAsync<TModel>(string path, ModelLoadOptions? options = null) where TModel : class, INeuralCodec
    {
        options ??= new ModelLoadOptions { ValidateModel = false };
        IModelConfig cfg = typeof(TModel) == typeof(CudaDAC) ? new DACConfig()
                         : typeof(TModel) == typeof(CudaSNAC) ? new SNACConfig()
                         : typeof(TModel) == typeof(CudaEncodec) ? new EncodecConfig()
                         : throw new LoadException($"{typeof(TModel).Name} is not a model of the Cuda backend");
        string local = IsLocalPath(path) ? path : await FetchAsync(path, cfg, options);
        return await CreateAndLoad(local, () => (TModel)(object)(cfg switch
        {
            DACConfig d => new CudaDAC(d), SNACConfig s => new CudaSNAC(s), EncodecConfig e => new CudaEncodec(e),
            _ => throw new LoadException("unreachable"),
        }), options, isFactory: false);
    }

    // ------------------------------------------------------------------ internals
    private void Report(string source, Exception ex) => OnError?.Invoke(this, new LoadErrorEventArgs(source, ex));

    private async Task<TModel> CreateAndLoad<TModel>(string path, Func<TModel> create, ModelLoadOptions options, bool isFactory)
        where TModel : class, INeuralCodec
    {
        if (!isFactory && !File.Exists(path))
        {
            if (path.Contains(".cache"))
            {
                _cache.ClearCache(null!);
                throw new CacheException($"Model file not found at {path}. Clearing Cache.");
            }
            throw new LoadException($"Model file not found at {path}");
        }
        TModel? model = null;
        try
        {
            model = create();
            using var cts = new CancellationTokenSource(WeightLoadTimeout);
            var m = model;
            await Task.Run(() => { m.LoadWeights(path); cts.Token.ThrowIfCancellationRequested(); }, cts.Token);
            if ((options.ValidateModel || isFactory) && _validators.TryGetValue(model.Config.GetType(), out var v))
            {
                var validator = v as IModelValidator<IModelConfig> ?? throw new ConfigurationException("Invalid model validator");
                var result = await validator.ValidateModel(model, model.Config);
                if (!result.IsValid) throw new LoadException($"Model validation failed: {string.Join(", ", result.Errors)}");
            }
            return model;
        }
        catch (Exception ex) when (isFactory || ex is not (LoadException or CacheException))
        {
            model?.Dispose();          // frees the native handle (device weights) of a half-loaded model
            Report(path, ex);
            throw new LoadException(isFactory ? $"Failed to load model using custom factory: {path}"
                                              : $"Failed to load local model: {path}. {ex.Message}", ex);
        }
    }

    private string ConfigPathFor(string modelPath)
    {
        var beside = Path.ChangeExtension(modelPath, ".json");
        if (File.Exists(beside)) return beside;
        var inDir = Path.Combine(Path.GetDirectoryName(modelPath) ?? "", "config.json");
        if (File.Exists(inDir)) return inDir;
        throw new FileNotFoundException($"Config file not found at {inDir}");
    }

    private async Task<TConfig> ReadConfig<TConfig>(string path) where TConfig : IModelConfig
    {
        TConfig cfg;
        try
        {
            cfg = JsonSerializer.Deserialize<TConfig>(await File.ReadAllTextAsync(path), Json) ?? throw new LoadException("Failed to deserialize config");
        }
        catch (Exception ex) when (ex is not LoadException) { throw new LoadException($"Failed to load config from {path}", ex); }
        if (_validators.TryGetValue(cfg.GetType(), out var v))
        {
            var validator = v as IModelValidator<IModelConfig> ?? throw new ConfigurationException("Invalid model validator");
            var result = validator.ValidateConfig(cfg);
            if (!result.IsValid) throw new ConfigurationException($"Invalid model configuration: {string.Join(", ", result.Errors)}");
        }
        return cfg;
    }

    private static IModelRepository RepositoryFor(string source)
    {
        if (Uri.TryCreate(source, UriKind.Absolute, out var uri))
        {
            if (uri.Host.Equals("github.com", StringComparison.OrdinalIgnoreCase)) return new GitHubRepository();
            var direct = new DirectUrlRepository();
            if (direct.CanHandleUrl(source)) return direct;
        }
        if (source.Count(ch => ch == '/') == 1 && !source.Contains(':')) return new HuggingFaceRepository();
        throw new InvalidDataException($"Unsupported model source: {source}");
    }

    /// Cache lookup, else download into a temp directory and move into the cache (TorchModelLoader.cs:385-450).
    private async Task<string> FetchAsync(string source, IModelConfig config, ModelLoadOptions options)
    {
        try
        {
            if (!options.ForceReload && await _cache.GetCachedPath(source, options.Revision) is { } hit) return hit;
            var repo = RepositoryFor(source);
            if (repo is GitHubRepository && !string.IsNullOrEmpty(config.Version)) options.Revision = config.Version;
            var meta = await repo.GetModelInfo(source, options.Revision);
            var tmp = Directory.CreateDirectory(Path.Combine(Path.GetTempPath(), $"neural_codecs_{Guid.NewGuid()}")).FullName;
            try
            {
                await repo.DownloadModel(source, tmp, new Progress<double>(p => OnProgress?.Invoke(this, new LoadProgressEventArgs(source, p))), options);
                return await _cache.CacheModel(meta.Source, tmp, options.Revision, meta.FileName, meta.ConfigFileName);
            }
            finally { try { Directory.Delete(tmp, recursive: true); } catch { /* best effort */ } }
        }
        catch (Exception ex) when (ex is not LoadException)
        {
            Report(source, ex);
            _cache.ClearCache(source);
            throw new LoadException($"Failed to load remote model: {source}. {ex.Message}", ex);
        }
    }
}
