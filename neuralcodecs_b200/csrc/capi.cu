// C ABI of libneuralcodecs_cuda.so (include/neuralcodecs_cuda.h).  No exception crosses the
// boundary; the last error message is kept per thread.
#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <new>

#include "ecdc.h"
#include "pth_reader.h"
#include "engine.h"

using namespace nc;

struct nc_handle_s {
  Engine* engine = nullptr;
  nc_codec_kind kind;
};

namespace {
thread_local std::string g_last_error;

nc_status fail(nc_status s, const std::string& msg) {
  g_last_error = msg;
  return s;
}

template <typename F>
nc_status guarded(F&& f) {
  try {
    f();
    g_last_error.clear();
    return NC_OK;
  } catch (const Error& e) {
    return fail(e.status, e.what());
  } catch (const std::bad_alloc&) {
    return fail(NC_OUT_OF_MEMORY, "host allocation failed");
  } catch (const std::exception& e) {
    return fail(NC_INTERNAL, e.what());
  } catch (...) {
    return fail(NC_INTERNAL, "unknown error");
  }
}

struct BusyGuard {
  Engine* e;
  explicit BusyGuard(Engine* eng) : e(eng) {
    if (e->busy.exchange(true, std::memory_order_acquire))
      throw Error(NC_INVALID_ARGUMENT, "handle is in use by another call (handles are not re-entrant)");
  }
  ~BusyGuard() { e->busy.store(false, std::memory_order_release); }
};

DacEngine* dac_of(nc_handle h) {
  if (!h || !h->engine) throw Error(NC_INVALID_ARGUMENT, "null handle");
  if (h->kind != NC_CODEC_DAC) throw Error(NC_INVALID_ARGUMENT, "handle is not a DAC codec");
  return static_cast<DacEngine*>(h->engine);
}

// Scoped device staging buffer for the host-pointer entry points.  cudaMalloc / cudaFree cost milliseconds and
// serialise the device, so freed blocks go to a small per-device cache and repeated calls with the same shapes (the
// steady state of a serving loop) allocate nothing.  A block returns to the cache only from the destructor, i.e.
// after the entry point synchronised on every use of it.
class StagingCache {
 public:
  void* take(int dev, size_t bytes) {
    std::lock_guard<std::mutex> g(mu_);
    auto& pool = pools_[dev];
    auto it = pool.lower_bound(bytes);
    if (it != pool.end() && it->first <= bytes + bytes / 4 + (1u << 20)) {
      void* p = it->second;
      cached_ -= it->first;
      sizes_[p] = it->first;
      pool.erase(it);
      return p;
    }
    return nullptr;
  }
  void remember(void* p, size_t bytes) { std::lock_guard<std::mutex> g(mu_); sizes_[p] = bytes; }
  // true when the block was cached (caller must not free it)
  bool give(int dev, void* p) {
    std::lock_guard<std::mutex> g(mu_);
    auto it = sizes_.find(p);
    if (it == sizes_.end()) return false;
    const size_t bytes = it->second;
    sizes_.erase(it);
    if (cached_ + bytes > kMaxCached) return false;
    pools_[dev].emplace(bytes, p);
    cached_ += bytes;
    return true;
  }
  // release every cached block of a device (nc_destroy: a destroyed model leaves no device memory behind)
  void trim(int dev) {
    std::lock_guard<std::mutex> g(mu_);
    auto it = pools_.find(dev);
    if (it == pools_.end()) return;
    for (auto& kv : it->second) { cached_ -= kv.first; cudaFree(kv.second); }
    it->second.clear();
  }
 private:
  static constexpr size_t kMaxCached = (size_t)12 << 30;
  std::mutex mu_;
  std::map<int, std::multimap<size_t, void*>> pools_;
  std::map<void*, size_t> sizes_;
  size_t cached_ = 0;
};
StagingCache& staging_cache() { static StagingCache* c = new StagingCache(); return *c; }   // never destroyed: outlives the CUDA context safely

struct DevMem {
  void* p = nullptr;
  int dev = 0;
  explicit DevMem(size_t bytes) {
    if (!bytes) return;
    cudaGetDevice(&dev);
    p = staging_cache().take(dev, bytes);
    if (!p) {
      NC_CUDA(cudaMalloc(&p, bytes));
      staging_cache().remember(p, bytes);
    }
  }
  ~DevMem() {
    if (!p) return;
    // unwinding from an entry point that already queued work on this block: drain the device before another handle
    // or thread can take the block from the cache
    if (std::uncaught_exceptions() > 0) cudaDeviceSynchronize();
    if (!staging_cache().give(dev, p)) cudaFree(p);
  }
  DevMem(const DevMem&) = delete;
  DevMem& operator=(const DevMem&) = delete;
  template <typename T>
  T* as() { return static_cast<T*>(p); }
};

// Host <-> device staging copies of the host-pointer entry points.  They are issued on the ENGINE's stream (created
// cudaStreamNonBlocking, so the legacy default stream gives no ordering against it): an H2D copy is stream-ordered
// before the kernels that read it, a D2H copy after the kernels that wrote it, and the call returns only after the
// stream drained, so the caller's buffers are complete and reusable.
void h2d(Engine* e, void* dst, const void* src, size_t bytes) {
  if (!bytes) return;
  NC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, e->stream()));
}
void d2h(Engine* e, void* dst, const void* src, size_t bytes) {
  if (!bytes) return;
  NC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, e->stream()));
  NC_CUDA(cudaStreamSynchronize(e->stream()));
}
void h2d_2d(Engine* e, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height) {
  if (!width || !height) return;
  NC_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice, e->stream()));
}
void d2h_2d(Engine* e, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height) {
  if (!width || !height) return;
  NC_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost, e->stream()));
  NC_CUDA(cudaStreamSynchronize(e->stream()));
}
}  // namespace

extern "C" {

const char* nc_version(void) { return "neuralcodecs_cuda 0.1.0 (sm_100a)"; }
const char* nc_last_error(void) { return g_last_error.c_str(); }

int nc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp p{};
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
  }
  return ok;
}

nc_status nc_create(nc_codec_kind kind, const void* cfg, size_t cfg_size, int device_index, nc_handle* out) {
  return guarded([&] {
    if (!out) throw Error(NC_INVALID_ARGUMENT, "out handle is null");
    *out = nullptr;
    Engine* e = create_engine(kind, cfg, cfg_size, device_index);
    nc_handle h = new nc_handle_s;
    h->engine = e;
    h->kind = kind;
    *out = h;
  });
}

nc_status nc_destroy(nc_handle h) {
  return guarded([&] {
    if (!h) return;
    if (h->engine) {
      h->engine->bind();
      int dev = 0;
      cudaGetDevice(&dev);
      staging_cache().trim(dev);
    }
    delete h->engine;
    delete h;
  });
}

nc_status nc_load_weights(nc_handle h, const char* path) {
  return guarded([&] {
    if (!h || !h->engine) throw Error(NC_INVALID_ARGUMENT, "null handle");
    if (!path) throw Error(NC_INVALID_ARGUMENT, "path is null");
    BusyGuard g(h->engine);
    h->engine->bind();
    h->engine->load_weights(path);
  });
}

nc_status nc_set_tensor(nc_handle h, const char* name, int dtype, int rank, const int64_t* shape, const void* data) {
  return guarded([&] {
    if (!h || !h->engine) throw Error(NC_INVALID_ARGUMENT, "null handle");
    if (!name || !data || rank < 0 || rank > 8 || (rank > 0 && !shape)) throw Error(NC_INVALID_ARGUMENT, "bad tensor arguments");
    HostTensor t;
    t.shape.assign(shape, shape + rank);
    for (auto d : t.shape)
      if (d < 0) throw Error(NC_INVALID_ARGUMENT, "negative dimension");
    const size_t n = t.numel();
    if (dtype == 0) {
      t.f32.assign(static_cast<const float*>(data), static_cast<const float*>(data) + n);
    } else if (dtype == 1) {
      t.is_int = true;
      t.i64.assign(static_cast<const int64_t*>(data), static_cast<const int64_t*>(data) + n);
    } else {
      throw Error(NC_INVALID_ARGUMENT, "dtype must be 0 (f32) or 1 (i64)");
    }
    BusyGuard g(h->engine);
    h->engine->set_tensor(name, std::move(t));
  });
}

nc_status nc_finalize_weights(nc_handle h) {
  return guarded([&] {
    if (!h || !h->engine) throw Error(NC_INVALID_ARGUMENT, "null handle");
    BusyGuard g(h->engine);
    h->engine->bind();
    h->engine->finalize_weights();
  });
}

nc_status nc_set_option(nc_handle h, const char* key, const char* value) {
  return guarded([&] {
    if (!h || !h->engine) throw Error(NC_INVALID_ARGUMENT, "null handle");
    if (!key || !value) throw Error(NC_INVALID_ARGUMENT, "null option");
    BusyGuard g(h->engine);
    h->engine->set_option(key, value);
  });
}

// ------------------------------------------------------------------------------------ DAC
nc_status nc_dac_query_shapes(nc_handle h, int64_t length, int64_t* padded_length, int64_t* frames, int32_t* latent_dim,
                              int32_t* n_codebooks, int32_t* codebook_dim) {
  return guarded([&] {
    DacEngine* e = dac_of(h);
    if (length < 0) throw Error(NC_INVALID_ARGUMENT, "length must be non-negative");
    if (padded_length) *padded_length = e->padded_length(length);
    if (frames) *frames = e->frames(length);
    if (latent_dim) *latent_dim = e->config().latent_dim;
    if (n_codebooks) *n_codebooks = e->config().n_codebooks;
    if (codebook_dim) *codebook_dim = e->config().codebook_dim;
  });
}

static void check_rate(DacEngine* e, int32_t sample_rate) {
  // Models/DAC.cs:143-149: ArgumentException on mismatch
  if (sample_rate != 0 && sample_rate != e->config().sample_rate)
    throw Error(NC_INVALID_ARGUMENT, "Input audio sample rate " + std::to_string(sample_rate) +
                                         "Hz does not match model sample rate " +
                                         std::to_string(e->config().sample_rate) + "Hz");
}

static int eff_nq(DacEngine* e, int nq) { return (nq <= 0 || nq > e->config().n_codebooks) ? e->config().n_codebooks : nq; }

nc_status nc_dac_encode(nc_handle h, const float* audio, int32_t batch, int64_t length, int32_t sample_rate,
                        int32_t n_quantizers, float* z, int64_t* codes, float* latents, int64_t* frames_out) {
  return guarded([&] {
    DacEngine* e = dac_of(h);
    if (!audio) throw Error(NC_INVALID_ARGUMENT, "audio is null");
    if (batch <= 0 || length <= 0) throw Error(NC_INVALID_ARGUMENT, "batch and length must be positive");
    if (n_quantizers < 0) throw Error(NC_INVALID_ARGUMENT, "n_quantizers must be >= 0");
    check_rate(e, sample_rate);
    BusyGuard g(e);
    e->bind();
    const int nq = eff_nq(e, n_quantizers);
    const int64_t T = e->frames(length);
    const auto& c = e->config();
    DevMem d_audio((size_t)batch * length * 4), d_z(z ? (size_t)batch * c.latent_dim * T * 4 : 0),
        d_codes(codes ? (size_t)batch * nq * T * 8 : 0), d_lat(latents ? (size_t)batch * nq * c.codebook_dim * T * 4 : 0);
    h2d(e, d_audio.p, audio, (size_t)batch * length * 4);
    e->encode_dev(d_audio.as<float>(), batch, length, nq, d_z.as<float>(), d_codes.as<int64_t>(), d_lat.as<float>());
    if (z) d2h(e, z, d_z.p, (size_t)batch * c.latent_dim * T * 4);
    if (codes) d2h(e, codes, d_codes.p, (size_t)batch * nq * T * 8);
    if (latents) d2h(e, latents, d_lat.p, (size_t)batch * nq * c.codebook_dim * T * 4);
    if (frames_out) *frames_out = T;
  });
}

nc_status nc_dac_decode(nc_handle h, const float* z, int32_t batch, int64_t frames, float* audio) {
  return guarded([&] {
    DacEngine* e = dac_of(h);
    if (!z || !audio) throw Error(NC_INVALID_ARGUMENT, "null buffer");
    if (batch <= 0 || frames <= 0) throw Error(NC_INVALID_ARGUMENT, "batch and frames must be positive");
    BusyGuard g(e);
    e->bind();
    const auto& c = e->config();
    const int64_t L = e->decoded_length(frames);
    DevMem d_z((size_t)batch * c.latent_dim * frames * 4), d_a((size_t)batch * L * 4);
    h2d(e, d_z.p, z, (size_t)batch * c.latent_dim * frames * 4);
    e->decode_dev(d_z.as<float>(), batch, frames, d_a.as<float>());
    d2h(e, audio, d_a.p, (size_t)batch * L * 4);
  });
}

nc_status nc_dac_from_codes(nc_handle h, const int64_t* codes, int32_t batch, int32_t n_quantizers, int64_t frames,
                            float* z) {
  return guarded([&] {
    DacEngine* e = dac_of(h);
    if (!codes || !z) throw Error(NC_INVALID_ARGUMENT, "null buffer");
    if (batch <= 0 || frames <= 0) throw Error(NC_INVALID_ARGUMENT, "batch and frames must be positive");
    BusyGuard g(e);
    e->bind();
    const auto& c = e->config();
    DevMem d_c((size_t)batch * n_quantizers * frames * 8), d_z((size_t)batch * c.latent_dim * frames * 4);
    h2d(e, d_c.p, codes, (size_t)batch * n_quantizers * frames * 8);
    e->from_codes_dev(d_c.as<int64_t>(), batch, n_quantizers, frames, d_z.as<float>());
    d2h(e, z, d_z.p, (size_t)batch * c.latent_dim * frames * 4);
  });
}

nc_status nc_dac_decode_codes(nc_handle h, const int64_t* codes, int32_t batch, int32_t n_quantizers, int64_t frames,
                              float* audio) {
  return guarded([&] {
    DacEngine* e = dac_of(h);
    if (!codes || !audio) throw Error(NC_INVALID_ARGUMENT, "null buffer");
    if (batch <= 0 || frames <= 0) throw Error(NC_INVALID_ARGUMENT, "batch and frames must be positive");
    BusyGuard g(e);
    e->bind();
    const int64_t L = e->decoded_length(frames);
    DevMem d_c((size_t)batch * n_quantizers * frames * 8), d_a((size_t)batch * L * 4);
    h2d(e, d_c.p, codes, (size_t)batch * n_quantizers * frames * 8);
    e->decode_codes_dev(d_c.as<int64_t>(), batch, n_quantizers, frames, d_a.as<float>());
    d2h(e, audio, d_a.p, (size_t)batch * L * 4);
  });
}

nc_status nc_dac_decode_dia(nc_handle h, const int64_t* generated, int32_t batch, int32_t steps, int32_t channels,
                            const int32_t* delay_pattern, const int64_t* lengths, float* audio, int64_t audio_stride) {
  return guarded([&] {
    DacEngine* e = dac_of(h);
    if (!generated || !delay_pattern || !lengths || !audio) throw Error(NC_INVALID_ARGUMENT, "null buffer");
    if (batch <= 0 || steps <= 0 || channels <= 0 || audio_stride <= 0) throw Error(NC_INVALID_ARGUMENT, "bad sizes");
    BusyGuard g(e);
    e->bind();
    const size_t n = (size_t)batch * steps * channels;
    DevMem d_g(n * 8), d_a((size_t)batch * audio_stride * 4);
    h2d(e, d_g.p, generated, n * 8);
    NC_CUDA(cudaMemsetAsync(d_a.p, 0, (size_t)batch * audio_stride * 4, e->stream()));
    e->decode_dia_dev(d_g.as<int64_t>(), batch, steps, channels, delay_pattern, lengths, d_a.as<float>(), audio_stride);
    d2h(e, audio, d_a.p, (size_t)batch * audio_stride * 4);
  });
}

nc_status nc_dac_forward(nc_handle h, const float* audio, int32_t batch, int64_t length, int32_t n_quantizers,
                         float* audio_out, int64_t* codes, float* z, int64_t* frames_out) {
  return guarded([&] {
    DacEngine* e = dac_of(h);
    if (!audio) throw Error(NC_INVALID_ARGUMENT, "audio is null");
    if (batch <= 0 || length <= 0) throw Error(NC_INVALID_ARGUMENT, "batch and length must be positive");
    if (n_quantizers < 0) throw Error(NC_INVALID_ARGUMENT, "n_quantizers must be >= 0");
    BusyGuard g(e);
    e->bind();
    const int nq = eff_nq(e, n_quantizers);
    const int64_t T = e->frames(length);
    const int64_t Lout = e->decoded_length(T);
    const auto& c = e->config();
    DevMem d_audio((size_t)batch * length * 4), d_out(audio_out ? (size_t)batch * Lout * 4 : 0),
        d_z(z ? (size_t)batch * c.latent_dim * T * 4 : 0), d_codes(codes ? (size_t)batch * nq * T * 8 : 0);
    // Chunks of one engine micro-batch: the H2D copy of chunk i+1 and the D2H copies of chunk i-1 run on the handle's copy
    // stream while chunk i computes (they overlap when the caller's buffers are pinned; pageable memory degrades to staged
    // synchronous copies, still correct).  forward_dev returns with the compute stream drained, so a chunk's outputs are
    // complete when their D2H is queued; a chunk's input is waited for explicitly.
    const int chunk = std::max(1, e->micro_batch(batch, e->padded_length(length)));
    cudaStream_t cs = e->copy_stream();
    cudaEvent_t ready = nullptr;
    NC_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    struct EventGuard { cudaEvent_t ev; ~EventGuard() { cudaEventDestroy(ev); } } eg{ready};
    auto in_bytes = [&](int n) { return (size_t)n * length * 4; };
    NC_CUDA(cudaMemcpyAsync(d_audio.p, audio, in_bytes(std::min(chunk, (int)batch)), cudaMemcpyHostToDevice, cs));
    for (int b0 = 0; b0 < batch; b0 += chunk) {
      const int nb = std::min(chunk, (int)batch - b0);
      NC_CUDA(cudaEventRecord(ready, cs));
      NC_CUDA(cudaEventSynchronize(ready));               // this chunk's input (and everything queued before it) has landed
      if (b0 + nb < batch) {                              // next chunk's input: flies during this chunk's kernels
        const int nn = std::min(chunk, (int)batch - b0 - nb);
        NC_CUDA(cudaMemcpyAsync(d_audio.as<float>() + (size_t)(b0 + nb) * length, audio + (size_t)(b0 + nb) * length, in_bytes(nn),
                                cudaMemcpyHostToDevice, cs));
      }
      e->forward_dev(d_audio.as<float>() + (size_t)b0 * length, nb, length, nq, audio_out ? d_out.as<float>() + (size_t)b0 * Lout : nullptr,
                     codes ? d_codes.as<int64_t>() + (size_t)b0 * nq * T : nullptr, z ? d_z.as<float>() + (size_t)b0 * c.latent_dim * T : nullptr);
      if (audio_out) NC_CUDA(cudaMemcpyAsync(audio_out + (size_t)b0 * Lout, d_out.as<float>() + (size_t)b0 * Lout, (size_t)nb * Lout * 4, cudaMemcpyDeviceToHost, cs));
      if (z) NC_CUDA(cudaMemcpyAsync(z + (size_t)b0 * c.latent_dim * T, d_z.as<float>() + (size_t)b0 * c.latent_dim * T, (size_t)nb * c.latent_dim * T * 4, cudaMemcpyDeviceToHost, cs));
      if (codes) NC_CUDA(cudaMemcpyAsync(codes + (size_t)b0 * nq * T, d_codes.as<int64_t>() + (size_t)b0 * nq * T, (size_t)nb * nq * T * 8, cudaMemcpyDeviceToHost, cs));
    }
    NC_CUDA(cudaStreamSynchronize(cs));
    if (frames_out) *frames_out = T;
  });
}

nc_status nc_dac_forward_dev(nc_handle h, const float* audio_dev, int32_t batch, int64_t length, int32_t n_quantizers,
                             float* audio_out_dev, int64_t* codes_dev, float* z_dev, int64_t* frames_out) {
  return guarded([&] {
    DacEngine* e = dac_of(h);
    if (!audio_dev) throw Error(NC_INVALID_ARGUMENT, "audio is null");
    if (batch <= 0 || length <= 0) throw Error(NC_INVALID_ARGUMENT, "batch and length must be positive");
    BusyGuard g(e);
    e->forward_dev(audio_dev, batch, length, eff_nq(e, n_quantizers), audio_out_dev, codes_dev, z_dev);
    if (frames_out) *frames_out = e->frames(length);
  });
}

nc_status nc_dac_decode_codes_dev(nc_handle h, const int64_t* codes_dev, int32_t batch, int32_t n_quantizers,
                                  int64_t frames, float* audio_dev) {
  return guarded([&] {
    DacEngine* e = dac_of(h);
    if (!codes_dev || !audio_dev) throw Error(NC_INVALID_ARGUMENT, "null buffer");
    BusyGuard g(e);
    e->decode_codes_dev(codes_dev, batch, n_quantizers, frames, audio_dev);
  });
}

// ------------------------------------------------------------------------------------ SNAC
static SnacEngine* snac_of(nc_handle h) {
  if (!h || !h->engine) throw Error(NC_INVALID_ARGUMENT, "null handle");
  if (h->kind != NC_CODEC_SNAC) throw Error(NC_INVALID_ARGUMENT, "handle is not a SNAC codec");
  return static_cast<SnacEngine*>(h->engine);
}

nc_status nc_snac_query_shapes(nc_handle h, int64_t length, int64_t* padded_length, int64_t* frames, int32_t* n_stages,
                               int64_t* code_lengths, int32_t* n_noise, int64_t* noise_lengths) {
  return guarded([&] {
    SnacEngine* e = snac_of(h);
    if (length < 0) throw Error(NC_INVALID_ARGUMENT, "length must be non-negative");
    const int64_t T = e->frames(length);
    if (padded_length) *padded_length = e->padded_length(length);
    if (frames) *frames = T;
    if (n_stages) *n_stages = e->n_stages();
    if (code_lengths)
      for (int i = 0; i < e->n_stages(); ++i) code_lengths[i] = T / e->config().vq_strides[i];
    const auto nl = e->noise_lengths(T);
    if (n_noise) *n_noise = e->config().noise ? (int32_t)nl.size() : 0;
    if (noise_lengths)
      for (size_t i = 0; i < nl.size(); ++i) noise_lengths[i] = nl[i];
  });
}

namespace {
// host arrays of host pointers -> device staging, with copy-back for outputs
struct SnacStaging {
  std::vector<DevMem*> mem;
  std::vector<int64_t*> codes_dev;
  std::vector<const float*> noise_dev;
  ~SnacStaging() {
    for (auto* m : mem) delete m;
  }
};
}  // namespace

static void snac_forward_host(SnacEngine* e, const float* audio, int32_t batch, int64_t length, const float* const* noise,
                              uint64_t seed, float* audio_out, int64_t* const* codes) {
  if (!audio) throw Error(NC_INVALID_ARGUMENT, "audio is null");
  if (batch <= 0 || length <= 0) throw Error(NC_INVALID_ARGUMENT, "batch and length must be positive");
  BusyGuard g(e);
  e->bind();
  const int64_t T = e->frames(length);
  const int ns = e->n_stages();
  SnacStaging st;
  DevMem d_audio((size_t)batch * length * 4), d_out(audio_out ? (size_t)batch * length * 4 : 0);
  h2d(e, d_audio.p, audio, (size_t)batch * length * 4);
  st.codes_dev.assign(ns, nullptr);
  if (codes)
    for (int i = 0; i < ns; ++i)
      if (codes[i]) {
        st.mem.push_back(new DevMem((size_t)batch * (T / e->config().vq_strides[i]) * 8));
        st.codes_dev[i] = st.mem.back()->as<int64_t>();
      }
  const auto nl = e->noise_lengths(T);
  st.noise_dev.assign(nl.size(), nullptr);
  if (noise && audio_out)
    for (size_t i = 0; i < nl.size(); ++i)
      if (noise[i]) {
        st.mem.push_back(new DevMem((size_t)batch * nl[i] * 4));
        h2d(e, st.mem.back()->p, noise[i], (size_t)batch * nl[i] * 4);
        st.noise_dev[i] = st.mem.back()->as<float>();
      }
  e->forward_dev(d_audio.as<float>(), batch, length, st.noise_dev.data(), seed, d_out.as<float>(),
                 codes ? st.codes_dev.data() : nullptr);
  if (audio_out) d2h(e, audio_out, d_out.p, (size_t)batch * length * 4);
  if (codes)
    for (int i = 0; i < ns; ++i)
      if (codes[i])
        d2h(e, codes[i], st.codes_dev[i], (size_t)batch * (T / e->config().vq_strides[i]) * 8);
}

nc_status nc_snac_encode(nc_handle h, const float* audio, int32_t batch, int64_t length, int64_t* const* codes) {
  return guarded([&] {
    SnacEngine* e = snac_of(h);
    if (!codes) throw Error(NC_INVALID_ARGUMENT, "codes is null");
    snac_forward_host(e, audio, batch, length, nullptr, 0, nullptr, codes);
  });
}

nc_status nc_snac_forward(nc_handle h, const float* audio, int32_t batch, int64_t length, const float* const* noise,
                          uint64_t seed, float* audio_out, int64_t* const* codes) {
  return guarded([&] { snac_forward_host(snac_of(h), audio, batch, length, noise, seed, audio_out, codes); });
}

nc_status nc_snac_decode(nc_handle h, const int64_t* const* codes, int32_t batch, int64_t frames, const float* const* noise,
                         uint64_t seed, float* audio) {
  return guarded([&] {
    SnacEngine* e = snac_of(h);
    if (!codes || !audio) throw Error(NC_INVALID_ARGUMENT, "Codes list cannot be empty or contain null arrays");   // SNAC.cs:177-180
    if (batch <= 0 || frames <= 0) throw Error(NC_INVALID_ARGUMENT, "batch and frames must be positive");
    BusyGuard g(e);
    e->bind();
    const int ns = e->n_stages();
    SnacStaging st;
    std::vector<const int64_t*> cdev(ns, nullptr);
    for (int i = 0; i < ns; ++i) {
      if (!codes[i]) throw Error(NC_INVALID_ARGUMENT, "Codes list cannot be empty or contain null arrays");
      const size_t bytes = (size_t)batch * (frames / e->config().vq_strides[i]) * 8;
      st.mem.push_back(new DevMem(bytes));
      h2d(e, st.mem.back()->p, codes[i], bytes);
      cdev[i] = st.mem.back()->as<int64_t>();
    }
    const auto nl = e->noise_lengths(frames);
    st.noise_dev.assign(nl.size(), nullptr);
    if (noise)
      for (size_t i = 0; i < nl.size(); ++i)
        if (noise[i]) {
          st.mem.push_back(new DevMem((size_t)batch * nl[i] * 4));
          h2d(e, st.mem.back()->p, noise[i], (size_t)batch * nl[i] * 4);
          st.noise_dev[i] = st.mem.back()->as<float>();
        }
    const int64_t L = e->decoded_length(frames);
    DevMem d_a((size_t)batch * L * 4);
    e->decode_dev(cdev.data(), batch, frames, st.noise_dev.data(), seed, d_a.as<float>());
    d2h(e, audio, d_a.p, (size_t)batch * L * 4);
  });
}

nc_status nc_snac_forward_dev(nc_handle h, const float* audio_dev, int32_t batch, int64_t length, const float* const* noise_dev,
                              uint64_t seed, float* audio_out_dev, int64_t* const* codes_dev) {
  return guarded([&] {
    SnacEngine* e = snac_of(h);
    if (!audio_dev) throw Error(NC_INVALID_ARGUMENT, "audio is null");
    BusyGuard g(e);
    e->forward_dev(audio_dev, batch, length, noise_dev, seed, audio_out_dev, codes_dev);
  });
}

// ------------------------------------------------------------------------------------ Encodec
static EncodecEngine* encodec_of(nc_handle h) {
  if (!h || !h->engine) throw Error(NC_INVALID_ARGUMENT, "null handle");
  if (h->kind != NC_CODEC_ENCODEC) throw Error(NC_INVALID_ARGUMENT, "handle is not an Encodec codec");
  return static_cast<EncodecEngine*>(h->engine);
}

nc_status nc_encodec_query_shapes(nc_handle h, int64_t length, float bandwidth_kbps, int64_t* frames, int32_t* n_q,
                                  int64_t* decoded_length) {
  return guarded([&] {
    EncodecEngine* e = encodec_of(h);
    if (length < 0) throw Error(NC_INVALID_ARGUMENT, "length must be non-negative");
    if (n_q) *n_q = e->n_q_for_bandwidth(bandwidth_kbps);
    if (e->simple()) {
      const int64_t T = e->frames(length);
      if (frames) *frames = T;
      if (decoded_length) *decoded_length = e->decoded_length(T);
    } else {   // segmented models: all segments' frames and the overlap-added length
      if (length == 0) throw Error(NC_INVALID_ARGUMENT, "length must be positive");
      const EncodecEngine::SegLayout lay = e->seg_layout(length);
      if (frames) *frames = lay.t_total;
      if (decoded_length) *decoded_length = lay.total_out;
    }
  });
}

nc_status nc_encodec_query_frames(nc_handle h, int64_t length, float bandwidth_kbps, int32_t* n_segments, int64_t* seg_frames,
                                  int32_t seg_frames_capacity, int64_t* total_frames, int32_t* n_q, int64_t* decoded_length) {
  return guarded([&] {
    EncodecEngine* e = encodec_of(h);
    if (length <= 0) throw Error(NC_INVALID_ARGUMENT, "length must be positive");
    const EncodecEngine::SegLayout lay = e->seg_layout(length);
    if (n_segments) *n_segments = lay.n_seg;
    if (seg_frames)
      for (int s = 0; s < lay.n_seg && s < seg_frames_capacity; ++s) seg_frames[s] = lay.frames[s];
    if (total_frames) *total_frames = lay.t_total;
    if (n_q) *n_q = e->n_q_for_bandwidth(bandwidth_kbps);
    if (decoded_length) *decoded_length = lay.total_out;
  });
}

// audio [B][C][length] host -> codes [B][nq][t_total], scales [B][n_seg], audio_out [B][C][length] (all nullable)
static void encodec_forward_host(EncodecEngine* e, const float* audio, int32_t batch, int64_t length, float bw, float* audio_out,
                                 int64_t* codes, float* scales) {
  if (!audio) throw Error(NC_INVALID_ARGUMENT, "audio is null");
  if (batch <= 0 || length <= 0) throw Error(NC_INVALID_ARGUMENT, "batch and length must be positive");
  BusyGuard g(e);
  e->bind();
  const int nq = e->n_q_for_bandwidth(bw);
  const int C = e->config().channels;
  const EncodecEngine::SegLayout lay = e->seg_layout(length);
  const size_t a_bytes = (size_t)batch * C * length * 4, c_bytes = (size_t)batch * nq * lay.t_total * 8;
  const bool want_scales = scales && e->config().normalize;
  DevMem d_audio(a_bytes), d_out(audio_out ? a_bytes : 0), d_codes(codes ? c_bytes : 0),
      d_scales(want_scales ? (size_t)batch * lay.n_seg * 4 : 0);
  h2d(e, d_audio.p, audio, a_bytes);
  if (e->simple())
    e->forward_dev(d_audio.as<float>(), batch, length, nq, d_out.as<float>(), d_codes.as<int64_t>());
  else
    e->forward_frames_dev(d_audio.as<float>(), batch, length, nq, d_out.as<float>(), d_codes.as<int64_t>(), d_scales.as<float>());
  if (audio_out) d2h(e, audio_out, d_out.p, a_bytes);
  if (codes) d2h(e, codes, d_codes.p, c_bytes);
  if (want_scales) d2h(e, scales, d_scales.p, (size_t)batch * lay.n_seg * 4);
}

nc_status nc_encodec_encode(nc_handle h, const float* audio, int32_t batch, int64_t length, float bandwidth_kbps, int64_t* codes) {
  return guarded([&] {
    if (!codes) throw Error(NC_INVALID_ARGUMENT, "codes is null");
    EncodecEngine* e = encodec_of(h);
    if (e->config().normalize)
      throw Error(NC_INVALID_ARGUMENT, "this Encodec model normalises every frame: use nc_encodec_encode_frames, which also returns the scales");
    encodec_forward_host(e, audio, batch, length, bandwidth_kbps, nullptr, codes, nullptr);
  });
}

nc_status nc_encodec_encode_frames(nc_handle h, const float* audio, int32_t batch, int64_t length, float bandwidth_kbps, int64_t* codes,
                                   float* scales) {
  return guarded([&] {
    if (!codes) throw Error(NC_INVALID_ARGUMENT, "codes is null");
    encodec_forward_host(encodec_of(h), audio, batch, length, bandwidth_kbps, nullptr, codes, scales);
  });
}

nc_status nc_encodec_forward(nc_handle h, const float* audio, int32_t batch, int64_t length, float bandwidth_kbps,
                             float* audio_out, int64_t* codes) {
  return guarded([&] {
    if (!audio_out) throw Error(NC_INVALID_ARGUMENT, "audio_out is null");
    encodec_forward_host(encodec_of(h), audio, batch, length, bandwidth_kbps, audio_out, codes, nullptr);
  });
}

nc_status nc_encodec_decode(nc_handle h, const int64_t* codes, int32_t batch, int32_t n_q, int64_t frames, float* audio) {
  return nc_encodec_decode_frames(h, codes, nullptr, batch, n_q, &frames, 1, audio);
}

nc_status nc_encodec_decode_frames(nc_handle h, const int64_t* codes, const float* scales, int32_t batch, int32_t n_q,
                                   const int64_t* seg_frames, int32_t n_segments, float* audio) {
  return guarded([&] {
    EncodecEngine* e = encodec_of(h);
    if (!codes || !audio) throw Error(NC_INVALID_ARGUMENT, "Invalid frame codes in Encodec Decode");   // Encodec.cs:438-442
    if (batch <= 0 || n_q <= 0 || n_segments <= 0 || !seg_frames) throw Error(NC_INVALID_ARGUMENT, "No frames provided to decode");
    BusyGuard g(e);
    e->bind();
    const EncodecEngine::SegLayout lay = e->seg_layout_from_frames(seg_frames, n_segments);
    const int C = e->config().channels;
    const size_t c_bytes = (size_t)batch * n_q * lay.t_total * 8, a_bytes = (size_t)batch * C * lay.total_out * 4;
    DevMem d_c(c_bytes), d_a(a_bytes), d_s(scales ? (size_t)batch * n_segments * 4 : 0);
    h2d(e, d_c.p, codes, c_bytes);
    if (scales) h2d(e, d_s.p, scales, (size_t)batch * n_segments * 4);
    e->decode_frames_dev(d_c.as<int64_t>(), scales ? d_s.as<float>() : nullptr, batch, n_q, seg_frames, n_segments, d_a.as<float>());
    d2h(e, audio, d_a.p, a_bytes);
  });
}

nc_status nc_encodec_query_decoded(nc_handle h, const int64_t* seg_frames, int32_t n_segments, int64_t* decoded_length) {
  return guarded([&] {
    EncodecEngine* e = encodec_of(h);
    const EncodecEngine::SegLayout lay = e->seg_layout_from_frames(seg_frames, n_segments);
    if (decoded_length) *decoded_length = lay.total_out;
  });
}

nc_status nc_encodec_forward_dev(nc_handle h, const float* audio_dev, int32_t batch, int64_t length, float bandwidth_kbps,
                                 float* audio_out_dev, int64_t* codes_dev) {
  return nc_encodec_forward_frames_dev(h, audio_dev, batch, length, bandwidth_kbps, audio_out_dev, codes_dev, nullptr);
}

nc_status nc_encodec_forward_frames_dev(nc_handle h, const float* audio_dev, int32_t batch, int64_t length, float bandwidth_kbps,
                                        float* audio_out_dev, int64_t* codes_dev, float* scales_dev) {
  return guarded([&] {
    EncodecEngine* e = encodec_of(h);
    if (!audio_dev) throw Error(NC_INVALID_ARGUMENT, "audio is null");
    BusyGuard g(e);
    const int nq = e->n_q_for_bandwidth(bandwidth_kbps);
    if (e->simple()) e->forward_dev(audio_dev, batch, length, nq, audio_out_dev, codes_dev);
    else e->forward_frames_dev(audio_dev, batch, length, nq, audio_out_dev, codes_dev, scales_dev);
  });
}

// ------------------------------------------------------------------------------------ weight files (host only)
nc_status nc_inspect_weights(const char* path, char* buf, size_t buf_size) {
  return guarded([&] {
    if (!path) throw Error(NC_INVALID_ARGUMENT, "path is null");
    if (!buf || buf_size == 0) throw Error(NC_INVALID_ARGUMENT, "null buffer");
    FILE* fp = std::fopen(path, "rb");
    if (!fp) throw Error(NC_FILE_NOT_FOUND, std::string("weights not found at ") + path);
    std::fclose(fp);
    TensorMap tm;
    std::string meta = "{}";
    const bool zip = is_torch_zip(path);
    if (zip) load_torch_zip(path, &tm, &meta); else load_safetensors(path, &tm);
    std::string s = std::string("{\"format\": \"") + (zip ? "torch_zip" : "safetensors") + "\", \"metadata\": " + meta + ", \"tensors\": {";
    bool first = true;
    for (auto& kv : tm) {
      s += first ? "\"" : ", \"";
      first = false;
      s += kv.first + "\": {\"dtype\": \"" + (kv.second.is_int ? "int64" : "float32") + "\", \"shape\": [";
      for (size_t i = 0; i < kv.second.shape.size(); ++i) s += (i ? ", " : "") + std::to_string(kv.second.shape[i]);
      // position-weighted checksum of the converted values: sum_i (i % 7 + 1) * x_i in double
      double cs = 0;
      if (kv.second.is_int) for (size_t i = 0; i < kv.second.i64.size(); ++i) cs += (double)(i % 7 + 1) * (double)kv.second.i64[i];
      else for (size_t i = 0; i < kv.second.f32.size(); ++i) cs += (double)(i % 7 + 1) * (double)kv.second.f32[i];
      char cbuf[40];
      std::snprintf(cbuf, sizeof cbuf, "%.17g", cs);
      s += std::string("], \"checksum\": ") + cbuf + "}";
    }
    s += "}}";
    if (s.size() + 1 > buf_size) throw Error(NC_INVALID_ARGUMENT, "inspect buffer too small");
    std::memcpy(buf, s.c_str(), s.size() + 1);
  });
}

// ------------------------------------------------------------------------------------ input conditioning
static int64_t resampled_length(int64_t length, int32_t src, int32_t dst, double* ratio) {
  if (src <= 0 || dst <= 0) throw Error(NC_INVALID_ARGUMENT, "sample rates must be positive");
  *ratio = (double)dst / (double)src;            // SNAC.cs:286
  return (int64_t)((double)length * *ratio);     // :287
}

nc_status nc_resample_linear(nc_handle h, const float* audio, int32_t batch, int64_t length, int32_t src_rate, int32_t dst_rate,
                             float* out, int64_t out_capacity, int64_t* out_length) {
  return guarded([&] {
    if (!h || !h->engine) throw Error(NC_INVALID_ARGUMENT, "null handle");
    Engine* e = h->engine;
    if (batch <= 0 || length <= 0) throw Error(NC_INVALID_ARGUMENT, "Audio data cannot be empty");
    double ratio;
    const int64_t n_out = resampled_length(length, src_rate, dst_rate, &ratio);
    if (out_length) *out_length = n_out;
    if (!out) return;
    if (!audio) throw Error(NC_INVALID_ARGUMENT, "audio is null");
    if (out_capacity < n_out) throw Error(NC_INVALID_ARGUMENT, "out_capacity is smaller than the resampled length");
    if (n_out == 0) return;
    BusyGuard g(e);
    e->bind();
    DevMem d_in((size_t)batch * length * 4), d_out((size_t)batch * n_out * 4);
    h2d(e, d_in.p, audio, (size_t)batch * length * 4);
    launch_resample_linear(d_in.as<float>(), length, length, d_out.as<float>(), n_out, n_out, ratio, batch, e->ctx());
    e->sync();
    d2h_2d(e, out, (size_t)out_capacity * 4, d_out.p, (size_t)n_out * 4, (size_t)n_out * 4, (size_t)batch);
  });
}

nc_status nc_convert_to_mono(nc_handle h, const float* interleaved, int64_t frames, int32_t channels, float* out) {
  return guarded([&] {
    if (!h || !h->engine) throw Error(NC_INVALID_ARGUMENT, "null handle");
    Engine* e = h->engine;
    if (!interleaved || !out) throw Error(NC_INVALID_ARGUMENT, "audio is null");
    if (frames <= 0 || channels <= 0) throw Error(NC_INVALID_ARGUMENT, "frames and channels must be positive");
    BusyGuard g(e);
    e->bind();
    DevMem d_in((size_t)frames * channels * 4), d_out((size_t)frames * 4);
    h2d(e, d_in.p, interleaved, (size_t)frames * channels * 4);
    launch_to_mono(d_in.as<float>(), d_out.as<float>(), frames, channels, e->ctx());
    e->sync();
    d2h(e, out, d_out.p, (size_t)frames * 4);
  });
}

nc_status nc_snac_process_audio(nc_handle h, const float* audio, int32_t batch, int64_t length, int32_t sample_rate,
                                const float* const* noise, uint64_t seed, float* audio_out, int64_t out_capacity,
                                int64_t* out_length) {
  return guarded([&] {
    SnacEngine* e = snac_of(h);
    if (batch <= 0 || length <= 0) throw Error(NC_INVALID_ARGUMENT, "Audio data cannot be empty");   // SNAC.cs:257-258
    double ratio = 1.0;
    const int model_rate = e->config().sample_rate;
    const int64_t n = sample_rate == model_rate ? length : resampled_length(length, sample_rate, model_rate, &ratio);
    if (out_length) *out_length = n;
    if (!audio_out) return;
    if (!audio) throw Error(NC_INVALID_ARGUMENT, "Audio data cannot be empty");
    if (n <= 0) throw Error(NC_INVALID_ARGUMENT, "Audio data cannot be empty");
    if (out_capacity < n) throw Error(NC_INVALID_ARGUMENT, "out_capacity is smaller than the output length");
    BusyGuard g(e);
    e->bind();
    DevMem d_in((size_t)batch * length * 4), d_rs(sample_rate == model_rate ? 0 : (size_t)batch * n * 4), d_out((size_t)batch * n * 4);
    h2d(e, d_in.p, audio, (size_t)batch * length * 4);
    const float* x = d_in.as<float>();
    if (sample_rate != model_rate) {
      launch_resample_linear(d_in.as<float>(), length, length, d_rs.as<float>(), n, n, ratio, batch, e->ctx());
      x = d_rs.as<float>();
    }
    const int64_t T = e->frames(n);
    const auto nl = e->noise_lengths(T);
    SnacStaging st;
    st.noise_dev.assign(nl.size(), nullptr);
    if (noise)
      for (size_t i = 0; i < nl.size(); ++i)
        if (noise[i]) {
          st.mem.push_back(new DevMem((size_t)batch * nl[i] * 4));
          h2d(e, st.mem.back()->p, noise[i], (size_t)batch * nl[i] * 4);
          st.noise_dev[i] = st.mem.back()->as<float>();
        }
    e->forward_dev(x, batch, n, st.noise_dev.data(), seed, d_out.as<float>(), nullptr);
    d2h_2d(e, audio_out, (size_t)out_capacity * 4, d_out.p, (size_t)n * 4, (size_t)n * 4, (size_t)batch);
  });
}

// ------------------------------------------------------------------------------------ .ecdc container
static EcdcMeta ecdc_meta_for(EncodecEngine* e, int64_t length, float bw) {
  EcdcMeta m;
  const EncodecConfig& c = e->config();
  m.model = c.sample_rate == 48000 ? "encodec_48khz" : "encodec_24khz";   // EncodecCompressor.cs:14-18 factory keys
  m.audio_length = length;
  m.n_codebooks = e->n_q_for_bandwidth(bw);
  m.use_lm = false;
  m.channels = c.channels;
  m.sample_rate = c.sample_rate;
  m.bandwidth = bw;
  m.has_bandwidth = true;
  return m;
}

// Payload of a segmented / normalised model (EncodecCompressor.cs:116-190): per segment [int32 BE count = 1][float32 BE
// scale] when the model normalises, then that segment's codes packed on their own.  off[s] = first byte of segment s.
struct EcdcSegs {
  std::vector<int64_t> frames, off;
  int64_t total = 0;
  int scale_bytes = 0;
};
static EcdcSegs ecdc_segs(EncodecEngine* e, const std::vector<int64_t>& frames, int nq) {
  EcdcSegs g;
  g.frames = frames;
  g.scale_bytes = e->config().normalize ? 8 : 0;
  for (int64_t T : frames) {
    g.off.push_back(g.total);
    g.total += g.scale_bytes + e->ecdc_payload_bytes(nq, T);
  }
  return g;
}
// the reader's frame counts (EncodecCompressor.cs:303-309): ceil(segment samples * FrameRate / SampleRate) -- NOT the
// encoder's count when a very short last segment was lengthened by Pad1d's short-input branch; mirrored as is.
static std::vector<int64_t> ecdc_reader_frames(EncodecEngine* e, int64_t al) {
  const EncodecConfig& c = e->config();
  const int64_t seg = e->segmented() ? c.segment_length() : al, stride = e->segmented() ? c.segment_stride() : al;
  const int frame_rate = (int)std::ceil((float)c.sample_rate / (float)c.hop());
  std::vector<int64_t> f;
  for (int64_t off = 0; off < al; off += stride) {
    const int64_t len = std::min(al - off, seg);
    f.push_back((int64_t)std::ceil((double)(len * frame_rate) / (double)c.sample_rate));
  }
  return f;
}
static void put_be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
static uint32_t get_be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }

nc_status nc_encodec_ecdc_size(nc_handle h, int64_t length, float bandwidth_kbps, int64_t* header_bytes, int64_t* stream_bytes) {
  return guarded([&] {
    EncodecEngine* e = encodec_of(h);
    if (length <= 0) throw Error(NC_INVALID_ARGUMENT, "length must be positive");
    const EcdcMeta m = ecdc_meta_for(e, length, bandwidth_kbps);
    const int64_t hb = (int64_t)ecdc_header(m).size();
    if (header_bytes) *header_bytes = hb;
    if (!stream_bytes) return;
    if (e->simple()) *stream_bytes = hb + e->ecdc_payload_bytes(m.n_codebooks, e->frames(length));
    else *stream_bytes = hb + ecdc_segs(e, e->seg_layout(length).frames, m.n_codebooks).total;
  });
}

nc_status nc_encodec_compress(nc_handle h, const float* audio, int32_t batch, int64_t length, float bandwidth_kbps, uint8_t* out,
                              int64_t out_stride, int64_t* stream_bytes) {
  return guarded([&] {
    EncodecEngine* e = encodec_of(h);
    if (!audio) throw Error(NC_INVALID_ARGUMENT, "audio is null");
    if (!out) throw Error(NC_INVALID_ARGUMENT, "out is null");
    if (batch <= 0 || length <= 0) throw Error(NC_INVALID_ARGUMENT, "batch and length must be positive");
    const EcdcMeta m = ecdc_meta_for(e, length, bandwidth_kbps);
    const std::string header = ecdc_header(m);
    if (!e->simple()) {   // one frame per segment, each with its scale block (EncodecCompressor.cs:116-190)
      const int nq = m.n_codebooks, C = e->config().channels;
      const EncodecEngine::SegLayout lay = e->seg_layout(length);
      const EcdcSegs sg = ecdc_segs(e, lay.frames, nq);
      const int64_t total = (int64_t)header.size() + sg.total;
      if (out_stride < total) throw Error(NC_INVALID_ARGUMENT, "out_stride is smaller than the stream (see nc_encodec_ecdc_size)");
      BusyGuard g(e);
      e->bind();
      const size_t a_bytes = (size_t)batch * C * length * 4;
      DevMem d_audio(a_bytes), d_codes((size_t)batch * nq * lay.t_total * 8), d_pay((size_t)batch * sg.total),
          d_scales(sg.scale_bytes ? (size_t)batch * lay.n_seg * 4 : 0);
      h2d(e, d_audio.p, audio, a_bytes);
      e->forward_frames_dev(d_audio.as<float>(), batch, length, nq, nullptr, d_codes.as<int64_t>(), d_scales.as<float>());
      for (int s = 0; s < lay.n_seg; ++s)
        e->ecdc_pack_segment_dev(d_codes.as<int64_t>(), lay.t_total, lay.col[s], lay.frames[s], nq,
                                 d_pay.as<uint8_t>() + sg.off[s] + sg.scale_bytes, sg.total, batch);
      for (int b = 0; b < batch; ++b) std::memcpy(out + (size_t)b * out_stride, header.data(), header.size());
      d2h_2d(e, out + header.size(), (size_t)out_stride, d_pay.p, (size_t)sg.total, (size_t)sg.total, (size_t)batch);
      if (sg.scale_bytes) {
        std::vector<float> sc((size_t)batch * lay.n_seg);
        d2h(e, sc.data(), d_scales.p, sc.size() * 4);
        for (int b = 0; b < batch; ++b)
          for (int s = 0; s < lay.n_seg; ++s) {
            uint8_t* p = out + (size_t)b * out_stride + header.size() + sg.off[s];
            uint32_t bits;
            std::memcpy(&bits, &sc[(size_t)b * lay.n_seg + s], 4);
            put_be32(p, 1u);          // one waveform per stream: scale.numel() == 1 (:137-139)
            put_be32(p + 4, bits);
          }
      }
      if (stream_bytes) *stream_bytes = total;
      return;
    }
    const int64_t payload = e->ecdc_payload_bytes(m.n_codebooks, e->frames(length));
    const int64_t total = (int64_t)header.size() + payload;
    if (out_stride < total) throw Error(NC_INVALID_ARGUMENT, "out_stride is smaller than the stream (see nc_encodec_ecdc_size)");
    BusyGuard g(e);
    e->bind();
    DevMem d_audio((size_t)batch * length * 4), d_pay((size_t)batch * payload);
    h2d(e, d_audio.p, audio, (size_t)batch * length * 4);
    e->compress_dev(d_audio.as<float>(), batch, length, m.n_codebooks, d_pay.as<uint8_t>(), payload);
    for (int b = 0; b < batch; ++b) std::memcpy(out + (size_t)b * out_stride, header.data(), header.size());
    d2h_2d(e, out + header.size(), (size_t)out_stride, d_pay.p, (size_t)payload, (size_t)payload, (size_t)batch);
    if (stream_bytes) *stream_bytes = total;
  });
}

nc_status nc_encodec_ecdc_info(const uint8_t* stream, int64_t stream_bytes, int64_t* audio_length, int32_t* n_q, int32_t* channels,
                               int32_t* sample_rate, float* bandwidth_kbps, int32_t* use_lm, int64_t* payload_offset) {
  return guarded([&] {
    if (!stream || stream_bytes <= 0) throw Error(NC_INVALID_ARGUMENT, "stream is null");
    EcdcMeta m;
    const size_t off = ecdc_read_header(stream, (size_t)stream_bytes, &m);
    if (m.sample_rate == 0) m.sample_rate = m.model.find("48khz") != std::string::npos ? 48000 : 24000;   // EncodecCompressor.cs:264-267
    if (audio_length) *audio_length = m.audio_length;
    if (n_q) *n_q = m.n_codebooks;
    if (channels) *channels = m.channels;
    if (sample_rate) *sample_rate = m.sample_rate;
    if (bandwidth_kbps) *bandwidth_kbps = m.has_bandwidth ? m.bandwidth : 0.f;
    if (use_lm) *use_lm = m.use_lm ? 1 : 0;
    if (payload_offset) *payload_offset = (int64_t)off;
  });
}

nc_status nc_encodec_decompress(nc_handle h, const uint8_t* streams, int32_t batch, int64_t stream_stride, int64_t stream_bytes,
                                float* audio, int64_t audio_capacity, int64_t* audio_length, int32_t* sample_rate) {
  return guarded([&] {
    EncodecEngine* e = encodec_of(h);
    if (!streams) throw Error(NC_INVALID_ARGUMENT, "stream is null");
    if (batch <= 0 || stream_bytes <= 0 || stream_stride < stream_bytes)
      throw Error(NC_INVALID_ARGUMENT, "batch, stream_bytes and stream_stride must describe a valid buffer");
    EcdcMeta m0;
    const size_t off = ecdc_read_header(streams, (size_t)stream_bytes, &m0);
    for (int b = 1; b < batch; ++b) {   // one launch decodes clips of one shape: all headers must agree
      EcdcMeta mb;
      const size_t ob = ecdc_read_header(streams + (size_t)b * stream_stride, (size_t)stream_bytes, &mb);
      if (ob != off || mb.audio_length != m0.audio_length || mb.n_codebooks != m0.n_codebooks || mb.use_lm != m0.use_lm ||
          mb.channels != m0.channels)
        throw Error(NC_INVALID_ARGUMENT, "batched decompress needs streams with identical metadata");
    }
    if (m0.use_lm) throw Error(NC_UNSUPPORTED, "ecdc streams written with the language-model entropy coder are not supported");
    const EncodecConfig& c = e->config();
    if (m0.channels != c.channels)                                                                          // EncodecCompressor.cs:282-286
      throw Error(NC_INVALID_ARGUMENT, "Model has " + std::to_string(c.channels) + " channels but compressed data has " +
                                           std::to_string(m0.channels) + " channels");
    if (m0.sample_rate != 0 && m0.sample_rate != c.sample_rate)
      throw Error(NC_INVALID_ARGUMENT, "Model " + m0.model + " not supported");
    if (audio_length) *audio_length = m0.audio_length;
    if (sample_rate) *sample_rate = c.sample_rate;
    if (!audio) return;   // size query
    if (audio_capacity < m0.audio_length) throw Error(NC_INVALID_ARGUMENT, "audio_capacity is smaller than the stored audio length");
    const int64_t payload = stream_bytes - (int64_t)off;
    if (!e->simple()) {   // frames with scale blocks, decoded and overlap-added (EncodecCompressor.cs:303-415)
      const int nq = m0.n_codebooks, C = c.channels;
      const int64_t al = m0.audio_length;
      if (al <= 0) throw Error(NC_INVALID_ARGUMENT, "Invalid ecdc payload");
      const std::vector<int64_t> fr = ecdc_reader_frames(e, al);
      const EcdcSegs sg = ecdc_segs(e, fr, nq);
      if (payload < sg.total) throw Error(NC_INVALID_ARGUMENT, "Stream ended too soon");   // :390-393
      const int n_seg = (int)fr.size();
      std::vector<float> sc;
      if (sg.scale_bytes) {
        sc.resize((size_t)batch * n_seg);
        for (int b = 0; b < batch; ++b)
          for (int s = 0; s < n_seg; ++s) {
            const uint8_t* p = streams + (size_t)b * stream_stride + off + sg.off[s];
            const int32_t count = (int32_t)get_be32(p);
            if (count <= 0 || count > 1000) throw Error(NC_INVALID_ARGUMENT, "Invalid scale count: " + std::to_string(count));   // :318-321
            if (count != 1) throw Error(NC_UNSUPPORTED, "ecdc frames with more than one scale value (a batch in one stream) are not supported");
            const uint32_t bits = get_be32(p + 4);
            std::memcpy(&sc[(size_t)b * n_seg + s], &bits, 4);
          }
      }
      BusyGuard g(e);
      e->bind();
      const EncodecEngine::SegLayout lay = e->seg_layout_from_frames(fr.data(), n_seg);
      DevMem d_pay((size_t)batch * sg.total), d_codes((size_t)batch * nq * lay.t_total * 8),
          d_audio((size_t)batch * C * lay.total_out * 4), d_sc(sc.empty() ? 0 : sc.size() * 4);
      h2d_2d(e, d_pay.p, (size_t)sg.total, streams + off, (size_t)stream_stride, (size_t)sg.total, (size_t)batch);
      if (!sc.empty()) h2d(e, d_sc.p, sc.data(), sc.size() * 4);
      for (int s = 0; s < n_seg; ++s)
        e->ecdc_unpack_segment_dev(d_pay.as<uint8_t>() + sg.off[s] + sg.scale_bytes, sg.total, d_codes.as<int64_t>(), lay.t_total,
                                   lay.col[s], fr[s], nq, batch);
      e->decode_frames_dev(d_codes.as<int64_t>(), sc.empty() ? nullptr : d_sc.as<float>(), batch, nq, fr.data(), n_seg, d_audio.as<float>());
      // rows = (clip, channel); only al samples of each are returned (:408-412)
      const int64_t keep = std::min<int64_t>(al, lay.total_out);
      d2h_2d(e, audio, (size_t)audio_capacity * 4, d_audio.p, (size_t)lay.total_out * 4, (size_t)keep * 4, (size_t)batch * C);
      if (audio_length) *audio_length = keep;
      return;
    }
    BusyGuard g(e);
    e->bind();
    DevMem d_pay((size_t)batch * std::max<int64_t>(payload, 1)), d_audio((size_t)batch * m0.audio_length * 4);
    if (payload > 0)
      h2d_2d(e, d_pay.p, (size_t)payload, streams + off, (size_t)stream_stride, (size_t)payload, (size_t)batch);
    e->decompress_dev(d_pay.as<uint8_t>(), payload, batch, m0.n_codebooks, m0.audio_length, d_audio.as<float>());
    d2h_2d(e, audio, (size_t)audio_capacity * 4, d_audio.p, (size_t)m0.audio_length * 4, (size_t)m0.audio_length * 4,
           (size_t)batch);
  });
}

nc_status nc_get_stream(nc_handle h, void** stream_out) {
  return guarded([&] {
    if (!h || !h->engine) throw Error(NC_INVALID_ARGUMENT, "null handle");
    if (!stream_out) throw Error(NC_INVALID_ARGUMENT, "stream_out is null");
    *stream_out = (void*)h->engine->stream();
  });
}

nc_status nc_describe(nc_handle h, char* buf, size_t buf_size) {
  return guarded([&] {
    if (!h || !h->engine) throw Error(NC_INVALID_ARGUMENT, "null handle");
    if (!buf || buf_size == 0) throw Error(NC_INVALID_ARGUMENT, "null buffer");
    const std::string s = h->engine->describe();
    if (s.size() + 1 > buf_size) throw Error(NC_INVALID_ARGUMENT, "describe buffer too small");
    std::memcpy(buf, s.c_str(), s.size() + 1);
  });
}

// ------------------------------------------------------------------------------------ instrumentation
uint64_t nc_launch_count(nc_handle h) { return (h && h->engine) ? h->engine->launches() : 0; }

nc_status nc_profile_report(nc_handle h, char* buf, size_t buf_size) {
  return guarded([&] {
    if (!h || !h->engine) throw Error(NC_INVALID_ARGUMENT, "null handle");
    if (!buf || buf_size == 0) throw Error(NC_INVALID_ARGUMENT, "null buffer");
    h->engine->bind();
    const std::string s = h->engine->profiler().report_json(true);
    if (s.size() + 1 > buf_size) throw Error(NC_INVALID_ARGUMENT, "profile buffer too small");
    std::memcpy(buf, s.c_str(), s.size() + 1);
  });
}

}  // extern "C"
