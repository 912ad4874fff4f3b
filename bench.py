#!/usr/bin/env python
"""bench.py -- audio-seconds per second of the DAC 44.1 kHz encode+decode hot path.

Workload (BASELINE.json configs[3], the configuration the metric is quoted on):
DAC 44.1 kHz, 9 codebooks, 512 clips x 30 s of synthetic mono audio, sharded by batch across
the ranks (one process per GPU, no collective on the data path; strong scaling: the global
batch is fixed).  One step = one encode -> RVQ -> decode pass over the rank's shard through
the C ABI of libneuralcodecs_cuda.so.

  value : whole-job audio-s/s, inputs and outputs resident in HBM (nc_dac_forward_dev)
  e2e   : the same pass through the host-buffer entry point (nc_dac_forward) with pinned
          host audio in and codes + audio out, copies inside the timed region
  roofline     : the tcgen05 implicit-GEMM conv kernel (tensor bound), timed live with CUDA
                 events on the engine's stream over one extra profiled pass
  cpu_baseline : the CPU oracle (PyTorch restatement of the reference's op stream) on a
                 bounded sample, rank 0, N=1 only

`--impl reference` times that CPU oracle alone (the reference is C#/TorchSharp and cannot run
here).  Nothing in the timed GPU path imports oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "audio_seconds_per_second_encode_decode"
UNIT = "audio-s/s"
SAMPLE_RATE = 44100
GFLOP_PER_AUDIO_S = 199.98        # conv/GEMM work of DAC-44.1k encode+decode
WORKLOADS = {
    # name: (global batch, seconds per clip, decode_only)
    "dac44k_b512x30s": (512, 30.0, False),       # BASELINE configs[3] (default: the config the metric is quoted on)
    "dac44k_b1x10s": (1, 10.0, False),           # BASELINE configs[0]
    "dia_dac_decode_b256x20s": (256, 20.004, True),  # BASELINE configs[4]
    "snac24k_b32x10s": (32, 10.0, False),        # BASELINE configs[1]
    "encodec24k_b64x10s": (64, 10.0, False),     # BASELINE configs[2]
    "encodec48k_b32x10s": (32, 10.0, False),     # the 48 kHz stereo preset (segments, GroupNorm, overlap-add)
}


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """Write the one JSON line to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def shard_range(rank: int, world: int, batch: int):
    """Contiguous shard [lo, hi) of the global batch owned by `rank` (static split, no collective)."""
    return rank * batch // world, (rank + 1) * batch // world


def codec_of(workload: str) -> str:
    if workload.startswith("encodec48"):
        return "encodec48"
    return "snac" if workload.startswith("snac") else "encodec" if workload.startswith("encodec") else "dac"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dac44k_b512x30s", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the global batch (debug)")
    ap.add_argument("--seconds", type=float, default=0.0, help="override seconds per clip (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--encoder-precision", default=None)
    ap.add_argument("--decoder-precision", default=None)
    ap.add_argument("--workspace-mb", type=int, default=0)
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); smax.append(float(p[1])); power.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_rate(budget_s: float, weights_path: str, threads: int):
    """audio-s/s of the CPU oracle (encode -> RVQ -> decode) on a bounded sample.  The ONLY place
    bench.py executes oracle/ (the checker, used here as the reported CPU baseline)."""
    import numpy as np
    import torch
    from oracle import dac as odac
    from neuralcodecs_b200 import synthetic

    torch.set_num_threads(threads)
    cfg = odac.DACConfig.dac_44khz()
    model = odac.load_hf_safetensors(weights_path, cfg)

    def run(seconds):
        x = torch.from_numpy(synthetic.synth_audio(1, int(round(seconds * SAMPLE_RATE)), SAMPLE_RATE)).unsqueeze(1)
        t0 = time.perf_counter()
        model.forward(x)
        return time.perf_counter() - t0

    run(0.5)                                   # warm-up (thread pool, oneDNN primitives)
    t1 = run(1.0)
    seconds = max(1.0, min(30.0, budget_s / max(t1, 1e-3)))
    seconds = float(int(seconds))
    dt = run(seconds)
    return seconds / dt, seconds, dt


def reference_arm(args, rank: int):
    """`--impl reference`: the reference's CPU implementation of the path.  The reference itself
    (C# + TorchSharp/libtorch) cannot run here; the oracle issues the same ATen op stream."""
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    wpath = ensure_weights()
    steps, warm = max(1, args.steps), max(0, args.warmup)
    budget = 150.0 / (steps + warm)
    torch.set_num_threads(threads)
    import numpy as np
    from oracle import dac as odac
    from neuralcodecs_b200 import synthetic
    model = odac.load_hf_safetensors(wpath, odac.DACConfig.dac_44khz())

    def run(seconds):
        x = torch.from_numpy(synthetic.synth_audio(1, int(round(seconds * SAMPLE_RATE)), SAMPLE_RATE)).unsqueeze(1)
        t0 = time.perf_counter()
        model.forward(x)
        return time.perf_counter() - t0

    run(0.5)
    t1 = run(1.0)
    seconds = float(int(max(1.0, min(30.0, budget / max(t1, 1e-3)))))
    for _ in range(warm):
        run(seconds)
    times = [run(seconds) for _ in range(steps)]
    total = sum(times)
    value = seconds * steps / total
    B, S, dec_only = WORKLOADS[args.workload]
    sample = f"1 clip x {seconds:.0f} s per step (of {B} x {S:g} s), fp32 PyTorch-CPU restatement of the reference op stream"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def parity_sample(model, weights_path: str, threads: int, seconds: float = 3.0):
    """Parity of THIS run's engine (the weights and precision policy being benchmarked) against the CPU oracle on a
    one-clip sample: RVQ code flips with near-tie accounting, decoder-only and end-to-end audio error.  The checker,
    not the measured path (runs after the timed regions)."""
    import numpy as np
    import torch
    from oracle import dac as odac
    from neuralcodecs_b200 import synthetic

    torch.set_num_threads(threads)
    o = odac.load_hf_safetensors(weights_path, odac.DACConfig.dac_44khz())
    x = synthetic.synth_audio(1, int(seconds * SAMPLE_RATE), SAMPLE_RATE, first_clip=11)
    xt = torch.from_numpy(x).unsqueeze(1)
    ref = o.forward(xt)
    z_e = o.encode_latent(xt)
    out = model.forward(x[:, None, :])
    rep = odac.near_tie_report(o, z_e, ref["codes"], torch.from_numpy(out["codes"]))
    flips = rep["uncascaded_flips"]

    def snr(r, t):
        r = r.astype(np.float64); t = t.astype(np.float64)
        return float(10 * np.log10((r ** 2).sum() / max(((r - t) ** 2).sum(), 1e-300)))

    a_ref = ref["audio"].numpy()
    a_dec = model.Decode(ref["z"].numpy())
    a_tf = o.decode(o.from_codes(torch.from_numpy(out["codes"]))).numpy()      # teacher-forced on the engine's codes
    return {"sample": f"1 clip x {seconds:g} s, bench weights, vs the fp32 CPU oracle (parity unpinned by the reference: it ships no fixtures)",
            "frames": int(rep["frames"]), "code_decisions": int(np.prod(out["codes"].shape)),
            "codes_equal": float((out["codes"] == ref["codes"].numpy()).mean()),
            "frames_flipped": int(rep["frames_flipped"]),
            "flips_at_or_above_1e-6": int(sum(abs(r["margin_scale"]) >= 1e-6 for r in flips)),
            "max_flip_margin": float(max([abs(r["margin_scale"]) for r in flips], default=0.0)),
            "decoder_snr_db": snr(a_ref, a_dec), "decoder_max_abs": float(np.abs(a_dec - a_ref).max()),
            "e2e_snr_db_teacher_forced": snr(a_tf, out["audio"]), "e2e_max_abs_teacher_forced": float(np.abs(out["audio"] - a_tf).max()),
            "gates": {"flip_margin": 1e-6, "max_abs": 1e-3, "snr_db": 60.0}}


# ------------------------------------------------------------------------------------------ weights
def ensure_weights() -> str:
    """Seeded random-init DAC-44.1k weights in the reference's HF safetensors layout."""
    from neuralcodecs_b200 import DACConfig, synthetic
    path = os.path.join(tempfile.gettempdir(), f"nc_bench_dac44_seed{synthetic.WEIGHT_SEED}.safetensors")
    if not os.path.exists(path):
        sd = synthetic.make_dac_weights_hf(DACConfig.DAC44kHz())
        # nn.Embedding-default N(0,1) rows are far from the latents of random-init encoders; scale them
        # into the latents' range so more than one code per stage is used (timing is data independent)
        for k in sd:
            if k.endswith("codebook.weight"):
                sd[k] = (0.05 * sd[k]).astype("float32")
        tmp = f"{path}.{os.getpid()}.tmp"
        synthetic.save_safetensors(sd, tmp)
        os.replace(tmp, path)
    return path


def synth_audio_cuda(torch, batch, length, first_clip, device):
    """Same recipe as neuralcodecs_b200.synthetic.synth_audio, generated on the device."""
    t = torch.arange(length, device=device, dtype=torch.float64) / SAMPLE_RATE
    out = torch.empty(batch, length, device=device, dtype=torch.float32)
    g = torch.Generator(device=device)
    for i in range(batch):
        b = first_clip + i
        f = 110.0 * 2.0 ** ((b % 48) / 12.0)
        g.manual_seed(1234 * 100003 + b)
        x = 0.30 * torch.sin(2 * torch.pi * f * t + 0.37 * b) + 0.15 * torch.sin(2 * torch.pi * 3.1 * f * t)
        x = x.float() + 0.05 * torch.randn(length, device=device, generator=g)
        out[i] = x.clamp_(-1.0, 1.0)
    return out


def codec_weights_path(codec: str) -> str:
    from neuralcodecs_b200 import synthetic
    tag = codec if codec.endswith("48") else codec + "24"
    return os.path.join(tempfile.gettempdir(), f"nc_bench_{tag}_seed{synthetic.WEIGHT_SEED}.safetensors")


def ensure_codec_weights(codec: str) -> str:
    """Seeded random-init SNAC 24 kHz / Encodec 24 kHz weights in the reference's key layouts."""
    import neuralcodecs_b200 as nc
    from neuralcodecs_b200 import synthetic
    wpath = codec_weights_path(codec)
    if not os.path.exists(wpath):
        if codec == "snac":
            sd = synthetic.make_snac_weights(nc.SNACConfig.SNAC24kHz())
        elif codec == "encodec48":
            sd = synthetic.make_encodec_weights(nc.EncodecConfig.Encodec48Khz())
        else:
            sd = synthetic.make_encodec_weights(nc.EncodecConfig.Encodec24Khz())
        for k in sd:   # N(0,1) codebooks scaled into the latents' range (timing is data independent)
            if k.endswith("codebook.weight") or k.endswith("codebook.embed"):
                sd[k] = (0.05 * sd[k]).astype("float32")
        tmp = f"{wpath}.{os.getpid()}.tmp"
        synthetic.save_safetensors(sd, tmp)
        os.replace(tmp, wpath)
    return wpath


def other_configs(dac_model, local_rank: int, steps: int = 3, warmup: int = 3):
    """Short device-resident measurements of BASELINE configs[0], [1], [2] and [4] inside the default run, so the
    driver observes them too (configs[3] is the bench line itself).  Same timing rules: W >= 3 warm-up passes, CUDA
    events on the engine's stream, inputs + activations far larger than L2 (except configs[0], one 10 s clip, where
    the activations of a layer still exceed L2).  value = audio-s/s, inputs resident in HBM."""
    import torch
    import neuralcodecs_b200 as nc
    from neuralcodecs_b200 import synthetic
    dev = torch.device("cuda", local_rank)
    out = {}

    def timed(model, fn, n):
        stream = torch.cuda.ExternalStream(model.stream_ptr(), device=dev)
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    try:
        # configs[0]: DAC 44.1 kHz, one 10 s clip
        L = 441000
        Lp, T = dac_model.query_shapes(L)
        a = synth_audio_cuda(torch, 1, L, 0, dev)
        ao = torch.empty(1, 1, Lp, device=dev); co = torch.empty(1, 9, T, device=dev, dtype=torch.int64)
        ms = timed(dac_model, lambda: dac_model.forward_dev(a.data_ptr(), 1, L, ao.data_ptr(), co.data_ptr()), 20)
        out["dac44k_b1x10s"] = {"value": 10.0 / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": 20, "warmup": warmup}
        # configs[4]: Dia decode stage, 256 x 20 s of codes -> audio
        B, T5 = 256, 1723
        g = torch.Generator(device=dev); g.manual_seed(99)
        codes = torch.randint(0, 1024, (B, 9, T5), device=dev, dtype=torch.int64, generator=g)
        ao = torch.empty(B, 1, T5 * 512, device=dev)
        ms = timed(dac_model, lambda: dac_model.decode_codes_dev(codes.data_ptr(), B, 9, T5, ao.data_ptr()), steps)
        out["dia_dac_decode_b256x20s"] = {"value": B * T5 * 512 / SAMPLE_RATE / (ms / 1e3), "unit": UNIT, "ms_per_step": ms,
                                          "steps": steps, "warmup": warmup}
        del codes, ao
        torch.cuda.empty_cache()
        for codec, B in (("snac", 32), ("encodec", 64)):
            cfg = nc.SNACConfig.SNAC24kHz() if codec == "snac" else nc.EncodecConfig.Encodec24Khz()
            cfg.device = nc.DeviceConfiguration.CUDA(local_rank)
            m = nc.SNAC(cfg) if codec == "snac" else nc.Encodec(cfg)
            m.LoadWeights(ensure_codec_weights(codec))
            L = 240000
            base = torch.from_numpy(synthetic.synth_audio(16, L, 24000)).to(dev)
            audio = base.repeat(B // 16, 1).contiguous()
            ao = torch.empty(B, L, device=dev)
            if codec == "snac":
                _, T, clens, _ = m.query_shapes(L)
                cs = [torch.empty(B, n, dtype=torch.int64, device=dev) for n in clens]
                fn = lambda: m.forward_dev(audio.data_ptr(), B, L, ao.data_ptr(), [c.data_ptr() for c in cs], None, 5)
            else:
                T, nq, _ = m.query_shapes(L)
                cs = torch.empty(B, nq, T, dtype=torch.int64, device=dev)
                fn = lambda: m.forward_dev(audio.data_ptr(), B, L, ao.data_ptr(), cs.data_ptr())
            ms = timed(m, fn, max(steps, 5))
            out[f"{codec}24k_b{B}x10s"] = {"value": B * 10.0 / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": max(steps, 5),
                                           "warmup": warmup}
            m.Dispose()
        # Encodec 48 kHz preset (stereo, GroupNorm, 1 s segments + overlap-add): 16 stereo clips x 10 s
        cfg = nc.EncodecConfig.Encodec48Khz()
        cfg.device = nc.DeviceConfiguration.CUDA(local_rank)
        m = nc.Encodec(cfg)
        m.LoadWeights(ensure_codec_weights("encodec48"))
        B, L = 16, 480000
        base = torch.from_numpy(synthetic.synth_audio(16, L, 48000)).to(dev)
        audio = torch.stack([base, base.flip(0) * 0.7], dim=1).contiguous()
        ao = torch.empty(B, 2, L, device=dev)
        seg, nq, _ = m.query_frames(L)
        cs = torch.empty(B, nq, sum(seg), dtype=torch.int64, device=dev)
        ms = timed(m, lambda: m.forward_dev(audio.data_ptr(), B, L, ao.data_ptr(), cs.data_ptr()), max(steps, 5))
        out["encodec48k_stereo_b16x10s"] = {"value": B * 10.0 / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": max(steps, 5),
                                            "warmup": warmup}
        m.Dispose()
    except Exception as e:   # a side measurement must never take the headline line down
        out["error"] = f"{type(e).__name__}: {e}"
    return out


# ------------------------------------------------------------------------------------------ SNAC / Encodec arms
def run_other_codec(args, codec: str, rank: int, local_rank: int, world: int):
    """BASELINE configs[1] (SNAC 24 kHz) and configs[2] (Encodec 24 kHz, 6 kbps): same JSON contract as the DAC arm."""
    import numpy as np
    import torch
    import neuralcodecs_b200 as nc
    from neuralcodecs_b200 import synthetic

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B, S, _ = WORKLOADS[args.workload]
    B = args.batch or B
    S = args.seconds or S
    sr = 48000 if codec == "encodec48" else 24000
    ch = 2 if codec == "encodec48" else 1
    L = int(round(S * sr))
    lo, hi = shard_range(rank, world, B)
    nb = hi - lo
    wpath = codec_weights_path(codec)
    if rank == 0:
        ensure_codec_weights(codec)
    if dist is not None:
        dist.barrier()
    if codec == "snac":
        cfg = nc.SNACConfig.SNAC24kHz()
        cfg.device = nc.DeviceConfiguration.CUDA(local_rank)
        model = nc.SNAC(cfg)
    else:
        cfg = nc.EncodecConfig.Encodec48Khz() if codec == "encodec48" else nc.EncodecConfig.Encodec24Khz()
        cfg.device = nc.DeviceConfiguration.CUDA(local_rank)
        model = nc.Encodec(cfg)
    model.LoadWeights(wpath)
    base = torch.from_numpy(synthetic.synth_audio(min(max(nb, 1), 16), L, sr, first_clip=lo)).to(dev)
    audio = base.repeat((max(nb, 1) + base.shape[0] - 1) // base.shape[0], 1)[:nb].contiguous()
    if ch == 2:   # planar stereo [nb, 2, L]: second channel = another clip at another level
        audio = torch.stack([audio, audio.flip(0) * 0.7], dim=1).contiguous()
    out = torch.empty(nb, ch * L, device=dev)
    if codec == "snac":
        _, T, clens, _ = model.query_shapes(L)
        codes = [torch.empty(nb, n, dtype=torch.int64, device=dev) for n in clens]
        step_dev = lambda: nb and model.forward_dev(audio.data_ptr(), nb, L, out.data_ptr(), [c.data_ptr() for c in codes], None, 5)
        code_bytes = sum(c.numel() for c in codes) * 8
    else:
        seg, nq, _ = model.query_frames(L)          # one segment for the 24 kHz preset
        T = sum(seg)
        codes = torch.empty(nb, nq, T, dtype=torch.int64, device=dev)
        step_dev = lambda: nb and model.forward_dev(audio.data_ptr(), nb, L, out.data_ptr(), codes.data_ptr())
        code_bytes = codes.numel() * 8
    stream = torch.cuda.ExternalStream(model.stream_ptr(), device=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        t = torch.tensor([e0.elapsed_time(e1), wall * 1e3], device=dev, dtype=torch.float64)
        barrier()
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = model.launch_count()
    dev_ms, wall_ms = timed(step_dev, args.steps)
    launches = model.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    audio_s = B * L / sr
    value = audio_s * args.steps / (dev_ms / 1e3)
    # e2e through the host-buffer entry point
    # pinned host buffers, raw-pointer host entry point (copies inside the call)
    h_in = torch.empty(max(nb, 1), ch * L, dtype=torch.float32).pin_memory()
    h_in[:nb].copy_(audio.reshape(nb, ch * L))
    h_out = torch.empty(max(nb, 1), ch * L, dtype=torch.float32).pin_memory()
    if codec == "snac":
        h_codes = [torch.empty(max(nb, 1), n, dtype=torch.int64).pin_memory() for n in clens]
        step_host = lambda: nb and model.forward_host(h_in.data_ptr(), nb, L, h_out.data_ptr(), [c.data_ptr() for c in h_codes], 5)
    else:
        h_codes = torch.empty(max(nb, 1), nq, T, dtype=torch.int64).pin_memory()
        step_host = lambda: nb and model.forward_host(h_in.data_ptr(), nb, L, h_out.data_ptr(), h_codes.data_ptr())
    step_host()
    _, e2e_ms = timed(step_host, args.steps)
    e2e = {"value": audio_s * args.steps / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(B * ch * L * 4),
           "d2h_bytes_per_step": int(B * ch * L * 4 + code_bytes * B // max(nb, 1)),
           "timer": "wall clock, max over ranks"}
    roofline = kernels = cpu = None
    if rank == 0 and nb:
        peaks, src = measured_peaks()
        model.set_option("profile", "1"); model.profile_report(); step_dev(); rep = model.profile_report()
        model.set_option("profile", "0")
        total_ms = sum(v["ms"] for v in rep.values())
        kernels = {k: {"launches": v["launches"], "ms": round(v["ms"], 3), "share": round(v["ms"] / total_ms, 4),
                       "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 2),
                       "gbs": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1)} for k, v in rep.items()}
        top = max(rep.items(), key=lambda kv: kv[1]["ms"])
        k, v = top
        # a tcgen05 conv launch can sit under either roof (SNAC's 1x1 / depthwise-fused units move far more bytes than
        # they multiply): report the roof it is closer to
        hbm_frac = v["bytes"] / v["ms"] / 1e6 / peaks["hbm_gbs"]
        tc_frac = v["flops"] / v["ms"] / 1e9 / peaks["bf16_tflops_sustained"]
        if (k.startswith("conv_umma") and tc_frac >= hbm_frac) or k.startswith("conv_simt") or k == "lstm_layer":
            peak = peaks["bf16_tflops_sustained"] if k.startswith("conv_umma") else 70.0
            ach = v["flops"] / v["ms"] / 1e9
            roofline = {"bound": "tensor", "kernel": k, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                        "traffic": None, "launches": v["launches"], "avg_launch_ms": v["ms"] / v["launches"],
                        "share_of_step": v["ms"] / total_ms,
                        "peak_note": f"{src} bf16_tflops_sustained" if k.startswith("conv_umma") else "fp32 FFMA nominal ~70 TFLOP/s (CUDA-core kernel)"}
        else:
            ach = v["bytes"] / v["ms"] / 1e6
            roofline = {"bound": "hbm", "kernel": k, "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": ach / peaks["hbm_gbs"], "traffic": None, "launches": v["launches"],
                        "avg_launch_ms": v["ms"] / v["launches"], "share_of_step": v["ms"] / total_ms,
                        "peak_note": f"{src} hbm_gbs (copy bandwidth)"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        if codec == "snac":
            from oracle import snac as om
            o = om.load_safetensors(wpath, om.SNACConfig.snac_24khz())
        else:
            from oracle import encodec as om
            o = om.load_safetensors(wpath, om.EncodecConfig.encodec_48khz() if codec == "encodec48" else om.EncodecConfig())
        xs = torch.from_numpy(synthetic.synth_audio(2 * ch, L, sr)).reshape(2, ch, L)
        fwd = o.forward_frames if codec == "encodec48" else o.forward
        fwd(xs[:1, :, : sr])
        t0 = time.perf_counter(); fwd(xs); dt = time.perf_counter() - t0
        cpu = {"value": 2 * S / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"2 clips x {S:g} s (of {B}) in {dt:.1f} s, fp32 PyTorch-CPU restatement of the reference op stream"}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 in/out; see config.precision", "data": "synthetic",
                "config": {"workload": args.workload, "codec": "SNAC 24 kHz" if codec == "snac" else "Encodec 48 kHz stereo 6 kbps" if codec == "encodec48" else "Encodec 24 kHz 6 kbps",
                           "global_batch": B, "clip_seconds": S, "clips_per_gpu": nb, "precision": model.describe().get("precision")
                           or f"encoder {model.describe().get('encoder_precision')}, decoder {model.describe().get('decoder_precision')}",
                           "l2": "inputs + activations far larger than L2; no flush", "weights": "random-init (seeded)"},
                "wall_ms_per_step": wall_ms / args.steps, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "roofline": roofline, "cpu_baseline": cpu, "kernels": kernels}
        emit(line)
    model.Dispose()
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse_args()
    # The contract is ONE JSON line on stdout, but libraries write there too (NCCL prints its version banner to stdout
    # at any NCCL_DEBUG level >= VERSION): point fd 1 at stderr for the whole run and keep the real stdout for emit().
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if codec_of(args.workload) != "dac":
        import torch
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback (use --impl reference)")
        run_other_codec(args, codec_of(args.workload), rank, local_rank, world)
        return

    import numpy as np
    import torch
    import neuralcodecs_b200 as nc

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, S, dec_only = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    if args.seconds:
        S = args.seconds
    L = int(round(S * SAMPLE_RATE))
    # contiguous shard of the global batch
    lo, hi = shard_range(rank, world, B)
    nb = hi - lo

    wpath = ensure_weights() if rank == 0 else None
    if dist is not None:
        dist.barrier()
        wpath = ensure_weights()
    cfg = nc.DACConfig.DAC44kHz()
    cfg.device = nc.DeviceConfiguration.CUDA(local_rank)
    opts = {}
    if args.encoder_precision:
        opts["encoder_precision"] = args.encoder_precision
    if args.decoder_precision:
        opts["decoder_precision"] = args.decoder_precision
    if args.workspace_mb:
        opts["max_workspace_mb"] = str(args.workspace_mb)
    model = nc.DAC(cfg, options=opts)
    model.LoadWeights(wpath)
    Lp, T = model.query_shapes(L)
    nq = cfg.num_codebooks

    # ---- inputs resident in HBM
    codes_in = None
    if dec_only:
        gen = torch.Generator(device=dev); gen.manual_seed(99 + lo)
        codes_in = torch.randint(0, cfg.codebook_size, (nb, nq, T), device=dev, dtype=torch.int64, generator=gen)
        audio = None
    else:
        audio = synth_audio_cuda(torch, nb, L, lo, dev) if nb else torch.empty(0, L, device=dev)
    audio_out = torch.empty(nb, 1, Lp, device=dev, dtype=torch.float32)
    codes_out = torch.empty(nb, nq, T, device=dev, dtype=torch.int64)
    stream = torch.cuda.ExternalStream(model.stream_ptr(), device=dev)

    def step_dev():
        if nb == 0:
            return
        if dec_only:
            model.decode_codes_dev(codes_in.data_ptr(), nb, nq, T, audio_out.data_ptr())
        else:
            model.forward_dev(audio.data_ptr(), nb, L, audio_out.data_ptr(), codes_out.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dev_ms = e0.elapsed_time(e1)
        barrier()
        t = torch.tensor([dev_ms, wall * 1e3], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = model.launch_count()
    dev_ms, wall_ms = timed(step_dev, args.steps)
    launches = model.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    audio_seconds_per_step = B * L / SAMPLE_RATE          # unpadded input length
    value = audio_seconds_per_step * args.steps / (dev_ms / 1e3)

    # ---- e2e: host buffers through the public host-pointer entry point
    e2e = None
    if not args.no_e2e:
        if dec_only:
            h_in = torch.empty(nb, nq, T, dtype=torch.int64).pin_memory()
            h_in.copy_(codes_in)
        else:
            h_in = torch.empty(nb, L, dtype=torch.float32).pin_memory()
            h_in.copy_(audio)
        h_audio = torch.empty(nb, 1, Lp, dtype=torch.float32).pin_memory()
        h_codes = torch.empty(nb, nq, T, dtype=torch.int64).pin_memory()

        def step_host():
            if nb == 0:
                return
            if dec_only:
                model.decode_codes_host(h_in.data_ptr(), nb, nq, T, h_audio.data_ptr())
            else:
                model.forward_host(h_in.data_ptr(), nb, L, h_audio.data_ptr(), h_codes.data_ptr())

        for _ in range(max(1, min(args.warmup, 2))):
            step_host()
        _, e2e_wall_ms = timed(step_host, args.steps)
        h2d = h_in.numel() * h_in.element_size()
        d2h = h_audio.numel() * 4 + (0 if dec_only else h_codes.numel() * 8)
        if dist is not None:
            t = torch.tensor([h2d, d2h], device=dev, dtype=torch.float64)
            dist.all_reduce(t)
            h2d, d2h = int(t[0]), int(t[1])
        e2e = {"value": audio_seconds_per_step * args.steps / (e2e_wall_ms / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "timer": "wall clock, max over ranks"}

    # ---- roofline of the dominant kernel: one extra pass with per-launch CUDA events
    roofline = None
    kernels = None
    if rank == 0 and nb:
        peaks, peak_src = measured_peaks()
        model.set_option("profile", "1")
        model.profile_report()
        step_dev()
        rep = model.profile_report()
        model.set_option("profile", "0")
        total_ms = sum(v["ms"] for v in rep.values())
        kernels = {k: {"launches": v["launches"], "ms": round(v["ms"], 3), "share": round(v["ms"] / total_ms, 4),
                       "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 2),
                       "gbs": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1)} for k, v in rep.items()}
        mma = {k: v for k, v in rep.items() if k.startswith(("conv_umma", "ru_fused", "conv_h16"))}
        if mma:
            # tensor-core passes per algorithmic FLOP and the peak of the operand kind actually issued
            def passes(k): return 1 if k.endswith(("_tf32", "_f16", "conv_h16")) else (2 if k.endswith("_f16x2") else 3)
            def kind_peak(k): return peaks["bf16_tflops_sustained"] / (1.0 if ("bf16" in k or "f16" in k or "h16" in k) else 2.0)
            fl = sum(v["flops"] for v in mma.values())
            ms = sum(v["ms"] for v in mma.values())
            n = sum(v["launches"] for v in mma.values())
            issued = sum(v["flops"] * passes(k) for k, v in mma.items())
            busy = sum(v["flops"] * passes(k) / kind_peak(k) for k, v in mma.items())   # seconds*1e12 at peak
            peak = peaks["bf16_tflops_sustained"]
            ach = fl / ms / 1e9
            traffic = None
            try:   # DRAM bytes per launch from the committed ncu capture of the same kernels, scaled by audio seconds
                tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_dram_traffic_dac.json")))
                if not dec_only:
                    traffic = tj["dram_bytes"] / (tj["clips"] * tj["clip_seconds"]) * (nb * L / SAMPLE_RATE) / n
            except Exception:
                traffic = None
            roofline = {"bound": "tensor", "kernel": "conv_umma_kernel + conv_ru_fused_kernel + conv_h16_kernel (tcgen05.mma, TMA-fed implicit-GEMM conv; all conv layers)",
                        "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                        "traffic": traffic, "traffic_note": "bytes per launch (avg), ncu dram read+write of these kernels (profiles/r01_dram_traffic_dac.json) scaled to this step; algorithmic bytes = 1.10x of it (re-reads hit L2)", "launches": n, "avg_launch_ms": ms / n, "share_of_step": ms / total_ms,
                        "issued_tflops": issued / ms / 1e9, "issued_frac": busy / ms / 1e9,
                        "peak_note": f"{peak_src} bf16_tflops_sustained (cuBLAS bf16 under the power cap). achieved = algorithmic "
                                     "conv FLOPs (2*MAC) per second; issued_* counts the MMAs actually issued per product (3 for the "
                                     "bf16x3 / 3xtf32 operand splits the code-parity gate requires, 1 for the fp16 wide decoder layers; tf32 kinds against peak/2)"}
            # tensor-pipe activity of the same launches, measured under ncu (time-weighted over one micro-batch): a
            # profiler number, reported beside the live ones with its source
            try:
                tp = json.load(open(os.path.join(ROOT, "profiles", "r02_tensor_pipe.json")))
                roofline["tensor_pipe_pct"] = tp["tensor_pipe_pct_time_weighted"]
                roofline["tensor_pipe_note"] = tp["note"]
            except Exception:
                roofline["tensor_pipe_pct"] = None
            # the fused narrow residual units (N = 64 / 96 / 128) are bound by the issue interval of their single MMA thread: one
            # thread sustains one tcgen05.mma per ~80 clk whatever N is (tools/probe_mma_rate.cu), against tensor-core floors of
            # 32 / 48 / 64 clk; the role timeline measures ~100 clk per MMA in the units
            fused = {k: v for k, v in rep.items() if k.startswith("ru_fused")}
            if fused:
                roofline["fused_units_mma_issue"] = {
                    "kernel": "conv_ru_fused_kernel", "bound": "tcgen05.mma issue interval of one thread (~80 clk per MMA, any N)",
                    "share_of_step": sum(v["ms"] for v in fused.values()) / total_ms,
                    "clk_per_mma_measured": 100, "clk_per_mma_issue_floor": 80, "tensor_floor_clk_by_width": {"C=64": 32, "C=96": 48, "C=128": 64},
                    "source": "profiles/r02_mma_issue_rate_probe.txt (two issuing warps double the rate; cta_group::2 halves the issue cost per "
                              "row) and profiles/r02_ru_fused_timeline.txt (clock64 at every role hand-off of CTA 0)"}
            # the other kernel BASELINE.json's metric names: the fused RVQ (algorithmic bytes per frame: read z, write
            # z_q, write codes = SURVEY 8d's 8.3 KB at the 44.1 kHz preset)
            rv = rep.get("rvq_encode")
            if rv and rv["ms"] > 0:
                gbs = rv["bytes"] / rv["ms"] / 1e6
                roofline["rvq"] = {"kernel": "rvq_encode_block_kernel", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                   "frac": gbs / peaks["hbm_gbs"], "launches": rv["launches"], "avg_launch_ms": rv["ms"] / rv["launches"],
                                   "share_of_step": rv["ms"] / total_ms, "fp32_tflops": rv["flops"] / rv["ms"] / 1e9,
                                   "note": "442 KFLOP of exact-fp32 work per 8.3 KB frame = 53 FLOP/B, 5x the fp32 ridge of the "
                                           "machine: the kernel's own ceiling is fp32 FMA issue / shared-memory bandwidth (about 20 % "
                                           "of HBM peak), see DESIGN.md 4"}

    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        parity = parity_sample(model, wpath, os.cpu_count() or 1)

    others = None
    if rank == 0 and world == 1 and not args.no_other_configs and args.workload == "dac44k_b512x30s" and not args.batch:
        del audio, audio_out, codes_out
        torch.cuda.empty_cache()
        others = other_configs(model, local_rank)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not dec_only:
        threads = os.cpu_count() or 1
        rate, sec, dt = cpu_oracle_rate(20.0, wpath, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"1 clip x {sec:.0f} s (of {B} x {S:g} s) in {dt:.1f} s, fp32 PyTorch-CPU restatement of the "
                         "reference op stream (reference is C#/TorchSharp: not runnable here)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32 in/out; tensor-core operands bf16x3-split (encoder, narrow decoder layers) / fp16 (wide decoder layers), f32 accumulate",
                "data": "synthetic",
                "config": {"workload": args.workload, "codec": "DAC 44.1 kHz 9 codebooks",
                           "global_batch": B, "clip_seconds": S, "clips_per_gpu": nb, "frames_per_clip": T,
                           "decode_only": dec_only, "precision": model.precision_summary(),
                           "l2": "inputs (%.1f MB/rank) larger than L2; no flush" % (nb * L * 4 / 1e6),
                           "weights": "random-init (seeded), HF DacModel safetensors layout"},
                "wall_ms_per_step": wall_ms / args.steps,
                "gflop_per_audio_s": GFLOP_PER_AUDIO_S if not dec_only else 138.6,
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "cpu_baseline": cpu, "parity": parity, "other_configs": others, "kernels": kernels}
        emit(line)
    model.Dispose()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
