"""Device-resident timing of SNAC / Encodec forward: python scripts/time_codec.py snac|encodec|encodec48 [batch] [seconds] [k=v ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neuralcodecs_b200 as nc
from neuralcodecs_b200 import synthetic
which = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 32; S = float(sys.argv[3]) if len(sys.argv) > 3 else 10.0
opts = dict(kv.split("=") for kv in sys.argv[4:])
prof = opts.pop("prof", "1")
dev = torch.device("cuda", 0)
import tempfile
if which == "snac":
    cfg = nc.SNACConfig.SNAC24kHz(); sr = 24000
    path = os.path.join(tempfile.gettempdir(), "snac24_bench.safetensors")
    if not os.path.exists(path): synthetic.save_safetensors(synthetic.make_snac_weights(cfg), path)
    m = nc.SNAC(cfg, options=opts)
elif which == "encodec48":
    cfg = nc.EncodecConfig.Encodec48Khz(); sr = 48000
    path = os.path.join(tempfile.gettempdir(), "encodec48_bench.safetensors")
    if not os.path.exists(path): synthetic.save_safetensors(synthetic.make_encodec_weights(cfg), path)
    m = nc.Encodec(cfg, options=opts)
else:
    cfg = nc.EncodecConfig.Encodec24Khz(); sr = 24000
    path = os.path.join(tempfile.gettempdir(), "encodec24_bench.safetensors")
    if not os.path.exists(path): synthetic.save_safetensors(synthetic.make_encodec_weights(cfg), path)
    m = nc.Encodec(cfg, options=opts)
m.LoadWeights(path)
L = int(S * sr)
x = torch.from_numpy(synthetic.synth_audio(min(B, 8), L, sr)).to(dev).repeat((B + 7) // 8, 1)[:B].contiguous()
out = torch.empty(B, L, device=dev)
if which == "encodec48":   # stereo: [B][2][L]
    x = torch.stack([x, x.flip(0) * 0.7], dim=1).contiguous()
    out = torch.empty(B, 2, L, device=dev)
    seg, nq, _ = m.query_frames(L)
    codes = torch.empty(B, nq, sum(seg), dtype=torch.int64, device=dev)
    run = lambda: m.forward_dev(x.data_ptr(), B, L, out.data_ptr(), codes.data_ptr())
elif which == "snac":
    _, T, clens, nlens = m.query_shapes(L)
    codes = [torch.empty(B, n, dtype=torch.int64, device=dev) for n in clens]
    run = lambda: m.forward_dev(x.data_ptr(), B, L, out.data_ptr(), [c.data_ptr() for c in codes], None, 5)
else:
    T, nq, _ = m.query_shapes(L)
    codes = torch.empty(B, nq, T, dtype=torch.int64, device=dev)
    run = lambda: m.forward_dev(x.data_ptr(), B, L, out.data_ptr(), codes.data_ptr())
for _ in range(3): run()
torch.cuda.synchronize(); t0 = time.perf_counter()
n = 5
for _ in range(n): run()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
print(f"# {which} B={B} x {S}s opts={opts}: {dt*1e3:.2f} ms/step -> {B*S/dt:.0f} audio-s/s; launches/step {m.launch_count()//(n+3)}")
if prof != "0":
    m.set_option("profile", prof); m.profile_report(); run(); rep = m.profile_report()
    tot = sum(v["ms"] for v in rep.values())
    for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])[:40]:
        print(f"{k:62s} n={v['launches']:3d} {v['ms']:8.3f} ms {v['ms']/tot:6.3f} {v['flops']/max(v['ms'],1e-9)/1e9:8.1f} TF/s {v['bytes']/max(v['ms'],1e-9)/1e6:7.0f} GB/s")
