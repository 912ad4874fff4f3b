"""Per-layer CUDA-event timing of one DAC forward (engine option profile=2).
usage: python scripts/layer_profile.py [batch] [seconds] [enc_prec] [dec_prec] > profiles/...txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neuralcodecs_b200 as nc
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
opts = {}
if len(sys.argv) > 3: opts["encoder_precision"] = sys.argv[3]
if len(sys.argv) > 4: opts["decoder_precision"] = sys.argv[4]
for kv in sys.argv[5:]:
    k, v = kv.split("="); opts[k] = v
L = int(S * 44100)
dev = torch.device("cuda", 0)
m = nc.DAC(nc.DACConfig.DAC44kHz(), options=opts)
m.LoadWeights(bench.ensure_weights())
Lp, T = m.query_shapes(L)
audio = bench.synth_audio_cuda(torch, B, L, 0, dev)
out = torch.empty(B, 1, Lp, device=dev); codes = torch.empty(B, 9, T, device=dev, dtype=torch.int64)
for _ in range(2):
    m.forward_dev(audio.data_ptr(), B, L, out.data_ptr(), codes.data_ptr())
m.set_option("profile", "2"); m.profile_report()
m.forward_dev(audio.data_ptr(), B, L, out.data_ptr(), codes.data_ptr())
rep = m.profile_report()
tot = sum(v["ms"] for v in rep.values())
print(f"# DAC 44.1k forward B={B} x {S}s, {m.precision_summary()}; total {tot:.2f} ms -> {B*S/tot*1e3:.1f} audio-s/s")
print(f"{'layer':58s} {'ms':>8s} {'share':>6s} {'TFLOP/s':>8s} {'GB/s':>7s}")
order = list(m.describe()["layers"].keys())
def key(k):
    n = k.split(" [")[0]
    return order.index(n) if n in order else -1
for k in sorted(rep, key=key):
    v = rep[k]
    print(f"{k:58s} {v['ms']:8.3f} {v['ms']/tot:6.3f} {v['flops']/max(v['ms'],1e-9)/1e9:8.1f} {v['bytes']/max(v['ms'],1e-9)/1e6:7.0f}")
