"""One DAC forward on device-resident input (for ncu): python scripts/one_forward.py [batch] [seconds] [k=v ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neuralcodecs_b200 as nc
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
opts = dict(kv.split("=") for kv in sys.argv[3:])
reps = int(opts.pop("reps", 1))
L = int(S * 44100)
dev = torch.device("cuda", 0)
m = nc.DAC(nc.DACConfig.DAC44kHz(), options=opts)
m.LoadWeights(bench.ensure_weights())
Lp, T = m.query_shapes(L)
audio = bench.synth_audio_cuda(torch, B, L, 0, dev)
out = torch.empty(B, 1, Lp, device=dev); codes = torch.empty(B, 9, T, device=dev, dtype=torch.int64)
for _ in range(reps):
    m.forward_dev(audio.data_ptr(), B, L, out.data_ptr(), codes.data_ptr())
print("done", m.launch_count())
