"""One DAC decode (codes -> audio) on device-resident input: the target of ncu captures of the decoder kernels.
usage: python scripts/one_decode.py [batch] [seconds] [k=v ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neuralcodecs_b200 as nc
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
S = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
opts = dict(kv.split("=") for kv in sys.argv[3:])
dev = torch.device("cuda", 0)
m = nc.DAC(nc.DACConfig.DAC44kHz(), options=opts)
m.LoadWeights(bench.ensure_weights())
L = int(S * 44100)
Lp, T = m.query_shapes(L)
g = torch.Generator(device=dev); g.manual_seed(99)
codes = torch.randint(0, 1024, (B, 9, T), device=dev, dtype=torch.int64, generator=g)
out = torch.empty(B, 1, Lp, device=dev)
for _ in range(2):
    m.decode_codes_dev(codes.data_ptr(), B, 9, T, out.data_ptr())
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
