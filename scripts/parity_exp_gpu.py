"""GPU-box side of the code-parity experiment: engine Encode of the same clips under several option sets;
codes + latents go to gpurun_out/parity_exp/<tag>.npz for scripts/parity_analyze.py (run in the container)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import neuralcodecs_b200 as nc
from neuralcodecs_b200 import synthetic
from scripts.exp_common import dac44_weights_file

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
S = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
FIRST = int(sys.argv[3]) if len(sys.argv) > 3 else 11
variants = json.loads(sys.argv[4]) if len(sys.argv) > 4 else {
    "default": {}, "tail1": {"encoder_tail_fp32": "1"}, "tail2": {"encoder_tail_fp32": "2"},
    "tail3": {"encoder_tail_fp32": "3"}, "tail5": {"encoder_tail_fp32": "5"}, "tail7": {"encoder_tail_fp32": "7"},
    "fp32": {"encoder_precision": "fp32"}, "f16x3": {"encoder_precision": "f16x3"}}
out_dir = os.path.join("gpurun_out", "parity_exp")
os.makedirs(out_dir, exist_ok=True)
path = dac44_weights_file()
ce = nc.DACConfig.DAC44kHz()
x = synthetic.synth_audio(N, int(S * 44100), 44100, first_clip=FIRST)
for tag, opts in variants.items():
    m = nc.DAC(ce, options=opts)
    m.LoadWeights(path)
    m.Encode(x[:1, None, :])
    t0 = time.time()
    z, codes, latents = m.Encode(x[:, None, :])
    dt = time.time() - t0
    np.savez(os.path.join(out_dir, f"{tag}_{N}x{int(S)}s_first{FIRST}.npz"), codes=codes, latents=latents)
    print(f"{tag}: {opts} encode {dt*1e3:.1f} ms", flush=True)
    m.Dispose()
