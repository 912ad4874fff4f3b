"""Decoder operand modes vs the fp32 oracle (DAC 44.1k, seeded weights): per-clip SNR / max-abs of Decode(z_ref).
python scripts/decoder_precision_exp.py [clips] [seconds] [mode ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth, dac as odac
import neuralcodecs_b200 as nc

def snr_db(ref, test):
    ref = ref.astype(np.float64); test = test.astype(np.float64)
    return 10 * np.log10((ref ** 2).sum() / max(((ref - test) ** 2).sum(), 1e-300))

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
S = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
modes = sys.argv[3:] or ["bf16x3", "f16x3", "f16x2", "f16", "tf32"]
co = odac.DACConfig.dac_44khz(); ce = nc.DACConfig.DAC44kHz()
path = os.path.join("tests", ".cache", "dac44_seed4321.safetensors")
if not os.path.exists(path):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    synth.save_safetensors(synth.make_dac_weights_hf(co, codebooks="data", codebook_seconds=10.0), path)
o = odac.load_hf_safetensors(path, co)
x = synth.synth_audio(B, int(S * 44100), 44100, first_clip=31)
x[1::3] *= 0.02          # quiet clips
x[2::3] *= 3.0           # hot clips
xt = torch.from_numpy(x).unsqueeze(1)
t0 = time.time()
ref = o.forward(xt)
print(f"oracle {time.time()-t0:.1f}s  frames {ref['codes'].shape[-1]} x {B}; audio rms per clip {np.sqrt((ref['audio'].numpy()**2).mean(axis=(1,2)))}")
a_ref = ref["audio"].numpy()
for dec in modes:
    m = nc.DAC(ce, options={"decoder_precision": dec})
    m.LoadWeights(path)
    a = m.Decode(ref["z"].numpy())
    per = [snr_db(a_ref[b], a[b]) for b in range(B)]
    print(f"dec={dec:7s}: snr all {snr_db(a_ref, a):.1f} dB, per clip min {min(per):.1f} max {max(per):.1f}; maxabs {np.abs(a - a_ref).max():.2e}")
    m.Dispose()
