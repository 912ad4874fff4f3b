"""fp16-operand decoder path (conv_h16.cu) against the fp32-activation path and the CPU oracle.
usage: python scripts/h16_check.py mid|full [batch] [seconds]   (env NC_H16_PAIR=0 forces single-CTA tiles)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neuralcodecs_b200 as nc
from neuralcodecs_b200 import synthetic
from oracle import dac as odac

def snr_db(ref, test):
    ref = ref.astype(np.float64); test = test.astype(np.float64)
    return 10 * np.log10((ref ** 2).sum() / max(((ref - test) ** 2).sum(), 1e-300))

which = sys.argv[1] if len(sys.argv) > 1 else "mid"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
S = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
if which == "mid":
    co = odac.DACConfig(sample_rate=16000, encoder_dim=32, decoder_dim=512, n_codebooks=4, codebook_size=256)
    ce = nc.DACConfig(sample_rate=16000, encoder_dim=32, decoder_dim=512, num_codebooks=4, codebook_size=256)
else:
    co, ce = odac.DACConfig.dac_44khz(), nc.DACConfig.DAC44kHz()
sd = synthetic.make_dac_weights_hf(ce)
path = f"/tmp/h16_check_{which}.safetensors"
synthetic.save_safetensors(sd, path)
x = synthetic.synth_audio(B, int(S * ce.sample_rate) + 37, ce.sample_rate, first_clip=3)
outs = {}
for tag, opts in (("h16", {}), ("f32act", {"decoder_h16": "0"}), ("bf16x3", {"decoder_precision": "bf16x3"})):
    m = nc.DAC(ce, options=opts)
    m.LoadWeights(path)
    if tag == "h16":
        z = m.EncodeAudio(x[:, None, :])
        d = m.describe()["layers"]
        print({k: v for k, v in d.items() if k.startswith("decoder")} if which == "mid" else "", flush=True)
    t0 = time.time()
    outs[tag] = m.Decode(z)
    print(f"{tag}: decode {1e3 * (time.time() - t0):.1f} ms, finite={np.isfinite(outs[tag]).all()}", flush=True)
    m.Dispose()
o = odac.load_hf_safetensors(path, co)
ref = o.decode(torch.from_numpy(z)).numpy()
for tag, a in outs.items():
    print(f"{tag:7s} vs oracle: snr {snr_db(ref, a):6.1f} dB  max-abs {np.abs(a - ref).max():.2e}")
print(f"h16 vs f32act: snr {snr_db(outs['f32act'], outs['h16']):.1f} dB  max-abs {np.abs(outs['h16'] - outs['f32act']).max():.2e}")
ok = snr_db(ref, outs["h16"]) >= 60 and np.abs(outs["h16"] - ref).max() <= 1e-3
print("H16_OK" if ok else "H16_FAIL")
sys.exit(0 if ok else 1)
