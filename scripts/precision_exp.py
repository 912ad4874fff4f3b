"""Code-flip / SNR experiment: encoder + decoder operand modes vs the fp32 oracle (DAC 44.1k, seeded weights)."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth, dac as odac
import neuralcodecs_b200 as nc

def snr_db(ref, test):
    ref = ref.astype(np.float64); test = test.astype(np.float64)
    return 10 * np.log10((ref ** 2).sum() / max(((ref - test) ** 2).sum(), 1e-300))

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
co = odac.DACConfig.dac_44khz(); ce = nc.DACConfig.DAC44kHz()
path = os.path.join("tests", ".cache", "dac44_seed4321.safetensors")
if not os.path.exists(path):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    synth.save_safetensors(synth.make_dac_weights_hf(co, codebooks="data", codebook_seconds=10.0), path)
o = odac.load_hf_safetensors(path, co)
x = synth.synth_audio(B, int(S * 44100), 44100, first_clip=11)
xt = torch.from_numpy(x).unsqueeze(1)
t0 = time.time()
ref = o.forward(xt); ze = o.encode_latent(xt)
print(f"oracle {time.time()-t0:.1f}s  frames {ref['codes'].shape[-1]} x {B}")
a_ref = ref["audio"].numpy()
for enc, dec, fs in [("3xtf32", "3xtf32", "-1"), ("3xtf32", "3xtf32", "1"), ("f16x3", "f16x3", "-1"), ("f16x3", "f16x3", "1"),
                     ("bf16x3", "bf16x3", "-1"), ("tf32", "tf32", "-1"), ("fp32", "fp32", "-1")]:
    m = nc.DAC(ce, options={"encoder_precision": enc, "decoder_precision": dec, "fast_sin": fs})
    m.LoadWeights(path)
    out = m.forward(x[:, None, :])
    rep = odac.near_tie_report(o, ze, ref["codes"], torch.from_numpy(out["codes"]))
    fl = rep["uncascaded_flips"]
    big = [r for r in fl if abs(r["margin_scale"]) >= 1e-6]
    a_dec = m.Decode(ref["z"].numpy())
    zerr = np.abs(out["z"] - ref["z"].numpy()).max()
    print(f"enc={enc:7s} dec={dec:7s} fast_sin={fs:>2s}: flipped frames {rep['frames_flipped']}/{rep['frames']} "
          f"uncascaded {len(fl)} (>=1e-6: {len(big)}; max |margin_scale| {max([abs(r['margin_scale']) for r in fl], default=0):.2e}) "
          f"| dec-only snr {snr_db(a_ref, a_dec):.1f} dB maxabs {np.abs(a_dec-a_ref).max():.2e} | e2e snr {snr_db(a_ref, out['audio']):.1f}")
    m.Dispose()
