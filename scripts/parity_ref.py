"""CPU-container side of the code-parity experiment: fp32 (and fp64) oracle encode of N x S-second clips.

Writes scratch/dac44_codebooks.npz (data-fitted codebooks, generated once) and
scratch/parity_ref_<N>x<S>s.npz {z_e, codes, latents, z_e64 (optional)}."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth, dac as odac
from scripts.exp_common import SCRATCH, LOCAL, CODEBOOKS

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
S = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
FIRST = int(sys.argv[3]) if len(sys.argv) > 3 else 11
F64 = (sys.argv[4] == "1") if len(sys.argv) > 4 else True
os.makedirs(SCRATCH, exist_ok=True)
co = odac.DACConfig.dac_44khz()
if not os.path.exists(CODEBOOKS):
    sd = synth.make_dac_weights_hf(co, codebooks="data", codebook_seconds=10.0)
    np.savez(CODEBOOKS, **{k: v for k, v in sd.items() if k.endswith("codebook.weight")})
else:
    from scripts.exp_common import dac44_state_dict
    sd = dac44_state_dict()
hf = {k: torch.from_numpy(v) for k, v in sd.items()}
o = odac.DACOracle(co, odac.convert_hf_state_dict(hf))
x = synth.synth_audio(N, int(S * co.sample_rate), co.sample_rate, first_clip=FIRST)
xt = torch.from_numpy(x).unsqueeze(1)
t0 = time.time()
ze = o.encode_latent(xt)
with torch.inference_mode():
    zq, codes, latents = o.rvq_forward(ze)
print(f"fp32 oracle encode {time.time() - t0:.1f}s", flush=True)
out = {"z_e": ze.numpy(), "codes": codes.numpy(), "latents": latents.numpy(), "z_q": zq.numpy()}
if F64:
    t0 = time.time()
    o64 = odac.DACOracle(co, odac.convert_hf_state_dict(hf), torch.float64)
    ze64 = o64.encode_latent(xt.double())
    with torch.inference_mode():
        _, codes64, latents64 = o64.rvq_forward(ze64)
    out.update({"z_e64": ze64.numpy(), "codes64": codes64.numpy(), "latents64": latents64.numpy()})
    print(f"fp64 oracle encode {time.time() - t0:.1f}s", flush=True)
os.makedirs(LOCAL, exist_ok=True)
np.savez(os.path.join(LOCAL, f"parity_ref_{N}x{int(S)}s_first{FIRST}.npz"), **out)
