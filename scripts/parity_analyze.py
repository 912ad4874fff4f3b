"""Container side: classify the engine's code flips against the oracle outputs of parity_ref.py."""
import os, sys, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import dac as odac
from scripts.exp_common import SCRATCH, dac44_state_dict

ref_path = sys.argv[1]
ref = np.load(ref_path)
suffix = os.path.basename(ref_path)[len("parity_ref_"):]
co = odac.DACConfig.dac_44khz()
o = odac.DACOracle(co, odac.convert_hf_state_dict({k: torch.from_numpy(v) for k, v in dac44_state_dict().items()}))
ze = torch.from_numpy(ref["z_e"]); rc = torch.from_numpy(ref["codes"])
lat = ref["latents"]; lat64 = ref["latents64"] if "latents64" in ref.files else None
def rel(a, b): return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))
if lat64 is not None:
    print(f"oracle fp32 vs fp64: stage-0 latent rel err {rel(lat[:, :8], lat64[:, :8]):.2e}; codes differ in "
          f"{int((ref['codes'] != ref['codes64']).any(1).sum())} frames")
for f in sorted(glob.glob(os.path.join("gpurun_out", "parity_exp", "*_" + suffix))):
    d = np.load(f)
    rep = odac.near_tie_report(o, ze, rc, torch.from_numpy(d["codes"]))
    fl = rep["uncascaded_flips"]
    big = [r for r in fl if abs(r["margin_scale"]) >= 1e-6]
    l0 = d["latents"][:, :8]
    s = f"{os.path.basename(f)[:-len(suffix)-1]:10s}: flipped frames {rep['frames_flipped']}/{rep['frames']} (>=1e-6: {len(big)}; max {max([abs(r['margin_scale']) for r in fl], default=0):.2e}) latent0 rel err vs fp32 {rel(l0, lat[:, :8]):.2e}"
    if lat64 is not None: s += f" vs fp64 {rel(l0, lat64[:, :8]):.2e}"
    print(s)
    for r in big: print("     ", r)
