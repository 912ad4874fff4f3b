"""Shared helpers of the GPU-box experiment scripts.

Oracle work is done in the CPU container (no GPU-minutes): `scripts/parity_ref.py` writes the data-fitted
codebooks and the oracle's outputs under scratch/ (git-ignored, shipped to the box by gpurun); the box
regenerates the seeded weights (numpy PCG64: same bytes everywhere), swaps in those codebooks and only
runs the engine.  Analysis (`scripts/parity_analyze.py`) happens back in the container."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SCRATCH = os.path.join(ROOT, "scratch")
LOCAL = os.path.join(ROOT, "scratch_local")   # oracle outputs: stay in the container
CODEBOOKS = os.path.join(SCRATCH, "dac44_codebooks.npz")


def dac44_state_dict():
    """Seeded DAC-44.1k weights (HF layout) with the codebooks fitted by parity_ref.py."""
    from neuralcodecs_b200 import synthetic
    import neuralcodecs_b200 as nc
    sd = synthetic.make_dac_weights_hf(nc.DACConfig.DAC44kHz())
    cb = np.load(CODEBOOKS)
    for k in cb.files:
        sd[k] = cb[k]
    return sd


def dac44_weights_file(path=None):
    from neuralcodecs_b200 import synthetic
    path = path or os.path.join("/tmp", "dac44_exp.safetensors")
    if not os.path.exists(path):
        synthetic.save_safetensors(dac44_state_dict(), path)
    return path
