"""Codes / latents / z of one DAC Encode, saved for a bit-exact comparison between RVQ kernels (env NC_RVQ_BLOCK=0|1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import neuralcodecs_b200 as nc
from neuralcodecs_b200 import synthetic
from scripts.exp_common import dac44_weights_file
m = nc.DAC(nc.DACConfig.DAC44kHz())
m.LoadWeights(dac44_weights_file())
x = synthetic.synth_audio(3, 5 * 44100 + 123, 44100, first_clip=5)
z, codes, lat = m.Encode(x[:, None, :])
np.savez(sys.argv[1], z=z, codes=codes, lat=lat)
print("saved", sys.argv[1], codes.shape)
