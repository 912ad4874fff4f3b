import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth, dac as odac
import neuralcodecs_b200 as nc
from scripts.gpu_check_dac import snr_db
co = odac.DACConfig.dac_44khz(); ce = nc.DACConfig.DAC44kHz()
sd = synth.make_dac_weights_hf(co, codebooks="data", codebook_seconds=2.0)
path = os.path.join(tempfile.gettempdir(), "dac_full.safetensors"); synth.save_safetensors(sd, path)
o = odac.load_hf_safetensors(path, co)
x = synth.synth_audio(1, 88237, 44100, first_clip=3)
ref = o.forward(torch.from_numpy(x).unsqueeze(1))
o64 = odac.load_hf_safetensors(path, co, dtype=torch.float64)
a64 = o64.decode(ref["z"].double()).numpy()
print("oracle fp32 vs fp64 decoder snr", snr_db(a64, ref["audio"].numpy()))
# emulate tf32 operand rounding in the oracle (weights + conv inputs), fp32 accumulate
def rna(t):
    u = t.contiguous().view(torch.int32); u = (u + 0x1000) & ~0x1FFF; return u.view(torch.float32)
import torch.nn.functional as F
class Emu(odac.DACOracle):
    def wnconv1d(self, name, x, stride=1, padding=0, dilation=1, groups=1):
        v, g = self.sd[name + ".weight_v"], self.sd[name + ".weight_g"]; b = self.sd.get(name + ".bias")
        w = (v / (v.pow(2).sum([1,2], keepdim=True).sqrt() + 1e-7)) * g
        return F.conv1d(rna(x), rna(w), b, stride, padding, dilation, groups)
    def wnconvtranspose1d(self, name, x, stride=1, padding=0, output_padding=0):
        v, g = self.sd[name + ".weight_v"], self.sd[name + ".weight_g"]; b = self.sd.get(name + ".bias")
        w = (v / (v.pow(2).sum([1,2], keepdim=True).sqrt() + 1e-7)) * g
        return F.conv_transpose1d(rna(x), rna(w), b, stride=stride, padding=padding, output_padding=output_padding)
emu = Emu(co, o.sd)
a_emu = emu.decode(ref["z"]).numpy()
print("emulated tf32-rounded-operand decoder (CPU) snr vs fp32 oracle", snr_db(ref["audio"].numpy(), a_emu))
for opt in [{"decoder_precision": "tf32", "fast_sin": "0"}, {"decoder_precision": "tf32", "fast_sin": "1"}, {"decoder_precision":"3xtf32","fast_sin":"1"}, {"decoder_precision":"3xtf32","fast_sin":"0"}]:
    m = nc.DAC(ce, options=dict(encoder_precision="fp32", profile="1", **opt)); m.LoadWeights(path)
    a = m.Decode(ref["z"].numpy())
    print(opt, "snr vs fp32 oracle", snr_db(ref["audio"].numpy(), a), "vs fp64", snr_db(a64, a), "vs emu", snr_db(a_emu, a))
    print({k: (v["launches"], round(v["ms"],3), round(v["flops"]/max(v["ms"],1e-9)/1e9,1)) for k, v in m.profile_report().items()})
    m.Dispose()
