"""Generates tests/golden/*.npz.  Run from the repo root:  python tests/golden/make_golden.py

The reference (C#/TorchSharp) cannot run in this environment and ships no fixtures, so the
oracle is pinned against `transformers` DacModel -- the same conv graph, paddings, stride
algebra and safetensors key layout the reference's DAC loads.  HF differs from
the reference ONLY in the VQ distance (HF L2-normalises, the reference does not) and in
sub-fp32-noise epsilons, so encoder / decoder / from_codes outputs are comparable and the
quantiser decisions are NOT taken from HF.

Weights come from oracle.synth (seeded, numpy PCG64) so the test can regenerate them.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dac as odac  # noqa: E402
from oracle import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def dac_small():
    from transformers import DacConfig as HFC, DacModel
    cfg = odac.DACConfig(sample_rate=16000, encoder_dim=16, decoder_dim=128, n_codebooks=4, codebook_size=64)
    sd = synth.make_dac_weights_hf(cfg, codebooks="normal")
    hf = DacModel(HFC(encoder_hidden_size=16, decoder_hidden_size=128, n_codebooks=4, codebook_size=64,
                      hidden_size=cfg.latent_dim, sampling_rate=16000)).eval()
    hf.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    x = torch.from_numpy(synth.synth_audio(2, 5000, 16000)).unsqueeze(1)
    rng = np.random.default_rng(5)
    with torch.inference_mode():
        xp = torch.nn.functional.pad(x, [0, (-x.shape[-1]) % cfg.hop_length])
        z = hf.encoder(xp)
        audio = hf.decoder(z)
        codes = torch.from_numpy(rng.integers(0, 64, size=(2, 4, z.shape[-1]), dtype=np.int64))
        zq = hf.quantizer.from_codes(codes)[0]
    np.savez_compressed(os.path.join(OUT, "dac_hf_small.npz"), audio_in=x.numpy(), encoder_out=z.numpy(),
                        decoder_out=audio.numpy(), codes=codes.numpy(), from_codes=zq.numpy())
    print("dac_hf_small.npz", z.shape, audio.shape)


if __name__ == "__main__":
    dac_small()


def encodec_small():
    """transformers.EncodecModel (24 kHz-style causal weight-norm config, shrunk) -> encoder / decoder / codes fixtures.
    HF's Encodec and the reference agree on everything on this path (conv stacks, causal reflect padding, LSTM skip,
    un-normalised Euclidean VQ), so -- unlike DAC -- the code decisions are pinned too."""
    from transformers import EncodecConfig as HFC, EncodecModel
    from oracle import encodec as oenc
    torch.manual_seed(7)
    hf = EncodecModel(HFC(num_filters=8, hidden_size=32, codebook_size=64, codebook_dim=32, upsampling_ratios=[4, 3, 2],
                          target_bandwidths=[1.5, 3.0, 6.0])).eval()
    sd = {k.replace(".parametrizations.weight.original0", ".weight_g").replace(".parametrizations.weight.original1", ".weight_v"):
          v.detach().clone().numpy() for k, v in hf.state_dict().items()}
    x = torch.from_numpy(synth.synth_audio(2, 2503, 24000)).unsqueeze(1)
    with torch.inference_mode():
        emb = hf.encoder(x)
        codes = hf.quantizer.encode(emb, 3.0).transpose(0, 1).contiguous()     # [B, nq, T]
        audio = hf.decoder(hf.quantizer.decode(codes.transpose(0, 1)))
    np.savez_compressed(os.path.join(OUT, "encodec_hf_small.npz"), audio_in=x.numpy(), encoder_out=emb.numpy(),
                        codes=codes.numpy(), decoder_out=audio.numpy(), **{"w/" + k: v for k, v in sd.items()})
    print("encodec_hf_small.npz", emb.shape, codes.shape, audio.shape)


if __name__ == "__main__" and "--encodec" in sys.argv:
    encodec_small()
