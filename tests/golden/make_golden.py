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


def encodec48_small():
    """transformers.EncodecModel with a shrunk 48 kHz-style config (stereo, non-causal, time_group_norm, normalize, segments):
    per-frame encode / decode of a full segment and of a shorter last segment, and the triangular overlap-add.  HF's own
    `encode` loop differs from the reference's (it requires pre-padded input and never emits a short last frame), so the
    pin is at the frame level; the segment loop is the reference's (Models/Encodec.cs:273-282), tested by hand in
    tests/test_oracle_snac_encodec.py."""
    from transformers import EncodecConfig as HFC, EncodecModel
    torch.manual_seed(11)
    hf = EncodecModel(HFC(num_filters=8, hidden_size=32, codebook_size=64, codebook_dim=32, upsampling_ratios=[4, 3, 2],
                          target_bandwidths=[24.0, 48.0, 96.0], sampling_rate=48000, audio_channels=2, normalize=True,
                          chunk_length_s=0.05, overlap=0.01, norm_type="time_group_norm", use_causal_conv=False)).eval()
    with torch.no_grad():                       # GroupNorm affine away from (1, 0) so the test sees it
        for n, p_ in hf.named_parameters():
            if ".norm." in n:
                p_.add_(0.2 * torch.randn_like(p_))
    sd = {k: v.detach().clone().numpy() for k, v in hf.state_dict().items()}
    seg = int(0.05 * 48000)                                            # 2400 samples = 100 frames of hop 24
    x = torch.from_numpy(synth.synth_audio(4, seg + 1013, 48000)).reshape(2, 2, -1) * torch.tensor([0.3, 1.7]).view(2, 1, 1)
    out = {}
    with torch.inference_mode():
        for name, fr in (("full", x[..., :seg]), ("tail", x[..., seg:])):
            codes, scale = hf._encode_frame(fr, 48.0)
            emb = hf.encoder(fr / scale.view(-1, 1, 1))
            audio = hf._decode_frame(codes, scale)
            out.update({f"{name}_codes": codes.numpy(), f"{name}_scale": scale.numpy(), f"{name}_emb": emb.numpy(),
                        f"{name}_audio": audio.numpy()})
        ola = hf._linear_overlap_add([torch.from_numpy(out["full_audio"]), torch.from_numpy(out["full_audio"]) * 0.5,
                                      torch.from_numpy(out["tail_audio"])], hf.config.chunk_stride)
    np.savez_compressed(os.path.join(OUT, "encodec48_hf_small.npz"), audio_in=x.numpy(), ola=ola.numpy(),
                        stride=np.int64(hf.config.chunk_stride), **out, **{"w/" + k: v for k, v in sd.items()})
    print("encodec48_hf_small.npz", {k: v.shape for k, v in out.items()}, ola.shape, hf.config.chunk_stride)


if __name__ == "__main__" and "--encodec48" in sys.argv:
    encodec48_small()
