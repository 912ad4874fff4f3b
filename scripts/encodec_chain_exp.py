"""Encodec 24 kHz encoder on tensor cores with short accumulation chains: code flips vs the fp32 CPU oracle and encode time.
usage: python scripts/encodec_chain_exp.py [clips] [seconds]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import encodec as oenc, synth
import neuralcodecs_b200 as nc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
co, ce = oenc.EncodecConfig(), nc.EncodecConfig.Encodec24Khz()
path = "/tmp/encodec24_exp.safetensors"
if not os.path.exists(path):
    synth.save_safetensors(synth.make_encodec_weights(co, codebook_clips=2, codebook_seconds=6.0), path)
o = oenc.load_safetensors(path, co)
x = synth.synth_audio(B, int(S * 24000), 24000, first_clip=9)[:, None, :]
xt = torch.from_numpy(x)
t0 = time.time(); ref = o.forward(xt); emb = o.encode_latent(xt); print(f"oracle {time.time()-t0:.1f}s", flush=True)
cr = ref["codes"]
for tag, opts in (("fp32", {"encoder_precision": "fp32"}), ("3xtf32 long chains", {"encoder_short_chains": "0"}),
                  ("3xtf32 default policy", {}), ("3xtf32 fold every layer", {"encoder_short_chains": "2"}),
                  ("f16x3 default policy", {"encoder_precision": "f16x3"}), ("bf16x3 default policy", {"encoder_precision": "bf16x3"})):
    m = nc.Encodec(ce, options=opts); m.LoadWeights(path)
    m.Encode(x[:1])
    t0 = time.time(); (codes, _), = m.Encode(x); dt = time.time() - t0
    ct = torch.from_numpy(codes)
    margins = []
    with torch.inference_mode():
        residual = emb.clone(); tainted = torch.zeros(cr.shape[0], cr.shape[2], dtype=torch.bool)
        for q in range(cr.shape[1]):
            flat = residual.transpose(1, 2).reshape(-1, 128)
            dist = o.vq_distances(q, flat).reshape(cr.shape[0], cr.shape[2], -1)
            E = o.sd[f"quantizer.layers.{q}.codebook.embed"]
            for b, t in ((cr[:, q] != ct[:, q]) & ~tainted).nonzero().tolist():
                scale = float(residual[b, :, t].pow(2).sum() + E[ct[b, q, t]].pow(2).sum())
                margins.append((q, float(dist[b, t, ct[b, q, t]] - dist[b, t, cr[b, q, t]]) / scale))
            tainted |= cr[:, q] != ct[:, q]
            quant, _ = o.vq_forward(q, residual); residual = residual - quant
    big = [m_ for m_ in margins if abs(m_[1]) >= 1e-6]
    print(f"{tag:22s}: encode {dt*1e3:7.1f} ms | match {(codes == cr.numpy()).mean():.5f} | un-cascaded flips {len(margins)} (>=1e-6: {len(big)}; "
          f"max {max([abs(v) for _, v in margins], default=0):.2e}; stages of the big ones {[q for q, _ in big][:12]})", flush=True)
    m.Dispose()
