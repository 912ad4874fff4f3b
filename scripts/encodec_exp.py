import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import encodec as oenc, synth
import neuralcodecs_b200 as nc
lstm = int(sys.argv[1]) if len(sys.argv) > 1 else 0
co, ce = oenc.EncodecConfig(num_lstm_layers=lstm), nc.EncodecConfig(num_lstm_layers=lstm)
path = f"tests/.cache/encodec_{'24_seed4321' if lstm else 'nolstm'}.safetensors".replace("encodec_24", "encodec24")
o = oenc.load_safetensors(path, co)
x = synth.synth_audio(3, 60001, 24000, first_clip=9)[:, None, :]
xt = torch.from_numpy(x)
ref = o.forward(xt); emb = o.encode_latent(xt)
for prec in ["fp32", "bf16x3", "3xtf32", "f16x3", "tf32"]:
    m = nc.Encodec(ce, options={"precision": prec}); m.LoadWeights(path)
    (codes, _), = m.Encode(x)
    ct = torch.from_numpy(codes); cr = ref["codes"]
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
    dec = m.Decode([(cr.numpy(), None)]); dref = o.decode(cr).numpy()
    err = dec.astype(np.float64) - dref
    snr = 10 * np.log10((dref.astype(np.float64) ** 2).sum() / max((err ** 2).sum(), 1e-300))
    print(f"{prec:7s}: match {(codes == cr.numpy()).mean():.5f} uncascaded flips {len(margins)} max margin {max([abs(m_) for _, m_ in margins], default=0):.2e} "
          f"stages {[q for q, _ in margins][:10]} | dec max-abs {np.abs(err).max():.2e} snr {snr:.1f} dB")
    m.Dispose()
