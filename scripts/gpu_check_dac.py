"""Ad-hoc GPU check: engine vs oracle for DAC (tiny + full config, all precisions)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth, dac as odac
import neuralcodecs_b200 as nc

def snr_db(ref, test):
    ref = ref.astype(np.float64); test = test.astype(np.float64)
    return 10 * np.log10((ref ** 2).sum() / max(((ref - test) ** 2).sum(), 1e-300))

def run(cfg_o, cfg_e, seconds, opts, tag, codebooks="data", B=2):
    t0 = time.time()
    sd = synth.make_dac_weights_hf(cfg_o, codebooks=codebooks, codebook_seconds=min(10.0, max(seconds, 2.0)))
    path = os.path.join(tempfile.gettempdir(), f"dac_{tag}.safetensors")
    synth.save_safetensors(sd, path)
    o = odac.load_hf_safetensors(path, cfg_o)
    L = int(seconds * cfg_o.sample_rate) + 37
    x = synth.synth_audio(B, L, cfg_o.sample_rate, first_clip=3)
    ref = o.forward(torch.from_numpy(x).unsqueeze(1))
    print(f"[{tag}] oracle ready {time.time()-t0:.1f}s  T={ref['codes'].shape[-1]}")
    for opt in opts:
        m = nc.DAC(cfg_e, options=opt)
        m.LoadWeights(path)
        t1 = time.time()
        out = m.forward(x[:, None, :])
        dt = time.time() - t1
        codes_match = (out["codes"] == ref["codes"].numpy()).mean()
        frames_ok = (out["codes"] == ref["codes"].numpy()).all(axis=1).mean()
        zerr = np.abs(out["z"] - ref["z"].numpy()).max()
        a_ref = ref["audio"].numpy()
        # decoder-only: decode the oracle's z
        a_dec = m.Decode(ref["z"].numpy())
        print(f"[{tag}] {opt}: codes match {codes_match:.5f} frames all-equal {frames_ok:.5f} | z maxerr {zerr:.3e} | "
              f"audio(full) maxabs {np.abs(out['audio']-a_ref).max():.3e} snr {snr_db(a_ref, out['audio']):.1f} dB | "
              f"audio(dec-only) maxabs {np.abs(a_dec-a_ref).max():.3e} snr {snr_db(a_ref, a_dec):.1f} dB | "
              f"|audio|max {np.abs(a_ref).max():.3f} | {dt:.2f}s launches {m.launch_count()}")
        zc = m.FromCodes(ref["codes"].numpy())
        zo = o.from_codes(ref["codes"]).numpy()
        ac = m.DecodeCodes(ref["codes"].numpy())
        print(f"[{tag}]    from_codes maxerr {np.abs(zc-zo).max():.3e}; decode_codes snr {snr_db(o.decode(torch.from_numpy(zo)).numpy(), ac):.1f} dB")
        m.Dispose()

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("tiny", "all"):
        co = odac.DACConfig(sample_rate=16000, encoder_dim=16, decoder_dim=128, n_codebooks=4, codebook_size=64)
        ce = nc.DACConfig(sample_rate=16000, encoder_dim=16, decoder_dim=128, num_codebooks=4, codebook_size=64)
        run(co, ce, 0.5, [{"encoder_precision": "fp32", "decoder_precision": "fp32"}], "tiny")
    if which in ("mid", "all"):
        # channels multiples of 32 so the tcgen05 path is taken, but small
        co = odac.DACConfig(sample_rate=16000, encoder_dim=32, decoder_dim=512, n_codebooks=4, codebook_size=256)
        ce = nc.DACConfig(sample_rate=16000, encoder_dim=32, decoder_dim=512, num_codebooks=4, codebook_size=256)
        run(co, ce, 1.0, [{"encoder_precision": "fp32", "decoder_precision": "fp32"},
                          {"encoder_precision": "tf32", "decoder_precision": "tf32"},
                          {"encoder_precision": "3xtf32", "decoder_precision": "3xtf32"}], "mid")
    if which in ("full", "all"):
        co = odac.DACConfig.dac_44khz(); ce = nc.DACConfig.DAC44kHz()
        run(co, ce, 2.0, [{"encoder_precision": "fp32", "decoder_precision": "fp32"},
                          {"encoder_precision": "3xtf32", "decoder_precision": "tf32"},
                          {"encoder_precision": "tf32", "decoder_precision": "tf32"},
                          {"encoder_precision": "3xtf32", "decoder_precision": "3xtf32"}], "full", B=1)
