"""Weight-file readers behind nc_load_weights / nc_inspect_weights (host only: no GPU needed): the torch.save zip
checkpoint of the official DAC `.pth` weights (reference: Config/DAC/DACUnpickler.cs:341-424) and safetensors."""
import collections

import numpy as np
import pytest
import torch

import neuralcodecs_b200 as nc


def _checksum(t: torch.Tensor) -> float:
    x = t.detach().contiguous().reshape(-1).to(torch.float64).numpy() if t.dtype.is_floating_point else \
        t.contiguous().reshape(-1).numpy().astype(np.float64)
    if t.dtype.is_floating_point:     # the reader converts every floating tensor to fp32 first
        x = t.detach().contiguous().reshape(-1).to(torch.float32).to(torch.float64).numpy()
    w = (np.arange(x.size) % 7 + 1).astype(np.float64)
    return float((w * x).sum())


def _descript_checkpoint():
    g = torch.Generator().manual_seed(3)
    r = lambda *s: torch.randn(*s, generator=g)
    sd = collections.OrderedDict()
    sd["encoder.block.0.weight_g"] = r(8, 1, 1)
    sd["encoder.block.0.weight_v"] = r(8, 1, 7)
    sd["encoder.block.0.bias"] = r(8)
    sd["encoder.block.1.block.0.block.0.alpha"] = r(1, 8, 1)
    sd["decoder.model.1.block.2.block.3.weight_v"] = r(8, 8, 1)
    sd["quantizer.quantizers.0.codebook.weight"] = r(16, 4).half()
    sd["bf"] = r(5, 3).bfloat16()
    sd["dbl"] = r(4).double()
    sd["strided"] = torch.arange(12, dtype=torch.float32).reshape(3, 4).t()       # non-contiguous view of its storage
    sd["offset_view"] = torch.arange(20, dtype=torch.float32)[5:11].reshape(2, 3)  # storage offset
    sd["ints"] = torch.arange(5) - 2
    sd["bools"] = torch.tensor([True, False, True])
    sd._metadata = {"": {"version": 1}}                                            # what Module.state_dict() attaches
    meta = {"kwargs": {"encoder_dim": 8, "encoder_rates": [2, 4], "decoder_dim": 64, "decoder_rates": [4, 2], "n_codebooks": 2,
                       "codebook_size": 16, "codebook_dim": 4, "sample_rate": 16000, "latent_dim": None, "quantizer_dropout": 0.5},
            "converted_from": "test", "flag": True}
    return {"state_dict": sd, "metadata": meta}


def test_torch_zip_checkpoint_tensors_metadata_and_values(tmp_path):
    ck = _descript_checkpoint()
    path = str(tmp_path / "weights.pth")
    torch.save(ck, path)
    info = nc.inspect_weights(path)
    assert info["format"] == "torch_zip"
    assert info["metadata"]["kwargs"]["encoder_rates"] == [2, 4] and info["metadata"]["kwargs"]["latent_dim"] is None
    assert info["metadata"]["flag"] is True and info["metadata"]["kwargs"]["quantizer_dropout"] == 0.5
    assert set(info["tensors"]) == set(ck["state_dict"])
    for name, t in ck["state_dict"].items():
        got = info["tensors"][name]
        assert got["shape"] == list(t.shape), name
        assert got["dtype"] == ("float32" if t.dtype.is_floating_point else "int64"), name
        assert got["checksum"] == pytest.approx(_checksum(t), rel=1e-12, abs=1e-12), name
    cfg = nc.DACConfig.FromWeights(path)                       # DACUnpickler.CreateConfigFromMetadata
    assert (cfg.sample_rate, cfg.encoder_dim, cfg.encoder_rates, cfg.decoder_dim, cfg.decoder_rates, cfg.num_codebooks,
            cfg.codebook_size, cfg.codebook_dim, cfg.latent_dim) == (16000, 8, [2, 4], 64, [4, 2], 2, 16, 4, None)


def test_bare_state_dict_and_legacy_and_corrupt_files(tmp_path):
    sd = {"a.weight": torch.ones(2, 3), "b": torch.arange(4, dtype=torch.int32)}
    p1 = str(tmp_path / "pytorch_model.bin")
    torch.save(sd, p1)                                          # bare state dict (SNAC / HF style)
    info = nc.inspect_weights(p1)
    assert info["metadata"] == {} and info["tensors"]["a.weight"]["checksum"] == pytest.approx(1 + 2 + 3 + 4 + 5 + 6)
    assert info["tensors"]["b"]["dtype"] == "int64" and info["tensors"]["b"]["shape"] == [4]
    p2 = str(tmp_path / "legacy.pth")
    torch.save(sd, p2, _use_new_zipfile_serialization=False)    # pre-1.6 format: the reference rejects it too
    with pytest.raises(RuntimeError):
        nc.inspect_weights(p2)
    p3 = str(tmp_path / "truncated.pth")
    data = open(p1, "rb").read()
    open(p3, "wb").write(data[: len(data) // 2])
    with pytest.raises(RuntimeError):
        nc.inspect_weights(p3)
    with pytest.raises(FileNotFoundError):
        nc.inspect_weights(str(tmp_path / "missing.pth"))


def test_safetensors_inspection_matches(tmp_path):
    from oracle import synth
    sd = {"x.weight": torch.arange(6, dtype=torch.float32).reshape(2, 3), "codes": torch.arange(3)}
    path = str(tmp_path / "w.safetensors")
    synth.save_safetensors(sd, path)
    info = nc.inspect_weights(path)
    assert info["format"] == "safetensors" and info["tensors"]["x.weight"]["shape"] == [2, 3]
    assert info["tensors"]["x.weight"]["checksum"] == pytest.approx(_checksum(sd["x.weight"]))
