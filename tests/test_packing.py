"""Packed export layout: the oracle's bit-stream packer (CPU) and the CUDA kernels against it."""
import numpy as np
import pytest
import torch

from oracle import quant_oracle as O


def test_oracle_layout_known_vectors():
    # 4 bits: eight codes per word, lowest nibble first
    assert O.pack_codes(np.array([[1, 2, 3, 4, 5, 6, 7, 8]], dtype=np.uint8), 4)[0, 0] == 0x87654321
    # 3 bits: code 10 straddles words 0 and 1 (bits 30..32)
    c = np.zeros((1, 32), dtype=np.uint8)
    c[0, 10] = 0b101
    w = O.pack_codes(c, 3)
    assert w.shape == (1, 3) and w[0, 0] == (0b01 << 30) and w[0, 1] == 0b1 and w[0, 2] == 0
    rng = np.random.default_rng(0)
    for b in range(1, 9):
        for K in (1, 37, 100, 128):
            codes = rng.integers(0, 1 << b, size=(5, K), dtype=np.uint8)
            assert np.array_equal(O.unpack_codes(O.pack_codes(codes, b), K, b), codes)


@pytest.mark.gpu
@pytest.mark.parametrize("b", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("N,K", [(3, 37), (64, 128), (5, 100), (17, 4096)])
def test_kernels_match_oracle(b, N, K):
    from b200q import export as E
    rng = np.random.default_rng(b * 1000 + K)
    codes = rng.integers(0, 1 << b, size=(N, K), dtype=np.uint8)
    d = torch.from_numpy(codes).cuda()
    packed = E.pack_codes(d, b)
    want = O.pack_codes(codes, b)
    assert np.array_equal(packed.cpu().numpy().view(np.uint32), want)
    assert torch.equal(E.unpack_codes(packed, K, b), d)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("b,G", [(4, 128), (3, 128), (8, 64), (4, -1)])
def test_uniform_export_round_trips_to_the_fake_quantized_weight(dtype, b, G):
    from b200q import export as E, ops
    g = torch.Generator().manual_seed(b + 7)
    W = (torch.randn(96, 512, generator=g) * 0.05).to(dtype).cuda()
    rec = E.export_uniform(W, b, G)
    assert rec["qweight"].shape == (96, 512 * b // 32) and rec["qweight"].dtype == torch.int32
    assert torch.equal(E.dequantize(rec), ops.group_fakequant(W, b, G))
    # the codes are the oracle's integers
    want = O.uniform_group_quant(W.cpu(), b, G)
    codes = E.unpack_codes(rec["qweight"], 512, b).cpu()
    assert torch.equal(codes.to(torch.int32), want["codes"].reshape(96, 512))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_gptq_export_round_trips(dtype):
    from b200q import export as E, ops
    g = torch.Generator().manual_seed(11)
    W = (torch.randn(200, 384, generator=g) * 0.05).to(dtype).cuda()
    rec = E.export_gptq_parity(W, 4)
    assert rec["bits"] == 5
    assert torch.equal(E.dequantize(rec), ops.gptq_parity_quant(W, 4))


def test_records_file_round_trip(tmp_path):
    """save_records / load_records keep every field of a record (host-only: no kernels involved)."""
    from b200q import export as E
    rec = {"scheme": "uniform_asym", "bits": 4, "group": 128, "shape": (8, 256), "dtype": "float16",
           "qweight": torch.randint(-2 ** 31, 2 ** 31 - 1, (8, 32), dtype=torch.int32),
           "scales": torch.rand(16), "zeros": torch.randint(0, 16, (16,)).float()}
    path = tmp_path / "model.b200q"
    E.save_records(path, {"layers.0.q_proj": rec})
    back = E.load_records(path)
    assert set(back) == {"layers.0.q_proj"}
    got = back["layers.0.q_proj"]
    assert got["scheme"] == "uniform_asym" and got["bits"] == 4 and tuple(got["shape"]) == (8, 256)
    for k in ("qweight", "scales", "zeros"):
        assert torch.equal(got[k], rec[k])
    (tmp_path / "other.pt").write_bytes(b"")
    torch.save({"format": "something else"}, tmp_path / "other.pt")
    with pytest.raises(ValueError):
        E.load_records(tmp_path / "other.pt")


@pytest.mark.gpu
def test_export_model_file_round_trip_dequantizes_to_the_quantizer_output(tmp_path):
    import torch.nn as nn
    import copy
    from b200q import export as E
    from quantization_utils import pseudo_quantize_tensor
    torch.manual_seed(3)
    net = nn.Sequential(nn.Linear(256, 128, bias=False), nn.ReLU(), nn.Linear(128, 384, bias=False))
    records = E.export_model(net, 4, 128)                   # host-resident weights
    assert set(records) == {"0", "2"}
    E.save_records(tmp_path / "net.b200q", records)
    loaded = E.load_records(tmp_path / "net.b200q", device="cuda")
    for name, lin in (("0", net[0]), ("2", net[2])):
        want = pseudo_quantize_tensor(lin.weight.data.clone(), 4, 128)
        assert torch.equal(E.dequantize(loaded[name]).cpu(), want.cpu())


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("b,G", [(4, 128), (3, 64), (4, -1)])
def test_pot_export_round_trips_and_codes_are_the_oracles(dtype, b, G):
    """POT exponent + sign plane.  dequantize(record) is pot_quantize_tensor's output
    bit for bit, and the exponents are the oracle's (pot_apot_quantizer.py:104-107)."""
    from b200q import export as E
    from pot_apot_quantizer import pot_quantize_tensor
    g = torch.Generator().manual_seed(b * 13 + 1)
    W = (torch.randn(48, 512, generator=g) * 0.05).to(dtype)
    W[3, 7] = 0.0                                       # sign(0) = 0: the extra "zero" symbol
    W[10] = 0.0                                         # an all-zero group
    rec = E.export_pot(W.cuda(), b, G)
    assert rec["qweight"].shape == (48, 512 * b // 32) and "zero_mask" in rec
    got = E.dequantize(rec)
    assert torch.equal(got, pot_quantize_tensor(W.cuda(), b, G))
    want = O.pot_quant(W, b, G)
    assert torch.equal(got.cpu(), want["out"])
    codes = E.unpack_codes(rec["qweight"], 512, b).cpu().to(torch.int32)
    nz = want["out"] != 0
    assert torch.equal((codes & ((1 << (b - 1)) - 1))[nz], want["exps"].reshape(48, 512).to(torch.int32)[nz])
    assert torch.equal((codes >> (b - 1)).bool()[nz], (want["out"] < 0)[nz])
    # without zeros no mask is stored
    assert "zero_mask" not in E.export_pot((torch.rand(8, 128, generator=g) + 0.1).to(dtype).cuda(), b, G)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("b,k,G", [(4, 2, 128), (8, 2, 128), (2, 1, 64), (4, 2, -1)])
def test_apot_export_round_trips_and_codes_are_the_oracles(dtype, b, k, G):
    """APOT level-index plane (pot_apot_quantizer.py:294-298)."""
    from b200q import export as E
    from pot_apot_quantizer import apot_quantize_tensor
    g = torch.Generator().manual_seed(b * 17 + k)
    W = (torch.randn(40, 512, generator=g) * 0.05).to(dtype)
    rec = E.export_apot(W.cuda(), b, G, k)
    assert rec["levels"].numel() <= 32 and rec["bits"] <= 5
    got = E.dequantize(rec)
    assert torch.equal(got, apot_quantize_tensor(W.cuda(), b, G, k))
    want = O.apot_quant(W, b, G, k)
    assert torch.equal(got.cpu(), want["out"])
    codes = E.unpack_codes(rec["qweight"], 512, rec["bits"]).cpu().to(torch.int32)
    assert torch.equal(codes, want["level_idx"].reshape(40, 512).to(torch.int32))


@pytest.mark.gpu
def test_export_model_pot_apot_files(tmp_path):
    import torch.nn as nn
    from b200q import export as E
    torch.manual_seed(5)
    net = nn.Sequential(nn.Linear(256, 64, bias=False), nn.Linear(128, 32, bias=False))
    for scheme in ("pot", "apot"):
        recs = E.export_model(net, 4, 128, scheme)
        E.save_records(tmp_path / f"{scheme}.pt", recs)
        back = E.load_records(tmp_path / f"{scheme}.pt", device="cuda")
        for name, rec in recs.items():
            assert torch.equal(E.dequantize(back[name]), E.dequantize(rec))
