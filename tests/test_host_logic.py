"""Host-side mirror of the reference interface: names, signatures, defaults, host tables, and the
fail-loud rule (no CPU fallback).  CPU only."""
import inspect

import numpy as np
import pytest
import torch

import awq_quantizer
import gptq_quantizer
import pot_apot_quantizer
import quantization_utils
import smooth_quant_quantizer

# the signatures benchmark_runner.py / test_quantization.py rely on
EXPECTED = {
    gptq_quantizer.gptq_quantize_model_weight:
        "(model, w_bit, q_group_size, input_feat, perp_damp=0.01, blocksize=128, nsamples=128, actorder=False, verbose=True)",
    gptq_quantizer._simple_quantize_layer: "(layer, n_bit, q_group_size)",
    gptq_quantizer._gptq_quantize_layer:
        "(layer, n_bit, q_group_size, input_feat, perp_damp=0.01, blocksize=128, nsamples=128, actorder=False, verbose=True)",
    gptq_quantizer.gptq_calibrate_hessian: "(model, calib_samples, nsamples=128, verbose=True)",
    awq_quantizer.awq_quantize_model_weight:
        "(model, w_bit, q_group_size, input_feat, protect_ratio=0.01, scale_factor=1.0)",
    awq_quantizer.awq_search_scale_factor:
        "(model, w_bit, q_group_size, input_feat, protect_ratio=0.01, scale_search_range=(1.0, 2.0), n_grid=20)",
    pot_apot_quantizer.pot_quantize_tensor: "(w, n_bit=4, q_group_size=-1)",
    pot_apot_quantizer.pot_quantize_model_weight: "(model, w_bit, q_group_size)",
    pot_apot_quantizer.apot_quantize_tensor: "(w, n_bit=4, q_group_size=-1, k=2)",
    pot_apot_quantizer.apot_quantize_model_weight: "(model, w_bit, q_group_size, k=2)",
    smooth_quant_quantizer.collect_act_scales: "(model, calib_samples, verbose=True)",
    smooth_quant_quantizer.smooth_weights: "(model, act_scales, alpha=0.5, verbose=True)",
    smooth_quant_quantizer.smooth_activations: "(model, calib_samples, alpha=0.5, verbose=True)",
    smooth_quant_quantizer.reverse_weight_smoothing: "(model, verbose=True)",
    smooth_quant_quantizer.smoothquant_quantize_model_weight:
        "(model, w_bit, q_group_size, act_scales, alpha=0.5, verbose=True)",
    smooth_quant_quantizer.smoothquant_search_alpha:
        "(model, calib_samples, act_scales, w_bit=8, q_group_size=-1, alpha_range=(0.0, 1.0), n_grid=20, verbose=True)",
    smooth_quant_quantizer.smoothquant_quantize_and_calibrate:
        "(model, w_bit, q_group_size, calib_samples, alpha=None, search_alpha=False, verbose=True)",
    quantization_utils.pseudo_quantize_tensor: "(w, n_bit=4, q_group_size=-1)",
    quantization_utils.get_calib_feat: "(model, tokenizer, calib_samples, verbose=True)",
    quantization_utils.get_model_size: "(model, data_width=16, group_size=-1, use_zero_point=True)",
}


def _sig(fn) -> str:
    s = inspect.signature(fn)
    parts = []
    for p in s.parameters.values():
        parts.append(p.name if p.default is inspect.Parameter.empty else f"{p.name}={p.default!r}")
    return "(" + ", ".join(parts) + ")"


def test_entry_points_keep_the_reference_signatures():
    for fn, want in EXPECTED.items():
        assert _sig(fn) == want, fn.__name__


def test_generate_apot_levels_device_kwarg_and_values():
    s = inspect.signature(pot_apot_quantizer.generate_apot_levels)
    assert list(s.parameters) == ["n", "k", "device"]
    lv = pot_apot_quantizer.generate_apot_levels(2, 2)
    want = [0, 1 / 32, 1 / 16, 3 / 32, 1 / 8, 3 / 16, 1 / 4, 9 / 32, 3 / 8, 1 / 2, 9 / 16, 3 / 4, 1,
            33 / 32, 9 / 8, 3 / 2]
    assert lv.dtype == torch.float32 and lv.tolist() == want
    assert pot_apot_quantizer.generate_apot_levels(1, 2).tolist() == [0.0, 0.25, 0.5, 1.0]


def test_signed_level_set_matches_oracle(golden):
    from oracle import quant_oracle as O
    for b, k in ((4, 2), (8, 2), (2, 1), (6, 3), (3, 2)):
        mine = pot_apot_quantizer._apot_signed_levels(b, k)
        assert torch.equal(mine, O.apot_level_set(b, k))
        assert mine.numel() <= 32 and torch.all(mine[1:] > mine[:-1])


def test_units_and_model_size():
    import torch.nn as nn
    assert (quantization_utils.Byte, quantization_utils.KiB) == (8, 8192)
    assert quantization_utils.GiB == 1024 * quantization_utils.MiB
    net = nn.Sequential(nn.Linear(10, 20), nn.Linear(20, 5))
    n = sum(p.numel() for p in net.parameters())
    assert quantization_utils.get_model_size(net, 16) == n * 16
    assert quantization_utils.get_model_size(net, 4, 128) == n * (4 + 16 / 128 + 4 / 128)
    assert quantization_utils.get_model_size(net, 4, 128, use_zero_point=False) == n * (4 + 16 / 128)
    assert [n for n, _ in quantization_utils.get_linear_layers(net)] == ["0", "1"]


def test_config_roundtrip(tmp_path):
    cfg = {"model_name": "x", "quantization_config": {"awq": {"w_bit": 4}}}
    p = tmp_path / "c.json"
    quantization_utils.save_config(cfg, str(p))
    assert quantization_utils.load_config(str(p)) == cfg


def test_out_of_scope_helpers_forward_to_a_reference_checkout(monkeypatch, tmp_path):
    monkeypatch.setenv("LLMQ_REFERENCE_DIR", str(tmp_path))          # empty dir: nothing to forward to
    monkeypatch.setattr(quantization_utils, "_ref_module", None)
    with pytest.raises(ImportError, match="reference checkout"):
        quantization_utils.evaluate_perplexity
    with pytest.raises(AttributeError):
        quantization_utils.no_such_name


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_gpu():
    """Without a CUDA device the entry points must raise, not compute on the CPU."""
    import torch.nn as nn
    w = torch.randn(4, 128)
    for call in (lambda: quantization_utils.pseudo_quantize_tensor(w, 4, 128),
                 lambda: pot_apot_quantizer.pot_quantize_tensor(w, 4, 128),
                 lambda: pot_apot_quantizer.apot_quantize_tensor(w, 4, 128),
                 lambda: gptq_quantizer._simple_quantize_layer(nn.Linear(128, 4), 4, 128),
                 lambda: awq_quantizer.awq_quantize_model_weight(
                     nn.Sequential(nn.Linear(128, 4)), 4, 128, {"0": [torch.rand(128)]}),
                 lambda: smooth_quant_quantizer.smooth_weights(
                     nn.Sequential(nn.Linear(128, 4)), {"0": torch.rand(128)}, verbose=False)):
        with pytest.raises(RuntimeError, match="no CUDA device"):
            call()


def test_shape_asserts_fire_before_any_device_work():
    with pytest.raises(AssertionError):
        quantization_utils.pseudo_quantize_tensor(torch.randn(4, 100), n_bit=4, q_group_size=32)
    with pytest.raises(AssertionError):
        pot_apot_quantizer.pot_quantize_tensor(torch.randn(2, 3, 8), 4, -1) if torch.cuda.is_available() \
            else pot_apot_quantizer._as_groups(torch.randn(2, 3, 8), -1)
    with pytest.raises(AssertionError):
        pot_apot_quantizer._as_groups(torch.randn(4, 100), 32)


def test_product_modules_do_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under llm-quantization_b200/ may reference it."""
    from pathlib import Path
    pkg = Path(awq_quantizer.__file__).resolve().parent
    for path in list(pkg.glob("*.py")) + list((pkg / "b200q").glob("*.py")):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path


def test_gptq_group_dealing_balances_the_inverse_work():
    """Row-sharded GPTQ deals the layers of a group to the ranks by K^3, longest first: with the
    Llama-2-7B mix (six 4096-wide and one 11008-wide Linear per block) no rank may end up with
    more than ~one 11008-wide layer over the mean."""
    import gptq_quantizer as gq
    block = [("q", 4096), ("k", 4096), ("v", 4096), ("o", 4096), ("gate", 4096), ("up", 4096),
             ("down", 11008)]
    layers = [(f"{i}.{n}", k) for i in range(5) for n, k in block][:32]
    owner = gq._deal_layers(layers, 8)
    assert set(owner) == {n for n, _ in layers} and set(owner.values()) <= set(range(8))
    load = [0.0] * 8
    for n, k in layers:
        load[owner[n]] += float(k) ** 3
    assert max(load) <= sum(load) / 8 + 11008.0 ** 3
    assert owner == gq._deal_layers(list(reversed(layers)), 8), "assignment must not depend on order"
    assert gq._deal_layers(layers, 1) == {n: 0 for n, _ in layers}


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs on the host only (the staged reference modules, or the oracle
    port where they are absent) and must print exactly one
    JSON line carrying the contract's keys -- checked here on the tiny model so it takes seconds."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    repo = Path(__file__).resolve().parent.parent
    res = subprocess.run([sys.executable, str(repo / "bench.py"), "--impl", "reference", "--model", "tiny",
                          "--method", "awq_fixed", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=str(repo))
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    for key in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert key in d, key
    assert d["vs_baseline"] is None and d["higher_is_better"] is True
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    # "reference" when the unmodified reference modules are staged in baseline/_ref (they are wherever
    # __graft_entry__.build() ran with /root/reference present), "port" (the oracle) otherwise
    staged = (repo / "baseline" / "_ref" / "awq_quantizer.py").exists()
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["ms_per_step"] < 60e3 and d["config"]["rows"] > 0
    assert "workload" in d["config"]


def test_gptq_layer_dealing_is_deterministic_and_balanced():
    """Row-sharded GPTQ deals the layers of a group to the ranks longest-first by K^3
    (gptq_quantizer._deal_layers): every rank derives the same assignment without talking."""
    import gptq_quantizer as gq
    layers = [(f"l{i}", 4096) for i in range(12)] + [("d0", 11008), ("d1", 11008)]
    a = gq._deal_layers(layers, 4)
    b = gq._deal_layers(list(reversed(layers)), 4)
    assert a == b and set(a.values()) == {0, 1, 2, 3}
    load = [0.0] * 4
    for n, k in layers:
        load[a[n]] += float(k) ** 3
    assert max(load) / min(load) < 4.0
    assert a["d0"] != a["d1"]                      # the two wide layers never share a rank
    assert gq._deal_layers(layers, 1) == {n: 0 for n, _ in layers}


def test_bench_sample_slices_and_syrk_fraction_are_fixed_functions():
    """The CPU arm's row slices are a function of the layer shape only (both arms and every step
    use the same ones), and the roofline's executed-MMA fraction is the SYRK tile count."""
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("_bench", Path(__file__).resolve().parent.parent / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.cpu_rows("awq", 4096, 4096) == bench.cpu_rows("awq", 4096, 4096) == 512
    assert bench.cpu_rows("gptq", 4096, 11008) == 704 and bench.cpu_rows("gptq", 768, 768) == 768
    assert bench.cpu_rows("pot", 32000, 4096) == 128
    assert bench.executed_fraction_syrk(4096) == 272 / 512
    assert abs(bench.executed_fraction_syrk(11008) - 1892 / 3698) < 1e-12
    assert [n for n, *_ in bench.MODELS["llama2-7b-block"]] == ["attn.qkvo", "mlp.gate_up", "mlp.down"]


def test_collective_helpers_are_identities_outside_row_sharding():
    from b200q import dist as D
    t = torch.ones(3)
    assert D.allreduce_sum(t) is t and D.allreduce_max(t) is t and D.broadcast(t, 0) is t
    assert D.allreduce_sum_async(t) is None and not D.is_sharded() and D.world_size() == 1
    with D.timed_wait(10), D.on_comm_stream():     # no-ops without a recorder / device
        pass
    assert D.WAIT_EVENTS is None
