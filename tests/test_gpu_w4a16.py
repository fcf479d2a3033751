"""W4A16 GEMM with the dequantisation fused into the tcgen05 operand pipeline (SURVEY 8f item 4:
the Linear of the reference's perplexity loop, quantization_utils.py:269-322, on the packed export)
against `x @ dequantize(record).T` in fp32."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def rel(got, want):
    return ((got.double() - want.double()).abs().max() / want.double().abs().max()).item()


@pytest.mark.parametrize("act", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K,G", [(128, 128, 256, 128), (300, 200, 512, 128), (1, 4096, 4096, 128),
                                     (2048, 1024, 4096, 128), (77, 264, 320, 64), (33, 128, 1024, -1),
                                     (512, 384, 768, 32)])
def test_matches_dequantized_reference(act, M, N, K, G):
    from b200q import export as E, qlinear as Q
    g = torch.Generator().manual_seed(M + N + K)
    W = (torch.randn(N, K, generator=g) * 0.03).to(act).cuda()        # weight quantised in the act dtype
    x = torch.randn(M, K, generator=g).to(act).cuda()
    rec = E.export_uniform(W, 4, G)
    Wd = E.dequantize(rec)                                             # values of dtype `act`
    want = x.float() @ Wd.float().T
    got = Q.w4a16_linear(x, rec, out_dtype=torch.float32)
    # same 16-bit operands, fp32 accumulation: only the summation order differs
    assert rel(got, want) < 2e-5, rel(got, want)
    y16 = Q.w4a16_linear(x, rec)
    assert y16.dtype == act and rel(y16.float(), want) < (2e-3 if act == torch.float16 else 1.6e-2)


@pytest.mark.parametrize("act", [torch.float16, torch.bfloat16])
def test_fp32_records_are_rounded_once_to_the_activation_dtype(act):
    from b200q import export as E, qlinear as Q
    g = torch.Generator().manual_seed(3)
    W = (torch.randn(256, 512, generator=g) * 0.03).cuda()            # fp32 weight -> fp32 scales
    x = torch.randn(64, 512, generator=g).to(act).cuda()
    rec = E.export_uniform(W, 4, 128)
    want = x.float() @ E.dequantize(rec).to(act).float().T
    got = Q.w4a16_linear(x, rec, out_dtype=torch.float32)
    assert rel(got, want) < 2e-5, rel(got, want)


def test_packed_model_forward_equals_fake_quantized_model():
    """pack_model on an MLP: the same function as the model fake-quantized by pseudo_quantize_tensor
    (quantization_utils.py:362-413) up to fp16 accumulation order, at a quarter of the weight bytes."""
    import copy
    from quantization_utils import pseudo_quantize_tensor
    from b200q import export as E, qlinear as Q
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(512, 1024, bias=True), nn.GELU(), nn.Linear(1024, 256, bias=False)).cuda().half()
    fake = copy.deepcopy(net)
    for m in fake:
        if isinstance(m, nn.Linear):
            m.weight.data = pseudo_quantize_tensor(m.weight.data, 4, 128)
    x = torch.randn(96, 512, device="cuda", dtype=torch.float16)
    want = fake(x).float()
    packed = Q.pack_model(net, 128)
    assert isinstance(packed[0], Q.QuantLinear) and isinstance(packed[2], Q.QuantLinear)
    assert torch.equal(E.dequantize(packed[0].record()), fake[0].weight.data)
    assert torch.equal(E.dequantize(packed[2].record()), fake[2].weight.data)
    got = packed(x).float()
    assert rel(got, want) < 5e-3
    n_packed = sum(b.numel() * b.element_size() for b in packed.buffers())
    assert n_packed < 0.35 * sum(m.weight.numel() * 2 for m in fake if isinstance(m, nn.Linear))


def test_smoothquant_hook_survives_packing():
    """smooth_weights installs a forward-pre-hook that multiplies the inputs by s
    (smooth_quant_quantizer.py:178-199): the packed module must keep it."""
    import copy
    import smooth_quant_quantizer as sq
    from quantization_utils import pseudo_quantize_tensor
    from b200q import qlinear as Q
    torch.manual_seed(1)
    net = nn.Sequential(nn.Linear(256, 512, bias=False)).cuda().half()
    act = (torch.rand(256) * 4 + 0.1).half()
    sq.smooth_weights(net, {"0": act}, alpha=0.5, verbose=False)
    fake = copy.deepcopy(net)
    fake[0].weight.data = pseudo_quantize_tensor(fake[0].weight.data, 4, 128)
    x = torch.randn(40, 256, device="cuda", dtype=torch.float16)
    want = fake(x).float()
    got = Q.pack_model(net, 128)(x).float()
    assert rel(got, want) < 5e-3


def test_reference_perplexity_loop_runs_on_the_packed_model():
    """SURVEY 8(f) item 4 end to end: the reference's own `evaluate_perplexity`
    (quantization_utils.py:269-322, reached through the drop-in module's pass-through) over a tiny
    random-init Llama -- once with every nn.Linear fake-quantized by pseudo_quantize_tensor (what
    the reference evaluates), once with every nn.Linear replaced by a QuantLinear that holds only
    the packed int4 record and multiplies through the dequant-fused tcgen05 GEMM."""
    import copy
    transformers = pytest.importorskip("transformers")
    import quantization_utils as qu
    from b200q import build as _build, qlinear as Q
    if _build.reference_dir() is None:
        pytest.skip("no reference checkout staged (baseline/_ref): evaluate_perplexity is the reference's code")
    torch.manual_seed(0)
    cfg = transformers.LlamaConfig(vocab_size=1000, hidden_size=256, intermediate_size=512,
                                   num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=4,
                                   max_position_embeddings=256, tie_word_embeddings=False)
    model = transformers.LlamaForCausalLM(cfg).half().cuda().eval()
    fake = copy.deepcopy(model)
    n_lin = 0
    for m in fake.modules():
        if isinstance(m, nn.Linear):
            m.weight.data = qu.pseudo_quantize_tensor(m.weight.data, 4, 128)
            n_lin += 1
    packed = Q.pack_model(copy.deepcopy(model), 128)
    assert sum(isinstance(m, Q.QuantLinear) for m in packed.modules()) == n_lin == 15
    assert not any(isinstance(m, nn.Linear) for m in packed.modules())
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, 1000, (1, 4 * 128), generator=g)
    ppl_fake = qu.evaluate_perplexity(fake, None, ids, n_samples=4, block_size=128, verbose=False)
    ppl_packed = qu.evaluate_perplexity(packed, None, ids, n_samples=4, block_size=128, verbose=False)
    ppl_raw = qu.evaluate_perplexity(model, None, ids, n_samples=4, block_size=128, verbose=False)
    assert ppl_fake > 1 and abs(ppl_packed - ppl_fake) / ppl_fake < 2e-3, (ppl_raw, ppl_fake, ppl_packed)
    # the packed model keeps a quarter of the Linear weight bytes
    lin_bytes = sum(m.weight.numel() * 2 for m in fake.modules() if isinstance(m, nn.Linear))
    packed_bytes = sum(b.numel() * b.element_size() for m in packed.modules() if isinstance(m, Q.QuantLinear)
                       for b in m.buffers())
    assert packed_bytes < 0.35 * lin_bytes
