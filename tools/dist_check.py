"""Row-sharded correctness on real GPUs (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
Every method quantizes its row shard inside b200q.dist.row_sharded(); rank 0 gathers the shards and
compares with the oracle on the UNSHARDED weight (bit-exact where the single-GPU path is)."""
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO / "llm-quantization_b200"), str(REPO)):
    sys.path.insert(0, p)
import torch
import torch.distributed as td
import torch.nn as nn

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
td.init_process_group("nccl", device_id=dev)
from b200q import dist as D, tensor_ops as T
from oracle import quant_oracle as O
import awq_quantizer, gptq_quantizer, pot_apot_quantizer, smooth_quant_quantizer

g = torch.Generator().manual_seed(0)
N, K = 4100, 512                     # 4100*512 > 500000: APOT's coarse grid only if numel is GLOBAL
W = torch.randn(N, K, generator=g) * 0.02
chan = torch.ones(K); chan[torch.randperm(K, generator=g)[:5]] = 20.0
acts = [(torch.randn(128, K, generator=g) * chan) for _ in range(8)]
stats = [a.abs().mean(0) for a in acts]
act_scale = torch.stack([a.abs().amax(0) for a in acts]).amax(0)
r0, r1 = D.shard_rows(N, world, rank)


def shard_model():
    net = nn.Sequential(nn.Linear(K, r1 - r0, bias=False)).to(dev)
    net[0].weight.data = W[r0:r1].clone().to(dev)
    return net


def gather(t):
    parts = [torch.empty((D.shard_rows(N, world, r)[1] - D.shard_rows(N, world, r)[0], K), device=dev)
             for r in range(world)]
    td.all_gather(parts, t.contiguous())
    return torch.cat(parts).cpu()


results = {}
with D.row_sharded():
    net = shard_model(); gptq_quantizer.gptq_quantize_model_weight(net, 4, 128, {"0": acts}, verbose=False)
    results["gptq parity (+H, H^-1 built, samples dealt, inverse broadcast)"] = \
        torch.equal(gather(net[0].weight.data), O.gptq_parity_quant(W, 4)["out"])
    net = shard_model(); awq_quantizer.awq_quantize_model_weight(net, 4, 128, {"0": stats}, 0.01, 2.0)
    results["awq"] = torch.equal(gather(net[0].weight.data), O.awq_layer(W, stats, 4, 128, 0.01, 2.0)["out"])
    net = shard_model(); pot_apot_quantizer.apot_quantize_model_weight(net, 4, 128)
    results["apot (global numel picks the grid)"] = \
        torch.equal(gather(net[0].weight.data), O.apot_quant(W, 4, 128, 2)["out"])
    small = W[:64]
    net = shard_model(); smooth_quant_quantizer.smoothquant_quantize_model_weight(net, 8, 128, {"0": act_scale}, 0.5, verbose=False)
    s = net[0].smoothing_scale.cpu()
    results["smoothquant (column max all-reduced)"] = \
        torch.equal(gather(net[0].weight.data), O.smoothquant_layer(W, None, 0.5, 8, 128, s=s)["out"]) and \
        torch.allclose(s, O.smooth_scale(act_scale, W, 0.5), rtol=3e-7, atol=0)
    # Hessian: every rank ends with the same H as a single rank would compute
    H = T.gptq_hessian(acts, K, dev, 0.01, 128)
with torch.no_grad():
    H1 = T.gptq_hessian(acts, K, dev, 0.01, 128)          # unsharded, this rank alone
    results["hessian all-reduce == unsharded"] = bool(((H - H1).abs().max() / H1.abs().max()) < 1e-5)
    net = shard_model()
    with D.row_sharded():
        best = awq_quantizer.awq_search_scale_factor(net, 4, 128, {"0": acts}, n_grid=10)
    full = nn.Sequential(nn.Linear(K, N, bias=False)).to(dev); full[0].weight.data = W.clone().to(dev)
    results[f"awq search argmin sharded == unsharded ({best})"] = \
        best == awq_quantizer.awq_search_scale_factor(full, 4, 128, {"0": acts}, n_grid=10)
    # three layers with different K: exercises the look-ahead (next layer's Gram matrix and its
    # all-reduce are started before the current layer's search) with raw 3-D activations
    Ks = [256, 512, 384]
    gg = torch.Generator().manual_seed(5)
    Ws = [torch.randn(512, k, generator=gg) * 0.02 for k in Ks]
    acts3 = {str(i): (torch.randn(4, 128, k, generator=gg)).to(torch.bfloat16).to(dev) for i, k in enumerate(Ks)}
    q0, q1 = D.shard_rows(512, world, rank)
    def stack(rows):
        net = nn.Sequential(*[nn.Linear(k, 1, bias=False) for k in Ks]).to(dev)
        for lin, w in zip(net, Ws):
            lin.weight.data = w[rows].clone().to(dev)
        return net
    with D.row_sharded():
        b_sh = awq_quantizer.awq_search_scale_factor(stack(slice(q0, q1)), 4, 128, acts3, n_grid=10)
    b_full = awq_quantizer.awq_search_scale_factor(stack(slice(0, 512)), 4, 128, acts3, n_grid=10)
    results[f"3-layer awq search with look-ahead: sharded {b_sh} == unsharded {b_full}"] = b_sh == b_full
    # five layers through the grouped GPTQ walker (world-size layers prepared at once, rank j
    # inverts layer j of the group): parity mode must stay bit-exact, the compensated loop must
    # agree with the single-GPU run up to the summation order of the all-reduced Hessian
    Ks5 = [256, 384, 256, 512, 384]
    W5 = [torch.randn(512, k, generator=gg) * 0.02 for k in Ks5]
    acts5 = {str(i): [torch.randn(128, k, generator=gg) for _ in range(8)] for i, k in enumerate(Ks5)}
    def stack5(rows):
        net = nn.Sequential(*[nn.Linear(k, 1, bias=False) for k in Ks5]).to(dev)
        for lin, w in zip(net, W5):
            lin.weight.data = w[rows].clone().to(dev)
        return net
    def gather5(t, k):
        parts = [torch.empty((D.shard_rows(512, world, r)[1] - D.shard_rows(512, world, r)[0], k), device=dev)
                 for r in range(world)]
        td.all_gather(parts, t.contiguous())
        return torch.cat(parts).cpu()
    for mode in ("parity", "compensated"):
        gptq_quantizer.MODE = mode
        sh = stack5(slice(q0, q1))
        with D.row_sharded():
            gptq_quantizer.gptq_quantize_model_weight(sh, 4, 128, acts5, actorder=True, verbose=False)
        full = stack5(slice(0, 512))
        gptq_quantizer.gptq_quantize_model_weight(full, 4, 128, acts5, actorder=True, verbose=False)
        outs = [(gather5(a.weight.data, k), b.weight.data.cpu()) for a, b, k in zip(sh, full, Ks5)]
        agree = min((x == y).float().mean().item() for x, y in outs)
        if mode == "parity":
            results[f"grouped gptq walker, parity: sharded vs unsharded agreement {agree:.5f}"] = agree == 1.0
        else:
            # act-order sorts diag(H); the all-reduced H differs from the single-rank one in its last
            # bits, which can swap neighbouring columns in that order -- a different, equally valid
            # run.  Codes then need not agree; the output error tr(dW H dW^T) must (2 %), and both
            # runs must beat round-to-nearest.
            worst, beats = 0.0, True
            for (x, y), w5, (i, k) in zip(outs, W5, enumerate(Ks5)):
                Hd = T.gptq_hessian(acts5[str(i)], k, dev, 0.01, 128).double().cpu()
                def out_err(Q, w5=w5, Hd=Hd):
                    D = Q.double() - w5.double()
                    return float(((D @ Hd) * D).sum())
                e_sh, e_full = out_err(x), out_err(y)
                e_rtn = out_err(O.uniform_group_quant(w5, 4, 128)["out"])
                worst = max(worst, abs(e_sh - e_full) / e_full)
                beats = beats and e_sh < e_rtn and e_full < e_rtn
            results[f"grouped gptq walker, compensated: code agreement {agree:.5f}, output error of sharded "
                    f"vs unsharded within {worst:.2e} (relative), both below round-to-nearest"] = \
                worst < 0.02 and beats
    gptq_quantizer.MODE = "parity"
    # Llama-3-8B k/v projection shape (BASELINE configs[2]): N = 1024 -> 128 rows per rank at 8
    # GPUs, K = 4096 (tensor-core inverse path), w3 act-order, parity and compensated
    Nkv, Kkv = 1024, 4096
    Wkv = torch.randn(Nkv, Kkv, generator=gg) * 0.02
    chan_kv = torch.ones(Kkv); chan_kv[torch.randperm(Kkv, generator=gg)[:40]] = 20.0
    acts_kv = {"0": [(torch.randn(256, Kkv, generator=gg) * chan_kv).to(torch.bfloat16).to(dev) for _ in range(8)]}
    k0, k1 = D.shard_rows(Nkv, world, rank)
    def kv(rows):
        net = nn.Sequential(nn.Linear(Kkv, 1, bias=False)).to(dev)
        net[0].weight.data = Wkv[rows].clone().to(dev)
        return net
    def gather_kv(t):
        parts = [torch.empty((D.shard_rows(Nkv, world, r)[1] - D.shard_rows(Nkv, world, r)[0], Kkv), device=dev)
                 for r in range(world)]
        td.all_gather(parts, t.contiguous())
        return torch.cat(parts).cpu()
    for mode in ("parity", "compensated"):
        gptq_quantizer.MODE = mode
        sh = kv(slice(k0, k1))
        with D.row_sharded():
            gptq_quantizer.gptq_quantize_model_weight(sh, 3, 128, acts_kv, actorder=True, verbose=False)
        full = kv(slice(0, Nkv))
        gptq_quantizer.gptq_quantize_model_weight(full, 3, 128, acts_kv, actorder=True, verbose=False)
        got = gather_kv(sh[0].weight.data)
        ref_out = full[0].weight.data.cpu()
        agree = (got == ref_out).float().mean().item()
        if mode == "parity":
            ok = agree == 1.0 and torch.equal(got, O.gptq_parity_quant(Wkv, 3)["out"])
            note = f"agreement {agree:.5f}"
        else:
            # The all-reduced Hessian differs from the single-rank one in the last bits (summation
            # order), which can swap neighbours in the act-order permutation argsort(diag(H)) and
            # with them the order two columns are quantised in: a different, equally valid GPTQ run
            # whose codes need not agree.  What must agree is the QUALITY: the output error
            # tr(dW H dW^T) of the two runs, and both must beat round-to-nearest.
            Hd = T.gptq_hessian(acts_kv["0"], Kkv, dev, 0.01, 128).double().cpu()
            def out_err(Q):
                D = Q.double() - Wkv.double()
                return float(((D @ Hd) * D).sum())
            e_sh, e_full = out_err(got), out_err(ref_out)
            e_rtn = out_err(O.uniform_group_quant(Wkv, 3, 128)["out"])
            ok = abs(e_sh - e_full) / e_full < 0.02 and e_sh < e_rtn and e_full < e_rtn
            note = (f"code agreement {agree:.5f}, output error sharded / unsharded / round-to-nearest "
                    f"{e_sh:.4e} / {e_full:.4e} / {e_rtn:.4e}")
        results[f"llama3 k/v 1024x4096 w3 act-order, {mode}: {k1 - k0} rows/rank, {note}"] = ok
    gptq_quantizer.MODE = "parity"
if rank == 0:
    print(f"dist_check: {world} ranks on {torch.cuda.get_device_name(0)}, NCCL {'.'.join(map(str, torch.cuda.nccl.version()))}")
    for k, v in results.items():
        print(("PASS " if v else "FAIL ") + k)
    print("ALL PASS" if all(results.values()) else "SOME FAILED")
td.destroy_process_group()
