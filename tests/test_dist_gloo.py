"""Row-sharded execution (b200q.dist) over gloo with world size 2, on the CPU.

The collectives and the sharding arithmetic are host logic: here they run for real across two
processes, with the ORACLE standing in for the kernels, and the sharded results must equal the
unsharded ones bit for bit (that is the property that lets 2/4/8 GPUs each take a block of output
rows: SURVEY.md section 8e)."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, ret):
    for p in (str(REPO / "llm-quantization_b200"), str(REPO)):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from b200q import dist as D
        from oracle import quant_oracle as O

        g = torch.Generator().manual_seed(42)
        N, K = 37, 256                       # uneven split: 19 + 18 rows
        W = torch.randn(N, K, generator=g) * 0.02
        act = torch.rand(K, generator=g) * 4

        # outside the context every helper is the identity
        t = torch.tensor([float(rank)])
        assert D.allreduce_max(t.clone()).item() == rank and not D.is_sharded()
        assert D.world_size() == 1 and D.global_numel(5, "cpu") == 5

        with D.row_sharded():
            assert D.is_sharded() and D.world_size() == world and D.rank() == rank
            r0, r1 = D.shard_rows(N)
            Ws = W[r0:r1]
            # GPTQ parity / SmoothQuant: column |max| spans ALL rows -> all-reduce MAX
            colmax = D.allreduce_max(Ws.abs().amax(0))
            assert torch.equal(colmax, W.abs().amax(0))
            qmax = 15
            scales = torch.clamp(colmax / qmax, min=1e-5)
            part = torch.clamp(torch.round(Ws / scales), -qmax - 1, qmax) * scales
            assert torch.equal(part, O.gptq_parity_quant(W, 4)["out"][r0:r1])
            s = torch.clamp(torch.pow(act.clamp(min=1e-5), 0.5) /
                            torch.pow(colmax.clamp(min=1e-5), 0.5), min=1e-5)
            assert torch.equal(Ws / s, O.smooth_layer(W, act, 0.5)["out"][r0:r1])
            # AWQ / pseudo-quant / POT: rows are independent, no exchange at all
            assert torch.equal(O.uniform_group_quant(Ws, 4, 128)["out"],
                               O.uniform_group_quant(W, 4, 128)["out"][r0:r1])
            assert torch.equal(O.pot_quant(Ws, 4, 128)["out"], O.pot_quant(W, 4, 128)["out"][r0:r1])
            # APOT picks its grid from the GLOBAL element count
            assert D.global_numel(Ws.numel(), "cpu") == W.numel()
            # Hessian: calibration samples dealt to the ranks, partial sums all-reduced
            feats = [torch.randn(16, K, generator=g) for _ in range(5)]
            lo, hi = D.shard_rows(len(feats))
            H = torch.zeros(K, K)
            for f in feats[lo:hi]:
                fn = f / (f.norm() + 1e-5)
                H += fn.T @ fn
            H = D.allreduce_sum(H) / len(feats) + 0.01 * torch.eye(K)
            torch.testing.assert_close(H, O.gptq_hessian(feats, K), rtol=1e-5, atol=1e-7)
            # the inverse is computed once and broadcast
            Hinv = O.gptq_hinv(H) if rank == 0 else torch.empty(K, K)
            D.broadcast(Hinv, 0)
            torch.testing.assert_close(Hinv @ (H + 1e-6 * torch.eye(K)), torch.eye(K), rtol=0, atol=1e-3)
            # AWQ search: per-candidate losses add up over row shards
            sal = torch.topk(act, 2)[1]
            cands = [1.0, 1.5, 2.0]
            loss = D.allreduce_sum(O.awq_search_losses(Ws, H, sal, 4, 128, cands).float())
            torch.testing.assert_close(loss.double(), O.awq_search_losses(W, H, sal, 4, 128, cands),
                                       rtol=1e-5, atol=0)
            # the packed symmetric exchange (tensor_ops._exchange_folded) in torch: every rank packs
            # the lower triangle of its partial, reduce-scatter -> each rank owns the sum of one
            # slice -> all-gather; unpacked and mirrored it is the all-reduced matrix
            part = torch.randn(K, K, generator=torch.Generator().manual_seed(100 + rank))
            part = part + part.T
            il = torch.tril_indices(K, K)
            packed = part[il[0], il[1]]
            per = (packed.numel() + world - 1) // world
            padded = torch.zeros(per * world)
            padded[:packed.numel()] = packed
            mine = D.reduce_scatter_sum(torch.empty(per), padded)
            whole = D.all_gather_into(torch.empty(per * world), mine)[:packed.numel()]
            full = torch.zeros(K, K)
            full[il[0], il[1]] = whole
            full = full + full.T - torch.diag(torch.diag(full))
            torch.testing.assert_close(full, D.allreduce_sum(part.clone()), rtol=1e-6, atol=1e-6)
            assert not D.backend_is_nccl()
        assert not D.is_sharded()
        ret[rank] = "ok"
    finally:
        td.destroy_process_group()


def test_row_sharding_over_gloo_world2():
    mp.set_start_method("spawn", force=True)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    procs = [mp.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_shard_rows_partitions_exactly():
    sys.path.insert(0, str(REPO / "llm-quantization_b200"))
    from b200q import dist as D
    for n in (0, 1, 7, 4096, 11008, 32000):
        for world in (1, 2, 4, 8):
            spans = [D.shard_rows(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_row_sharded_requires_process_group():
    sys.path.insert(0, str(REPO / "llm-quantization_b200"))
    from b200q import dist as D
    with pytest.raises(RuntimeError):
        with D.row_sharded():
            pass
