"""N > 1 host logic on CPU: world_size-2 gloo run of the sharding + single all-reduce used by bench.py / engine."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deftet_b200 import dist as ddist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    B, V = 5, 40
    data = {"pos": torch.randn(B, V, 3), "w": torch.rand(B)}
    shard = ddist.shard_batch(data, rank, world)
    delta = torch.zeros(V, 3, requires_grad=True)
    bias = torch.zeros(3, requires_grad=True)
    loss = (shard["w"].reshape(-1, 1, 1) * (shard["pos"] + delta + bias) ** 2).sum()
    loss.backward()
    ddist.GradBucket([delta, bias]).all_reduce()
    if rank == 0:
        torch.save({"delta": delta.grad, "bias": bias.grad}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradient_equals_single_process(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    B, V = 5, 40
    pos, w = torch.randn(B, V, 3), torch.rand(B)
    delta = torch.zeros(V, 3, requires_grad=True)
    bias = torch.zeros(3, requires_grad=True)
    (w.reshape(-1, 1, 1) * (pos + delta + bias) ** 2).sum().backward()
    assert torch.allclose(got["delta"], delta.grad, rtol=1e-6, atol=1e-6)
    assert torch.allclose(got["bias"], bias.grad, rtol=1e-6, atol=1e-5)


def test_shard_ranges_cover_batch():
    for n in (1, 4, 7, 8):
        for world in (1, 2, 4, 8):
            spans = [ddist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def _worker_none_grad(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = torch.zeros(6, requires_grad=True)
    if rank == 1:                       # rank 0 contributes nothing: its grad stays None until the all-reduce
        (p * torch.arange(6.0)).sum().backward()
    ddist.GradBucket([p]).all_reduce()
    if rank == 0:
        torch.save(p.grad, out)
    dist.barrier()
    dist.destroy_process_group()


def test_single_parameter_bucket_with_missing_grad(tmp_path):
    """the single-parameter fast path must hand the reduced sum to a rank whose grad was None."""
    out = str(tmp_path / "g1.pt")
    mp.spawn(_worker_none_grad, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    assert got is not None and torch.equal(got, torch.arange(6.0))
