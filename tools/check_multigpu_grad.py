"""T3: a batch sharded over N ranks + ONE all-reduce gives the same gradient as the whole batch on
one GPU.  Run: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_multigpu_grad.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from bench import Step, analytic_scene
from deftet_b200 import dist as ddist
from deftet_b200.engine import GeometryEngine
from deftet_b200.grid import acute_lattice_grid


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    grid = acute_lattice_grid(24)
    B = 2 * world
    eng = GeometryEngine(grid.centred(), grid.tets, max_boundary_faces=4096, device=dev)
    full = analytic_scene(grid, B, 20000, 20000, 5, dev)             # identical on every rank (same seed)
    gen = torch.Generator(device=dev).manual_seed(1)
    u = torch.sqrt(torch.rand(B, 4096, 20, device=dev, generator=gen))
    v = torch.rand(B, 4096, 20, device=dev, generator=gen)
    lo, hi = ddist.shard_range(B, rank, world)
    shard = {k: t[lo:hi].contiguous() for k, t in full.items()}
    step = Step(eng, None, 4096, 20)
    step.forward_backward(shard, u[lo:hi].contiguous(), v[lo:hi].contiguous())
    ddist.GradBucket([step.delta]).all_reduce()
    g_sharded = step.delta.grad.clone()
    ok = True
    if rank == 0:
        step.delta.grad = None
        step.forward_backward(full, u, v)
        g_full = step.delta.grad
        err = float((g_sharded - g_full).abs().max() / g_full.abs().max())
        ok = err < 1e-5
        print("multi-GPU gradient check: world=%d batch=%d max rel err %.3e -> %s" % (world, B, err, "OK" if ok else "FAIL"))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
