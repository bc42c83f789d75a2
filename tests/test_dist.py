"""world_size-2 gloo test (CPU) of the data-parallel plumbing: bucket all-reduce averages gradients, state broadcast
makes replicas identical, and the averaged gradient equals the gradient of the global-batch mean loss."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from sg2b200 import dist as sdist
    r, l, w = sdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)                      # different init per rank ...
    lin = torch.nn.Linear(5, 3)
    sdist.broadcast_state([lin])                       # ... made identical
    w0 = lin.weight.detach().clone()
    gathered = [torch.empty_like(w0) for _ in range(world)]
    dist.all_gather(gathered, w0)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    # per-rank shard of a global batch; mean of per-replica mean-loss gradients == global-batch mean gradient
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 5, generator=g)
    shard = x[rank * 4:(rank + 1) * 4]
    lin(shard).pow(2).mean().backward()
    flat = torch.cat([p.grad.reshape(-1) for p in lin.parameters()])
    red = sdist.GradAllReducer()
    red(flat)
    lin.zero_grad()
    lin(x).pow(2).mean().backward()
    ref = torch.cat([p.grad.reshape(-1) for p in lin.parameters()])
    if rank == 0:
        out.put((same, float((flat - ref).abs().max()), red.bytes))
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_matches_global_batch_gradient():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    same, err, nbytes = q.get()
    assert same and err < 1e-6 and nbytes == (15 + 3) * 4
