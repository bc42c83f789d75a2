"""Data-parallel plumbing: one process per GPU, per-replica BatchNorm (the reference's nn.DataParallel / DDP
without SyncBN semantics, trainer.py:165-171,191-196), gradients averaged with one all-reduce per flat bucket
(D64 23 MB, D128 75 MB, D256 285 MB, G 85 MB in fp32) over NCCL / NVLink. Works with gloo on CPU for tests."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, local_rank, world_size)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


class GradAllReducer:
    """callable(flat_grad): in-place mean over ranks. Mean of per-replica mean losses == the global-batch mean for
    equal shards."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bytes = 0

    def __call__(self, flat):
        if self.world == 1:
            return flat
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.mul_(1.0 / self.world)
        self.bytes += flat.numel() * flat.element_size()
        return flat


def broadcast_state(modules, src=0):
    """Identical initial weights and BN buffers on every rank (what DDP's constructor does)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src=src)


def shutdown():
    """End of a benchmark process. destroy_process_group() blocked forever here when CUDA graphs that captured NCCL
    collectives were still alive (2-GPU bench: result printed, process never exited), so the process group is left to
    the interpreter's teardown; stdout is flushed first because the caller may hard-exit."""
    import sys
    sys.stdout.flush()
    sys.stderr.flush()
