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
    """callable(flat_grad, chan=0): in-place mean over ranks. Mean of per-replica mean losses == the global-batch mean
    for equal shards.

    channels > 1 (NCCL only): one communicator per `chan` (the trainer passes the network's index). Collectives on ONE
    communicator execute in host issue order, so with a single communicator the early, large gradient slices of the
    biggest discriminator would queue behind the smaller networks' reductions, which are issued earlier by the host but
    become ready later on the device; separate communicators let the networks' branches of the step reduce
    independently. NCCL averages in the collective (ReduceOp.AVG); gloo sums and scales."""

    def __init__(self, group=None, channels=1):
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bytes = 0
        self.groups = [group]
        self.avg = False
        if self.world > 1 and dist.get_backend(group) == "nccl":
            self.avg = True
            if group is None and channels > 1:
                self.groups += [dist.new_group(backend="nccl") for _ in range(channels - 1)]

    def __call__(self, flat, chan=0):
        if self.world == 1:
            return flat
        g = self.groups[chan % len(self.groups)]
        if self.avg:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=g)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=g)
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
