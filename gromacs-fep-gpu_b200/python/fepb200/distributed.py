"""Multi-GPU driver of the perturbed-pair path: one process per GPU, the FEP pair list split by
i-entry over the ranks (shard.py), coordinates and parameters replicated, and two collectives per
step over NCCL/NVLink:

  * forces + shift forces: all-reduce of the fp32 result block [3*nTouched + 135] -- every rank
    numbers the touched atoms identically, so the block is handed to NCCL as is;
  * Vc/Vvdw per energy-group pair, dV/dlambda, foreign energies: all-reduce of the small fp64 block.

torch.distributed is plumbing only (process group, NCCL communicator, stream); the tensors it
reduces are zero-copy views of the library's device result block.
"""
from __future__ import annotations

import torch

from .lib import FepContext


class _DeviceMemory:
    """Minimal __cuda_array_interface__ holder so torch can view foreign device memory."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = dict(shape=(n,), typestr=typestr, data=(ptr, False), version=2)


def result_tensors(ctx: FepContext) -> tuple[torch.Tensor, torch.Tensor]:
    """(fp32 block, fp64 block) of the context's device result block as torch views (no copy)."""
    lay = ctx.layout()
    p32, p64 = ctx.result_device_ptrs()
    dev = torch.device("cuda", ctx.device)
    f32 = torch.as_tensor(_DeviceMemory(p32, int(lay.f32_words), "<f4"), device=dev)
    f64 = torch.as_tensor(_DeviceMemory(p64, int(lay.f64_words), "<f8"), device=dev)
    return f32, f64


class ShardedFep:
    """The per-rank object: holds this rank's shard and reduces results over the group."""

    def __init__(self, problem, device: int, rank: int, world: int, group=None):
        self.rank, self.world, self.group = rank, world, group
        torch.cuda.set_device(device)
        self.ctx = FepContext(device)
        # all work of the context goes to one torch stream so that kernels, NCCL collectives and
        # copies are ordered without host synchronisation
        self.stream = torch.cuda.Stream(device)
        self.ctx.set_stream(self.stream.cuda_stream)
        self.ctx.set_problem(problem, rank=rank, nranks=world)
        self.f32, self.f64 = result_tensors(self.ctx)

    def launch(self, flags: int) -> None:
        """Kernels of this rank's shard, then the two all-reduces, all asynchronous on self.stream."""
        import torch.distributed as dist

        self.ctx.launch(flags)
        if self.world > 1:
            with torch.cuda.stream(self.stream):
                dist.all_reduce(self.f64, group=self.group)
                dist.all_reduce(self.f32, group=self.group)

    def step(self, x, shiftvec, flags: int, out: dict | None = None) -> dict:
        """Host buffers in, reduced host buffers out (every rank receives the full result)."""
        self.ctx.upload_x(x, shiftvec)
        self.launch(flags)
        return self.ctx.download(flags, out)

    def close(self) -> None:
        self.ctx.close()
