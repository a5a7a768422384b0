"""Multi-GPU driver of the perturbed-pair path: one process per GPU, the FEP pair list split by
i-entry over the ranks (shard.py), coordinates and parameters replicated, and two collectives per
step over NCCL/NVLink:

  * forces + shift forces: sum of the fp32 result block [3*nTouched + 135] -- every rank
    numbers the touched atoms identically, so the blocks add up element by element;
  * Vc/Vvdw per energy-group pair, dV/dlambda, foreign energies: sum of the small fp64 block.
Both sums are done by ONE kernel of libfepb200 over NVLink peer memory (ShardedFep.reduction ==
"p2p"), or by two ncclAllReduce calls ("nccl").

torch.distributed is plumbing only (process group, NCCL communicator, stream); the tensors it
reduces are zero-copy views of the library's device result block.
"""
from __future__ import annotations

import os

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
    """The per-rank object: holds this rank's shard and reduces results over the group.

    reduction = "p2p" (default when available): every rank publishes its result block in
    symmetric memory (torch.distributed._symmetric_memory: CUDA VMM allocations every rank of the
    node has mapped over NVLink), passes a device-side barrier, and libfepb200's own kernel reads
    all blocks through the peer pointers and sums them in rank order (fepb200_reduce_peers).
    reduction = "nccl": two ncclAllReduce calls on zero-copy views of the result block.
    """

    def __init__(self, problem, device: int, rank: int, world: int, group=None, reduction: str | None = None):
        self.rank, self.world, self.group = rank, world, group
        torch.cuda.set_device(device)
        self.ctx = FepContext(device)
        # all work of the context goes to one torch stream so that kernels, collectives and copies
        # are ordered without host synchronisation
        self.stream = torch.cuda.Stream(device)
        self.ctx.set_stream(self.stream.cuda_stream)
        self.ctx.set_problem(problem, rank=rank, nranks=world)
        self.f32, self.f64 = result_tensors(self.ctx)
        self.reduction = "none"
        self._p2p_error = None  # why symmetric memory was not used, if it was asked for
        self._step = 0
        if world > 1:
            want = reduction or os.environ.get("FEPB200_REDUCTION", "p2p")
            self.reduction = "nccl"
            if want == "p2p":
                try:
                    self._setup_p2p(device)
                    self.reduction = "p2p"
                except Exception as exc:  # no symmetric memory on this system: NCCL does the same job
                    self._p2p_error = repr(exc)

    def _setup_p2p(self, device: int) -> None:
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        group = self.group if self.group is not None else dist.group.WORLD
        self.block_bytes = (self.ctx.result_block_bytes() + 255) // 256 * 256
        with torch.cuda.stream(self.stream):
            # [slot 0 | slot 1 | flag array (16 x uint32 sequence numbers)]
            self._sym = symm_mem.empty(2 * self.block_bytes + 256, dtype=torch.uint8, device=torch.device("cuda", device))
            self._sym.zero_()
            self._hdl = symm_mem.rendezvous(self._sym, group)
            self._hdl.barrier(channel=0)  # everybody's flags are zero before anybody announces a step
        base = [int(p) for p in self._hdl.buffer_ptrs]
        # two alternating slots: a rank that runs ahead writes the other slot, and cannot come back
        # to this one before everybody has announced the next step
        self._slots = [[b + k * self.block_bytes for b in base] for k in (0, 1)]
        self._flags = [b + 2 * self.block_bytes for b in base]
        self.stream.synchronize()

    def launch(self, flags: int) -> None:
        """Kernels of this rank's shard, then the reduction over ranks, all asynchronous on self.stream."""
        if self.reduction == "p2p":
            k = self._step & 1
            self._step += 1
            # the epilogue writes this rank's partial result straight into its symmetric slot
            self.ctx.set_partial_result_block(self._slots[k][self.rank])
            self.ctx.launch(flags)
            # one kernel: announce the step to all peers, wait for theirs, sum all blocks over NVLink
            self.ctx.reduce_peers(self._slots[k], self._flags, self.rank, self._step)
            return
        self.ctx.launch(flags)
        if self.reduction == "nccl":
            import torch.distributed as dist

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
