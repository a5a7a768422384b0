"""Multi-GPU driver of the perturbed-pair path: one process per GPU, the FEP pair list split over
the ranks, coordinates and parameters replicated.  Ways to combine the ranks' results
(SURVEY.md section 8e), selected by ShardedFep.reduction:

  "fused": no separate collective.  Every rank holds the layout of the full list and
      evaluates its share of the pairs with the unchanged pair kernels, results staying in its own
      exchange buffer (symmetric memory); after a cross-GPU barrier inside the epilogue kernel each
      rank reads over NVLink, from whichever rank produced them, the contributions of the atoms it
      owns (forces: reduce-scatter) and all scalar inputs (all-reduce).
      fepb200_set_peer_exchange().
  "p2p": the list is split by i-entry (shard.py), every rank computes the result block of its
      shard (its epilogue visits only the atoms its shard touches) and ONE kernel of libfepb200 does the
      force REDUCE-SCATTER over NVLink peer memory -- each rank sums the atoms it owns over all ranks'
      blocks -- together with the all-reduce of shift forces and scalars (fepb200_reduce_scatter_peers).
  "p2p-push" (default): the same split and the same result, the data travelling the other way: the epilogue stores
      the force of an atom straight into a receive block on the rank that owns it, shift forces and scalars into
      receive blocks on all ranks (fepb200_set_push_targets), and the same reduction kernel sums a rank's OWN
      receive blocks behind its barrier -- local loads only.  46.9 against 48.5 us per C5 step on 2 B200, 39.3
      against 44.0 us on 4.
  "p2p-allreduce": the same split, every rank sums everything (fepb200_reduce_peers; round-1 default).
  "nccl": same split, two ncclAllReduce calls on zero-copy views of the result block.
See DESIGN.md section 5 for the measurements.

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

    reduction = "fused": see the module docstring; step() returns the
    forces of the atoms this rank owns (zeros elsewhere: the sum over ranks is the full force array)
    and the full scalars on every rank.
    reduction = "p2p": every rank writes its result block into
    symmetric memory (torch.distributed._symmetric_memory: CUDA VMM allocations every rank of the
    node has mapped over NVLink), passes a device-side barrier, and libfepb200's own kernel sums, through
    the peer pointers and in rank order, the forces of the atoms this rank owns and all scalars
    (fepb200_reduce_scatter_peers); step() returns what "fused" returns.  "p2p-allreduce": every rank sums
    everything (fepb200_reduce_peers).  "p2p-push" (default when available): the same result as "p2p", but the epilogue stores its sums
    straight into receive blocks on the ranks that need them (fepb200_set_push_targets) and the reduction kernel
    reads local memory only: one one-way NVLink trip per step.
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
        self.reduction = "none"
        self._p2p_error = None  # why symmetric memory was not used, if it was asked for
        self._step = 0
        # default: push (measured faster than the pull on 2 and 4 B200); the push targets are kernel arguments for at most 8 ranks
        want = (reduction or os.environ.get("FEPB200_REDUCTION", "p2p-push" if world <= 8 else "p2p")) if world > 1 else "none"
        if want == "fused":
            try:
                self.ctx.set_problem(problem)  # the full list on every rank
                self._setup_fused(device)
                self.reduction = "fused"
            except Exception as exc:  # no symmetric memory on this system
                self._p2p_error = repr(exc)
                want = "p2p"
        if self.reduction != "fused":
            self.ctx.set_problem(problem, rank=rank, nranks=world)
        self.f32, self.f64 = result_tensors(self.ctx)
        if world > 1 and self.reduction != "fused":
            self.reduction = "nccl"
            if want in ("p2p", "p2p-allreduce", "p2p-push"):
                try:
                    self._setup_p2p(device, blocks_per_slot=world if want == "p2p-push" else 1)
                    self.reduction = want
                except Exception as exc:  # no symmetric memory on this system: NCCL does the same job
                    self._p2p_error = repr(exc)

    def _setup_fused(self, device: int) -> None:
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        group = self.group if self.group is not None else dist.group.WORLD
        nbytes = (self.ctx.exchange_bytes(self.world) + 4095) // 4096 * 4096
        with torch.cuda.stream(self.stream):
            # two exchange slots + the barrier's sequence flags, zero before anybody announces a step
            self._sym = symm_mem.empty(nbytes, dtype=torch.uint8, device=torch.device("cuda", device))
            self._sym.zero_()
            self._hdl = symm_mem.rendezvous(self._sym, group)
            self._hdl.barrier(channel=0)
        self.stream.synchronize()
        self.ctx.set_peer_exchange(self.world, self.rank, [int(p) for p in self._hdl.buffer_ptrs], nbytes)
        self.exchange_bytes = nbytes

    def _setup_p2p(self, device: int, blocks_per_slot: int = 1) -> None:
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        group = self.group if self.group is not None else dist.group.WORLD
        self.block_bytes = (self.ctx.result_block_bytes() + 255) // 256 * 256
        slot_bytes = blocks_per_slot * self.block_bytes
        with torch.cuda.stream(self.stream):
            # [slot 0 | slot 1 | flag array (16 x uint32 sequence numbers)]; a slot is one block ("p2p": the rank's own
            # partial result, pulled by the others) or one receive block per rank ("p2p-push": written by that rank)
            self._sym = symm_mem.empty(2 * slot_bytes + 256, dtype=torch.uint8, device=torch.device("cuda", device))
            self._sym.zero_()
            self._hdl = symm_mem.rendezvous(self._sym, group)
            self._hdl.barrier(channel=0)  # everybody's flags and blocks are zero before anybody announces or pushes
        base = [int(p) for p in self._hdl.buffer_ptrs]
        # two alternating slots: a rank that runs ahead writes the other slot, and cannot come back
        # to this one before everybody has announced the next step
        self._slots = [[b + k * slot_bytes for b in base] for k in (0, 1)]
        self._flags = [b + 2 * slot_bytes for b in base]
        if blocks_per_slot > 1:
            # my receive block on rank r, and the blocks the ranks write on me
            from .shard import push_block_addresses

            self._push, self._recv, flags = push_block_addresses(base, self.rank, self.block_bytes)
            assert flags == self._flags
        self.stream.synchronize()

    def launch(self, flags: int) -> None:
        """Kernels of this rank's shard, then the reduction over ranks, all asynchronous on self.stream."""
        if self.reduction == "fused":
            self.ctx.launch(flags)  # the exchange is part of the pair kernels and the epilogue
            return
        if self.reduction in ("p2p", "p2p-allreduce"):
            k = self._step & 1
            self._step += 1
            # the epilogue writes this rank's partial result straight into its symmetric slot
            self.ctx.set_partial_result_block(self._slots[k][self.rank])
            self.ctx.launch(flags)
            # one kernel: announce the step to all peers, wait for theirs, sum over NVLink -- the atoms this rank
            # owns and the scalars ("p2p"), or everything ("p2p-allreduce")
            if self.reduction == "p2p":
                self.ctx.reduce_scatter_peers(self._slots[k], self._flags, self.rank, self._step)
            else:
                self.ctx.reduce_peers(self._slots[k], self._flags, self.rank, self._step)
            return
        if self.reduction == "p2p-push":
            k = self._step & 1
            self._step += 1
            # the epilogue stores this rank's sums straight into the ranks that need them; the reduction kernel behind
            # the barrier adds up the blocks the ranks have written HERE
            self.ctx.set_push_targets(self._push[k])
            self.ctx.launch(flags)
            self.ctx.reduce_scatter_peers(self._recv[k], self._flags, self.rank, self._step)
            return
        if self.reduction == "nccl":
            import torch.distributed as dist

            with torch.cuda.stream(self.stream):
                # the all-reduce works in place: the words of atoms this rank's shard does not touch (which the epilogue
                # never writes) would keep the previous step's sums
                self.f32.zero_()
                self.ctx.launch(flags)
                dist.all_reduce(self.f64, group=self.group)
                dist.all_reduce(self.f32, group=self.group)
            return
        self.ctx.launch(flags)

    def align(self) -> bool:
        """Device-side barrier over the ranks on self.stream (the symmetric-memory handle's own barrier kernel, the one
        the set-up uses): every rank's stream continues only when all ranks' streams have arrived.  For measurements:
        bench.py puts it between the L2 flush and the start event of a timed step, so that the ranks' flush times do
        not show up as waiting time inside the step's cross-GPU barrier.  Not part of a step.  Returns False when
        there is nothing to align with (one rank, or no symmetric memory: the "nccl" reduction)."""
        hdl = getattr(self, "_hdl", None)
        if self.world == 1 or hdl is None:
            return False
        with torch.cuda.stream(self.stream):
            hdl.barrier(channel=1)
        return True

    def step(self, x, shiftvec, flags: int, out: dict | None = None) -> dict:
        """Host buffers in, reduced host buffers out ("fused" / "p2p": the forces of the atoms this rank owns and
        all scalars; "p2p-allreduce" / "nccl": every rank receives the full result)."""
        if self.world == 1:
            # one GPU: the public call itself (fepb200_compute: the epilogue writes the result block
            # straight into pinned host memory, no device-to-host copy)
            return self.ctx.compute(x, shiftvec, flags, out)
        self.ctx.upload_x(x, shiftvec)
        self.launch(flags)
        return self.ctx.download(flags, out)

    def close(self) -> None:
        self.ctx.close()
