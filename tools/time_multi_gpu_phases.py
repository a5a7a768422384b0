"""Per-rank device-time breakdown of one sharded step (run under torchrun):
kernels of the shard (pass / foreign / epilogue), the peer-memory reduction kernel, and the whole
step, so that the cost of the exchange can be told from skew between the ranks.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29533 tools/time_multi_gpu_phases.py [C5]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gromacs-fep-gpu_b200", "python"), ROOT]
import numpy as np
import torch
import torch.distributed as dist

from fepb200 import params as P
from fepb200.distributed import ShardedFep
from fepb200.synth import make_system

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "C5"
steps = 60
problem = make_system(name)
flags = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA
sh = ShardedFep(problem, local, rank, world)
ctx = sh.ctx
lay = ctx.layout()
torch.cuda.set_stream(sh.stream)
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
ctx.upload_x(np.ascontiguousarray(problem.x), problem.shiftvec)
for _ in range(5):
    flush.zero_()
    sh.launch(flags)
torch.cuda.synchronize()


def ev():
    return torch.cuda.Event(enable_timing=True)


# whole step, back to back as bench.py does it; with an event between the shard's kernels and the reduction
whole, local_part, red_part = [], [], []
e = [[ev(), ev(), ev()] for _ in range(steps)]
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
for i in range(steps):
    flush.zero_()
    e[i][0].record()
    if sh.reduction == "p2p":
        k = sh._step & 1
        sh._step += 1
        ctx.set_partial_result_block(sh._slots[k][rank])
        ctx.launch(flags)
        e[i][1].record()
        ctx.reduce_scatter_peers(sh._slots[k], sh._flags, rank, sh._step)
    elif sh.reduction == "p2p-push":
        k = sh._step & 1
        sh._step += 1
        ctx.set_push_targets(sh._push[k])
        ctx.launch(flags)
        e[i][1].record()
        ctx.reduce_scatter_peers(sh._recv[k], sh._flags, rank, sh._step)
    elif sh.reduction == "p2p-allreduce":
        k = sh._step & 1
        sh._step += 1
        ctx.set_partial_result_block(sh._slots[k][rank])
        ctx.launch(flags)
        e[i][1].record()
        ctx.reduce_peers(sh._slots[k], sh._flags, rank, sh._step)
    else:
        sh.launch(flags)
        e[i][1].record()
    e[i][2].record()
torch.cuda.synchronize()
whole = np.array([a.elapsed_time(c) for a, b, c in e]) * 1e3
local_part = np.array([a.elapsed_time(b) for a, b, c in e]) * 1e3
red_part = np.array([b.elapsed_time(c) for a, b, c in e]) * 1e3

# the reduction kernel on its own, the ranks aligned on the device right before it (ShardedFep.align): the cost of the
# exchange itself (barrier inside the kernel + the pull over NVLink), without any waiting for a slower rank
red_alone = None
if sh.reduction in ("p2p", "p2p-allreduce", "p2p-push"):
    for cold in (True, False):
        ea = [[ev(), ev()] for _ in range(steps)]
        dist.barrier()
        torch.cuda.synchronize()
        for i in range(steps):
            if cold:
                flush.zero_()
            sh.align()
            ea[i][0].record()
            k = sh._step & 1
            sh._step += 1
            if sh.reduction == "p2p":
                ctx.reduce_scatter_peers(sh._slots[k], sh._flags, rank, sh._step)
            elif sh.reduction == "p2p-push":
                ctx.reduce_scatter_peers(sh._recv[k], sh._flags, rank, sh._step)
            else:
                ctx.reduce_peers(sh._slots[k], sh._flags, rank, sh._step)
            ea[i][1].record()
        torch.cuda.synchronize()
        t = np.array([a.elapsed_time(b) for a, b in ea]) * 1e3
        red_alone = (red_alone or "") + f" {'L2 flushed' if cold else 'warm'} {t.mean():.1f} (min {t.min():.1f})"
    # ... and an empty-ish reference: the alignment barrier kernel alone between two events
    ea = [[ev(), ev()] for _ in range(steps)]
    for i in range(steps):
        sh.align()
        ea[i][0].record()
        sh.align()
        ea[i][1].record()
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) for a, b in ea]) * 1e3
    red_alone += f" | torch symm-mem barrier kernel alone {t.mean():.1f} (min {t.min():.1f})"

# kernels one by one
ctx.set_profiling(True)
kms = []
for i in range(20):
    flush.zero_()
    sh.launch(flags)
    torch.cuda.synchronize()
    kms.append(ctx.kernel_ms())
ctx.set_profiling(False)
kms = np.array(kms) * 1e3

msg = (f"rank {rank}/{world} {name} red={sh.reduction} pairs {int(lay.nrj)} entries {int(lay.nri)} touched {int(lay.ntouched)} | "
       f"step mean {whole.mean():.1f} min {whole.min():.1f} us | shard kernels {local_part.mean():.1f} (min {local_part.min():.1f}) | "
       f"reduction {red_part.mean():.1f} (min {red_part.min():.1f}) | reduction kernel alone, ranks aligned:{red_alone} | alone: pass {kms[:,0].mean():.1f} foreign {kms[:,1].mean():.1f} "
       f"epilogue {kms[:,2].mean():.1f} us")
if world > 1:
    allmsg = [None] * world
    dist.all_gather_object(allmsg, msg)
    t = torch.tensor([whole.sum()], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("\n".join(allmsg))
        print(f"max over ranks of the step: {t.item() / steps:.1f} us")
    dist.destroy_process_group()
else:
    print(msg)
sh.close()
