"""Where the time of the fused exchange's epilogue goes (run under torchrun, N >= 2): global-timer
stamps of every epilogue block (fepb200_epilogue_trace) for steps with and without the L2 flush
between them.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29533 tools/trace_exchange.py [C5]
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
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "C5"
problem = make_system(name)
flags = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA
sh = ShardedFep(problem, local, rank, world, reduction="fused")
assert sh.reduction == "fused", sh._p2p_error
ctx = sh.ctx
torch.cuda.set_stream(sh.stream)
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
ctx.upload_x(np.ascontiguousarray(problem.x), problem.shiftvec)
ctx.epilogue_trace(True)
lines = []
for mode in ("flush", "noflush", "flush+hostsync"):
    rows = []
    for i in range(12):
        if mode.startswith("flush"):
            flush.zero_()
        if mode.endswith("hostsync"):
            torch.cuda.synchronize()
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sh.launch(flags)
        e1.record()
        torch.cuda.synchronize()
        t = ctx.epilogue_trace(True).astype(np.int64)
        t = t[t[:, 0] > 0]
        if i < 2 or len(t) == 0:
            continue
        role = t[:, 0] & 3
        z = (t[:, 0] & ~3).min()
        row = [e0.elapsed_time(e1) * 1e3, (t[:, 1].max() - z) * 1e-3, (t[:, 2].min() - z) * 1e-3, (t[:, 2].max() - z) * 1e-3,
               (t[:, 3].max() - z) * 1e-3]
        for ro in range(4):
            m = role == ro
            d = (t[m, 3] - t[m, 2]) * 1e-3 if m.any() else np.zeros(1)
            row += [m.sum(), np.median(d), d.max(), ((t[m, 3].max() - z) * 1e-3) if m.any() else 0.0]
        rows.append(row)
    r = np.median(np.array(rows), axis=0)
    msg = (f"rank {rank} {mode:15s} step {r[0]:6.1f} us | since first block entry: pair kernels done {r[1]:5.1f}, barrier passed "
           f"first {r[2]:5.1f} last {r[3]:5.1f}, last block done {r[4]:5.1f} us |")
    for ro, nm in enumerate(("jobs", "scalars", "heavy", "atoms")):
        n, med, mx, end = r[5 + 4 * ro: 9 + 4 * ro]
        msg += f" {nm}: {int(n)} blocks, after barrier median {med:4.1f} max {mx:4.1f}, last done {end:5.1f};"
    lines.append(msg)
allmsg = [None] * world
dist.all_gather_object(allmsg, "\n".join(lines))
if rank == 0:
    print("\n".join(allmsg))
dist.destroy_process_group()
sh.close()
