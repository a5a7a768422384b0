"""The N>1 path on real GPUs: one process per GPU; "fused" = the pair kernels scatter over NVLink
and every rank sums the atoms it owns (fepb200_set_peer_exchange), "p2p" = one reduction kernel over
peer memory, "nccl" = ncclAllReduce of the result block.  Skipped when the box has a single GPU (the
fused exchange then runs with several contexts on one device, test_gpu_peer_exchange.py; the
host-side logic is covered on CPU by test_sharding_cpu.py)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, root, result_file, reduction):
    for p in (os.path.join(root, "gromacs-fep-gpu_b200", "python"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    from fepb200 import params as P
    from fepb200.distributed import ShardedFep
    from fepb200.synth import make_system, scaled_spec
    from oracle import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    flags = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA
    prob = make_system(scaled_spec("C4", 4.2, 2, 25, n_foreign=5))
    sh = ShardedFep(prob, rank, rank, world, reduction=reduction)
    used = sh.reduction
    lay = sh.ctx.layout()
    out = sh.step(prob.x, prob.shiftvec, flags)
    out2 = sh.step(prob.x, prob.shiftvec, flags | P.CLEAR_OUTPUTS, out={k: v.copy() for k, v in out.items()})
    want = oracle.run_best(prob, flags)
    f_full = out["f"]
    if used in ("fused", "p2p", "p2p-push"):
        # every rank returns the forces of the atoms it owns: the sum over ranks is the force array
        p0, p1, a0, a1 = sh.ctx.peer_ranges()
        owned = np.zeros(prob.natoms, bool)
        owned[sh.ctx.touched_atoms()[a0:a1]] = True
        split_ok = 0 < a1 - a0 < len(sh.ctx.touched_atoms()) and not np.any(out["f"][~owned])
        split_ok = split_ok and (0 < p1 - p0 < prob.nblist.nrj if used == "fused" else 0 < int(lay.nri) < prob.nblist.nri)
        t = torch.from_numpy(out["f"].copy()).cuda()
        dist.all_reduce(t)
        f_full = t.cpu().numpy()
    else:
        split_ok = 0 < int(lay.nri) < prob.nblist.nri
    rms = np.sqrt(np.mean((f_full - want["f"]) ** 2) / np.mean(want["f"] ** 2))
    ok = rms < 1e-5 and int(lay.nri_total) == prob.nblist.nri and split_ok
    for k in ("Vc", "Vv", "dvdl", "foreign_energy", "foreign_dvdl"):
        scale = np.maximum(np.abs(want[k]), 1e-2 * np.max(np.abs(want[k])))
        ok = ok and bool(np.all(np.abs(out[k] - want[k]) <= 1e-4 * scale))
    ok = ok and all(np.array_equal(out[k], out2[k]) for k in out)  # deterministic, incl. the collective
    if used == "fused":
        # Many steps with moving atoms against an unsplit context on the same device: the exchange
        # slots are reused every other step, so a stale read (or a missing barrier) would show; the
        # forces, shift forces and Vc/Vv of the fused path are bit-identical to the single-GPU ones.
        from fepb200.lib import FepContext

        rng = np.random.default_rng(7)
        with FepContext(rank) as one:
            one.set_problem(prob)
            for _ in range(25):
                x = (prob.x + 2e-3 * rng.standard_normal(prob.x.shape)).astype(np.float32)
                got = sh.step(x, prob.shiftvec, flags | P.CLEAR_OUTPUTS, out=sh.ctx.new_outputs())
                ref = one.compute(x, prob.shiftvec, flags)
                t = torch.from_numpy(got["f"].copy()).cuda()
                dist.all_reduce(t)
                ok = ok and np.array_equal(t.cpu().numpy(), ref["f"])
                ok = ok and all(np.array_equal(got[k], ref[k]) for k in ("fshift", "Vc", "Vv"))
                ok = ok and np.allclose(got["foreign_energy"], ref["foreign_energy"], rtol=1e-6)
    if used in ("p2p", "p2p-push"):
        # Many steps with moving atoms against an unsplit context on the same device: the two slots are reused every
        # other step, so a stale block, a push that arrives after the sums, or a missing barrier would show.  The
        # order of the additions differs from the single-GPU one, so forces agree to rounding, not bit for bit.
        from fepb200.lib import FepContext

        rng = np.random.default_rng(11)
        with FepContext(rank) as one:
            one.set_problem(prob)
            for _ in range(25):
                x = (prob.x + 2e-3 * rng.standard_normal(prob.x.shape)).astype(np.float32)
                got = sh.step(x, prob.shiftvec, flags | P.CLEAR_OUTPUTS, out=sh.ctx.new_outputs())
                ref = one.compute(x, prob.shiftvec, flags)
                t = torch.from_numpy(got["f"].copy()).cuda()
                dist.all_reduce(t)
                ok = ok and np.allclose(t.cpu().numpy(), ref["f"], rtol=0.0, atol=2e-5 * np.max(np.abs(ref["f"])))
                ok = ok and np.allclose(got["fshift"], ref["fshift"], rtol=0.0, atol=2e-5 * np.max(np.abs(ref["fshift"])))
                ok = ok and np.allclose(got["foreign_energy"], ref["foreign_energy"], rtol=1e-5)
                ok = ok and np.allclose(got["dvdl"], ref["dvdl"], rtol=1e-5, atol=1e-5 * np.max(np.abs(ref["dvdl"])))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        with open(result_file, "w") as fh:
            fh.write(f"ok {used}" if int(flag.item()) == 1 else f"mismatch rms={rms:.3e} {used}")
    sh.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("reduction", ["fused", "p2p", "p2p-push", "p2p-allreduce", "nccl"])
def test_two_ranks_match_oracle(tmp_path, reduction):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    result = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), root, result, reduction), nprocs=2, join=True)
    got = open(result).read()
    assert got.startswith("ok"), got
    if reduction == "nccl":
        assert got == "ok nccl"
    if reduction in ("fused", "p2p", "p2p-push", "p2p-allreduce"):
        assert got == "ok " + reduction  # symmetric memory is available on an NVLink box: no silent downgrade
    print(got)
