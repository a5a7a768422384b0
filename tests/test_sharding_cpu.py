"""Host-side logic of the multi-GPU path, on CPU: the i-entry split, the result-block layout and the
reduction over ranks (torch.distributed, gloo backend, world_size 2 and 3).  Each rank evaluates its
shard with the CPU oracle (test infrastructure), packs it into the result-block layout the library
uses, all-reduces, and every rank must end up with the full-list result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fepb200 import params as P
from fepb200.shard import ResultLayout, balanced_ranges, touched_atoms
from fepb200.synth import make_system, random_problem, scaled_spec

ALL = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA


def test_balanced_ranges_are_a_partition_and_balanced():
    prob = make_system(scaled_spec("C4", 4.2, 2, 25, n_foreign=2))
    jindex = prob.nblist.jindex
    for n in (1, 2, 3, 8, 64):
        r = balanced_ranges(jindex, n)
        assert r[0][0] == 0 and r[-1][1] == prob.nblist.nri
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        pairs = [int(jindex[e1] - jindex[e0]) for e0, e1 in r]
        assert sum(pairs) == prob.nblist.nrj
        if n <= 8:
            target = -(-prob.nblist.nrj // n)
            assert max(pairs) <= target + 64  # never more than one i-entry (<= 64 pairs) over target


def test_balanced_ranges_edge_cases():
    assert balanced_ranges(np.array([0]), 4) == [(0, 0)] * 4
    assert balanced_ranges(np.array([0, 5]), 3) == [(0, 1), (1, 1), (1, 1)]
    assert balanced_ranges(np.array([0, 0, 0, 7]), 2)[0][0] == 0


def test_result_layout_round_trip():
    rng = np.random.default_rng(0)
    prob = random_problem(1, P.make_params(), natoms=50, nri=8, n_foreign=3)
    touched = touched_atoms(prob.nblist)
    lay = ResultLayout(len(touched), prob.nenergrp_pairs, prob.n_foreign)
    out = dict(f=np.zeros((50, 3), np.float32), fshift=rng.normal(size=(45, 3)).astype(np.float32),
               Vc=rng.normal(size=4), Vv=rng.normal(size=4), dvdl=rng.normal(size=2),
               foreign_energy=rng.normal(size=4), foreign_dvdl=rng.normal(size=(4, 2)))
    out["f"][touched] = rng.normal(size=(len(touched), 3))
    back = lay.unpack(*lay.pack(out, touched), touched, 50)
    for k in out:
        assert np.array_equal(back[k], out[k]), k


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, root, result_file):
    for p in (os.path.join(root, "gromacs-fep-gpu_b200", "python"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    import copy

    from oracle import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = make_system(scaled_spec("C4", 4.2, 2, 25, n_foreign=3))
    touched = touched_atoms(prob.nblist)
    lay = ResultLayout(len(touched), prob.nenergrp_pairs, prob.n_foreign)
    e0, e1 = balanced_ranges(prob.nblist.jindex, world)[rank]
    shard = copy.copy(prob)
    shard.nblist = prob.nblist.slice_entries(e0, e1)
    mine = oracle.run_port(shard, ALL)
    f32, f64 = lay.pack(mine, touched)
    t32, t64 = torch.from_numpy(f32), torch.from_numpy(f64)
    dist.all_reduce(t64)
    dist.all_reduce(t32)
    got = lay.unpack(t32.numpy(), t64.numpy(), touched, prob.natoms)
    want = oracle.run_port(prob, ALL)
    ok = True
    for k in ("f", "fshift", "Vc", "Vv", "dvdl", "foreign_energy", "foreign_dvdl"):
        scale = max(np.max(np.abs(want[k])), 1e-12)
        tol = 2e-6 if k in ("f", "fshift") else 1e-12  # the fp32 block carries forces in float
        ok = ok and np.max(np.abs(got[k] - want[k])) <= tol * scale
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        with open(result_file, "w") as fh:
            fh.write("ok" if int(flag.item()) == 1 else "mismatch")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_reduction_matches_full_list_gloo(world, tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    result = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(world, _free_port(), root, result), nprocs=world, join=True)
    assert open(result).read() == "ok"
