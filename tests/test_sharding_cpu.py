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
from fepb200.shard import ResultLayout, balanced_ranges, owned_atom_ranges, push_block_addresses, touched_atoms
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


# ---- the fused peer exchange (fepb200_set_peer_exchange): host-side mirror of its split ----------
def test_trip_layout_and_peer_ranges_are_partitions():
    from fepb200.shard import peer_atom_ranges, peer_trip_ranges, trip_layout

    prob = make_system(scaled_spec("C4", 4.2, 2, 25, n_foreign=2))
    nb = prob.nblist
    lay = trip_layout(nb, prob.nenergrp_pairs)
    touched, atom_ptr = lay["touched"], lay["atom_ptr"]
    assert np.array_equal(touched, touched_atoms(nb)) and atom_ptr[0] == 0
    assert np.all(np.diff(atom_ptr) > 0)  # every touched atom receives at least one contribution
    assert atom_ptr[-1] == nb.nrj + lay["n_segments"]  # one per pair (partner) + one per segment (owner)
    assert lay["run_trips"] == 1 and lay["n_segments"] == lay["n_trips"]  # a list this small: every trip its own run
    for rt in (2, 4, 8):
        l2 = trip_layout(nb, prob.nenergrp_pairs, rt)
        last = l2["trip_last"]
        assert last[-1] and np.all(last[rt - 1 :: rt])  # every run ends a segment
        key = np.stack([l2["trip_owner"], l2["trip_gid"], l2["trip_shift"], l2["trip_flipped"]], 1)
        change = np.any(key[1:] != key[:-1], axis=1)
        assert np.all(last[:-1][change])  # so does every change of group
        assert lay["n_trips"] / rt <= l2["n_segments"] <= lay["n_segments"]
        assert l2["atom_ptr"][-1] == nb.nrj + l2["n_segments"]
    # every pair sits in exactly one trip of at most 32 pairs, and a trip is uniform in (owner, gid, shift, side)
    top = lay["trip_of_pair"]
    sizes = np.bincount(top, minlength=lay["n_trips"])
    assert sizes.min() >= 1 and sizes.max() <= 32
    ent = np.repeat(np.arange(nb.nri), np.diff(nb.jindex))
    compact = np.full(prob.natoms, -1)
    compact[touched] = np.arange(len(touched))
    ci, cj = compact[nb.iinr[ent]], compact[nb.jjnr]
    own = lay["trip_owner"][top]
    flipped = lay["trip_flipped"][top]
    assert np.all(np.where(flipped, cj, ci) == own)
    assert np.array_equal(nb.gid[ent], lay["trip_gid"][top]) and np.array_equal(nb.shift[ent], lay["trip_shift"][top])
    # the regrouping pays: far fewer trips than the i-entry-major layout has runs, and they are nearly full
    assert lay["n_trips"] < 1.35 * nb.nrj / 32 + 8 * prob.nenergrp_pairs
    for n in (1, 2, 3, 8):
        pr = peer_trip_ranges(lay["n_trips"], n)
        assert pr[0][0] == 0 and pr[-1][1] == lay["n_trips"] and all(a[1] == b[0] for a, b in zip(pr, pr[1:]))
        assert max(b - a for a, b in pr) - min(b - a for a, b in pr[:-1] or pr) <= n
        ar = peer_atom_ranges(atom_ptr, n)
        assert ar[0][0] == 0 and ar[-1][1] == len(touched) and all(a[1] == b[0] for a, b in zip(ar, ar[1:]))
        cost = [atom_ptr[b] - atom_ptr[a] + 8 * (b - a) for a, b in ar]
        heaviest = int(np.max(np.diff(atom_ptr))) + 8
        assert max(cost) - min(cost) <= 2 * heaviest  # balanced up to one atom at either end
    # degenerate inputs
    assert peer_trip_ranges(0, 4) == [(0, 0)] * 4
    assert peer_trip_ranges(5, 4) == [(0, 2), (2, 4), (4, 5), (5, 5)]
    assert peer_trip_ranges(21, 3, 4) == [(0, 8), (8, 16), (16, 21)]


def test_select_pairs_keeps_order_and_reassembles():
    prob = make_system(scaled_spec("C2", 3.2, 1, 20, n_foreign=1))
    nb = prob.nblist
    rng = np.random.default_rng(3)
    part = rng.integers(0, 3, nb.nrj)
    ent = np.repeat(np.arange(nb.nri), np.diff(nb.jindex))
    seen = np.zeros(nb.nrj, int)
    for r in range(3):
        s = nb.select_pairs(part == r)
        assert s.nrj == int(np.sum(part == r)) and np.all(np.diff(s.jindex) > 0)
        assert np.array_equal(s.jjnr, nb.jjnr[part == r]) and np.array_equal(s.excl_fep, nb.excl_fep[part == r])
        assert np.array_equal(np.repeat(s.iinr, np.diff(s.jindex)), nb.iinr[ent[part == r]])
        seen[part == r] += 1
    assert np.all(seen == 1)
    assert nb.select_pairs(np.zeros(nb.nrj, bool)).nri == 0


def _fused_worker(rank, world, port, root, result_file):
    """Every rank evaluates ITS pairs (CPU oracle = test infrastructure), the owner of an atom range
    ends up with the sum of all ranks' contributions to its atoms (reduce-scatter), every rank with
    the summed scalars (all-reduce): the data flow of the fused exchange, on gloo."""
    for p in (os.path.join(root, "gromacs-fep-gpu_b200", "python"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    import copy

    from fepb200.shard import peer_atom_ranges, peer_trip_ranges, trip_layout
    from oracle import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = make_system(scaled_spec("C4", 4.2, 2, 25, n_foreign=3))
    lay = trip_layout(prob.nblist, prob.nenergrp_pairs)
    touched, atom_ptr = lay["touched"], lay["atom_ptr"]
    t0, t1 = peer_trip_ranges(lay["n_trips"], world)[rank]
    a0, a1 = peer_atom_ranges(atom_ptr, world)[rank]
    shard = copy.copy(prob)
    shard.nblist = prob.nblist.select_pairs((lay["trip_of_pair"] >= t0) & (lay["trip_of_pair"] < t1))
    mine = oracle.run_port(shard, ALL)
    want = oracle.run_port(prob, ALL)
    # forces: contributions of every rank to the atoms this rank owns
    f = torch.from_numpy(np.asarray(mine["f"], np.float64).copy())
    dist.all_reduce(f)
    owned = np.zeros(prob.natoms, bool)
    owned[touched[a0:a1]] = True
    f_owned = np.where(owned[:, None], f.numpy(), 0.0)
    gathered = torch.from_numpy(f_owned.copy())
    dist.all_reduce(gathered)  # the union of the owners' slices is the full force array
    ok = np.max(np.abs(gathered.numpy() - want["f"])) <= 1e-10 * np.max(np.abs(want["f"]))
    ok = ok and not np.any(f_owned[~owned])
    for k in ("fshift", "Vc", "Vv", "dvdl", "foreign_energy", "foreign_dvdl"):
        t = torch.from_numpy(np.asarray(mine[k], np.float64).copy())
        dist.all_reduce(t)
        ok = ok and np.max(np.abs(t.numpy() - want[k])) <= 1e-10 * max(np.max(np.abs(want[k])), 1e-12)
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        with open(result_file, "w") as fh:
            fh.write("ok" if int(flag.item()) == 1 else "mismatch")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_fused_exchange_data_flow_gloo(world, tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    result = str(tmp_path / "result.txt")
    mp.spawn(_fused_worker, args=(world, _free_port(), root, result), nprocs=world, join=True)
    assert open(result).read() == "ok"


def test_owned_atom_ranges_and_push_addresses():
    """Host mirrors of the push reduction's rules: the owned ranges are a partition on 4-atom boundaries, the owner of an
    atom is a // per, and the block rank s pushes into on rank r is the block rank r reads as 'written by s'."""
    for nT, n in ((75818, 8), (75818, 2), (10, 4), (3, 8), (0, 2), (4097, 3)):
        rg = owned_atom_ranges(nT, n)
        assert rg[0][0] == 0 and rg[-1][1] == nT and all(a1 == b0 for (_, a1), (b0, _) in zip(rg, rg[1:]))
        assert all(a0 % 4 == 0 or a0 == nT for a0, _ in rg)
        per = max(4, ((nT + n - 1) // n + 3) // 4 * 4)
        for a in range(0, nT, max(1, nT // 97)):
            r = min(a // per, n - 1)
            assert rg[r][0] <= a < rg[r][1]
    world, bb = 4, 4096
    # base[r] as mapped on rank q: every rank sees the same buffers at its own virtual addresses
    mapped = [[(q + 1) * 10**9 + r * 10**6 for r in range(world)] for q in range(world)]
    plans = [push_block_addresses(mapped[q], q, bb) for q in range(world)]
    for s_ in range(world):          # the writer
        for r in range(world):       # the owner of the memory
            for k in (0, 1):
                off_w = plans[s_][0][k][r] - mapped[s_][r]   # where s writes inside r's buffer
                off_r = plans[r][1][k][s_] - mapped[r][r]    # where r reads 'the block of s' inside its own buffer
                assert off_w == off_r
                assert 0 <= off_w and off_w + bb <= 2 * world * bb
    # no two writers share a block, the two slots do not overlap, the flags follow the slots
    offs = sorted(plans[s_][0][k][0] - mapped[s_][0] for s_ in range(world) for k in (0, 1))
    assert offs == [i * bb for i in range(2 * world)]
    assert plans[0][2][1] - mapped[0][1] == 2 * world * bb


def _push_worker(rank, world, port, root, result_file):
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
    nT = len(touched)
    lay = ResultLayout(nT, prob.nenergrp_pairs, prob.n_foreign)
    e0, e1 = balanced_ranges(prob.nblist.jindex, world)[rank]
    shard = copy.copy(prob)
    shard.nblist = prob.nblist.slice_entries(e0, e1)
    f32, f64 = lay.pack(oracle.run_port(shard, ALL), touched)
    ranges = owned_atom_ranges(nT, world)
    # what the epilogue does: to rank r the forces of r's atoms (nothing else of the force part), shift forces and
    # scalars to everybody; the receive blocks start zeroed
    send = []
    for r in range(world):
        b32 = np.zeros_like(f32)
        a0, a1 = ranges[r]
        b32[3 * a0 : 3 * a1] = f32[3 * a0 : 3 * a1]
        b32[3 * nT :] = f32[3 * nT :]
        send.append((torch.from_numpy(b32), torch.from_numpy(f64.copy())))
    recv32 = [torch.zeros(f32.size, dtype=torch.float32) for _ in range(world)]
    recv64 = [torch.zeros(f64.size, dtype=torch.float64) for _ in range(world)]
    for s_ in range(world):  # rank s_ "pushes": every rank receives its block from s_
        dist.scatter(recv32[s_], [t[0] for t in send] if rank == s_ else None, src=s_)
        dist.scatter(recv64[s_], [t[1] for t in send] if rank == s_ else None, src=s_)
    # the reduction kernel on the rank's own receive blocks: sums in rank order
    sum32, sum64 = np.zeros_like(f32), np.zeros_like(f64)
    for s_ in range(world):
        sum32 += recv32[s_].numpy()
        sum64 += recv64[s_].numpy()
    got = lay.unpack(sum32, sum64, touched, prob.natoms)
    want = oracle.run_port(prob, ALL)
    a0, a1 = ranges[rank]
    owned = np.zeros(prob.natoms, bool)
    owned[touched[a0:a1]] = True
    ok = not np.any(got["f"][~owned])
    scale = max(np.max(np.abs(want["f"])), 1e-12)
    ok = ok and np.max(np.abs(got["f"][owned] - want["f"][owned])) <= 2e-6 * scale
    for k in ("fshift", "Vc", "Vv", "dvdl", "foreign_energy", "foreign_dvdl"):
        sc = max(np.max(np.abs(want[k])), 1e-12)
        ok = ok and np.max(np.abs(got[k] - want[k])) <= (2e-6 if k == "fshift" else 1e-12) * sc
    # the ranks' owned forces add up to the whole array
    t = torch.from_numpy(got["f"].astype(np.float64))
    dist.all_reduce(t)
    ok = ok and np.max(np.abs(t.numpy() - want["f"])) <= 2e-6 * scale
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        with open(result_file, "w") as fh:
            fh.write("ok" if int(flag.item()) == 1 else "mismatch")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_push_reduction_data_flow_gloo(world, tmp_path):
    """The data flow of the push reduction (fepb200_set_push_targets + fepb200_reduce_scatter_peers on the receive blocks)
    played on CPU: shards evaluated by the oracle, blocks moved with gloo, owner rule and block layout from fepb200.shard."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    result = str(tmp_path / "result.txt")
    mp.spawn(_push_worker, args=(world, _free_port(), root, result), nprocs=world, join=True)
    assert open(result).read() == "ok"
