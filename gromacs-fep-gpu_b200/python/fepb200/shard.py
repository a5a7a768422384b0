"""Split of the FEP pair list over ranks (one rank = one GPU).

Contiguous ranges of i-entries with about equal pair counts, found with the rule the reference
uses to balance its FEP lists over OpenMP threads (src/gromacs/nbnxm/pairlist.cpp:2786-2838):
walk the entries once and move on to the next destination when adding the entry would overshoot
the per-destination target by more than the current shortfall.  libfepb200 applies the same rule
inside fepb200_set_list(); this module exists so that host code and tests can predict the shards.
"""
from __future__ import annotations

import numpy as np


def balanced_ranges(jindex, nranks: int) -> list[tuple[int, int]]:
    jindex = np.asarray(jindex, np.int64)
    nri = jindex.shape[0] - 1
    first = [nri] * (nranks + 1)
    first[0] = 0
    total = int(jindex[-1]) if nri > 0 else 0
    target = (total + nranks - 1) // nranks
    dest, have = 0, 0
    for n in range(nri):
        nrj = int(jindex[n + 1] - jindex[n])
        if dest + 1 < nranks and have > 0 and have + nrj - target > target - have:
            dest += 1
            first[dest] = n
            have = 0
        have += nrj
    return [(first[r], first[r + 1]) for r in range(nranks)]


def touched_atoms(nblist) -> np.ndarray:
    """Ascending atom indices that occur anywhere in the FULL list -- the compact numbering every
    rank uses for the force part of the result block (same rule as fepb200_set_list(), which
    mirrors setReductionMaskFromFepPairlist, src/gromacs/nbnxm/freeenergydispatch.cpp:74-89)."""
    return np.unique(np.concatenate([np.asarray(nblist.iinr, np.int64), np.asarray(nblist.jjnr, np.int64)])).astype(np.int32)


# ---- the fused peer exchange (fepb200_set_peer_exchange): pairs by warp, atoms by contributions ----
PEER_ATOM_COST = 8  # the library's fixed share per atom when it balances the atom ranges


MAX_RUN_TRIPS = 8
PASS_WARPS_PER_SM = 32  # resident warps of the force-pass kernel: 8 CTAs of 4 warps


def run_trips_for(n_trips: int, sms: int = 148) -> int:
    """Trips per run, the rule of choose_run_trips() in csrc/fepb200_api.cu: the smallest power of two with which one
    wave of the pass kernel's resident warps covers the trips, at most MAX_RUN_TRIPS (148 SMs on a B200)."""
    r = 1
    while r < MAX_RUN_TRIPS and n_trips > sms * PASS_WARPS_PER_SM * r:
        r *= 2
    return r


def trip_layout(nblist, ngrp: int = 1, run_trips: int | None = None) -> dict:
    """Host-side mirror of the regrouping fepb200_set_list() does on the GPU (csrc/fep_list_build.cu): every pair is
    given to its OWNER, the end that takes part in more pairs of the list (ties: the i atom); pairs are grouped by
    (owner, energy-group pair, shift index, flipped = the owner was the j atom) with a stable sort, and every group is
    cut into TRIPS of at most 32 pairs, the unit of work of one warp.  The trips are cut into runs of `run_trips`
    (default: run_trips_for(n_trips)); a SEGMENT is a maximal sequence of trips of one group inside a run, and
    carries ONE force contribution to its owner.  Returns
      touched      compact -> atom
      trip_of_pair trip of every pair of the list (list order)
      n_trips, run_trips, n_segments
      trip_owner, trip_gid, trip_shift, trip_flipped   per trip (shift = the index the list gave)
      trip_last    the trip ends a segment
      atom_ptr     atom k of the compact numbering owns the force contributions [atom_ptr[k], atom_ptr[k+1]) of the
                   atom-sorted buffer: one per pair it is the partner of, one per segment it owns."""
    nrj = int(nblist.nrj)
    touched = touched_atoms(nblist)
    nt = len(touched)
    empty = np.zeros(0, np.int64)
    if nrj == 0:
        return dict(touched=touched, trip_of_pair=empty, n_trips=0, run_trips=run_trips or 1, n_segments=0, trip_owner=empty,
                    trip_gid=empty, trip_shift=empty, trip_flipped=empty.astype(bool), trip_last=empty.astype(bool),
                    atom_ptr=np.zeros(nt + 1, np.int64))
    natoms = int(touched[-1]) + 1
    compact = np.full(natoms, -1, np.int64)
    compact[touched] = np.arange(nt)
    ent = np.repeat(np.arange(nblist.nri), np.diff(np.asarray(nblist.jindex, np.int64)))
    ci = compact[np.asarray(nblist.iinr, np.int64)[ent]]
    cj = compact[np.asarray(nblist.jjnr, np.int64)]
    deg = np.bincount(ci, minlength=nt) + np.bincount(cj, minlength=nt)
    flip = deg[cj] > deg[ci]
    owner = np.where(flip, cj, ci)
    other = np.where(flip, ci, cj)
    gid = np.asarray(nblist.gid, np.int64)[ent]
    shift = np.asarray(nblist.shift, np.int64)[ent]
    key = ((owner * ngrp + gid) * 64 + shift) * 2 + flip
    order = np.argsort(key, kind="stable")
    ks = key[order]
    gstart = np.zeros(nrj, np.int64)
    heads = np.flatnonzero(np.concatenate([[True], ks[1:] != ks[:-1]]))
    gstart[heads] = heads
    gstart = np.maximum.accumulate(gstart)
    trip_head = (np.arange(nrj) - gstart) % 32 == 0
    trip_sorted = np.cumsum(trip_head) - 1
    n_trips = int(trip_sorted[-1]) + 1
    trip_of_pair = np.empty(nrj, np.int64)
    trip_of_pair[order] = trip_sorted
    first = order[trip_head]  # the pair that opens each trip
    if run_trips is None:
        run_trips = run_trips_for(n_trips)
    # a segment starts with every run and with every group
    group_head = np.zeros(nrj, bool)
    group_head[heads] = True
    seg_first = (np.arange(n_trips) % run_trips == 0) | group_head[np.flatnonzero(trip_head)]
    trip_last = np.concatenate([seg_first[1:], [True]])
    cnt = np.bincount(other, minlength=nt) + np.bincount(owner[first][trip_last], minlength=nt)
    return dict(touched=touched, trip_of_pair=trip_of_pair, n_trips=n_trips, run_trips=run_trips,
                n_segments=int(trip_last.sum()), trip_owner=owner[first], trip_gid=gid[first], trip_shift=shift[first],
                trip_flipped=flip[first], trip_last=trip_last,
                atom_ptr=np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64))


def peer_trip_ranges(n_trips: int, nranks: int, run_trips: int = 1) -> list[tuple[int, int]]:
    """Rank r evaluates the trips [begin, end): equal shares in whole runs, the rule of fepb200_set_peer_exchange()
    (which picks run_trips = run_trips_for(a rank's share))."""
    tpr = (n_trips + nranks - 1) // nranks
    tpr = (tpr + run_trips - 1) // run_trips * run_trips
    return [(min(r * tpr, n_trips), min((r + 1) * tpr, n_trips)) for r in range(nranks)]


def contribution_ranges(nblist, ngrp: int = 1) -> tuple[np.ndarray, np.ndarray]:
    """(touched atoms, atom_ptr) of trip_layout()."""
    lay = trip_layout(nblist, ngrp)
    return lay["touched"], lay["atom_ptr"]


def peer_atom_ranges(atom_ptr, nranks: int) -> list[tuple[int, int]]:
    """Rank r owns the compact atoms [begin, end): contiguous ranges with equal shares of
    (contributions + PEER_ATOM_COST per atom), the rule of fepb200_set_peer_exchange()."""
    atom_ptr = np.asarray(atom_ptr, np.int64)
    nt = atom_ptr.shape[0] - 1
    total = int(atom_ptr[-1]) + PEER_ATOM_COST * nt
    cost = atom_ptr[:-1] + PEER_ATOM_COST * np.arange(nt)  # cost of everything before atom a
    bounds = [0]
    for r in range(1, nranks):
        bounds.append(max(int(np.searchsorted(cost, total * r // nranks, side="left")), bounds[-1]))
    bounds.append(nt)
    return [(bounds[r], bounds[r + 1]) for r in range(nranks)]


def owned_atom_ranges(ntouched: int, nranks: int) -> list[tuple[int, int]]:
    """The compact atoms every rank owns after the force reduce-scatter: equal ranges starting on multiples of four atoms
    (so that a range of 3-word forces starts on a 16-byte boundary) -- the rule of fepb200_reduce_scatter_peers() and of
    the push targets (fepb200_set_push_targets: the epilogue sends the force of atom a to rank a // per)."""
    per = max(4, ((ntouched + nranks - 1) // nranks + 3) // 4 * 4)
    return [(min(ntouched, per * r), min(ntouched, per * (r + 1))) for r in range(nranks)]


def push_block_addresses(base: list[int], rank: int, block_bytes: int) -> tuple[list[list[int]], list[list[int]], list[int]]:
    """Push reduction over symmetric memory: every rank's buffer (base[r] = its address as mapped HERE) holds two slots of
    len(base) receive blocks -- block s of a slot is written by rank s only -- followed by the flag array.  Returns
    (push, recv, flags): push[k][r] = MY block of slot k on rank r, recv[k][s] = the block rank s writes in my slot k,
    flags[r] = rank r's flag array."""
    world = len(base)
    slot_bytes = world * block_bytes
    push = [[b + k * slot_bytes + rank * block_bytes for b in base] for k in (0, 1)]
    recv = [[base[rank] + k * slot_bytes + s * block_bytes for s in range(world)] for k in (0, 1)]
    flags = [b + 2 * slot_bytes for b in base]
    return push, recv, flags


class ResultLayout:
    """Host-side mirror of `struct fepb200_layout`: where forces, shift forces, energy-group
    energies, dV/dlambda and foreign-lambda terms sit in the fp32 / fp64 result blocks."""

    def __init__(self, ntouched: int, nenergrp: int, nforeign: int):
        self.ntouched, self.nenergrp, self.nforeign = ntouched, nenergrp, nforeign
        self.f32_words = 3 * ntouched + 3 * 45
        self.off_fshift = 3 * ntouched
        self.off_vc, self.off_vv = 0, nenergrp
        self.off_dvdl = 2 * nenergrp
        self.off_foreign_e = 2 * nenergrp + 2
        self.off_foreign_dvdl = 2 * nenergrp + 2 + (nforeign + 1)
        self.f64_words = 2 * nenergrp + 2 + 3 * (nforeign + 1)

    def pack(self, out: dict, touched: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        """Result dict (full-size arrays) -> (fp32 block, fp64 block)."""
        f32 = np.zeros(self.f32_words, np.float32)
        f64 = np.zeros(self.f64_words, np.float64)
        f32[: self.off_fshift] = np.asarray(out["f"], np.float32)[touched].ravel()
        f32[self.off_fshift :] = np.asarray(out["fshift"], np.float32).ravel()
        f64[self.off_vc : self.off_vc + self.nenergrp] = out["Vc"]
        f64[self.off_vv : self.off_vv + self.nenergrp] = out["Vv"]
        f64[self.off_dvdl : self.off_dvdl + 2] = out["dvdl"]
        f64[self.off_foreign_e : self.off_foreign_e + self.nforeign + 1] = out["foreign_energy"]
        f64[self.off_foreign_dvdl :] = np.asarray(out["foreign_dvdl"]).ravel()
        return f32, f64

    def unpack(self, f32: np.ndarray, f64: np.ndarray, touched: np.ndarray, natoms: int) -> dict:
        f = np.zeros((natoms, 3), np.float32)
        f[touched] = f32[: self.off_fshift].reshape(-1, 3)
        return dict(
            f=f,
            fshift=f32[self.off_fshift :].reshape(45, 3).copy(),
            Vc=f64[self.off_vc : self.off_vc + self.nenergrp].copy(),
            Vv=f64[self.off_vv : self.off_vv + self.nenergrp].copy(),
            dvdl=f64[self.off_dvdl : self.off_dvdl + 2].copy(),
            foreign_energy=f64[self.off_foreign_e : self.off_foreign_e + self.nforeign + 1].copy(),
            foreign_dvdl=f64[self.off_foreign_dvdl :].reshape(-1, 2).copy(),
        )
