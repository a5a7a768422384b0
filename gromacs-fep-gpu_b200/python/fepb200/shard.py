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


def peer_pair_ranges(nrj: int, nranks: int) -> list[tuple[int, int]]:
    """Rank r evaluates the pairs [begin, end) of the flat pair space: equal shares of its 32-pair warps."""
    n_warps = (nrj + 31) // 32
    wpr = (n_warps + nranks - 1) // nranks
    return [(min(r * wpr * 32, nrj), min((r + 1) * wpr * 32, nrj)) for r in range(nranks)]


def contribution_ranges(nblist) -> tuple[np.ndarray, np.ndarray]:
    """(touched atoms, atom_ptr): atom k of the compact numbering owns the force contributions
    [atom_ptr[k], atom_ptr[k+1]) of the atom-sorted buffer -- one per pair it is the j atom of, and
    one per segment (maximal run of pairs of one i-entry inside one 32-pair warp) it is the i atom of."""
    nrj = int(nblist.nrj)
    touched = touched_atoms(nblist)
    if nrj == 0:
        return touched, np.zeros(len(touched) + 1, np.int64)
    ent = np.repeat(np.arange(nblist.nri), np.diff(np.asarray(nblist.jindex, np.int64)))
    warp = np.arange(nrj) // 32
    head = np.ones(nrj, bool)
    head[1:] = (ent[1:] != ent[:-1]) | (warp[1:] != warp[:-1])
    natoms = int(max(np.max(nblist.iinr), np.max(nblist.jjnr))) + 1
    cnt = np.bincount(np.asarray(nblist.jjnr, np.int64), minlength=natoms)
    cnt = cnt + np.bincount(np.asarray(nblist.iinr, np.int64)[ent[head]], minlength=natoms)
    return touched, np.concatenate([[0], np.cumsum(cnt[touched])]).astype(np.int64)


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
