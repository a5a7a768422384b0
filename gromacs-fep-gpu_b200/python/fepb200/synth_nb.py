"""Synthetic inputs for the non-perturbed neighbour of the FEP path: the atoms of a
`Problem` put into the reference's GPU cluster layout, and a cluster pair list in the reference's GPU format.

Pair search is NOT part of the product (out of scope) -- this is the input generator of the tests and of
bench.py, the counterpart of `synth.build_fep_list`.  It produces what the reference's search hands to its GPU kernels:

  * atoms in grid order, padded with filler atoms to whole 8-atom clusters / 64-atom super-clusters: columns in xy,
    sorted along z, every 64 atoms split 2 x 2 x 2 along z, y, x (the scheme of src/gromacs/nbnxm/grid.cpp);
    `atom_index[slot]` = original atom or -1 (GridSet::atomIndices());
  * perturbed atoms masked: charge 0, type ntype-1 (nbnxn_atomdata_mask_fep, src/gromacs/nbnxm/atomdata.cpp:930-964);
  * `sci[nsci]` {sci, shift, cjPackedBegin, cjPackedEnd}, `cj[ncj]` {cj[4], {imask, excl_ind}[2]}, `excl[nexcl]`
    {pair[32]} with the bit conventions of src/gromacs/nbnxm/pairlist.h:195-280: imask bit jm*8+im = i-cluster im of the
    super-cluster interacts with j-cluster jm of the packed entry; exclusion word (jj%4)*8+ii of half jj/4, same bit;
    excl[0] = all pairs interact.  Half list: central shift with cj >= ci, otherwise only shift indices > 22.
  * topology exclusions (water molecules, bonded neighbours in the blobs) and every pair that is in the FEP list have
    their interaction bit cleared, as make_fep_list does (src/gromacs/nbnxm/pairlist.cpp:1776-1942).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

CL = 8  # atoms per cluster (pairlistparams.h:65)
NCL_SC = 8  # clusters per super-cluster (pairlist.h:174)
JGROUP = 4  # j-clusters per packed entry (pairlist.h:180)
CENTRAL = 22
FAR = 1.0e4  # nm; filler atoms sit far outside every cut-off, each at its own place

SCI_DTYPE = np.dtype([("sci", "<i4"), ("shift", "<i4"), ("cj_begin", "<i4"), ("cj_end", "<i4")])
CJ_DTYPE = np.dtype([("cj", "<i4", (JGROUP,)), ("imask0", "<u4"), ("excl0", "<i4"), ("imask1", "<u4"), ("excl1", "<i4")])
EXCL_DTYPE = np.dtype([("pair", "<u4", (32,))])


@dataclass
class ClusterSystem:
    natoms: int  # padded, multiple of 64
    atom_index: np.ndarray  # int32 [natoms]: original atom or -1
    slot_of_atom: np.ndarray  # int32 [n original atoms]
    xq: np.ndarray  # float32 [natoms, 4], charges masked
    type: np.ndarray  # int32 [natoms], types masked
    ntype: int
    nbfp: np.ndarray  # float32 [2 ntype^2]
    shiftvec: np.ndarray  # float32 [45, 3]
    sci: np.ndarray  # SCI_DTYPE
    cj: np.ndarray  # CJ_DTYPE
    excl: np.ndarray  # EXCL_DTYPE
    rlist: float
    perturbed_slots: np.ndarray  # int32, slots of the perturbed atoms
    q_unmasked: np.ndarray  # float32 [natoms]
    type_unmasked: np.ndarray  # int32 [natoms]
    pairs_in_cutoff: int = 0  # atom pairs of the list inside rcoulomb (real atoms, each pair once): the unit of the flop count

    @property
    def n_cluster_pairs(self) -> int:
        m = self.cj["imask0"].astype(np.uint64)
        return int(sum(int(((m >> b) & 1).sum()) for b in range(32)))


def grid_order(x: np.ndarray, box: float, density_hint: float | None = None):
    """Slots of the atoms in a GROMACS-like cluster grid.  Returns atom_index [natoms_padded] (-1 = filler)."""
    n = x.shape[0]
    dens = density_hint if density_hint else n / box**3
    a = (CL * NCL_SC / dens) ** (1.0 / 3.0)
    nc = max(1, int(round(box / a)))
    cx = np.minimum((x[:, 0] / box * nc).astype(np.int64), nc - 1)
    cy = np.minimum((x[:, 1] / box * nc).astype(np.int64), nc - 1)
    col = cx * nc + cy
    order = np.lexsort((x[:, 2], col))
    counts = np.bincount(col, minlength=nc * nc)
    padded = (counts + CL * NCL_SC - 1) // (CL * NCL_SC) * (CL * NCL_SC)
    start_p = np.concatenate([[0], np.cumsum(padded)])
    start = np.concatenate([[0], np.cumsum(counts)])
    natoms = int(start_p[-1])
    atom_index = np.full(natoms, -1, np.int64)
    col_sorted = col[order]
    rank_in_col = np.arange(n) - start[col_sorted]
    atom_index[start_p[col_sorted] + rank_in_col] = order
    # inside every 64 atoms (sorted along z): halves along z, each sorted along y and halved, each sorted along x
    big = np.float64(1e30)
    xs = np.where(atom_index[:, None] >= 0, x[np.maximum(atom_index, 0)].astype(np.float64), big)
    for dim, width in ((1, 32), (0, 16)):
        blk = atom_index.reshape(-1, width)
        key = xs[:, dim].reshape(-1, width)
        o = np.argsort(key, axis=1, kind="stable")
        atom_index = np.take_along_axis(blk, o, axis=1).reshape(-1)
        xs = np.take_along_axis(xs.reshape(-1, width, 3), o[:, :, None], axis=1).reshape(-1, 3)
    return atom_index.astype(np.int32)


def _shift_index(k):
    return 5 * (3 * (k[:, 2] + 1) + (k[:, 1] + 1)) + (k[:, 0] + 2)


def build_cluster_system(problem, rlist: float = 1.1, extra_excluded_pairs=None, max_cj_groups_per_sci: int = 0,
                         water_first_atom: int | None = None) -> ClusterSystem:
    """`problem`: a fepb200 Problem from synth.make_system (solute atoms first, then 3-site waters)."""
    from scipy.spatial import cKDTree

    box = float(problem.box[0, 0])
    x = np.asarray(problem.x, np.float32)
    n = x.shape[0]
    perturbed = np.asarray(problem.perturbed, np.int64)
    n_sol = int(perturbed.size) if water_first_atom is None else int(water_first_atom)
    atom_index = grid_order(x.astype(np.float64), box)
    natoms = atom_index.shape[0]
    real = atom_index >= 0
    slot_of_atom = np.empty(n, np.int32)
    slot_of_atom[atom_index[real]] = np.nonzero(real)[0]

    xq = np.zeros((natoms, 4), np.float32)
    fill = np.nonzero(~real)[0]
    xq[fill, 0] = -FAR - 3.0 * np.arange(fill.size)
    xq[fill, 1] = -FAR
    xq[fill, 2] = -FAR
    xq[real, :3] = x[atom_index[real]]
    xq[real, 3] = np.asarray(problem.qA, np.float32)[atom_index[real]]
    typ = np.full(natoms, problem.ntype - 1, np.int32)
    typ[real] = problem.typeA[atom_index[real]]
    q_unmasked, type_unmasked = xq[:, 3].copy(), typ.copy()
    pslots = slot_of_atom[perturbed].astype(np.int32)
    xq[pslots, 3] = 0.0
    typ[pslots] = problem.ntype - 1

    # ---- cluster pairs within rlist (exact atom distances), half list with shifts -------------------------------
    ncl = natoms // CL
    xc = xq[:, :3].astype(np.float64).reshape(ncl, CL, 3)
    rc = real.reshape(ncl, CL)
    cnt = rc.sum(axis=1)
    live = np.nonzero(cnt > 0)[0]
    centre = np.where(rc[live, :, None], xc[live], 0.0).sum(axis=1) / cnt[live, None]
    radius = np.sqrt(np.where(rc[live], ((xc[live] - centre[:, None, :]) ** 2).sum(axis=2), 0.0).max(axis=1))
    cw = np.mod(centre, box)
    cw[cw >= box] = 0.0
    tree = cKDTree(cw, boxsize=box)
    cand = tree.query_pairs(rlist + 2.0 * float(radius.max()), output_type="ndarray")
    c1 = np.concatenate([live[cand[:, 0]], live])  # plus every cluster against itself
    c2 = np.concatenate([live[cand[:, 1]], live])
    ctr = np.zeros((ncl, 3))
    ctr[live] = centre
    k = np.rint((ctr[c2] - ctr[c1]) / box).astype(np.int64)  # image of c1 closest to c2
    keep = np.zeros(c1.shape[0], bool)
    rl2 = rlist * rlist
    rc2 = float(problem.params.rcoulomb) ** 2
    upper = np.triu(np.ones((CL, CL), bool), 1)
    n_in_cutoff = 0
    for lo in range(0, c1.shape[0], 1 << 18):
        hi = min(lo + (1 << 18), c1.shape[0])
        d = xc[c1[lo:hi]][:, :, None, :] + (k[lo:hi] * box)[:, None, None, :] - xc[c2[lo:hi]][:, None, :, :]
        r2 = (d * d).sum(axis=3)
        ok = rc[c1[lo:hi]][:, :, None] & rc[c2[lo:hi]][:, None, :]
        r2 = np.where(ok, r2, np.inf)
        keep[lo:hi] = r2.min(axis=(1, 2)) < rl2
        same = (c1[lo:hi] == c2[lo:hi])[:, None, None]
        n_in_cutoff += int(((r2 < rc2) & (~same | upper[None])).sum())
    c1, c2, k = c1[keep], c2[keep], k[keep]
    sh = _shift_index(k)
    # orientation: central shift -> cj >= ci; otherwise the shift index must be > 22
    swap = (sh < CENTRAL) | ((sh == CENTRAL) & (c2 < c1))
    ci = np.where(swap, c2, c1)
    cjn = np.where(swap, c1, c2)
    sh = np.where(sh < CENTRAL, 2 * CENTRAL - sh, sh)

    # ---- group by (super-cluster, shift, j-cluster): 8-bit i-cluster masks --------------------------------------
    key = (ci // NCL_SC * 45 + sh) * ncl + cjn
    order = np.argsort(key, kind="stable")
    key, ci = key[order], ci[order]
    first = np.concatenate([[True], key[1:] != key[:-1]])
    starts = np.nonzero(first)[0]
    mask8 = np.bitwise_or.reduceat((1 << (ci % NCL_SC)).astype(np.uint32), starts)
    ukey = key[starts]
    e_key = ukey // ncl  # sci * 45 + shift
    e_cj = (ukey % ncl).astype(np.int32)
    e_first = np.concatenate([[True], e_key[1:] != e_key[:-1]])
    e_id = np.cumsum(e_first) - 1
    e_start = np.nonzero(e_first)[0]
    pos = np.arange(ukey.shape[0]) - e_start[e_id]  # position of the j-cluster in its entry
    n_in_entry = np.diff(np.concatenate([e_start, [ukey.shape[0]]]))
    ngrp = (n_in_entry + JGROUP - 1) // JGROUP
    grp_begin = np.concatenate([[0], np.cumsum(ngrp)])
    ncj = int(grp_begin[-1])
    cj = np.zeros(ncj, CJ_DTYPE)
    g = grp_begin[e_id] + pos // JGROUP
    jm = pos % JGROUP
    cj["cj"][g, jm] = e_cj
    np.bitwise_or.at(cj["imask0"], g, (mask8.astype(np.uint32) << (jm * NCL_SC).astype(np.uint32)))
    # padding slots of a packed entry repeat a valid cluster index with an empty mask
    padslot = cj["imask0"][:, None] >> (np.arange(JGROUP, dtype=np.uint32) * NCL_SC)[None, :] & 0xFF
    cj["cj"] = np.where(padslot != 0, cj["cj"], cj["cj"][:, :1])
    cj["imask1"] = cj["imask0"]

    ek = e_key[e_start]
    if max_cj_groups_per_sci > 0:
        pieces = [(kk, b + o, min(b + o + max_cj_groups_per_sci, e))
                  for kk, b, e in zip(ek, grp_begin[:-1], grp_begin[1:]) for o in range(0, e - b, max_cj_groups_per_sci)]
        ek = np.array([p[0] for p in pieces])
        gb = np.array([p[1] for p in pieces])
        ge = np.array([p[2] for p in pieces])
    else:
        gb, ge = grp_begin[:-1], grp_begin[1:]
    sci = np.zeros(ek.shape[0], SCI_DTYPE)
    sci["sci"], sci["shift"], sci["cj_begin"], sci["cj_end"] = ek // 45, ek % 45, gb, ge

    # ---- interaction bits ----------------------------------------------------------------------------------------
    pairs = []
    if n > n_sol:  # the three pairs inside every water molecule
        w0 = np.arange(n_sol, n, 3)
        pairs += [np.stack([w0, w0 + 1], 1), np.stack([w0, w0 + 2], 1), np.stack([w0 + 1, w0 + 2], 1)]
    fl = problem.nblist  # everything the FEP kernel computes is excluded here
    if fl.nrj:
        pairs.append(np.stack([np.repeat(fl.iinr, np.diff(fl.jindex)), fl.jjnr], 1))
    if extra_excluded_pairs is not None and len(extra_excluded_pairs):
        pairs.append(np.asarray(extra_excluded_pairs, np.int64).reshape(-1, 2))
    ex = np.concatenate(pairs).astype(np.int64) if pairs else np.zeros((0, 2), np.int64)
    ex = ex[ex[:, 0] != ex[:, 1]]
    sa, sb = slot_of_atom[ex[:, 0]].astype(np.int64), slot_of_atom[ex[:, 1]].astype(np.int64)
    xa, xb = xq[sa, :3].astype(np.float64), xq[sb, :3].astype(np.float64)
    kk = np.rint((xb - xa) / box).astype(np.int64)
    shx = _shift_index(kk)
    ca, cb = sa // CL, sb // CL
    swap = (shx < CENTRAL) | ((shx == CENTRAL) & ((cb < ca) | ((cb == ca) & (sb < sa))))
    si, sj = np.where(swap, sb, sa), np.where(swap, sa, sb)
    shx = np.where(shx < CENTRAL, 2 * CENTRAL - shx, shx)
    # the intra-cluster lower triangle of the central cell (real lists clear it; the kernels skip it anyway)
    diag = np.nonzero((sci["shift"] == CENTRAL))[0]
    want = ((si // CL // NCL_SC) * 45 + shx) * ncl + sj // CL
    idx = np.searchsorted(ukey, want)
    idx = np.minimum(idx, ukey.shape[0] - 1)
    hit = ukey[idx] == want
    si, sj, idx = si[hit], sj[hit], idx[hit]
    grp = grp_begin[e_id[idx]] + pos[idx] // JGROUP
    bit = (pos[idx] % JGROUP) * NCL_SC + (si // CL) % NCL_SC
    # only cluster pairs that are in the list (their imask bit is set) can carry exclusions
    listed = (cj["imask0"][grp] >> bit.astype(np.uint32)) & 1 == 1
    si, sj, grp, bit = si[listed], sj[listed], grp[listed], bit[listed]
    excl_groups = np.unique(grp)
    excl = np.zeros(1 + 2 * excl_groups.shape[0], EXCL_DTYPE)
    excl["pair"][:] = 0xFFFFFFFF
    slot_of_group = np.full(ncj, -1, np.int64)
    slot_of_group[excl_groups] = np.arange(excl_groups.shape[0])
    cj["excl0"][excl_groups] = 1 + 2 * np.arange(excl_groups.shape[0])
    cj["excl1"][excl_groups] = 2 + 2 * np.arange(excl_groups.shape[0])
    ii, jj = si % CL, sj % CL
    e_idx = 1 + 2 * slot_of_group[grp] + jj // (CL // 2)
    word = (jj % (CL // 2)) * CL + ii
    flat = excl["pair"].reshape(-1)
    np.bitwise_and.at(flat, e_idx * 32 + word, ~(np.uint32(1) << bit.astype(np.uint32)))
    del diag

    return ClusterSystem(natoms=natoms, atom_index=atom_index, slot_of_atom=slot_of_atom, xq=xq, type=typ,
                         ntype=problem.ntype, nbfp=np.asarray(problem.nbfp, np.float32), shiftvec=problem.shiftvec,
                         sci=sci, cj=cj, excl=excl, rlist=rlist, perturbed_slots=pslots, q_unmasked=q_unmasked,
                         type_unmasked=type_unmasked, pairs_in_cutoff=n_in_cutoff)


def fep_list_in_slots(problem, cs: ClusterSystem):
    """The problem's FEP list with atom indices replaced by grid slots (the index space of the fork's GPU route,
    nbnxm_gpu_data_mgmt.cpp:763-787), and charges / types of both states in slot order."""
    from .problem import FepList

    fl = problem.nblist
    s = cs.slot_of_atom
    out = FepList(s[fl.iinr], fl.gid, fl.shift, fl.jindex, s[fl.jjnr], fl.excl_fep)
    real = cs.atom_index >= 0
    ai = cs.atom_index[real]

    def slots(a, fill):
        o = np.full(cs.natoms, fill, np.asarray(a).dtype)
        o[real] = np.asarray(a)[ai]
        return o

    return out, slots(problem.qA, 0), slots(problem.qB, 0), slots(problem.typeA, problem.ntype - 1), slots(
        problem.typeB, problem.ntype - 1)


def brute_force(cs: ClusterSystem, params, excluded_slot_pairs, box: float, want_ewald: bool):
    """All-pairs reference for SMALL systems (minimum image, box > 2 rc): forces and energies of the masked system
    with the given excluded pairs, to validate list + kernel together.  float64."""
    real = np.nonzero(cs.atom_index >= 0)[0]
    x = cs.xq[real, :3].astype(np.float64)
    q = cs.xq[real, 3].astype(np.float64)
    t = cs.type[real]
    m = real.shape[0]
    pos = np.full(cs.natoms, -1, np.int64)
    pos[real] = np.arange(m)
    exm = np.zeros((m, m), bool)
    if len(excluded_slot_pairs):
        e = np.asarray(excluded_slot_pairs, np.int64)
        exm[pos[e[:, 0]], pos[e[:, 1]]] = True
        exm[pos[e[:, 1]], pos[e[:, 0]]] = True
    nb = np.asarray(cs.nbfp, np.float64).reshape(cs.ntype, cs.ntype, 2)
    f = np.zeros((m, 3))
    vc = vv = 0.0
    from scipy.special import erf

    beta = float(params.ewaldcoeff_q)
    for i in range(m - 1):
        d = x[i] - x[i + 1:]
        d -= box * np.rint(d / box)
        r2 = (d * d).sum(axis=1)
        inr = r2 < float(params.rcoulomb) ** 2
        if not inr.any():
            continue
        j = np.nonzero(inr)[0] + i + 1
        d, r2 = d[inr], np.maximum(r2[inr], 3.82e-7)
        bit = (~exm[i, j]).astype(np.float64)
        rinv = 1.0 / np.sqrt(r2)
        r = r2 * rinv
        qq = float(params.epsfac) * q[i] * q[j]
        if want_ewald:
            flr = erf(beta * r) / r2 - 2 * beta / np.sqrt(np.pi) * np.exp(-(beta * r) ** 2) / r
            fs = qq * (bit * rinv * rinv - flr) * rinv
            vel = qq * ((bit - erf(beta * r)) * rinv - bit * float(params.sh_ewald))
        else:
            krf, crf = float(params.reactionFieldCoefficient), float(params.reactionFieldShift)
            fs = qq * (bit * rinv - 2 * krf * r2) * rinv * rinv
            vel = qq * (bit * rinv + krf * r2 - crf)
        c6, c12 = nb[t[i], t[j], 0], nb[t[i], t[j], 1]
        inv = (r2 < float(params.rvdw) ** 2).astype(np.float64)
        r6 = bit * rinv**6 * inv
        fs += (c12 * r6 * r6 - c6 * r6) * rinv * rinv
        vv += (((c12 * r6 * r6 + bit * c12 * float(params.repulsion_shift_cpot)) / 12
                - (c6 * r6 + bit * c6 * float(params.dispersion_shift_cpot)) / 6) * inv).sum()
        vc += (vel * inv).sum()
        fv = fs[:, None] * d
        f[i] += fv.sum(axis=0)
        np.subtract.at(f, j, fv)
    q2 = (q * q).sum()
    vc += (-float(params.epsfac) * beta / np.sqrt(np.pi) * q2) if want_ewald else (
        -float(params.epsfac) * 0.5 * float(params.reactionFieldShift) * q2)
    out = np.zeros((cs.natoms, 3))
    out[real] = f
    return out, vc, vv
