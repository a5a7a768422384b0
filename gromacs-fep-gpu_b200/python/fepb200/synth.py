"""Synthetic solvated FEP systems of the shapes BASELINE.json names (C1..C5) and small random
stress problems, all seeded and reproducible (NumPy PCG64).

Recipe: TIP3P-like 3-site water on a jittered cubic lattice at 33.4
molecules/nm^3 with random orientations; compact blobs of perturbed Lennard-Jones sites;
geometric combination of c6/c12; cubic box with the 45 shift vectors of the reference
(`problem.shift_vectors`).  The FEP pair list holds every pair with at least one perturbed atom
within r_list = 1.1 nm, as a half list (the atom that comes first in a Morton-type spatial
order is the i atom, like the grid order of the reference's pair search), i-entries split at
64 j (reference: src/gromacs/nbnxm/pairlist.cpp:1509,1710-1714) and at energy-group-pair
changes (:1698-1707), perturbed atoms carry an excluded self pair (the i == j half-weight
case of nb_free_energy.cpp:1035-1040,1079-1084), and bonded neighbours inside a blob are
excluded pairs (excl_fep = 0).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .params import LAMBDA_COUL, LAMBDA_VDW, NUM_LAMBDA_COMPONENTS, Params, make_params
from .problem import FepList, Problem, nbfp_from_c6c12, nbfp_grid_geometric, shift_vectors

WATER_DENSITY = 33.4  # molecules / nm^3
R_OH = 0.09572
ANGLE_HOH = np.deg2rad(104.52)
Q_OW, Q_HW = -0.834, 0.417
C6_OW, C12_OW = 2.489e-3, 2.435e-6
MAX_NRJ_FEP = 64  # pairlist.cpp:1509
R_LIST_FEP = 1.1


@dataclass
class SystemSpec:
    name: str
    box: float  # nm, cubic
    n_blobs: int
    blob_size: int
    coulombtype: str = "pme"
    vdw_modifier: str = "potshift"
    vdwtype: str = "cut"
    softcore: str = "beutler"
    sc_alpha: float = 0.5
    sc_power: int = 1
    sc_coul: bool = False
    lambda_coul: float = 0.5
    lambda_vdw: float = 0.5
    n_foreign: int = 0
    n_energy_groups: int = 1
    transform: bool = False  # A->B transform (both states interacting) instead of decoupling
    neutral_solute: bool = False  # q = 0 in both states (methane-like)
    n_adversarial: int = 0


# The five configurations of BASELINE.json / SURVEY 8d.
SPECS = {
    "C1": SystemSpec("C1", 3.0, 1, 5, neutral_solute=True, n_adversarial=2),
    "C2": SystemSpec("C2", 6.3, 1, 50, n_foreign=20, transform=True, n_adversarial=8),
    "C3": SystemSpec("C3", 10.0, 4, 50, softcore="gapsys", lambda_coul=0.7, lambda_vdw=0.3, n_adversarial=16),
    "C4": SystemSpec("C4", 13.56, 10, 50, coulombtype="rf", sc_coul=True, lambda_coul=0.4, lambda_vdw=0.4,
                     n_foreign=40, n_energy_groups=4, n_adversarial=32),
    "C5": SystemSpec("C5", 21.5, 40, 50, n_foreign=20, n_adversarial=64),
}  # fmt: skip


def _random_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    r = np.empty((n, 3, 3))
    r[:, 0, 0] = 1 - 2 * (y * y + z * z)
    r[:, 0, 1] = 2 * (x * y - z * w)
    r[:, 0, 2] = 2 * (x * z + y * w)
    r[:, 1, 0] = 2 * (x * y + z * w)
    r[:, 1, 1] = 1 - 2 * (x * x + z * z)
    r[:, 1, 2] = 2 * (y * z - x * w)
    r[:, 2, 0] = 2 * (x * z - y * w)
    r[:, 2, 1] = 2 * (y * z + x * w)
    r[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return r


def _water_box(rng, box):
    n_target = int(round(WATER_DENSITY * box**3))
    m = int(np.ceil(n_target ** (1.0 / 3.0)))
    a = box / m
    grid = np.stack(np.meshgrid(*(np.arange(m),) * 3, indexing="ij"), axis=-1).reshape(-1, 3)
    pick = rng.permutation(grid.shape[0])[:n_target]
    pick.sort()
    o = (grid[pick] + 0.5) * a + rng.uniform(-0.15, 0.15, size=(n_target, 3)) * max(a - 0.25, 0.0)
    # canonical molecule in its own frame
    h1 = np.array([R_OH * np.sin(ANGLE_HOH / 2), 0.0, R_OH * np.cos(ANGLE_HOH / 2)])
    h2 = np.array([-R_OH * np.sin(ANGLE_HOH / 2), 0.0, R_OH * np.cos(ANGLE_HOH / 2)])
    rot = _random_rotations(rng, n_target)
    mol = np.stack([o, o + rot @ h1, o + rot @ h2], axis=1)  # [n,3,3]
    return mol


def _blob(rng, n, centre):
    """A compact self-avoiding chain of n sites: bonds of 0.15 nm, 1-3 distances >= 0.23 nm and
    every other intra-blob distance >= 0.30 nm (1-2 and 1-3 pairs become exclusions)."""
    radius = 0.20 * n ** (1.0 / 3.0) + 0.15
    while True:
        pts = [np.zeros(3)]
        stuck = False
        while len(pts) < n and not stuck:
            arr = np.asarray(pts)
            for _ in range(400):
                d = rng.normal(size=3)
                cand = arr[-1] + d * (0.15 / np.linalg.norm(d))
                if np.linalg.norm(cand - arr.mean(axis=0)) > radius:
                    continue
                if len(pts) >= 2 and np.linalg.norm(cand - arr[-2]) < 0.23:
                    continue
                if len(pts) >= 3 and np.any(np.linalg.norm(arr[:-2] - cand, axis=1) < 0.30):
                    continue
                pts.append(cand)
                break
            else:
                stuck = True
        if not stuck:
            break
        radius *= 1.05
    arr = np.asarray(pts)
    return arr - arr.mean(axis=0) + centre


def _morton_rank(x, box, cell=0.55):
    n = max(1, int(box / cell))
    c = np.minimum((x / box * n).astype(np.int64), n - 1)
    key = np.zeros(x.shape[0], np.int64)
    for b in range(11):
        for d in range(3):
            key |= ((c[:, d] >> b) & 1) << (3 * b + d)
    order = np.argsort(key, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    return rank


def build_fep_list(x, box, perturbed, excluded_pairs, group_of_atom, n_groups, r_list=R_LIST_FEP,
                   max_nrj=MAX_NRJ_FEP) -> FepList:
    """All pairs with >= 1 perturbed atom within r_list under the minimum-image convention."""
    from scipy.spatial import cKDTree

    n = x.shape[0]
    xw = np.mod(x.astype(np.float64), box)
    xw[xw >= box] = 0.0
    tree = cKDTree(xw, boxsize=box)
    perturbed = np.asarray(perturbed, np.int64)
    is_pert = np.zeros(n, bool)
    is_pert[perturbed] = True
    neigh = tree.query_ball_point(xw[perturbed], r=r_list)
    a = np.repeat(perturbed, [len(v) for v in neigh])
    b = np.concatenate([np.asarray(v, np.int64) for v in neigh]) if len(neigh) else np.zeros(0, np.int64)
    # drop self hits; keep perturbed-perturbed pairs once
    keep = (a != b) & ~(is_pert[b] & (b < a))
    a, b = a[keep], b[keep]
    # the kd-tree tests with its own rounding; apply the list criterion on the float32 data
    rank = _morton_rank(xw, box)
    swap = rank[b] < rank[a]
    i = np.where(swap, b, a)
    j = np.where(swap, a, b)
    # excluded self pairs of perturbed atoms
    i = np.concatenate([i, perturbed])
    j = np.concatenate([j, perturbed])
    # shift of the i atom: x_i + k*box is the image closest to x_j
    xf = x.astype(np.float64)
    k = np.rint((xf[j] - xf[i]) / box).astype(np.int64)
    k = np.clip(k, [-2, -1, -1], [2, 1, 1])
    shift = 5 * (3 * (k[:, 2] + 1) + (k[:, 1] + 1)) + (k[:, 0] + 2)
    gid = group_of_atom[i] * n_groups + group_of_atom[j]
    excl_set = {(min(p, q), max(p, q)) for p, q in excluded_pairs}
    included = np.ones(i.shape[0], np.int32)
    included[i == j] = 0
    if excl_set:
        both = is_pert[i] & is_pert[j]
        for idx in np.nonzero(both)[0]:
            if (min(i[idx], j[idx]), max(i[idx], j[idx])) in excl_set:
                included[idx] = 0
    order = np.lexsort((rank[j], gid, shift, rank[i]))
    i, j, shift, gid, included = i[order], j[order], shift[order], gid[order], included[order]
    # entry boundaries: change of (i, shift, gid) or 64 pairs
    m = i.shape[0]
    new_entry = np.ones(m, bool)
    if m > 1:
        new_entry[1:] = (i[1:] != i[:-1]) | (shift[1:] != shift[:-1]) | (gid[1:] != gid[:-1])
    seg_id = np.cumsum(new_entry) - 1
    seg_start = np.nonzero(new_entry)[0]
    pos_in_seg = np.arange(m) - seg_start[seg_id]
    new_entry |= (pos_in_seg % max_nrj) == 0
    starts = np.nonzero(new_entry)[0]
    jindex = np.concatenate([starts, [m]])
    return FepList(i[starts], gid[starts], shift[starts], jindex, j, included)


def make_system(spec: SystemSpec | str, seed: int | None = None, dtype=np.float32) -> Problem:
    if isinstance(spec, str):
        spec = SPECS[spec]
    seed = 20261018 + sum(map(ord, spec.name)) if seed is None else seed
    rng = np.random.default_rng(seed)
    box = float(spec.box)
    water = _water_box(rng, box)  # [nmol,3,3]

    # blobs at well separated random centres
    centres = []
    while len(centres) < spec.n_blobs:
        c = rng.uniform(0.0, box, size=3)
        if all(np.linalg.norm((c - o + box / 2) % box - box / 2) > min(2.6, box / 2.2) for o in centres):
            centres.append(c)
        elif len(centres) and rng.random() < 0.001:
            centres.append(c)  # give up on separation in crowded boxes
    blobs = [_blob(rng, spec.blob_size, c) for c in centres]
    solute = np.concatenate(blobs) if blobs else np.zeros((0, 3))

    # remove waters whose oxygen is within 0.28 nm of a solute site
    from scipy.spatial import cKDTree

    wo = np.mod(water[:, 0, :], box)
    wo[wo >= box] = 0.0
    tree = cKDTree(wo, boxsize=box)
    clash = set()
    for hits in tree.query_ball_point(np.mod(solute, box), r=0.28):
        clash.update(hits)
    keep = np.ones(water.shape[0], bool)
    keep[list(clash)] = False
    water = water[keep]

    n_sol = solute.shape[0]
    n_wat = water.shape[0] * 3
    x = np.concatenate([solute, water.reshape(-1, 3)])
    n = x.shape[0]

    # atom types: 0 = OW, 1 = HW, 2..2+k-1 solute types, last = dummy
    k_sol = 4
    sig = rng.uniform(0.30, 0.37, size=k_sol)
    eps = rng.uniform(0.2, 0.8, size=k_sol)
    c6_t = np.concatenate([[C6_OW, 0.0], 4 * eps * sig**6, [0.0]])
    c12_t = np.concatenate([[C12_OW, 0.0], 4 * eps * sig**12, [0.0]])
    ntype = c6_t.shape[0]
    dummy = ntype - 1
    c6 = np.sqrt(c6_t[:, None] * c6_t[None, :])
    c12 = np.sqrt(c12_t[:, None] * c12_t[None, :])
    nbfp = nbfp_from_c6c12(c6, c12)
    nbfp_grid = nbfp_grid_geometric(c6_t)

    typeA = np.empty(n, np.int32)
    typeB = np.empty(n, np.int32)
    qA = np.empty(n)
    qB = np.empty(n)
    typeA[:n_sol] = rng.integers(2, 2 + k_sol, size=n_sol)
    if spec.neutral_solute:
        qA[:n_sol] = 0.0
    else:
        q = rng.uniform(-0.5, 0.5, size=n_sol)
        for b in range(spec.n_blobs):
            s = slice(b * spec.blob_size, (b + 1) * spec.blob_size)
            q[s] -= q[s].mean()
        qA[:n_sol] = np.round(q, 4)
    if spec.transform:
        typeB[:n_sol] = rng.integers(2, 2 + k_sol, size=n_sol)
        # a few sites appear / disappear so that soft-core pairs exist as well
        vanish = rng.random(n_sol) < 0.2
        typeB[:n_sol][vanish] = dummy
        q = rng.uniform(-0.5, 0.5, size=n_sol)
        q -= q.mean()
        qB[:n_sol] = np.round(q, 4)
        qB[:n_sol][vanish] = 0.0
    else:
        typeB[:n_sol] = dummy
        qB[:n_sol] = 0.0
    typeA[n_sol:] = np.tile([0, 1, 1], water.shape[0])
    typeB[n_sol:] = typeA[n_sol:]
    qA[n_sol:] = np.tile([Q_OW, Q_HW, Q_HW], water.shape[0])
    qB[n_sol:] = qA[n_sol:]

    # adversarial placements (only next to sites that are soft-cored, i.e. vanish in state B):
    # a water hydrogen (no LJ) deep inside the soft core of a perturbed atom, and a water
    # molecule moved so that a hydrogen sits at the cut-off distance from a perturbed atom
    soft_sites = np.nonzero(typeB[:n_sol] == dummy)[0]
    for a in range(spec.n_adversarial if soft_sites.size else 0):
        mol = rng.integers(0, water.shape[0])
        tgt = soft_sites[rng.integers(0, soft_sites.size)]
        first = n_sol + 3 * mol
        d = rng.normal(size=3)
        if a % 2 == 0:
            d *= rng.uniform(0.05, 0.08) / np.linalg.norm(d)
            move = x[tgt] + d - x[first + 1]
        else:
            d *= 1.0 / np.linalg.norm(d)
            move = x[tgt] + d - x[first + 2]
        x[first : first + 3] += move

    x = np.mod(x, box)
    x32 = x.astype(np.float32)
    x32[x32 >= np.float32(box)] = 0.0

    # energy groups: solute = 0, water slabs along z = 1..n-1
    ng = spec.n_energy_groups
    group = np.zeros(n, np.int64)
    if ng > 1:
        slab = np.minimum((x32[n_sol:, 2] / box * (ng - 1)).astype(np.int64), ng - 2) + 1
        # whole molecules share the group of their oxygen
        slab = np.repeat(slab.reshape(-1, 3)[:, 0], 3)
        group[n_sol:] = slab

    # exclusions: 1-2 and 1-3 neighbours along each blob chain
    excl = []
    for b in range(spec.n_blobs):
        base = b * spec.blob_size
        for a in range(spec.blob_size):
            for d in (1, 2):
                if a + d < spec.blob_size:
                    excl.append((base + a, base + a + d))

    perturbed = np.arange(n_sol)
    nblist = build_fep_list(x32, box, perturbed, excl, group, ng)

    params = make_params(
        coulombtype=spec.coulombtype,
        vdwtype=spec.vdwtype,
        vdw_modifier=spec.vdw_modifier,
        softcore=spec.softcore,
        sc_alpha=spec.sc_alpha,
        sc_power=spec.sc_power,
        sc_coul=spec.sc_coul,
    )
    lam = np.zeros(NUM_LAMBDA_COMPONENTS, np.float32)
    lam[:] = spec.lambda_vdw
    lam[LAMBDA_COUL] = spec.lambda_coul
    lam[LAMBDA_VDW] = spec.lambda_vdw
    if spec.n_foreign > 0:
        grid = np.linspace(0.0, 1.0, spec.n_foreign)
        if spec.lambda_coul != spec.lambda_vdw:
            # separate paths: coulomb is switched off first, then vdw
            alc = np.clip(2 * grid, 0, 1)
            alv = np.clip(2 * grid - 1, 0, 1)
        else:
            alc = alv = grid
    else:
        alc = alv = np.zeros(0)
    box_m = np.diag([box, box, box])
    return Problem(
        name=spec.name,
        params=params,
        ntype=ntype,
        nbfp=nbfp,
        nbfp_grid=nbfp_grid,
        x=x32,
        qA=qA,
        qB=qB,
        typeA=typeA,
        typeB=typeB,
        shiftvec=shift_vectors(box_m),
        nblist=nblist,
        nenergrp_pairs=ng * ng,
        lambda_=lam,
        all_lambda_coul=alc,
        all_lambda_vdw=alv,
        box=box_m,
        perturbed=perturbed,
        real_dtype=dtype,
    )


def scaled_spec(name: str, box: float, n_blobs: int, blob_size: int | None = None, **over) -> SystemSpec:
    """A smaller system of the same kind as SPECS[name] (for tests)."""
    base = SPECS[name]
    d = dict(base.__dict__)
    d.update(name=f"{name}s", box=box, n_blobs=n_blobs)
    if blob_size is not None:
        d["blob_size"] = blob_size
    d.update(over)
    return SystemSpec(**d)


def random_problem(seed: int, params: Params, *, natoms=96, nri=40, max_j=70, ntype=5, n_groups=2,
                   box=1.6, n_foreign=0, lambda_coul=0.35, lambda_vdw=0.6, frac_excluded=0.15,
                   frac_self=0.05, frac_overlap=0.03, frac_cutoff=0.03, min_overlap=1e-7, dtype=np.float32) -> Problem:
    """A small unphysical problem that hits every branch: random types (some with zero c6/c12 in
    one or both states), random charges (some zero), all 45 shift vectors, several energy-group
    pairs, excluded pairs inside and beyond the cut-off, self pairs, overlapping atoms and pairs
    placed (to float32 resolution) on the cut-off."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, box, size=(natoms, 3))
    sig = rng.uniform(0.25, 0.36, size=ntype)
    eps = rng.uniform(0.2, 1.0, size=ntype)
    c6_t = 4 * eps * sig**6
    c12_t = 4 * eps * sig**12
    c6_t[-1] = c12_t[-1] = 0.0  # dummy type
    c6 = np.sqrt(c6_t[:, None] * c6_t[None, :])
    c12 = np.sqrt(c12_t[:, None] * c12_t[None, :])
    if ntype > 3:
        c12[1, 2] = c12[2, 1] = 0.0  # a pair type with c6 > 0 but c12 == 0
    typeA = rng.integers(0, ntype, size=natoms)
    typeB = np.where(rng.random(natoms) < 0.5, typeA, rng.integers(0, ntype, size=natoms))
    qA = np.round(rng.uniform(-1, 1, size=natoms), 3)
    qA[rng.random(natoms) < 0.2] = 0.0
    qB = np.where(rng.random(natoms) < 0.4, qA, np.round(rng.uniform(-1, 1, size=natoms), 3))
    qB[rng.random(natoms) < 0.3] = 0.0
    box_m = np.diag([box, box * 1.1, box * 0.9])
    sv = shift_vectors(box_m)

    iinr, gid, shift, jindex, jjnr, excl = [], [], [], [0], [], []
    x32 = x.astype(np.float32)
    for _ in range(nri):
        i = int(rng.integers(0, natoms))
        nj = int(rng.integers(1, max_j + 1))
        s = int(rng.integers(0, 45)) if rng.random() < 0.5 else 22
        js = rng.integers(0, natoms, size=nj)
        ex = (rng.random(nj) >= frac_excluded).astype(np.int32)
        for k in range(nj):
            u = rng.random()
            if u < frac_self:
                js[k] = i
                ex[k] = 0
            elif u < frac_self + frac_overlap:
                # move atom j on top of the shifted i atom
                d = rng.normal(size=3)
                d *= rng.uniform(min_overlap, max(2e-3, 2 * min_overlap)) / np.linalg.norm(d)
                if js[k] != i:
                    x32[js[k]] = (x32[i].astype(np.float64) + sv[s] + d).astype(np.float32)
            elif u < frac_self + frac_overlap + frac_cutoff:
                d = rng.normal(size=3)
                d *= params.rcoulomb / np.linalg.norm(d)
                if js[k] != i:
                    x32[js[k]] = (x32[i].astype(np.float64) + sv[s] + d).astype(np.float32)
        ex[js == i] = 0  # an included i == j pair (r = 0 without exclusion) is not a physical input
        iinr.append(i)
        gid.append(int(rng.integers(0, n_groups * n_groups)))
        shift.append(s)
        jjnr.extend(js.tolist())
        excl.extend(ex.tolist())
        jindex.append(len(jjnr))
    # Excluded pairs far outside the pair-search range do not occur in the reference (its
    # ExclusionChecker aborts, pairlist.cpp:4456-4467) and its rational Ewald-correction fits
    # are only valid for beta*r <~ 4 (simd_math.h:1560-1610): keep exclusions within 1.2 r_c.
    ii = np.repeat(np.asarray(iinr), np.diff(jindex))
    jj = np.asarray(jjnr)
    ss = np.repeat(np.asarray(shift), np.diff(jindex))
    dist = np.linalg.norm(x32[ii].astype(np.float64) + sv[ss] - x32[jj], axis=1)
    excl = np.asarray(excl, np.int32)
    excl[(excl == 0) & (dist > 1.2 * params.rcoulomb)] = 1
    lam = np.full(NUM_LAMBDA_COMPONENTS, lambda_vdw, np.float32)
    lam[LAMBDA_COUL] = lambda_coul
    lam[LAMBDA_VDW] = lambda_vdw
    alc = rng.uniform(0, 1, size=n_foreign)
    alv = rng.uniform(0, 1, size=n_foreign)
    if n_foreign >= 2:
        alc[0], alv[0] = 0.0, 0.0
        alc[1], alv[1] = 1.0, 1.0
    return Problem(
        name=f"random{seed}",
        params=params,
        ntype=ntype,
        nbfp=nbfp_from_c6c12(c6, c12),
        nbfp_grid=nbfp_grid_geometric(c6_t),
        x=x32,
        qA=qA,
        qB=qB,
        typeA=typeA,
        typeB=typeB,
        shiftvec=sv,
        nblist=FepList(iinr, gid, shift, jindex, jjnr, excl),
        nenergrp_pairs=n_groups * n_groups,
        lambda_=lam,
        all_lambda_coul=alc,
        all_lambda_vdw=alv,
        box=box_m,
        real_dtype=dtype,
    )
