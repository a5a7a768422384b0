"""The perturbed cases of the reference's own test of the 1-4 pair interactions, rebuilt from
/root/reference/src/gromacs/listed_forces/tests/pairs.cpp (:160-190 soft-core and forcerec set-up,
:326-334 atoms, :440-447 the LJ14 parameters, :456-458 coordinates, :301-303 unit box), plus seeded
random perturbed 1-4 lists."""
from __future__ import annotations

import numpy as np

from fepb200 import params as P
from fepb200.pairs14 import PBC_NONE, PBC_XY, PBC_XYZ, Pairs14Problem

PBC = {"no": PBC_NONE, "xy": PBC_XY, "xyz": PBC_XYZ}


def _params(softcore: str, dtype) -> P.Params:
    p = P.Params()
    p.eeltype, p.vdwtype, p.vdw_modifier = P.EEL_CUT, P.VDW_CUT, P.MOD_NONE
    p.epsfac = 1.0  # interaction_const_t default (interaction_const.h:167)
    p.rcoulomb = 1.0  # default (:155); only the Gapsys linearisation point looks at it
    p.rvdw = 1.0
    p.softcoreType = P.SC_BEUTLER if softcore == "beutler" else P.SC_GAPSYS
    p.alphaVdw = 0.3
    p.alphaCoulomb = 0.3  # bScCoul = true
    p.lambdaPower = 1
    p.sigma6WithInvalidSigma = 0.3**6
    p.sigma6Minimum = 0.3**6
    p.gapsysScaleLinpointVdW = 0.85
    p.gapsysScaleLinpointCoul = 0.3
    p.gapsysSigma6VdW = 0.3**6
    return p.rounded() if dtype == np.float32 else p


def kat_pairs14(case: dict, dtype=np.float64) -> Pairs14Problem:
    lam = np.full(P.NUM_LAMBDA_COMPONENTS, case["lam"])
    return Pairs14Problem(
        params=_params(case["softcore"], dtype), fudgeQQ=0.5,
        iatoms=[[0, 1, 2], [0, 0, 2]],
        c6A=[0.001458], c12A=[1.0062882e-6], c6B=[0.0], c12B=[0.0],
        x=[[0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [1.1, 1.2, 1.3]],
        qA=[1.0, -0.5, -0.5], qB=[0.0, 0.0, 0.0],
        box_diag=[1.0, 1.0, 1.0], pbc_type=PBC[case["pbc"]], lambda_=lam, real_dtype=dtype)


def random_pairs14(seed: int, softcore: str, *, natoms=600, npairs=300, ntypes=12, n_groups=2, lam_c=0.35, lam_v=0.6,
                   sc_power=1, sc_coul=True, pbc_type=PBC_XYZ, dtype=np.float32) -> Pairs14Problem:
    rng = np.random.default_rng(seed)
    box = np.array([2.0, 2.3, 1.9])
    x = rng.uniform(0, 1, size=(natoms, 3)) * box
    ai = rng.integers(0, natoms, size=npairs)
    aj = (ai + rng.integers(1, natoms, size=npairs)) % natoms
    # 1-4 partners sit 0.25 .. 0.4 nm apart (a few at 0.12 .. 0.2 nm, inside the soft core), across the
    # box too.  Every atom is the j atom of at most one pair, so that placing it does not move an
    # earlier pair's partner to an arbitrary (possibly overlapping) distance.
    aj = rng.permutation(natoms)[:npairs] if npairs <= natoms else aj
    ai = np.where(ai == aj, (ai + 1) % natoms, ai)
    d = rng.normal(size=(npairs, 3))
    d *= (rng.uniform(0.25, 0.4, size=npairs) / np.linalg.norm(d, axis=1))[:, None]
    close = rng.random(npairs) < 0.1
    d[close] *= 0.5
    x0 = x.copy()
    for k in range(npairs):
        x[aj[k]] = np.mod(x0[ai[k]] + d[k], box) if pbc_type != PBC_NONE else x0[ai[k]] + d[k]
    moved = np.zeros(natoms, bool)
    moved[aj] = True
    keep = ~moved[ai]  # pairs whose i atom kept its original position have exactly the intended distance
    ai, aj = ai[keep], aj[keep]
    npairs = int(keep.sum())
    sig, eps = rng.uniform(0.25, 0.36, size=ntypes), rng.uniform(0.2, 1.0, size=ntypes)
    c6a, c12a = 4 * eps * sig**6, 4 * eps * sig**12
    c6b, c12b = c6a.copy(), c12a.copy()
    vanish = rng.random(ntypes) < 0.5
    c6b[vanish] = c12b[vanish] = 0.0
    appear = (~vanish) & (rng.random(ntypes) < 0.3)
    c6a[appear] = c12a[appear] = 0.0
    other = (~vanish) & (~appear)
    c6b[other] *= rng.uniform(0.5, 1.5, size=other.sum())
    qA = np.round(rng.uniform(-1, 1, size=natoms), 3)
    qB = np.where(rng.random(natoms) < 0.5, 0.0, np.round(rng.uniform(-1, 1, size=natoms), 3))
    p = _params(softcore, dtype)
    p.lambdaPower = sc_power
    p.epsfac = P.ONE_4PI_EPS0
    if not sc_coul:
        p.alphaCoulomb, p.sigma6Minimum = 0.0, 0.0
    if dtype == np.float32:
        p = p.rounded()
    lam = np.full(P.NUM_LAMBDA_COMPONENTS, lam_v)
    lam[P.LAMBDA_COUL], lam[P.LAMBDA_VDW] = lam_c, lam_v
    return Pairs14Problem(params=p, fudgeQQ=0.8333, iatoms=np.stack([rng.integers(0, ntypes, size=npairs), ai, aj], 1),
                          c6A=c6a, c12A=c12a, c6B=c6b, c12B=c12b, x=x, qA=qA, qB=qB, box_diag=box, pbc_type=pbc_type,
                          gid=rng.integers(0, n_groups * n_groups, size=npairs), nenergrp_pairs=n_groups * n_groups,
                          lambda_=lam, real_dtype=dtype)
