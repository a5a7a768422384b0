"""The 72 known-answer cases of the reference's own unit test of this path, rebuilt from the
parameters in the reference test source
  /root/reference/src/gromacs/gmxlib/nonbonded/tests/nb_free_energy.cpp
(atoms and list :306-362, constants :163-201 and :224-240, coordinates :516-518, parameter
grid :503-527).  gtest `Combine` varies the LAST parameter fastest, so

    index = ((((sc*3 + inter)*1 + coord)*3 + lam)*2 + alpha)*2 + scCoulIdx

with softcore type sc in (Beutler, Gapsys), interaction in (Cut/Cut/None, Cut/Cut/PotSwitch,
Pme/Pme/None), lambda in (0, 0.5, 1), alpha in (0, 0.3), scCoul in (True, False).

The expected values live in tests/golden/nb_free_energy_kat.json, extracted verbatim from the
reference's refdata XML files by tests/golden/make_kat_golden.py.
"""
from __future__ import annotations

import numpy as np

from fepb200 import params as P
from fepb200.problem import FepList, Problem, nbfp_from_c6c12

NUM_CASES = 72
_INTERACTIONS = [
    (P.EEL_CUT, P.VDW_CUT, P.MOD_NONE),
    (P.EEL_CUT, P.VDW_CUT, P.MOD_POTSWITCH),
    (P.EEL_PME, P.VDW_PME, P.MOD_NONE),
]
_LAMBDAS = [0.0, 0.5, 1.0]
_ALPHAS = [0.0, 0.3]
_SCCOUL = [True, False]


def decode(index: int):
    sccoul_i = index % 2
    index //= 2
    alpha_i = index % 2
    index //= 2
    lam_i = index % 3
    index //= 3
    inter_i = index % 3
    sc = index // 3
    return sc, inter_i, _LAMBDAS[lam_i], _ALPHAS[alpha_i], _SCCOUL[sccoul_i]


def kat_problem(index: int, real_dtype=np.float64) -> Problem:
    sc, inter_i, lam, alpha, sccoul = decode(index)
    eel, vdw, mod = _INTERACTIONS[inter_i]

    # test :163-201 (InteractionConstHelper) -- note the deliberately odd constants
    p = P.Params()
    p.eeltype, p.vdwtype, p.vdw_modifier = eel, vdw, mod
    p.epsfac = P.ONE_4PI_EPS0 * 0.25
    p.reactionFieldCoefficient = 0.0
    p.reactionFieldShift = 1.0
    p.ewaldcoeff_q = P.calc_ewaldcoeff_q(1.0, 1.0e-5)
    p.ewaldcoeff_lj = P.calc_ewaldcoeff_lj(1.0, 1.0e-5)
    p.sh_ewald = 1.0e-5
    p.sh_lj_ewald = -1.0
    p.dispersion_shift_cpot = -1.0
    p.repulsion_shift_cpot = -1.0
    p.rcoulomb = 1.0  # interaction_const_t defaults (interaction_const.h:141-157)
    p.rvdw = 1.0
    p.rvdw_switch = 0.0
    # test :224-240 (ForcerecHelper) through SoftCoreParameters (interaction_const.cpp:50-63)
    p.softcoreType = sc
    p.alphaVdw = alpha
    p.alphaCoulomb = alpha if sccoul else 0.0
    p.lambdaPower = 1
    p.sigma6WithInvalidSigma = 0.3**6
    p.sigma6Minimum = 0.3**6 if sccoul else 0.0
    p.gapsysScaleLinpointVdW = alpha
    p.gapsysScaleLinpointCoul = alpha
    p.gapsysSigma6VdW = 0.3**6
    if real_dtype == np.float32:
        p = p.rounded()

    # test :306-362 (AtomData)
    c6 = np.zeros((3, 3))
    c12 = np.zeros((3, 3))
    for ti, tj in ((0, 0), (0, 2), (2, 0), (2, 2)):
        c6[ti, tj] = 0.001458
        c12[ti, tj] = 1.0062882e-6
    nbfp = np.stack([6.0 * c6, 12.0 * c12], axis=-1).ravel()
    # makeLJPmeC6GridCorrectionParameters with geometric mixing of the diagonal
    diag = np.array([c6[0, 0], c6[1, 1], c6[2, 2]])
    grid = np.zeros((3, 3, 2))
    grid[..., 0] = 6.0 * np.sqrt(diag[:, None] * diag[None, :])

    shiftvec = np.zeros((P.NUM_SHIFT_VECTORS, 3))
    prob = Problem(
        name=f"kat{index}",
        params=p,
        ntype=3,
        nbfp=nbfp,
        nbfp_grid=grid.ravel(),
        x=[[1.0, 1.0, 1.0], [1.1, 1.15, 1.2], [0.9, 0.85, 0.8], [1.1, 1.15, 0.8]],
        qA=[1.0, -1.0, -1.0, 1.0],
        qB=[1.0, 0.0, 0.0, 1.0],
        typeA=[0, 0, 0, 0],
        typeB=[0, 1, 2, 1],
        shiftvec=shiftvec,
        # the test's single shift vector is index 0 (and it is the zero vector)
        nblist=FepList([0], [0], [0], [0, 4], [0, 1, 2, 3], [0, 1, 1, 1]),
        nenergrp_pairs=1,
        real_dtype=real_dtype,
    )
    prob.set_lambda(lam)
    return prob


KAT_FLAGS = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL
