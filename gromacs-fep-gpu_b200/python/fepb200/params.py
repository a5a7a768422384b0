"""Host-side mirror of the constants the perturbed-pair kernel reads.

`Params` mirrors, field for field, `fepb200_params` of include/fepb200.h, which in turn
mirrors what the reference kernel reads from `interaction_const_t` and its
`SoftCoreParameters` (reference: src/gromacs/gmxlib/nonbonded/nb_free_energy.cpp:323-396,
src/gromacs/mdtypes/interaction_const.h:111-184).

`make_params()` derives the dependent constants the way the reference's
`init_interaction_const()` does (src/gromacs/mdtypes/interaction_const.cpp:106-149,248-345;
reaction field: src/gromacs/mdlib/rf_util.cpp:50-61; Ewald coefficients:
src/gromacs/ewald/ewald_utils.cpp:43-112) so that synthetic problems use physically
consistent numbers.  All values are rounded to float32, the `real` of the mixed-precision
reference build, because that is what crosses the C-ABI.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass, asdict

import numpy as np

# enum integers of the reference (api/legacy/include/gromacs/mdtypes/md_enums.h)
EEL_CUT, EEL_RF, EEL_PME, EEL_EWALD, EEL_RFZERO = 0, 1, 3, 4, 16
VDW_CUT, VDW_PME = 0, 5
MOD_POTSHIFT, MOD_NONE, MOD_POTSWITCH, MOD_FORCESWITCH = 1, 2, 3, 5
SC_BEUTLER, SC_GAPSYS = 0, 1

# kernel flags (src/gromacs/gmxlib/nonbonded/nonbonded.h:38-42)
DO_FORCE = 1 << 1
DO_SHIFTFORCE = 1 << 2
DO_FOREIGNLAMBDA = 1 << 3
DO_POTENTIAL = 1 << 4
DO_SR = 1 << 5
CLEAR_OUTPUTS = 1 << 16
ATOMIC_OUTPUTS = 1 << 17  # add_forces_device / export_scalars_device: atomic adds (another stream adds into the same buffers)

NUM_SHIFT_VECTORS = 45
CENTRAL_SHIFT_INDEX = 22
NUM_LAMBDA_COMPONENTS = 7
LAMBDA_COUL, LAMBDA_VDW = 2, 3

# api/legacy/include/gromacs/math/units.h:84-103
_E_CHARGE = 1.602176634e-19
_AVOGADRO = 6.02214076e23
_EPS0_SI = 8.8541878128e-12
EPSILON0 = (_EPS0_SI * 1e-9 * 1e3) / (_E_CHARGE * _E_CHARGE * _AVOGADRO)
ONE_4PI_EPS0 = 1.0 / (4.0 * math.pi * EPSILON0)  # 138.935458 kJ mol^-1 nm e^-2


class CParams(ctypes.Structure):
    """ctypes image of `struct fepb200_params` (include/fepb200.h)."""

    _fields_ = [
        ("eeltype", ctypes.c_int),
        ("vdwtype", ctypes.c_int),
        ("vdw_modifier", ctypes.c_int),
        ("epsfac", ctypes.c_float),
        ("rcoulomb", ctypes.c_float),
        ("rvdw", ctypes.c_float),
        ("rvdw_switch", ctypes.c_float),
        ("reactionFieldCoefficient", ctypes.c_float),
        ("reactionFieldShift", ctypes.c_float),
        ("sh_ewald", ctypes.c_float),
        ("sh_lj_ewald", ctypes.c_float),
        ("ewaldcoeff_q", ctypes.c_float),
        ("ewaldcoeff_lj", ctypes.c_float),
        ("dispersion_shift_cpot", ctypes.c_float),
        ("repulsion_shift_cpot", ctypes.c_float),
        ("softcoreType", ctypes.c_int),
        ("alphaVdw", ctypes.c_float),
        ("alphaCoulomb", ctypes.c_float),
        ("lambdaPower", ctypes.c_int),
        ("sigma6WithInvalidSigma", ctypes.c_float),
        ("sigma6Minimum", ctypes.c_float),
        ("gapsysScaleLinpointVdW", ctypes.c_float),
        ("gapsysScaleLinpointCoul", ctypes.c_float),
        ("gapsysSigma6VdW", ctypes.c_float),
    ]


@dataclass
class Params:
    eeltype: int = EEL_CUT
    vdwtype: int = VDW_CUT
    vdw_modifier: int = MOD_NONE
    epsfac: float = 1.0
    rcoulomb: float = 1.0
    rvdw: float = 1.0
    rvdw_switch: float = 0.0
    reactionFieldCoefficient: float = 0.0
    reactionFieldShift: float = 0.0
    sh_ewald: float = 0.0
    sh_lj_ewald: float = 0.0
    ewaldcoeff_q: float = 0.0
    ewaldcoeff_lj: float = 0.0
    dispersion_shift_cpot: float = 0.0
    repulsion_shift_cpot: float = 0.0
    softcoreType: int = SC_BEUTLER
    alphaVdw: float = 0.0
    alphaCoulomb: float = 0.0
    lambdaPower: int = 1
    sigma6WithInvalidSigma: float = 0.0
    sigma6Minimum: float = 0.0
    gapsysScaleLinpointVdW: float = 0.0
    gapsysScaleLinpointCoul: float = 0.0
    gapsysSigma6VdW: float = 0.0

    def rounded(self) -> "Params":
        """Round every real to float32 (what the C-ABI carries)."""
        d = asdict(self)
        for k, v in d.items():
            if isinstance(v, float):
                d[k] = float(np.float32(v))
        return Params(**d)

    def to_c(self) -> CParams:
        c = CParams()
        for name, _ in CParams._fields_:
            setattr(c, name, getattr(self, name))
        return c

    def to_dict(self) -> dict:
        return asdict(self)

    @staticmethod
    def from_dict(d: dict) -> "Params":
        return Params(**d)

    # convenience predicates, same meaning as md_enums.h:290-353
    @property
    def elec_ewald(self) -> bool:
        return self.eeltype in (EEL_PME, EEL_EWALD, 5, 13, 14, 15)

    @property
    def vdw_ewald(self) -> bool:
        return self.vdwtype == VDW_PME


def calc_ewaldcoeff_q(rc: float, rtol: float) -> float:
    """beta with erfc(beta*rc) == rtol by bisection (ewald_utils.cpp:43-71)."""
    beta = 5.0
    i = 0
    while True:
        i += 1
        beta *= 2
        if not math.erfc(beta * rc) > rtol:
            break
    lo, hi = 0.0, beta
    for _ in range(i + 60):
        beta = (lo + hi) / 2
        if math.erfc(beta * rc) > rtol:
            lo = beta
        else:
            hi = beta
    return beta


def _lj_ewald_function(beta: float, rc: float) -> float:
    x2 = (beta * rc) ** 2
    return math.exp(-x2) * (1 + x2 + x2 * x2 / 2.0)


def calc_ewaldcoeff_lj(rc: float, rtol: float) -> float:
    """LJ-PME beta by bisection (ewald_utils.cpp:84-112)."""
    beta = 5.0
    i = 0
    while True:
        i += 1
        beta *= 2.0
        if not _lj_ewald_function(beta, rc) > rtol:
            break
    lo, hi = 0.0, beta
    for _ in range(i + 60):
        beta = (lo + hi) / 2.0
        if _lj_ewald_function(beta, rc) > rtol:
            lo = beta
        else:
            hi = beta
    return beta


def make_params(
    *,
    coulombtype: str = "pme",  # "pme" | "rf" | "cut"
    vdwtype: str = "cut",  # "cut" | "pme"
    vdw_modifier: str = "potshift",  # "potshift" | "none" | "potswitch" | "forceswitch"
    coulomb_potshift: bool = True,
    rcoulomb: float = 1.0,
    rvdw: float = 1.0,
    rvdw_switch: float = 0.0,
    epsilon_r: float = 1.0,
    epsilon_rf: float = 78.0,
    ewald_rtol: float = 1e-5,
    ewald_rtol_lj: float = 1e-3,
    softcore: str = "beutler",  # "beutler" | "gapsys"
    sc_alpha: float = 0.5,
    sc_power: int = 1,
    sc_sigma: float = 0.3,
    sc_sigma_min: float | None = None,
    sc_coul: bool = False,
    gapsys_scale_lj: float = 0.85,
    gapsys_scale_q: float = 0.3,
    gapsys_sigma_lj: float = 0.3,
) -> Params:
    p = Params()
    p.eeltype = {"pme": EEL_PME, "ewald": EEL_EWALD, "rf": EEL_RF, "cut": EEL_CUT}[coulombtype]
    p.vdwtype = {"cut": VDW_CUT, "pme": VDW_PME}[vdwtype]
    p.vdw_modifier = {
        "potshift": MOD_POTSHIFT,
        "none": MOD_NONE,
        "potswitch": MOD_POTSWITCH,
        "forceswitch": MOD_FORCESWITCH,
    }[vdw_modifier]
    p.rcoulomb, p.rvdw, p.rvdw_switch = rcoulomb, rvdw, rvdw_switch
    p.epsfac = ONE_4PI_EPS0 / epsilon_r if epsilon_r != 0 else 0.0
    # interaction_const.cpp:273-299
    if p.vdw_modifier == MOD_POTSHIFT:
        p.dispersion_shift_cpot = -1.0 / rvdw**6
        p.repulsion_shift_cpot = -1.0 / rvdw**12
    elif p.vdw_modifier == MOD_FORCESWITCH:
        # force_switch_constants(): only cpot reaches the FEP kernel (SURVEY 8a note)
        # (interaction_const.cpp:215-230)
        def cpot(power: float) -> float:
            rsw, rc = rvdw_switch, rvdw
            c2 = ((power + 1) * rsw - (power + 4) * rc) / (rc ** (power + 2) * (rc - rsw) ** 2)
            c3 = -((power + 1) * rsw - (power + 3) * rc) / (rc ** (power + 2) * (rc - rsw) ** 3)
            return -(rc**-power) + power * c2 / 3 * (rc - rsw) ** 3 + power * c3 / 4 * (rc - rsw) ** 4

        p.dispersion_shift_cpot = cpot(6.0)
        p.repulsion_shift_cpot = cpot(12.0)
    # interaction_const.cpp:320-341 and rf_util.cpp:50-61
    if p.eeltype in (EEL_RF, EEL_RFZERO):
        if epsilon_rf == 0:
            krf = 1.0 / (2 * rcoulomb**3)
        else:
            krf = (epsilon_rf - epsilon_r) / (2 * epsilon_rf + epsilon_r) / rcoulomb**3
        p.reactionFieldCoefficient = krf
        p.reactionFieldShift = 1.0 / rcoulomb + krf * rcoulomb**2
    else:
        p.reactionFieldCoefficient = 0.0
        p.reactionFieldShift = 1.0 / rcoulomb if coulomb_potshift else 0.0
    # interaction_const.cpp:106-149
    if p.elec_ewald:
        p.ewaldcoeff_q = calc_ewaldcoeff_q(rcoulomb, ewald_rtol)
        p.sh_ewald = math.erfc(p.ewaldcoeff_q * rcoulomb) / rcoulomb if coulomb_potshift else 0.0
    if p.vdw_ewald:
        p.ewaldcoeff_lj = calc_ewaldcoeff_lj(rvdw, ewald_rtol_lj)
        if p.vdw_modifier == MOD_POTSHIFT:
            crc2 = (p.ewaldcoeff_lj * rvdw) ** 2
            p.sh_lj_ewald = (math.exp(-crc2) * (1 + crc2 + 0.5 * crc2 * crc2) - 1) / rvdw**6
    # SoftCoreParameters(fepvals), interaction_const.cpp:50-63
    p.softcoreType = {"beutler": SC_BEUTLER, "gapsys": SC_GAPSYS}[softcore]
    p.alphaVdw = sc_alpha
    p.alphaCoulomb = sc_alpha if sc_coul else 0.0
    p.lambdaPower = sc_power
    p.sigma6WithInvalidSigma = sc_sigma**6
    smin = sc_sigma if sc_sigma_min is None else sc_sigma_min
    p.sigma6Minimum = smin**6 if sc_coul else 0.0
    p.gapsysScaleLinpointVdW = gapsys_scale_lj
    p.gapsysScaleLinpointCoul = gapsys_scale_q
    p.gapsysSigma6VdW = gapsys_sigma_lj**6
    return p.rounded()
