"""TEST INFRASTRUCTURE -- loaders for the two CPU checkers of the FEP perturbed-pair path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (gromacs-fep-gpu_b200/) never does.

  * `run_port(problem, ...)`  -- oracle/fep_oracle.c, our plain-C double-precision restatement
    of the reference algorithm (built by oracle/Makefile into oracle/libfep_oracle.so).
  * `run_ref(problem, ...)`   -- oracle/_ref/libfepref_*.so, the reference's own
    nb_free_energy.cpp compiled in place by oracle/ref_build/Makefile (present only when it was
    built in a container that has /root/reference; the .so files travel to the GPU box).

  * `run_fork_cuda(problem, ...)` -- oracle/_ref/libfepfork_cuda.so, the reference fork's own CUDA
    FEP kernels compiled in place for sm_100a (oracle/ref_build/fork_cuda/): the GPU kernels "to
    beat", run on the same device and problem; needs a GPU.  Mixed precision, fewer features than
    the CPU kernel -- a performance baseline and a loose cross-check, not the parity oracle.

All return the same dict: f [N,3], fshift [45,3], Vc [G], Vv [G], dvdl [2] (coul, vdw),
foreign_energy [L+1], foreign_dvdl [L+1,2], all float64, for ONE step starting from zeroed
outputs, plus `seconds` = (best current-lambda pass, best foreign sweep) wall time.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
PORT_LIB = os.path.join(HERE, "libfep_oracle.so")

_DP = ctypes.POINTER(ctypes.c_double)
_IP = ctypes.POINTER(ctypes.c_int)


class CParamsD(ctypes.Structure):
    """`struct fepref_params` / `struct fep_oracle_params`: fepb200_params with double reals."""

    _fields_ = [
        ("eeltype", ctypes.c_int),
        ("vdwtype", ctypes.c_int),
        ("vdw_modifier", ctypes.c_int),
        ("epsfac", ctypes.c_double),
        ("rcoulomb", ctypes.c_double),
        ("rvdw", ctypes.c_double),
        ("rvdw_switch", ctypes.c_double),
        ("reactionFieldCoefficient", ctypes.c_double),
        ("reactionFieldShift", ctypes.c_double),
        ("sh_ewald", ctypes.c_double),
        ("sh_lj_ewald", ctypes.c_double),
        ("ewaldcoeff_q", ctypes.c_double),
        ("ewaldcoeff_lj", ctypes.c_double),
        ("dispersion_shift_cpot", ctypes.c_double),
        ("repulsion_shift_cpot", ctypes.c_double),
        ("softcoreType", ctypes.c_int),
        ("alphaVdw", ctypes.c_double),
        ("alphaCoulomb", ctypes.c_double),
        ("lambdaPower", ctypes.c_int),
        ("sigma6WithInvalidSigma", ctypes.c_double),
        ("sigma6Minimum", ctypes.c_double),
        ("gapsysScaleLinpointVdW", ctypes.c_double),
        ("gapsysScaleLinpointCoul", ctypes.c_double),
        ("gapsysSigma6VdW", ctypes.c_double),
    ]


def _params_d(params) -> CParamsD:
    c = CParamsD()
    for name, _ in CParamsD._fields_:
        setattr(c, name, getattr(params, name))
    return c


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr_d(a):
    return a.ctypes.data_as(_DP)


def _ptr_i(a):
    return a.ctypes.data_as(_IP)


_DISPATCH_ARGTYPES = [
    ctypes.POINTER(CParamsD), ctypes.c_int, ctypes.c_int, ctypes.c_int, _DP, _DP, ctypes.c_int,
    _DP, _DP, _DP, _IP, _IP, _DP,
    ctypes.c_int, _IP, _IP, _IP, _IP, _IP, _IP, ctypes.c_int,
    ctypes.c_int, _DP, ctypes.c_int, _DP, _DP,
    _DP, _DP, _DP, _DP, _DP, _DP, _DP, ctypes.c_int, _DP,
]  # fmt: skip


def _call_dispatch(fn, problem, flags, nthreads, use_simd, repeats):
    p = problem
    n, g, l = p.natoms, p.nenergrp_pairs, p.n_foreign
    nb = p.nblist
    x, qa, qb = _d(p.x), _d(p.qA), _d(p.qB)
    nbfp, grid, sv = _d(p.nbfp), _d(p.nbfp_grid), _d(p.shiftvec)
    lam, alc, alv = _d(p.lambda_), _d(p.all_lambda_coul), _d(p.all_lambda_vdw)
    ta, tb_ = _i(p.typeA), _i(p.typeB)
    out = dict(
        f=np.zeros((n, 3)),
        fshift=np.zeros((45, 3)),
        Vc=np.zeros(g),
        Vv=np.zeros(g),
        dvdl=np.zeros(2),
        foreign_energy=np.zeros(l + 1),
        foreign_dvdl=np.zeros((l + 1, 2)),
    )
    seconds = np.zeros(2)
    cp = _params_d(p.params)
    rc = fn(
        ctypes.byref(cp), int(use_simd), int(nthreads), p.ntype, _ptr_d(nbfp), _ptr_d(grid), n,
        _ptr_d(x), _ptr_d(qa), _ptr_d(qb), _ptr_i(ta), _ptr_i(tb_), _ptr_d(sv),
        nb.nri, _ptr_i(nb.iinr), _ptr_i(nb.gid), _ptr_i(nb.shift), _ptr_i(nb.jindex),
        _ptr_i(nb.jjnr), _ptr_i(nb.excl_fep), g,
        int(flags), _ptr_d(lam), l, _ptr_d(alc), _ptr_d(alv),
        _ptr_d(out["f"]), _ptr_d(out["fshift"]), _ptr_d(out["Vc"]), _ptr_d(out["Vv"]),
        _ptr_d(out["dvdl"]), _ptr_d(out["foreign_energy"]), _ptr_d(out["foreign_dvdl"]),
        int(repeats), _ptr_d(seconds),
    )  # fmt: skip
    if rc != 0:
        raise RuntimeError(f"oracle dispatch failed with code {rc}")
    out["seconds"] = (float(seconds[0]), float(seconds[1]))
    return out


# ---------------------------------------------------------------------------------------------
# the reference itself (oracle/_ref)
# ---------------------------------------------------------------------------------------------
_ref_libs: dict[str, ctypes.CDLL] = {}


def host_has_avx512() -> bool:
    try:
        with open("/proc/cpuinfo") as fh:
            txt = fh.read()
        return all(f in txt for f in ("avx512f", "avx512dq", "avx512bw", "avx512vl", "avx512cd"))
    except OSError:
        return False


def host_has_avx2() -> bool:
    try:
        with open("/proc/cpuinfo") as fh:
            txt = fh.read()
        return "avx2" in txt and "fma" in txt
    except OSError:
        return False


def ref_variant(precision: str) -> str | None:
    """Name of the oracle/_ref library usable on this host for 'sp' or 'dp', or None."""
    if precision == "dp":
        cands = ["dp"] if host_has_avx2() else []
    else:
        cands = (["sp_avx512"] if host_has_avx512() else []) + (["sp_avx2"] if host_has_avx2() else [])
    for c in cands:
        if os.path.exists(os.path.join(REF_DIR, f"libfepref_{c}.so")):
            return c
    return None


def have_ref(precision: str = "dp") -> bool:
    return ref_variant(precision) is not None


def _load_ref(variant: str) -> ctypes.CDLL:
    if variant not in _ref_libs:
        lib = ctypes.CDLL(os.path.join(REF_DIR, f"libfepref_{variant}.so"))
        lib.fepref_dispatch.argtypes = _DISPATCH_ARGTYPES
        lib.fepref_dispatch.restype = ctypes.c_int
        lib.fepref_real_bytes.restype = ctypes.c_int
        lib.fepref_simd_string.restype = ctypes.c_char_p
        _ref_libs[variant] = lib
    return _ref_libs[variant]


def run_ref(problem, flags, *, precision="dp", nthreads=1, use_simd=True, repeats=1):
    variant = ref_variant(precision)
    if variant is None:
        raise RuntimeError(f"oracle/_ref has no usable '{precision}' library on this host")
    lib = _load_ref(variant)
    out = _call_dispatch(lib.fepref_dispatch, problem, flags, nthreads, use_simd, repeats)
    out["variant"] = variant
    out["simd"] = lib.fepref_simd_string().decode() if use_simd else "scalar"
    return out


# ---------------------------------------------------------------------------------------------
# the fork's CUDA FEP kernels (oracle/_ref/libfepfork_cuda.so)
# ---------------------------------------------------------------------------------------------
FORK_CUDA_LIB = os.path.join(REF_DIR, "libfepfork_cuda.so")
_fork_lib = None

FORK_CUDA_UNSUPPORTED = {1: "electrostatics type", 2: "Gapsys soft-core", 3: "LJ-PME / potential-switch / force-switch",
                         4: "energy groups", 5: "rcoulomb != rvdw"}


def have_fork_cuda() -> bool:
    return os.path.exists(FORK_CUDA_LIB)


def _load_fork_cuda() -> ctypes.CDLL:
    global _fork_lib
    if _fork_lib is None:
        lib = ctypes.CDLL(FORK_CUDA_LIB)
        lib.fepfork_dispatch.argtypes = _DISPATCH_ARGTYPES
        lib.fepfork_dispatch.restype = ctypes.c_int
        lib.fepfork_supports.argtypes = [ctypes.POINTER(CParamsD), ctypes.c_int]
        lib.fepfork_supports.restype = ctypes.c_int
        lib.fepfork_describe.restype = ctypes.c_char_p
        _fork_lib = lib
    return _fork_lib


def fork_cuda_unsupported(problem) -> str | None:
    """Why the fork's GPU kernels cannot run this problem (None if they can)."""
    cp = _params_d(problem.params)
    code = _load_fork_cuda().fepfork_supports(ctypes.byref(cp), problem.nenergrp_pairs)
    return FORK_CUDA_UNSUPPORTED.get(code, f"code {code}") if code else None


def run_fork_cuda(problem, flags, *, repeats=1):
    """One step on the current CUDA device; `seconds` = device time of (current-lambda kernel,
    foreign-lambda kernel), best of `repeats`."""
    lib = _load_fork_cuda()
    out = _call_dispatch(lib.fepfork_dispatch, problem, flags, 1, 0, repeats)
    out["variant"] = "fork_cuda"
    return out


# ---------------------------------------------------------------------------------------------
# our C restatement (oracle/fep_oracle.c)
# ---------------------------------------------------------------------------------------------
_port_lib = None


def build_port(force: bool = False) -> str:
    src = os.path.join(HERE, "fep_oracle.c")
    if force or not os.path.exists(PORT_LIB) or os.path.getmtime(PORT_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "libfep_oracle.so"], stdout=subprocess.DEVNULL)
    return PORT_LIB


def _load_port() -> ctypes.CDLL:
    global _port_lib
    if _port_lib is None:
        build_port()
        lib = ctypes.CDLL(PORT_LIB)
        lib.fep_oracle_dispatch.argtypes = _DISPATCH_ARGTYPES
        lib.fep_oracle_dispatch.restype = ctypes.c_int
        _port_lib = lib
    return _port_lib


def run_port(problem, flags, *, nthreads=1, repeats=1):
    lib = _load_port()
    out = _call_dispatch(lib.fep_oracle_dispatch, problem, flags, nthreads, 0, repeats)
    out["variant"] = "port"
    return out


def run_best(problem, flags, **kw):
    """The strongest double-precision checker available: the reference itself if
    oracle/_ref was built, else the C restatement."""
    if have_ref("dp"):
        return run_ref(problem, flags, precision="dp", **kw)
    kw.pop("use_simd", None)
    return run_port(problem, flags, **kw)


# ---------------------------------------------------------------------------------------------
# perturbed 1-4 pairs (oracle/fep_oracle.c: fep_oracle_pairs14)
# ---------------------------------------------------------------------------------------------
def run_pairs14(problem, compute_virial=True):
    """Double-precision restatement of do_pairs(F_LJ14) for perturbed pairs; returns f [N,3],
    fshift [45,3], Vc [G] (Coulomb-14), Vv [G] (LJ-14), dvdl [2]."""
    lib = _load_port()
    fn = lib.fep_oracle_pairs14
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.POINTER(CParamsD), ctypes.c_double, ctypes.c_int, _IP, _DP, _DP, _DP, _DP, _DP, _DP, _DP, _DP,
                   ctypes.c_int, _IP, ctypes.c_double, ctypes.c_double, _DP, _DP, _DP, _DP, _DP]
    p = problem
    out = dict(f=np.zeros((p.natoms, 3)), fshift=np.zeros((45, 3)), Vc=np.zeros(p.nenergrp_pairs),
               Vv=np.zeros(p.nenergrp_pairs), dvdl=np.zeros(2))
    cp = _params_d(p.params)
    ia, gid = _i(p.iatoms), _i(p.gid)
    arrs = [_d(a) for a in (p.c6A, p.c12A, p.c6B, p.c12B, p.x, p.qA, p.qB, p.box_diag)]
    lam = _d(p.lambda_)
    rc = fn(ctypes.byref(cp), float(p.fudgeQQ), p.npairs, _ptr_i(ia), *[_ptr_d(a) for a in arrs], int(p.pbc_type),
            _ptr_i(gid), float(lam[2]), float(lam[3]), _ptr_d(out["f"]),
            _ptr_d(out["fshift"]) if compute_virial else None, _ptr_d(out["Vc"]), _ptr_d(out["Vv"]), _ptr_d(out["dvdl"]))
    if rc != 0:
        raise RuntimeError(f"pairs14 oracle failed with code {rc}")
    return out
