"""TEST INFRASTRUCTURE -- loaders for the two CPU checkers of the non-perturbed cluster-pair kernel.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

  * `run_port(cs, params, ...)` -- oracle/nb_oracle.c, our plain-C double-precision restatement of
    nbnxn_kernel_gpu_ref (built by oracle/Makefile into oracle/libnb_oracle.so).
  * `run_ref(cs, params, ...)`  -- oracle/_ref/libnbref_{dp,sp}.so, the reference's own kernel_gpu_ref.cpp compiled in
    place by oracle/ref_build/Makefile (present only when built in a container that has /root/reference).

Both take a `fepb200.synth_nb.ClusterSystem` and return dict(f [natoms,3], fshift [45,3], vc, vvdw) as float64.
`table`: (scale, n) -> the Ewald force table the reference kernel interpolates is filled with the analytical
function at that spacing; None -> the port evaluates the analytical function itself (the reference cannot).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(HERE, "libnb_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
MIN_RSQ_FLOAT = 3.82e-07  # c_nbnxnMinDistanceSquared, mixed precision (pairlist.h:167)
MIN_RSQ_DOUBLE = 1.0e-36  # double precision (pairlist.h:162)

_DP = ctypes.POINTER(ctypes.c_double)


class PortParams(ctypes.Structure):
    _fields_ = [("eeltype", ctypes.c_int)] + [(n, ctypes.c_double) for n in (
        "epsfac", "rcoulomb", "rvdw", "rlist", "k_rf", "c_rf", "sh_ewald", "beta", "disp_cpot", "rep_cpot",
        "min_rsq", "tab_scale")] + [("tab_size", ctypes.c_int), ("tableF", _DP), ("vdw_switch_kind", ctypes.c_int),
                                     ("rvdw_switch", ctypes.c_double)]  # fmt: skip


class RefParams(ctypes.Structure):
    _fields_ = [("eeltype", ctypes.c_int)] + [(n, ctypes.c_double) for n in (
        "epsfac", "rcoulomb", "rvdw", "rlist", "reactionFieldCoefficient", "reactionFieldShift", "sh_ewald",
        "ewaldcoeff_q", "dispersion_shift_cpot", "repulsion_shift_cpot", "tab_scale")] + [
        ("tab_size", ctypes.c_int), ("tableF", _DP)]  # fmt: skip


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _pd(a):
    return a.ctypes.data_as(_DP)


def build_port(force: bool = False) -> str:
    src = os.path.join(HERE, "nb_oracle.c")
    if force or not os.path.exists(PORT_LIB) or os.path.getmtime(PORT_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "libnb_oracle.so"], stdout=subprocess.DEVNULL)
    return PORT_LIB


_port = None
_refs: dict[str, ctypes.CDLL] = {}


def _load_port():
    global _port
    if _port is None:
        build_port()
        _port = ctypes.CDLL(PORT_LIB)
        _port.nbo_fill_table.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int, _DP]
        _port.nbo_ewald_force_lr.restype = ctypes.c_double
        _port.nbo_ewald_force_lr.argtypes = [ctypes.c_double, ctypes.c_double]
    return _port


def have_ref(precision: str = "dp") -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"libnbref_{precision}.so"))


def _load_ref(precision: str):
    if precision not in _refs:
        _refs[precision] = ctypes.CDLL(os.path.join(REF_DIR, f"libnbref_{precision}.so"))
    return _refs[precision]


def ewald_table(beta: float, scale: float, n: int) -> np.ndarray:
    tab = np.zeros(n)
    _load_port().nbo_fill_table(beta, scale, n, _pd(tab))
    return tab


def _table_for(params, table):
    if table is None:
        return 0.0, 0, None
    scale, n = table
    return float(scale), int(n), ewald_table(float(params.ewaldcoeff_q), float(scale), int(n))


def _common(cs):
    xq = _d(cs.xq).reshape(-1)
    typ = np.ascontiguousarray(cs.type, np.int32)
    nbfp = _d(cs.nbfp)
    sv = _d(cs.shiftvec).reshape(-1)
    sci, cj, excl = np.ascontiguousarray(cs.sci), np.ascontiguousarray(cs.cj), np.ascontiguousarray(cs.excl)
    return xq, typ, nbfp, sv, sci, cj, excl


def _outputs(natoms):
    return np.zeros((natoms, 3)), np.zeros((45, 3)), ctypes.c_double(0), ctypes.c_double(0)


def switch_kind(params) -> int:
    """0: none / potential shift, 1: force switch, 2: potential switch (InteractionModifiers 5 / 3) -- the modifiers of the
    reference's CUDA kernels; nbnxn_kernel_gpu_ref itself ignores them."""
    return {5: 1, 3: 2}.get(int(params.vdw_modifier), 0)


def run_port(cs, params, *, energy=True, table=None, min_rsq=MIN_RSQ_FLOAT, repeats=1, cuda_modifiers=False):
    """cuda_modifiers: apply the Lennard-Jones force / potential switch the way the reference's CUDA kernels do."""
    lib = _load_port()
    xq, typ, nbfp, sv, sci, cj, excl = _common(cs)
    scale, n, tab = _table_for(params, table)
    p = PortParams(int(params.eeltype), params.epsfac, params.rcoulomb, params.rvdw, cs.rlist,
                   params.reactionFieldCoefficient, params.reactionFieldShift, params.sh_ewald, params.ewaldcoeff_q,
                   params.dispersion_shift_cpot, params.repulsion_shift_cpot, min_rsq, scale, n,
                   _pd(tab) if tab is not None else None, switch_kind(params) if cuda_modifiers else 0, params.rvdw_switch)
    f, fsh, vc, vv = _outputs(cs.natoms)
    best = np.inf
    for _ in range(repeats):
        t0 = time.perf_counter()
        rc = lib.nbo_run(cs.natoms, _pd(xq), typ.ctypes.data_as(ctypes.c_void_p), cs.ntype, _pd(nbfp), ctypes.byref(p),
                         sci.shape[0], sci.ctypes.data_as(ctypes.c_void_p), cj.shape[0],
                         cj.ctypes.data_as(ctypes.c_void_p), excl.shape[0], excl.ctypes.data_as(ctypes.c_void_p),
                         _pd(sv), int(energy), _pd(f), _pd(fsh), ctypes.byref(vc), ctypes.byref(vv))
        best = min(best, time.perf_counter() - t0)
        assert rc == 0
    return dict(f=f, fshift=fsh, vc=vc.value, vvdw=vv.value, seconds=best)


def run_ref(cs, params, *, energy=True, table=(2000.0, 4096), precision="dp", repeats=1):
    lib = _load_ref(precision)
    xq, typ, nbfp, sv, sci, cj, excl = _common(cs)
    scale, n, tab = _table_for(params, table)
    p = RefParams(int(params.eeltype), params.epsfac, params.rcoulomb, params.rvdw, cs.rlist,
                  params.reactionFieldCoefficient, params.reactionFieldShift, params.sh_ewald, params.ewaldcoeff_q,
                  params.dispersion_shift_cpot, params.repulsion_shift_cpot, scale, n,
                  _pd(tab) if tab is not None else None)
    f, fsh, vc, vv = _outputs(cs.natoms)
    best = np.inf
    for _ in range(repeats):
        t0 = time.perf_counter()
        rc = lib.nbref_run(cs.natoms, _pd(xq), typ.ctypes.data_as(ctypes.c_void_p), cs.ntype, _pd(nbfp), ctypes.byref(p),
                           sci.shape[0], sci.ctypes.data_as(ctypes.c_void_p), cj.shape[0],
                           cj.ctypes.data_as(ctypes.c_void_p), excl.shape[0], excl.ctypes.data_as(ctypes.c_void_p),
                           _pd(sv), int(energy), _pd(f), _pd(fsh), ctypes.byref(vc), ctypes.byref(vv))
        best = min(best, time.perf_counter() - t0)
        assert rc == 0
    return dict(f=f, fshift=fsh, vc=vc.value, vvdw=vv.value, seconds=best)


def ref_struct_sizes(precision="dp"):
    out = (ctypes.c_int * 7)()
    _load_ref(precision).nbref_struct_sizes(out)
    return list(out)


def mask_perturbed(xq, typ, ntype, atoms):
    """nbo_mask_perturbed on copies (atomdata.cpp:930-964)."""
    lib = _load_port()
    xq = _d(xq).copy()
    typ = np.ascontiguousarray(typ, np.int32).copy()
    atoms = np.ascontiguousarray(atoms, np.int32)
    lib.nbo_mask_perturbed(_pd(xq.reshape(-1)), typ.ctypes.data_as(ctypes.c_void_p), int(ntype), int(atoms.shape[0]),
                           atoms.ctypes.data_as(ctypes.c_void_p))
    return xq, typ


# ---- the reference's own CUDA cluster-pair kernels, compiled in place for sm_100a (GPU baseline; needs a GPU) ----
FORK_CUDA_LIB = os.path.join(REF_DIR, "libnbfork_cuda.so")
_FP = ctypes.POINTER(ctypes.c_float)


class ForkParams(ctypes.Structure):
    _fields_ = [("eeltype", ctypes.c_int)] + [(n, ctypes.c_double) for n in (
        "epsfac", "rcoulomb", "rvdw", "krf", "crf", "sh_ewald", "ewaldcoeff_q", "dispersion_cpot", "repulsion_cpot")] + [
        ("vdw_switch_kind", ctypes.c_int), ("rvdw_switch", ctypes.c_double)]  # fmt: skip


def have_fork_cuda() -> bool:
    return os.path.exists(FORK_CUDA_LIB)


def run_fork_cuda(cs, params, *, energy=False, repeats=5):
    """oracle/_ref/libnbfork_cuda.so: nbnxn_kernel_Elec{Ew,RF}_VdwLJ_{F,VF}_cuda of the reference on the same list.
    Returns dict(f, fshift, vc, vvdw, ms) -- ms = device time of the kernel alone, warm caches, best of `repeats`."""
    lib = ctypes.CDLL(FORK_CUDA_LIB)
    p = ForkParams(int(params.eeltype), params.epsfac, params.rcoulomb, params.rvdw, params.reactionFieldCoefficient,
                   params.reactionFieldShift, params.sh_ewald, params.ewaldcoeff_q, params.dispersion_shift_cpot,
                   params.repulsion_shift_cpot, switch_kind(params), params.rvdw_switch)
    xq = np.ascontiguousarray(cs.xq, np.float32)
    typ = np.ascontiguousarray(cs.type, np.int32)
    nbfp = _d(cs.nbfp)
    sv = np.ascontiguousarray(cs.shiftvec, np.float32)
    sci, cj, excl = np.ascontiguousarray(cs.sci), np.ascontiguousarray(cs.cj), np.ascontiguousarray(cs.excl)
    f = np.zeros((cs.natoms, 3), np.float32)
    fsh = np.zeros((45, 3), np.float32)
    e = np.zeros(2, np.float32)
    ms = ctypes.c_float(0)
    rc = lib.nbfork_run(ctypes.byref(p), cs.ntype, _pd(nbfp), cs.natoms, xq.ctypes.data_as(_FP),
                        typ.ctypes.data_as(ctypes.c_void_p), sci.shape[0], sci.ctypes.data_as(ctypes.c_void_p), cj.shape[0],
                        cj.ctypes.data_as(ctypes.c_void_p), excl.shape[0], excl.ctypes.data_as(ctypes.c_void_p),
                        sv.ctypes.data_as(_FP), int(energy), int(repeats), f.ctypes.data_as(_FP), fsh.ctypes.data_as(_FP),
                        e.ctypes.data_as(_FP), ctypes.byref(ms))
    if rc != 0:
        raise RuntimeError(f"nbfork_run failed ({rc})")
    return dict(f=f.astype(np.float64), fshift=fsh.astype(np.float64), vvdw=float(e[0]), vc=float(e[1]), ms=float(ms.value))
