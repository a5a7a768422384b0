"""Perturbed 1-4 pair interactions: the input bundle of the reference's
`do_pairs(F_LJ14, ...)` perturbed branch (src/gromacs/listed_forces/pairs.cpp:516-835) and the
ctypes binding of the `fepb200_pairs14_*` entry points of libfepb200.so."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np

from .params import NUM_LAMBDA_COMPONENTS, Params

PBC_NONE, PBC_XYZ, PBC_XY = 0, 1, 2


@dataclass
class Pairs14Problem:
    params: Params  # epsfac, rcoulomb and the soft-core parameters are used
    fudgeQQ: float
    iatoms: np.ndarray  # int32 [npairs,3]: 1-4 type, ai, aj  (t_iatom layout of the reference)
    c6A: np.ndarray  # per 1-4 type (t_iparams::lj14)
    c12A: np.ndarray
    c6B: np.ndarray
    c12B: np.ndarray
    x: np.ndarray  # [N,3]
    qA: np.ndarray
    qB: np.ndarray
    box_diag: np.ndarray  # rectangular box
    pbc_type: int = PBC_XYZ
    gid: np.ndarray | None = None  # energy-group pair per 1-4 pair
    nenergrp_pairs: int = 1
    lambda_: np.ndarray = field(default_factory=lambda: np.zeros(NUM_LAMBDA_COMPONENTS, np.float32))
    real_dtype: type = np.float32

    def __post_init__(self):
        r = self.real_dtype
        self.iatoms = np.ascontiguousarray(self.iatoms, np.int32).reshape(-1, 3)
        for k in ("c6A", "c12A", "c6B", "c12B", "qA", "qB", "box_diag", "lambda_"):
            setattr(self, k, np.ascontiguousarray(getattr(self, k), r))
        self.x = np.ascontiguousarray(self.x, r).reshape(-1, 3)
        if self.gid is None:
            self.gid = np.zeros(self.npairs, np.int32)
        self.gid = np.ascontiguousarray(self.gid, np.int32)

    @property
    def npairs(self) -> int:
        return int(self.iatoms.shape[0])

    @property
    def natoms(self) -> int:
        return int(self.x.shape[0])


class Pairs14Context:
    """fepb200_pairs14_*: perturbed 1-4 pairs on the GPU (one context per list)."""

    def __init__(self, device: int = 0):
        from .lib import FepError, load_library

        self._lib = load_library()
        h = ctypes.c_void_p()
        rc = self._lib.fepb200_pairs14_create(ctypes.byref(h), int(device))
        if rc != 0:
            raise FepError(rc, self._lib.fepb200_last_error(None).decode())
        self._h = h

    def _check(self, rc):
        from .lib import FepError

        if rc != 0:
            raise FepError(rc, self._lib.fepb200_pairs14_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fepb200_pairs14_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_problem(self, p: Pairs14Problem) -> None:
        f32 = lambda a: np.ascontiguousarray(a, np.float32)  # noqa: E731
        fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))  # noqa: E731
        ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))  # noqa: E731
        c = p.params.to_c()
        self._check(self._lib.fepb200_pairs14_set_params(self._h, ctypes.byref(c), ctypes.c_float(p.fudgeQQ)))
        qa, qb = f32(p.qA), f32(p.qB)
        c6a, c12a, c6b, c12b = f32(p.c6A), f32(p.c12A), f32(p.c6B), f32(p.c12B)
        ia, gid = np.ascontiguousarray(p.iatoms, np.int32), np.ascontiguousarray(p.gid, np.int32)
        self._check(self._lib.fepb200_pairs14_set_pairs(self._h, p.natoms, fp(qa), fp(qb), p.npairs, ip(ia),
                                                        int(c6a.shape[0]), fp(c6a), fp(c12a), fp(c6b), fp(c12b),
                                                        ip(gid), int(p.nenergrp_pairs)))
        self._natoms, self._ngrp = p.natoms, p.nenergrp_pairs

    def compute(self, p: Pairs14Problem, flags: int, out: dict | None = None) -> dict:
        f32 = lambda a: np.ascontiguousarray(a, np.float32)  # noqa: E731
        fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))  # noqa: E731
        dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))  # noqa: E731
        if out is None:
            out = dict(f=np.zeros((self._natoms, 3), np.float32), fshift=np.zeros((45, 3), np.float32),
                       Vc=np.zeros(self._ngrp), Vv=np.zeros(self._ngrp), dvdl=np.zeros(2))
        x, box, lam = f32(p.x), f32(p.box_diag), f32(p.lambda_)
        self._check(self._lib.fepb200_pairs14_compute(self._h, fp(x), fp(box), int(p.pbc_type), fp(lam), int(flags),
                                                      fp(out["f"]), fp(out["fshift"]), dp(out["Vc"]), dp(out["Vv"]),
                                                      dp(out["dvdl"])))
        return out

    def compute_foreign(self, p: Pairs14Problem, lambda_coul, lambda_vdw) -> tuple[np.ndarray, np.ndarray]:
        """(energy[n], dvdl[n, 2]) of the energy-only evaluations at all the given lambda points, one library call
        (fepb200_pairs14_compute_foreign)."""
        f32 = lambda a: np.ascontiguousarray(a, np.float32)  # noqa: E731
        fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))  # noqa: E731
        dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))  # noqa: E731
        lc, lv = f32(lambda_coul), f32(lambda_vdw)
        assert lc.shape == lv.shape and lc.ndim == 1
        e, d = np.zeros(lc.shape[0]), np.zeros((lc.shape[0], 2))
        x, box = f32(p.x), f32(p.box_diag)
        self._check(self._lib.fepb200_pairs14_compute_foreign(self._h, fp(x), fp(box), int(p.pbc_type), int(lc.shape[0]), fp(lc),
                                                              fp(lv), dp(e), dp(d)))
        return e, d
