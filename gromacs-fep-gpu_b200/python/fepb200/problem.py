"""A self-contained input set for the perturbed-pair kernel ("problem").

It bundles exactly the arguments `gmx_nb_free_energy_kernel()` takes
(reference: src/gromacs/gmxlib/nonbonded/nb_free_energy.h:53-72) plus the foreign-lambda
table that `dispatchFreeEnergyKernel()` loops over (src/gromacs/nbnxm/freeenergydispatch.cpp:236-306):

  coordinates rvec[N], chargeA/B real[N], typeA/B int[N], nbfp real[2 T^2] (+ nbfp_grid),
  shiftvec rvec[45], the FEP t_nblist (iinr, gid, shift, jindex, jjnr, excl_fep; reference:
  src/gromacs/mdtypes/nblist.h:41-55), lambda[7], all_lambda_coul/vdw[L], interaction constants.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field

import numpy as np

from .params import (
    CENTRAL_SHIFT_INDEX,
    LAMBDA_COUL,
    LAMBDA_VDW,
    NUM_LAMBDA_COMPONENTS,
    NUM_SHIFT_VECTORS,
    Params,
)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


@dataclass
class FepList:
    """The FEP `t_nblist` in CSR form; all arrays int32 (nblist.h:41-55)."""

    iinr: np.ndarray
    gid: np.ndarray
    shift: np.ndarray
    jindex: np.ndarray
    jjnr: np.ndarray
    excl_fep: np.ndarray

    def __post_init__(self):
        self.iinr = _i32(self.iinr)
        self.gid = _i32(self.gid)
        self.shift = _i32(self.shift)
        self.jindex = _i32(self.jindex)
        self.jjnr = _i32(self.jjnr)
        self.excl_fep = _i32(self.excl_fep)
        assert self.jindex.shape[0] == self.iinr.shape[0] + 1
        assert self.jjnr.shape[0] == self.excl_fep.shape[0]
        assert self.nri == 0 or int(self.jindex[-1]) == self.jjnr.shape[0]

    @property
    def nri(self) -> int:
        return int(self.iinr.shape[0])

    @property
    def nrj(self) -> int:
        return int(self.jjnr.shape[0])

    def slice_entries(self, e0: int, e1: int) -> "FepList":
        """Entries [e0, e1) as an independent list (used for sharding)."""
        j0, j1 = int(self.jindex[e0]), int(self.jindex[e1])
        return FepList(
            self.iinr[e0:e1],
            self.gid[e0:e1],
            self.shift[e0:e1],
            self.jindex[e0 : e1 + 1] - j0,
            self.jjnr[j0:j1],
            self.excl_fep[j0:j1],
        )


    def select_pairs(self, keep) -> "FepList":
        """The pairs with keep[k] true as an independent list; entries keep their order and are cut where pairs
        drop out, entries left without pairs disappear (the fused multi-GPU exchange splits the list by trips of
        the device layout, which gather pairs from many i-entries)."""
        keep = np.asarray(keep, bool)
        ent = np.repeat(np.arange(self.nri), np.diff(np.asarray(self.jindex, np.int64)))
        cnt = np.bincount(ent[keep], minlength=self.nri)
        live = cnt > 0
        jindex = np.concatenate([[0], np.cumsum(cnt[live])]).astype(self.jindex.dtype)
        return FepList(self.iinr[live], self.gid[live], self.shift[live], jindex, self.jjnr[keep], self.excl_fep[keep])


@dataclass
class Problem:
    name: str
    params: Params
    ntype: int
    nbfp: np.ndarray  # float32 [2*T*T]  (6*C6, 12*C12)
    nbfp_grid: np.ndarray  # float32 [2*T*T]  (6*C6grid, 0)
    x: np.ndarray  # float32 [N,3]
    qA: np.ndarray
    qB: np.ndarray
    typeA: np.ndarray
    typeB: np.ndarray
    shiftvec: np.ndarray  # float32 [45,3]
    nblist: FepList
    nenergrp_pairs: int = 1
    lambda_: np.ndarray = field(default_factory=lambda: np.zeros(NUM_LAMBDA_COMPONENTS, np.float32))
    all_lambda_coul: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    all_lambda_vdw: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    box: np.ndarray | None = None
    perturbed: np.ndarray | None = None  # indices of perturbed atoms (informational)
    # float32 is what crosses the C-ABI; float64 is only used to feed the double-precision
    # oracle the exact inputs of the reference's unit test (tests/golden)
    real_dtype: type = np.float32

    def __post_init__(self):
        def _r(a):
            return np.ascontiguousarray(a, dtype=self.real_dtype)

        self.nbfp = _r(self.nbfp).ravel()
        self.nbfp_grid = _r(self.nbfp_grid).ravel()
        self.x = _r(self.x).reshape(-1, 3)
        self.qA, self.qB = _r(self.qA), _r(self.qB)
        self.typeA, self.typeB = _i32(self.typeA), _i32(self.typeB)
        self.shiftvec = _r(self.shiftvec).reshape(NUM_SHIFT_VECTORS, 3)
        self.lambda_ = _r(self.lambda_)
        self.all_lambda_coul = _r(self.all_lambda_coul)
        self.all_lambda_vdw = _r(self.all_lambda_vdw)
        assert self.nbfp.shape[0] == 2 * self.ntype * self.ntype
        assert self.lambda_.shape[0] == NUM_LAMBDA_COMPONENTS
        assert self.all_lambda_coul.shape == self.all_lambda_vdw.shape

    @property
    def natoms(self) -> int:
        return int(self.x.shape[0])

    @property
    def n_foreign(self) -> int:
        return int(self.all_lambda_coul.shape[0])

    def set_lambda(self, coul: float, vdw: float | None = None) -> None:
        vdw = coul if vdw is None else vdw
        lam = np.full(NUM_LAMBDA_COMPONENTS, coul, self.real_dtype)
        lam[LAMBDA_COUL], lam[LAMBDA_VDW] = coul, vdw
        self.lambda_ = lam

    # ---- persistence (flat .npz, the "problem file" of SURVEY section 7) -----------------
    def save(self, path: str) -> None:
        meta = dict(
            name=self.name,
            params=self.params.to_dict(),
            ntype=self.ntype,
            nenergrp_pairs=self.nenergrp_pairs,
        )
        extra = {}
        if self.box is not None:
            extra["box"] = np.asarray(self.box, np.float32)
        if self.perturbed is not None:
            extra["perturbed"] = _i32(self.perturbed)
        np.savez_compressed(
            path,
            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
            nbfp=self.nbfp,
            nbfp_grid=self.nbfp_grid,
            x=self.x,
            qA=self.qA,
            qB=self.qB,
            typeA=self.typeA,
            typeB=self.typeB,
            shiftvec=self.shiftvec,
            iinr=self.nblist.iinr,
            gid=self.nblist.gid,
            shift=self.nblist.shift,
            jindex=self.nblist.jindex,
            jjnr=self.nblist.jjnr,
            excl_fep=self.nblist.excl_fep,
            lambda_=self.lambda_,
            all_lambda_coul=self.all_lambda_coul,
            all_lambda_vdw=self.all_lambda_vdw,
            **extra,
        )

    @staticmethod
    def load(path: str) -> "Problem":
        z = np.load(path)
        meta = json.loads(bytes(z["meta"]).decode())
        return Problem(
            name=meta["name"],
            params=Params.from_dict(meta["params"]),
            ntype=meta["ntype"],
            nbfp=z["nbfp"],
            nbfp_grid=z["nbfp_grid"],
            x=z["x"],
            qA=z["qA"],
            qB=z["qB"],
            typeA=z["typeA"],
            typeB=z["typeB"],
            shiftvec=z["shiftvec"],
            nblist=FepList(z["iinr"], z["gid"], z["shift"], z["jindex"], z["jjnr"], z["excl_fep"]),
            nenergrp_pairs=meta["nenergrp_pairs"],
            lambda_=z["lambda_"],
            all_lambda_coul=z["all_lambda_coul"],
            all_lambda_vdw=z["all_lambda_vdw"],
            box=z["box"] if "box" in z else None,
            perturbed=z["perturbed"] if "perturbed" in z else None,
        )


def shift_vectors(box: np.ndarray) -> np.ndarray:
    """The 45 shift vectors of a (triclinic) box, index n = 5*(3*(m+1)+(l+1))+(k+2),
    vec = k*box[0] + l*box[1] + m*box[2] (reference: src/gromacs/pbcutil/pbc.cpp:1218-1233,
    api/legacy/include/gromacs/pbcutil/ishift.h:41-54)."""
    box = np.asarray(box, np.float64).reshape(3, 3)
    out = np.zeros((NUM_SHIFT_VECTORS, 3))
    n = 0
    for m in (-1, 0, 1):
        for l in (-1, 0, 1):
            for k in (-2, -1, 0, 1, 2):
                out[n] = k * box[0] + l * box[1] + m * box[2]
                n += 1
    assert np.all(out[CENTRAL_SHIFT_INDEX] == 0)
    return out.astype(np.float32)


def nbfp_from_c6c12(c6: np.ndarray, c12: np.ndarray) -> np.ndarray:
    """nbfp[2*(T*ti+tj)+{0,1}] = {6*C6, 12*C12} (reference: src/gromacs/mdlib/forcerec.cpp:115-152)."""
    c6 = np.asarray(c6, np.float64)
    c12 = np.asarray(c12, np.float64)
    out = np.empty(c6.shape + (2,), np.float64)
    out[..., 0] = 6.0 * c6
    out[..., 1] = 12.0 * c12
    return out.astype(np.float32).ravel()


def nbfp_grid_geometric(c6_diag: np.ndarray) -> np.ndarray:
    """LJ-PME grid C6 with geometric mixing: grid[2*(T*i+j)] = 6*sqrt(c6_ii*c6_jj), odd slots 0
    (reference: src/gromacs/mdlib/forcerec.cpp:154-190)."""
    c6_diag = np.asarray(c6_diag, np.float64)
    t = c6_diag.shape[0]
    # the reference computes in `real`; do the sqrt in float32 like the mixed-precision build
    c6 = np.sqrt((c6_diag[:, None] * c6_diag[None, :]).astype(np.float32))
    out = np.zeros((t, t, 2), np.float32)
    out[..., 0] = c6 * np.float32(6.0)
    return out.ravel()
