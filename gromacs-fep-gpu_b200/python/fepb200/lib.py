"""ctypes binding of libfepb200.so (the C-ABI declared in include/fepb200.h) and `FepContext`,
the host-side object that plays the role `FreeEnergyDispatch` plays in the reference
(src/gromacs/nbnxm/freeenergydispatch.cpp:63-459): it is told the constants once, the atoms and the
FEP pair list at search steps, the lambdas when they change, and is asked once per step for forces,
shift forces, per-energy-group-pair Vc/Vvdw, dV/dlambda and the foreign-lambda energies.

There is no fallback: if the shared library is missing, or no sm_100 device is present, creation
raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .params import CParams, NUM_LAMBDA_COMPONENTS, NUM_SHIFT_VECTORS

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FEPB200_LIB") or os.path.normpath(os.path.join(_HERE, "..", "..", "lib", "libfepb200.so"))

_FP = ctypes.POINTER(ctypes.c_float)
_DP = ctypes.POINTER(ctypes.c_double)
_IP = ctypes.POINTER(ctypes.c_int)
_VP = ctypes.c_void_p


class CListView(ctypes.Structure):
    """struct fepb200_list_view"""

    _fields_ = [("nri", ctypes.c_int), ("iinr", _IP), ("gid", _IP), ("shift", _IP), ("jindex", _IP), ("jjnr", _IP),
                ("excl_fep", _IP)]


class CLayout(ctypes.Structure):
    """`struct fepb200_layout`."""

    _fields_ = [
        ("natoms", ctypes.c_int),
        ("ntouched", ctypes.c_int),
        ("nri", ctypes.c_int),
        ("nrj", ctypes.c_longlong),
        ("nri_total", ctypes.c_int),
        ("nrj_total", ctypes.c_longlong),
        ("nenergrp", ctypes.c_int),
        ("nforeign", ctypes.c_int),
        ("f32_words", ctypes.c_longlong),
        ("f64_words", ctypes.c_longlong),
        ("off_fshift", ctypes.c_longlong),
        ("off_vc", ctypes.c_longlong),
        ("off_vv", ctypes.c_longlong),
        ("off_dvdl", ctypes.c_longlong),
        ("off_foreign_e", ctypes.c_longlong),
        ("off_foreign_dvdl", ctypes.c_longlong),
    ]


# every symbol include/fepb200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "fepb200_create": (ctypes.c_int, [ctypes.POINTER(_VP), ctypes.c_int]),
    "fepb200_destroy": (ctypes.c_int, [_VP]),
    "fepb200_last_error": (ctypes.c_char_p, [_VP]),
    "fepb200_describe": (ctypes.c_char_p, [_VP]),
    "fepb200_set_stream": (ctypes.c_int, [_VP, _VP]),
    "fepb200_set_params": (ctypes.c_int, [_VP, ctypes.POINTER(CParams)]),
    "fepb200_set_nbfp": (ctypes.c_int, [_VP, ctypes.c_int, _FP, _FP]),
    "fepb200_set_atoms": (ctypes.c_int, [_VP, ctypes.c_int, _FP, _FP, _IP, _IP]),
    "fepb200_set_list": (ctypes.c_int, [_VP, ctypes.c_int, _IP, _IP, _IP, _IP, _IP, _IP, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int]),
    "fepb200_set_lists": (ctypes.c_int, [_VP, ctypes.c_int, _VP, _IP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "fepb200_get_list": (ctypes.c_int, [_VP, _IP, _IP, _IP, _IP, _IP, _IP, _IP]),
    "fepb200_touched_atoms": (ctypes.c_int, [_VP, _IP]),
    "fepb200_result_layout": (ctypes.c_int, [_VP, ctypes.POINTER(CLayout)]),
    "fepb200_set_lambdas": (ctypes.c_int, [_VP, _FP, ctypes.c_int, _FP, _FP]),
    "fepb200_compute": (ctypes.c_int, [_VP, _FP, _FP, ctypes.c_int, _FP, _FP, _DP, _DP, _DP, _DP, _DP]),
    "fepb200_upload_x": (ctypes.c_int, [_VP, _FP, _FP]),
    "fepb200_gather_x_device": (ctypes.c_int, [_VP, _VP, _FP]),
    "fepb200_gather_xq_device": (ctypes.c_int, [_VP, _VP, _FP]),
    "fepb200_launch": (ctypes.c_int, [_VP, ctypes.c_int, _VP]),
    "fepb200_add_forces_device": (ctypes.c_int, [_VP, _VP, ctypes.c_int]),
    "fepb200_export_scalars_device": (ctypes.c_int, [_VP, ctypes.c_int] + [_VP] * 9),
    "fepb200_wait": (ctypes.c_int, [_VP]),
    "fepb200_result_device_ptrs": (ctypes.c_int, [_VP, ctypes.POINTER(_VP), ctypes.POINTER(_VP)]),
    "fepb200_result_block_bytes": (ctypes.c_size_t, [_VP]),
    "fepb200_publish_result": (ctypes.c_int, [_VP, _VP]),
    "fepb200_set_partial_result_block": (ctypes.c_int, [_VP, _VP]),
    "fepb200_set_push_targets": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.POINTER(_VP)]),
    "fepb200_reduce_peers": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.POINTER(_VP), ctypes.POINTER(_VP), ctypes.c_int,
                                            ctypes.c_uint]),
    "fepb200_reduce_scatter_peers": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.POINTER(_VP), ctypes.POINTER(_VP), ctypes.c_int,
                                            ctypes.c_uint]),
    "fepb200_exchange_bytes": (ctypes.c_size_t, [_VP, ctypes.c_int]),
    "fepb200_set_peer_exchange": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_VP), ctypes.c_size_t]),
    "fepb200_peer_ranges": (ctypes.c_int, [_VP, _IP, _IP, _IP, _IP]),
    "fepb200_epilogue_trace": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]),
    "fepb200_download": (ctypes.c_int, [_VP, ctypes.c_int, _FP, _FP, _DP, _DP, _DP, _DP, _DP]),
    "fepb200_launch_count": (ctypes.c_longlong, [_VP]),
    "fepb200_last_launch_ms": (ctypes.c_int, [_VP, _FP]),
    "fepb200_set_profiling": (ctypes.c_int, [_VP, ctypes.c_int]),
    "fepb200_kernel_ms": (ctypes.c_int, [_VP, _FP]),
    "fepb200_pairs14_create": (ctypes.c_int, [ctypes.POINTER(_VP), ctypes.c_int]),
    "fepb200_pairs14_destroy": (ctypes.c_int, [_VP]),
    "fepb200_pairs14_last_error": (ctypes.c_char_p, [_VP]),
    "fepb200_pairs14_set_params": (ctypes.c_int, [_VP, ctypes.POINTER(CParams), ctypes.c_float]),
    "fepb200_pairs14_set_pairs": (ctypes.c_int, [_VP, ctypes.c_int, _FP, _FP, ctypes.c_int, _IP, ctypes.c_int, _FP, _FP,
                                                 _FP, _FP, _IP, ctypes.c_int]),
    "fepb200_pairs14_compute": (ctypes.c_int, [_VP, _FP, _FP, ctypes.c_int, _FP, ctypes.c_int, _FP, _FP, _DP, _DP, _DP]),
    "fepb200_pairs14_compute_foreign": (ctypes.c_int, [_VP, _FP, _FP, ctypes.c_int, ctypes.c_int, _FP, _FP, _DP, _DP]),
}  # fmt: skip

_lib = None


class FepError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"fepb200 error {code}: {message}")
        self.code = code


def load_library(path: str | None = None) -> ctypes.CDLL:
    """Loads libfepb200.so and declares every entry point; raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found: build it with `make -C gromacs-fep-gpu_b200/csrc` "
            "(or __graft_entry__.build()); there is no fallback implementation"
        )
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SYMBOLS.items():
        if os.environ.get("FEPB200_LIB") and not hasattr(lib, name):
            continue  # A/B timing against an older build of the library (tools/): newer entry points are absent
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _pf(a):
    return a.ctypes.data_as(_FP) if a is not None else None


def _pd(a):
    return a.ctypes.data_as(_DP) if a is not None else None


def _pi(a):
    return a.ctypes.data_as(_IP) if a is not None else None


class FepContext:
    """One context = one GPU, one stream, one shard of the FEP pair list."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        h = _VP()
        rc = self._lib.fepb200_create(ctypes.byref(h), int(device))
        if rc != 0:
            raise FepError(rc, self._lib.fepb200_last_error(None).decode())
        self._h = h
        self.device = device
        self._problem_natoms = 0

    # ---- plumbing -------------------------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise FepError(rc, self._lib.fepb200_last_error(self._h).decode())

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.fepb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def describe(self) -> str:
        return self._lib.fepb200_describe(self._h).decode()

    def set_stream(self, stream: int | None) -> None:
        """Adopt an external cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""
        self._check(self._lib.fepb200_set_stream(self._h, _VP(stream) if stream else None))

    # ---- inputs ---------------------------------------------------------------------------
    def set_params(self, params) -> None:
        c = params.to_c()
        self._check(self._lib.fepb200_set_params(self._h, ctypes.byref(c)))

    def set_nbfp(self, ntype: int, nbfp, nbfp_grid=None) -> None:
        nbfp = _f32(nbfp)
        grid = _f32(nbfp_grid) if nbfp_grid is not None else None
        self._check(self._lib.fepb200_set_nbfp(self._h, int(ntype), _pf(nbfp), _pf(grid)))

    def set_atoms(self, qA, qB, typeA, typeB) -> None:
        qA, qB, typeA, typeB = _f32(qA), _f32(qB), _i32(typeA), _i32(typeB)
        self._problem_natoms = int(qA.shape[0])
        self._check(self._lib.fepb200_set_atoms(self._h, self._problem_natoms, _pf(qA), _pf(qB), _pi(typeA), _pi(typeB)))

    def set_list(self, nblist, nenergrp_pairs: int = 1, rank: int = 0, nranks: int = 1) -> None:
        self._check(
            self._lib.fepb200_set_list(
                self._h, nblist.nri, _pi(nblist.iinr), _pi(nblist.gid), _pi(nblist.shift), _pi(nblist.jindex),
                _pi(nblist.jjnr), _pi(nblist.excl_fep) if nblist.excl_fep is not None else None,
                int(nenergrp_pairs), int(rank), int(nranks),
            )
        )  # fmt: skip

    def set_lists(self, nblists, nenergrp_pairs: int = 1, rank: int = 0, nranks: int = 1, atom_map=None) -> None:
        """The per-thread lists as the reference holds them (fepb200_set_lists): concatenated, mapped through `atom_map`
        (list atom index -> set_atoms index) and checked on the device."""
        views = (CListView * max(len(nblists), 1))()
        keep = []
        for v, nb in zip(views, nblists):
            arrs = [np.ascontiguousarray(a, np.int32) for a in (nb.iinr, nb.gid, nb.shift, nb.jindex, nb.jjnr)]
            excl = np.ascontiguousarray(nb.excl_fep, np.int32) if nb.excl_fep is not None else None
            keep += arrs + [excl]
            v.nri = int(nb.nri)
            v.iinr, v.gid, v.shift, v.jindex, v.jjnr = [_pi(a) for a in arrs]
            v.excl_fep = _pi(excl) if excl is not None else None
        amap = np.ascontiguousarray(atom_map, np.int32) if atom_map is not None else None
        self._check(self._lib.fepb200_set_lists(self._h, len(nblists), ctypes.cast(views, _VP),
                                                _pi(amap) if amap is not None else None, int(amap.shape[0]) if amap is not None else 0,
                                                int(nenergrp_pairs), int(rank), int(nranks)))

    def set_lambdas(self, lambda_, all_lambda_coul=(), all_lambda_vdw=()) -> None:
        lam = _f32(lambda_)
        assert lam.shape[0] == NUM_LAMBDA_COMPONENTS
        alc, alv = _f32(all_lambda_coul), _f32(all_lambda_vdw)
        assert alc.shape == alv.shape
        self._check(self._lib.fepb200_set_lambdas(self._h, _pf(lam), int(alc.shape[0]), _pf(alc), _pf(alv)))

    def set_problem(self, problem, rank: int = 0, nranks: int = 1) -> None:
        """Everything `init_nb_verlet` + a search step would hand over, from a `Problem`."""
        self.set_params(problem.params)
        self.set_nbfp(problem.ntype, problem.nbfp, problem.nbfp_grid)
        self.set_atoms(problem.qA, problem.qB, problem.typeA, problem.typeB)
        self.set_lambdas(problem.lambda_, problem.all_lambda_coul, problem.all_lambda_vdw)
        self.set_list(problem.nblist, problem.nenergrp_pairs, rank, nranks)

    # ---- queries --------------------------------------------------------------------------
    def layout(self) -> CLayout:
        lay = CLayout()
        self._check(self._lib.fepb200_result_layout(self._h, ctypes.byref(lay)))
        return lay

    def touched_atoms(self) -> np.ndarray:
        out = np.empty(self.layout().ntouched, np.int32)
        self._check(self._lib.fepb200_touched_atoms(self._h, _pi(out)))
        return out

    def get_list(self):
        """(first_entry, FepList) of the shard this context holds, read back from the device."""
        from .problem import FepList

        lay = self.layout()
        first = ctypes.c_int(0)
        iinr = np.empty(lay.nri, np.int32)
        gid = np.empty(lay.nri, np.int32)
        shift = np.empty(lay.nri, np.int32)
        jindex = np.empty(lay.nri + 1, np.int32)
        jjnr = np.empty(lay.nrj, np.int32)
        excl = np.empty(lay.nrj, np.int32)
        self._check(
            self._lib.fepb200_get_list(self._h, ctypes.byref(first), _pi(iinr), _pi(gid), _pi(shift), _pi(jindex),
                                       _pi(jjnr), _pi(excl))
        )  # fmt: skip
        if lay.nri == 0:
            jindex[:] = 0
        return first.value, FepList(iinr, gid, shift, jindex, jjnr, excl)

    # ---- the hot call ---------------------------------------------------------------------
    def new_outputs(self) -> dict:
        lay = self.layout()
        return dict(
            f=np.zeros((lay.natoms, 3), np.float32),
            fshift=np.zeros((NUM_SHIFT_VECTORS, 3), np.float32),
            Vc=np.zeros(lay.nenergrp),
            Vv=np.zeros(lay.nenergrp),
            dvdl=np.zeros(2),
            foreign_energy=np.zeros(lay.nforeign + 1),
            foreign_dvdl=np.zeros((lay.nforeign + 1, 2)),
        )

    def compute(self, x, shiftvec, flags: int, out: dict | None = None) -> dict:
        """fepb200_compute(): host buffers in, host buffers out (accumulated into `out`)."""
        out = self.new_outputs() if out is None else out
        x, sv = _f32(x), _f32(shiftvec)
        self._check(
            self._lib.fepb200_compute(self._h, _pf(x), _pf(sv), int(flags), _pf(out["f"]), _pf(out["fshift"]),
                                      _pd(out["Vc"]), _pd(out["Vv"]), _pd(out["dvdl"]), _pd(out["foreign_energy"]),
                                      _pd(out["foreign_dvdl"]))
        )  # fmt: skip
        return out

    # ---- device-resident variants ---------------------------------------------------------
    def upload_x(self, x, shiftvec) -> None:
        x, sv = _f32(x), _f32(shiftvec)
        self._check(self._lib.fepb200_upload_x(self._h, _pf(x), _pf(sv)))

    def gather_x_device(self, d_x_ptr: int, shiftvec) -> None:
        sv = _f32(shiftvec)
        self._check(self._lib.fepb200_gather_x_device(self._h, _VP(d_x_ptr), _pf(sv)))

    def gather_xq_device(self, d_xq_ptr: int, shiftvec) -> None:
        """Coordinates from a device-resident float4[natoms] {x, y, z, q} array (the nbnxm GPU atom data)."""
        sv = _f32(shiftvec)
        self._check(self._lib.fepb200_gather_xq_device(self._h, _VP(d_xq_ptr), _pf(sv)))

    def launch(self, flags: int, stream: int | None = None) -> None:
        self._check(self._lib.fepb200_launch(self._h, int(flags), _VP(stream) if stream else None))

    def add_forces_device(self, d_f_ptr: int, flags: int = 0) -> None:
        """Add the forces of the last launch() into a device-resident float[natoms][3] array."""
        self._check(self._lib.fepb200_add_forces_device(self._h, _VP(d_f_ptr), int(flags)))

    def export_scalars_device(self, flags: int, **ptrs: int) -> None:
        """Add the scalars of the last launch() into float device buffers laid out like the fork's NBAtomDataGpu
        outputs; keyword arguments eLJ, eElec, dvdlLJ, dvdlElec, eLJForeign, eElecForeign, dvdlLJForeign,
        dvdlElecForeign, fShift are device pointers (omitted = NULL)."""
        names = ("eLJ", "eElec", "dvdlLJ", "dvdlElec", "eLJForeign", "eElecForeign", "dvdlLJForeign", "dvdlElecForeign", "fShift")
        unknown = set(ptrs) - set(names)
        if unknown:
            raise TypeError(f"unknown buffers {sorted(unknown)}")
        args = [_VP(ptrs[n]) if ptrs.get(n) else None for n in names]
        self._check(self._lib.fepb200_export_scalars_device(self._h, int(flags), *args))

    def wait(self) -> None:
        self._check(self._lib.fepb200_wait(self._h))

    def download(self, flags: int, out: dict | None = None) -> dict:
        out = self.new_outputs() if out is None else out
        self._check(
            self._lib.fepb200_download(self._h, int(flags), _pf(out["f"]), _pf(out["fshift"]), _pd(out["Vc"]),
                                       _pd(out["Vv"]), _pd(out["dvdl"]), _pd(out["foreign_energy"]),
                                       _pd(out["foreign_dvdl"]))
        )  # fmt: skip
        return out

    def result_device_ptrs(self) -> tuple[int, int]:
        a, b = _VP(), _VP()
        self._check(self._lib.fepb200_result_device_ptrs(self._h, ctypes.byref(a), ctypes.byref(b)))
        return int(a.value or 0), int(b.value or 0)

    def launch_count(self) -> int:
        return int(self._lib.fepb200_launch_count(self._h))

    def last_launch_ms(self) -> float:
        ms = ctypes.c_float(0)
        self._check(self._lib.fepb200_last_launch_ms(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def set_profiling(self, on: bool) -> None:
        self._check(self._lib.fepb200_set_profiling(self._h, int(bool(on))))

    def kernel_ms(self) -> tuple[float, float, float]:
        """Device ms of (pass kernel, foreign kernel, epilogue kernel) of the last profiled launch."""
        ms = (ctypes.c_float * 3)()
        self._check(self._lib.fepb200_kernel_ms(self._h, ms))
        return float(ms[0]), float(ms[1]), float(ms[2])

    # ---- peer-memory reduction ------------------------------------------------------------
    def result_block_bytes(self) -> int:
        return int(self._lib.fepb200_result_block_bytes(self._h))

    def publish_result(self, d_block: int) -> None:
        self._check(self._lib.fepb200_publish_result(self._h, _VP(d_block)))

    def set_partial_result_block(self, d_block: int | None) -> None:
        self._check(self._lib.fepb200_set_partial_result_block(self._h, _VP(d_block) if d_block else None))

    def reduce_peers(self, peer_blocks: list[int], peer_flags: list[int] | None = None, rank: int = 0,
                     seq: int = 0) -> None:
        arr = (_VP * len(peer_blocks))(*[_VP(p) for p in peer_blocks])
        flg = (_VP * len(peer_flags))(*[_VP(p) for p in peer_flags]) if peer_flags else None
        self._check(self._lib.fepb200_reduce_peers(self._h, len(peer_blocks), arr, flg, int(rank), int(seq) & 0xFFFFFFFF))

    def set_push_targets(self, peer_blocks: list[int] | None) -> None:
        """This rank's receive block on every rank (peer-mapped addresses, rank order), or None: off."""
        if not peer_blocks:
            self._check(self._lib.fepb200_set_push_targets(self._h, 0, None))
            return
        arr = (_VP * len(peer_blocks))(*[_VP(p) for p in peer_blocks])
        self._check(self._lib.fepb200_set_push_targets(self._h, len(peer_blocks), arr))

    def reduce_scatter_peers(self, peer_blocks: list[int], peer_flags: list[int] | None = None, rank: int = 0,
                             seq: int = 0) -> None:
        """Force reduce-scatter + all-reduce of the scalars over peer memory: afterwards this context holds the forces of
        the atoms it owns (peer_ranges()) and all scalars."""
        arr = (_VP * len(peer_blocks))(*[_VP(p) for p in peer_blocks])
        flg = (_VP * len(peer_flags))(*[_VP(p) for p in peer_flags]) if peer_flags else None
        self._check(self._lib.fepb200_reduce_scatter_peers(self._h, len(peer_blocks), arr, flg, int(rank), int(seq) & 0xFFFFFFFF))

    # ---- fused peer exchange (pair kernels scatter over NVLink, every rank sums its atoms) ----
    def exchange_bytes(self, nranks: int) -> int:
        return int(self._lib.fepb200_exchange_bytes(self._h, int(nranks)))

    def set_peer_exchange(self, nranks: int, rank: int, peer_bufs: list[int] | None, nbytes: int = 0) -> None:
        arr = (_VP * len(peer_bufs))(*[_VP(p) for p in peer_bufs]) if peer_bufs else None
        self._check(self._lib.fepb200_set_peer_exchange(self._h, int(nranks), int(rank), arr, int(nbytes)))

    def peer_ranges(self) -> tuple[int, int, int, int]:
        """(pair_begin, pair_end, atom_begin, atom_end) of this context."""
        v = [ctypes.c_int(0) for _ in range(4)]
        self._check(self._lib.fepb200_peer_ranges(self._h, *[ctypes.byref(x) for x in v]))
        return tuple(int(x.value) for x in v)

    def epilogue_trace(self, enable: bool = True, max_blocks: int = 4096) -> np.ndarray:
        """Global-timer stamps [blocks, 4] (entry, pair kernels done, barrier passed, sums done) of the
        epilogue blocks of the last launch (empty before tracing was enabled); see fepb200.h."""
        buf = np.zeros((max_blocks, 4), np.uint64)
        n = self._lib.fepb200_epilogue_trace(self._h, int(bool(enable)), buf.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong)),
                                             int(max_blocks))
        if n < 0:
            self._check(n)
        return buf[:n]
