"""ctypes binding of the non-perturbed cluster-pair part of libfepb200.so (include/fepb200_nb.h) and `NbContext`,
the host-side object that plays the role of the reference's nbnxm GPU module for one locality
(src/gromacs/nbnxm/nbnxm_gpu_data_mgmt.cpp gpu_init_atomdata / gpu_init_pairlist, cuda/nbnxm_cuda.cu gpu_launch_kernel):
atoms and the cluster pair list at search steps, one launch per step.  No fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import lib as L
from .params import CParams, DO_FORCE, DO_POTENTIAL, DO_SHIFTFORCE, CLEAR_OUTPUTS, NUM_SHIFT_VECTORS

_FP, _DP, _IP, _VP = L._FP, L._DP, L._IP, L._VP
NB_Q_FROM_XQ = 1 << 20
NB_SHIFTVEC_ON_DEVICE = 1 << 21

# every symbol include/fepb200_nb.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "fepb200_nb_create": (ctypes.c_int, [ctypes.POINTER(_VP), ctypes.c_int]),
    "fepb200_nb_destroy": (ctypes.c_int, [_VP]),
    "fepb200_nb_last_error": (ctypes.c_char_p, [_VP]),
    "fepb200_nb_set_stream": (ctypes.c_int, [_VP, _VP]),
    "fepb200_nb_set_params": (ctypes.c_int, [_VP, ctypes.POINTER(CParams)]),
    "fepb200_nb_set_nbfp": (ctypes.c_int, [_VP, ctypes.c_int, _FP]),
    "fepb200_nb_set_atoms": (ctypes.c_int, [_VP, ctypes.c_int, _IP, _FP]),
    "fepb200_nb_mask_perturbed": (ctypes.c_int, [_VP, ctypes.c_int, _IP]),
    "fepb200_nb_get_atoms": (ctypes.c_int, [_VP, _IP, _FP]),
    "fepb200_nb_set_pairlist": (ctypes.c_int, [_VP, ctypes.c_int, _VP, ctypes.c_int, _VP, ctypes.c_int, _VP]),
    "fepb200_nb_use_device_list": (ctypes.c_int, [_VP, _VP, _VP, _VP]),
    "fepb200_nb_compute": (ctypes.c_int, [_VP, _FP, _FP, ctypes.c_int, _FP, _FP, _DP, _DP]),
    "fepb200_nb_compute_xyzq": (ctypes.c_int, [_VP, _FP, _FP, ctypes.c_int, _FP, _FP, _DP, _DP]),
    "fepb200_nb_launch_device": (ctypes.c_int, [_VP, _VP, _FP, ctypes.c_int, _VP, _VP, _VP]),
    "fepb200_nb_launch_device_float_energies": (ctypes.c_int, [_VP, _VP, _FP, ctypes.c_int, _VP, _VP, _VP, _VP]),
    "fepb200_nb_export_energies_device": (ctypes.c_int, [_VP, _VP, _VP]),
    "fepb200_nb_wait": (ctypes.c_int, [_VP]),
    "fepb200_nb_launch_count": (ctypes.c_longlong, [_VP]),
    "fepb200_nb_last_kernel_ms": (ctypes.c_int, [_VP, _FP]),
    "fepb200_nb_cluster_pairs": (ctypes.c_longlong, [_VP]),
}  # fmt: skip

_declared = False


def load_library() -> ctypes.CDLL:
    global _declared
    lib = L.load_library()
    if not _declared:
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = restype, argtypes
        _declared = True
    return lib


class NbContext:
    """One handle = one GPU, one stream, one cluster pair list."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        h = _VP()
        rc = self._lib.fepb200_nb_create(ctypes.byref(h), int(device))
        if rc != 0:
            raise L.FepError(rc, self._lib.fepb200_nb_last_error(None).decode())
        self._h = h
        self.natoms = 0

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise L.FepError(rc, self._lib.fepb200_nb_last_error(self._h).decode())

    def close(self) -> None:
        if self._h:
            self._lib.fepb200_nb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream) -> None:
        self._check(self._lib.fepb200_nb_set_stream(self._h, _VP(stream) if stream else None))

    def set_params(self, params) -> None:
        c = params.to_c()
        self._check(self._lib.fepb200_nb_set_params(self._h, ctypes.byref(c)))

    def set_nbfp(self, ntype: int, nbfp) -> None:
        nbfp = L._f32(nbfp)
        self._check(self._lib.fepb200_nb_set_nbfp(self._h, int(ntype), L._pf(nbfp)))

    def set_atoms(self, type_, charge) -> None:
        type_, charge = L._i32(type_), L._f32(charge)
        self.natoms = int(type_.shape[0])
        self._check(self._lib.fepb200_nb_set_atoms(self._h, self.natoms, L._pi(type_), L._pf(charge)))

    def mask_perturbed(self, atoms) -> None:
        atoms = L._i32(atoms)
        self._check(self._lib.fepb200_nb_mask_perturbed(self._h, int(atoms.shape[0]), L._pi(atoms)))

    def get_atoms(self):
        t, q = np.empty(self.natoms, np.int32), np.empty(self.natoms, np.float32)
        self._check(self._lib.fepb200_nb_get_atoms(self._h, L._pi(t), L._pf(q)))
        return t, q

    def set_pairlist(self, sci, cj, excl) -> None:
        sci, cj, excl = np.ascontiguousarray(sci), np.ascontiguousarray(cj), np.ascontiguousarray(excl)
        assert sci.dtype.itemsize == 16 and cj.dtype.itemsize == 32 and excl.dtype.itemsize == 128
        self._check(self._lib.fepb200_nb_set_pairlist(self._h, int(sci.shape[0]), sci.ctypes.data_as(_VP), int(cj.shape[0]),
                                                      cj.ctypes.data_as(_VP), int(excl.shape[0]), excl.ctypes.data_as(_VP)))

    def use_device_list(self, d_sci: int, d_cj: int, d_excl: int) -> None:
        self._check(self._lib.fepb200_nb_use_device_list(self._h, _VP(d_sci) if d_sci else None, _VP(d_cj) if d_cj else None,
                                                         _VP(d_excl) if d_excl else None))

    def setup(self, cs, params, mask: bool = True) -> None:
        """Everything of a `synth_nb.ClusterSystem`: constants, unmasked atoms, the mask, the list."""
        self.set_params(params)
        self.set_nbfp(cs.ntype, cs.nbfp)
        self.set_atoms(cs.type_unmasked, cs.q_unmasked)
        if mask:
            self.mask_perturbed(cs.perturbed_slots)
        self.set_pairlist(cs.sci, cs.cj, cs.excl)

    def compute(self, x, shiftvec, flags=DO_FORCE | DO_SHIFTFORCE | DO_POTENTIAL, out=None):
        """Host buffers; returns dict(f, fshift, vc, vvdw); `out` = arrays to accumulate into."""
        x, sv = L._f32(x).reshape(-1, 3), L._f32(shiftvec).reshape(NUM_SHIFT_VECTORS, 3)
        if out is None:
            out = dict(f=np.zeros((self.natoms, 3), np.float32), fshift=np.zeros((NUM_SHIFT_VECTORS, 3), np.float32))
            vc, vv = ctypes.c_double(0), ctypes.c_double(0)
        else:
            vc, vv = ctypes.c_double(out.get("vc", 0.0)), ctypes.c_double(out.get("vvdw", 0.0))
        self._check(self._lib.fepb200_nb_compute(self._h, L._pf(x), L._pf(sv), int(flags), L._pf(out["f"]),
                                                 L._pf(out["fshift"]), ctypes.byref(vc), ctypes.byref(vv)))
        out["vc"], out["vvdw"] = vc.value, vv.value
        return out

    def compute_xyzq(self, xq, shiftvec, flags=DO_FORCE | DO_SHIFTFORCE | DO_POTENTIAL):
        """Host buffers, coordinates as float4 {x, y, z, q} (nbat->x() of the reference, nbatXYZQ)."""
        xq, sv = L._f32(xq).reshape(-1, 4), L._f32(shiftvec).reshape(NUM_SHIFT_VECTORS, 3)
        out = dict(f=np.zeros((self.natoms, 3), np.float32), fshift=np.zeros((NUM_SHIFT_VECTORS, 3), np.float32))
        vc, vv = ctypes.c_double(0), ctypes.c_double(0)
        self._check(self._lib.fepb200_nb_compute_xyzq(self._h, L._pf(xq), L._pf(sv), int(flags), L._pf(out["f"]),
                                                      L._pf(out["fshift"]), ctypes.byref(vc), ctypes.byref(vv)))
        out["vc"], out["vvdw"] = vc.value, vv.value
        return out

    def launch_device(self, d_xq: int, shiftvec, flags, d_f: int, d_fshift: int = 0, d_energies: int = 0) -> None:
        sv = L._f32(shiftvec).reshape(NUM_SHIFT_VECTORS, 3)
        self._check(self._lib.fepb200_nb_launch_device(self._h, _VP(d_xq), L._pf(sv), int(flags), _VP(d_f),
                                                       _VP(d_fshift) if d_fshift else None,
                                                       _VP(d_energies) if d_energies else None))

    def launch_device_raw(self, d_xq: int, d_shiftvec: int, flags, d_f: int, d_fshift: int = 0, d_energies: int = 0) -> None:
        """As launch_device, with the shift vectors on the device (flags must carry NB_SHIFTVEC_ON_DEVICE)."""
        self._check(self._lib.fepb200_nb_launch_device(self._h, _VP(d_xq), ctypes.cast(_VP(d_shiftvec), _FP), int(flags), _VP(d_f),
                                                       _VP(d_fshift) if d_fshift else None,
                                                       _VP(d_energies) if d_energies else None))

    def launch_device_float_energies(self, d_xq: int, d_shiftvec: int, flags, d_f: int, d_fshift: int, d_elj: int, d_eelec: int) -> None:
        self._check(self._lib.fepb200_nb_launch_device_float_energies(
            self._h, _VP(d_xq), ctypes.cast(_VP(d_shiftvec), _FP), int(flags), _VP(d_f), _VP(d_fshift) if d_fshift else None,
            _VP(d_elj) if d_elj else None, _VP(d_eelec) if d_eelec else None))

    def export_energies_device(self, d_elj: int, d_eelec: int) -> None:
        self._check(self._lib.fepb200_nb_export_energies_device(self._h, _VP(d_elj), _VP(d_eelec)))

    def wait(self) -> None:
        self._check(self._lib.fepb200_nb_wait(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.fepb200_nb_launch_count(self._h))

    @property
    def cluster_pairs(self) -> int:
        return int(self._lib.fepb200_nb_cluster_pairs(self._h))

    def last_kernel_ms(self) -> float:
        ms = ctypes.c_float(0)
        self._check(self._lib.fepb200_nb_last_kernel_ms(self._h, ctypes.byref(ms)))
        return float(ms.value)


__all__ = ["NbContext", "SYMBOLS", "NB_Q_FROM_XQ", "DO_FORCE", "DO_POTENTIAL", "DO_SHIFTFORCE", "CLEAR_OUTPUTS"]
