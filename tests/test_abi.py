"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/fepb200.h declares, and fails loudly (no fallback) when there is no GPU.  No compute
calls are made here."""
import ctypes
import os
import re

import pytest

from fepb200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "fepb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fepb200_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(L.SYMBOLS)


def test_constants_of_header_and_python_mirror_agree():
    """Every integer macro of include/fepb200.h (flags, enum values, error codes) that fepb200.params mirrors has
    the same value there; the flags and extension bits must all be mirrored."""
    from fepb200 import params as P

    txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "fepb200.h")).read(), flags=re.S)
    macros = {m.group(1): eval(m.group(2), {"__builtins__": {}})  # noqa: S307 -- "(1 << 4)", "(-3)", "45" of our own header
              for m in re.finditer(r"^#define\s+FEPB200_([A-Z0-9_]+)\s+(\(?-?\s*[0-9][0-9 <()]*\)?)\s*$", txt, flags=re.M)}
    assert len(macros) >= 25
    mirrored = [k for k in macros if hasattr(P, k)]
    for k in mirrored:
        assert getattr(P, k) == macros[k], k
    for k in ("DO_FORCE", "DO_SHIFTFORCE", "DO_FOREIGNLAMBDA", "DO_POTENTIAL", "DO_SR", "CLEAR_OUTPUTS", "ATOMIC_OUTPUTS",
              "EEL_PME", "VDW_PME", "SC_GAPSYS", "LAMBDA_COUL", "LAMBDA_VDW"):
        assert k in mirrored, k


def _declared_nb_symbols():
    txt = open(os.path.join(ROOT, "include", "fepb200_nb.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fepb200_nb_[a-z0-9_]+)\s*\(", txt)))


def test_nb_header_binding_and_library_agree():
    """include/fepb200_nb.h (the non-perturbed cluster-pair kernel): every declared entry point is bound in
    fepb200.nb and exported by the library; the list structures have the reference's sizes."""
    from fepb200 import nb as NB
    from fepb200 import synth_nb

    assert _declared_nb_symbols() == sorted(NB.SYMBOLS)
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in _declared_nb_symbols():
        assert hasattr(lib, name), name
    assert (synth_nb.SCI_DTYPE.itemsize, synth_nb.CJ_DTYPE.itemsize, synth_nb.EXCL_DTYPE.itemsize) == (16, 32, 128)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), name


def test_params_struct_matches_header():
    txt = open(os.path.join(ROOT, "include", "fepb200.h")).read()
    body = txt[txt.index("typedef struct fepb200_params") : txt.index("} fepb200_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(?:int|float)\s+([A-Za-z0-9_]+)\s*;", body)
    from fepb200.params import CParams

    assert fields == [n for n, _ in CParams._fields_]


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L.load_library()
    with pytest.raises(L.FepError) as ei:
        L.FepContext(0)
    assert ei.value.code == -3  # FEPB200_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gromacs-fep-gpu_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "fep_oracle" not in txt and "libfepref" not in txt, f
