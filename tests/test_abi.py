"""C-ABI surface: the header, the ctypes prototype table and both libraries agree."""
import ctypes as C
import os
import re
import subprocess

import pytest

import harness as H
from hpsim_b200 import _capi as K

HEADER = os.path.join(H.ROOT, "include", "hpsim_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(hps_[a-z0-9_]+)\s*\(", src))


def test_header_matches_prototype_table():
    assert header_functions() == set(K.PROTOTYPES)


def test_product_library_exports_every_symbol():
    assert os.path.exists(K.LIB_PATH), "build the product first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = K.load()  # binds every prototype; raises AttributeError on a missing export
    assert lib.hps_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", K.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (hps_[a-z0-9_]+)", out))
    assert header_functions() <= exported


def test_product_has_no_cpu_fallback():
    """Without a usable sm_100 device the product must fail loudly, never compute on the CPU."""
    import shutil
    if shutil.which("nvidia-smi") and subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0:
        pytest.skip("a GPU is present")
    lib = K.load()
    assert not lib.hps_ctx_create(0)
    assert b"CUDA" in lib.hps_last_error(None) or b"device" in lib.hps_last_error(None)
    from hpsim_b200 import HPSim
    HPSim._use_library(None)
    with pytest.raises(K.HpsError):
        HPSim.Beam(mass=939.294, charge=-1.0, current=0.015, num=16)


def test_oracle_implements_the_same_abi(oracle):
    assert oracle.hps_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", H.ORACLE_SO], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (hps_[a-z0-9_]+)", out))
    assert header_functions() <= exported


def test_elem_struct_layout():
    from hpsim_b200 import lattice as L
    assert C.sizeof(K.Elem) == L.ELEM_DTYPE.itemsize == 248
    assert K.Elem.p.offset == L.ELEM_DTYPE.fields["p"][1] == 40
