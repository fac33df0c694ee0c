"""CPU-only checks of the drop-in boundary: libphare_b200.so loads, exports every symbol declared in
include/phare_b200.h, the pure-host geometry helpers agree with the oracle, and the product fails loudly
(never falls back) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from phare_b200 import abi
from util import ALL_DIM_INTERP, small_layout

HEADER = os.path.join(abi.ROOT, "include", "phare_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(phb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = abi.load()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/phare_b200.h but not exported"
    # and the ctypes prototypes cover the header
    assert set(names) == set(abi.EXPORTED)


@pytest.mark.parametrize("dim,interp", ALL_DIM_INTERP)
def test_field_shapes_match_oracle(cpu_oracle, dim, interp):
    lib = abi.load()
    L = small_layout(dim, interp)
    for qty in range(14):
        s = (C.c_uint32 * 3)()
        n = lib.phb_field_shape(C.byref(L), qty, s)
        want = cpu_oracle.field_shape(L, qty)
        assert tuple(s[d] for d in range(dim)) == want and n == int(np.prod(want))
    assert lib.phb_field_ghosts(interp) == (2 if interp == 1 else 4)
    assert lib.phb_particle_ghosts(interp) == (1 if interp == 1 else 2)


def test_aos_stride_matches_reference_particle_layout():
    lib = abi.load()
    assert [lib.phb_aos_stride(d) for d in (1, 2, 3)] == [56, 64, 80]  # sizeof(Particle<d>), SURVEY §2 row 4


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = abi.load()
    h = C.c_void_p()
    rc = lib.phb_create(0, 3, 1, C.byref(h))
    assert rc == abi.PHB_ERR_NO_DEVICE and not h.value
    assert b"no CPU fallback" in lib.phb_last_error(None)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(abi.ROOT, "phare_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(root, f)).read()
                # no import of the oracle package, no dlopen / link of its libraries
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "libphare_ref" not in src and "phare_oracle.h" not in src, f
    for root, _, files in os.walk(os.path.join(abi.ROOT, "include")):
        for f in files:
            assert "import oracle" not in open(os.path.join(root, f)).read()


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference tree (dev container only)")
def test_bridge_header_against_reference_types():
    """include/phare_b200/bridge.hpp (the b200:: helpers of INTEGRATION.md) instantiated against the UNMODIFIED reference
    GridLayout / Field / VecField headers: the member names it relies on exist (syntax check, nothing is linked or run)"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["g++", "-std=c++20", "-fsyntax-only", "-w", "-I" + os.path.join(root, "oracle/ref"),
                        "-I/root/reference/src", "-I/root/reference", os.path.join(root, "oracle/ref/ref_bridge_check.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
