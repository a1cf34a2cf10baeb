"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol include/sipp.h
declares, mirrors the header's struct layouts, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import sipp
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    sipp.build()
    return sipp.lib()


def test_exports_every_declared_symbol(L):
    hdr = open(os.path.join(ROOT, "include", "sipp.h")).read()
    declared = sorted(set(re.findall(r"\b(sipp_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    assert sorted(sipp.ABI_SYMBOLS) == declared
    for name in declared:
        assert hasattr(L, name), name
    assert L.sipp_abi_version() == 3


def test_struct_layouts_match_between_product_and_oracle_bindings():
    assert C.sizeof(sipp.MapDesc) == C.sizeof(orc.MapDesc) == 360
    assert C.sizeof(sipp.EvalParams) == C.sizeof(orc.EvalParams)
    for a, b in zip(sipp.MapDesc._fields_, orc.MapDesc._fields_):
        assert a[0] == b[0] and getattr(sipp.MapDesc, a[0]).offset == getattr(orc.MapDesc, b[0]).offset


def test_no_cpu_fallback(L):
    """Without a CUDA device sipp_create must fail loudly; with one this test is a no-op."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sipp.SippError) as e:
        sipp.Context(sipp.make_desc(64, 64, 4.0, 4.0, (0.5, 0.5), (3.5, 3.5), 30, 120, 30))
    assert e.value.code == sipp.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_create_argument_validation(L):
    h = C.c_void_p()
    assert L.sipp_create(None, 0, C.byref(h)) == sipp.ERR_INVALID_ARG
    d = sipp.make_desc(0, 64, 4.0, 4.0, (0.5, 0.5), (3.5, 3.5), 30, 120, 30)
    assert L.sipp_create(C.byref(d), 0, C.byref(h)) == sipp.ERR_INVALID_ARG
    d = sipp.make_desc(64, 64, 4.0, 9.0, (0.5, 0.5), (3.5, 3.5), 30, 120, 30)   # non-square cells (map_server_planexp.cpp:97)
    assert L.sipp_create(C.byref(d), 0, C.byref(h)) == sipp.ERR_INVALID_ARG
    d = sipp.make_desc(64, 64, 4.0, 4.0, (0.5, 0.5), (4.5, 3.5), 30, 120, 30)   # goal off the map (:104)
    assert L.sipp_create(C.byref(d), 0, C.byref(h)) == sipp.ERR_INVALID_ARG
    assert b"outside the map" in L.sipp_last_error(None)


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under scouting-ipp_b200/ may import, link or mention it."""
    pkg = os.path.join(ROOT, "scouting-ipp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".h", ".cuh", ".py", ".cpp", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), os.path.join(dirpath, f)


def test_synthetic_inputs_are_deterministic():
    from sipp import synth
    a = synth.make_labels(320, 240, 5)
    b = synth.make_labels(320, 240, 5)
    assert np.array_equal(a, b) and a[8, 8] == 30 and a[240 - 8, 320 - 8] == 30 and (a == 0).any()
    o = synth.make_observed(320, 240, 5, 10)
    assert o[:240, :320].all() or o.sum() > 0
    xy, cost = synth.make_candidates(320, 240, 5, 100)
    assert xy.shape == (100, 2) and (cost >= 0.2).all() and (cost <= 5.0).all()
