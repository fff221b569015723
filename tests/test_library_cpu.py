"""CPU tests of the drop-in boundary: libtrp.so loads, exports every symbol of include/tr_prover.h, fails loudly
without a GPU (no CPU fallback), and the host-side mirror validates arguments like halo2's assert_eq! panics."""
import ctypes
import os
import re

import numpy as np
import pytest

import __graft_entry__ as ge

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    ge.build()
    return ge.load_package()


def test_header_symbols_are_exported(pkg):
    header = open(os.path.join(ROOT, "include", "tr_prover.h")).read()
    declared = set(re.findall(r"\b(trp_[a-z0-9_]+)\s*\(", header))
    from tiny_ram_halo2_b200._lib import EXPORTED_SYMBOLS
    assert declared == set(EXPORTED_SYMBOLS), declared ^ set(EXPORTED_SYMBOLS)
    lib = pkg.load_library()
    for s in declared:
        assert hasattr(lib, s), s
    assert b"sm_100a" in lib.trp_version()


def test_library_is_sm100a_only(pkg):
    out = os.popen(f"cuobjdump -lelf {pkg.lib_path()} 2>/dev/null").read()
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.TrpError) as ei:
        pkg.Context(0, pkg.VESTA)
    assert ei.value.code == -4
    h = ctypes.c_void_p()
    assert pkg.load_library().trp_ctx_create(ctypes.byref(h), 0, 1) == -4 and not h.value


def test_product_package_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "tiny-ram-halo2_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "liboracle" not in src and "pasta_model" not in src.replace(
                    "oracle/pasta_model.py", ""), f


def test_synthetic_seed_points_are_on_curve(pkg):
    from tiny_ram_halo2_b200 import synthetic
    from util import O, pm
    for cid, C in ((O.PALLAS, pm.Pallas), (O.VESTA, pm.Vesta)):
        for pt in synthetic.SEEDS[cid]:
            x, y = O.limbs_to_ints(O.from_mont(O.BASE_FIELD[cid], pt.reshape(2, 4)))
            assert C.on_curve((x, y))
    s = synthetic.random_scalars(1000, 1)
    assert s.shape == (1000, 4) and int(s[:, 3].max()) < (1 << 62)
