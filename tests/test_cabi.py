"""The C-ABI shared library loads and exports every symbol include/emagls_cuda.h declares.
No compute calls here (no GPU in the CPU suite)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "emagls_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(emagls_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from emagls_b200 import build
    path = build.build()
    return ctypes.CDLL(path)


def test_header_declares_the_reference_entry_points():
    syms = _declared_symbols()
    for name in ("emagls_design_ls", "emagls_design_magls", "emagls_design_emagls", "emagls_design_emagls2",
                 "emagls_design_ema_ch", "emagls_design_ema_sh", "emagls_design_from_atf",
                 "emagls_smair_matrix", "emagls_binaural_decode"):
        assert name in syms


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_covers_every_declared_symbol():
    from emagls_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    _lib.load()


def test_config_defaults_match_reference_constants(lib):
    from emagls_b200._lib import Config
    cfg = Config()
    lib.emagls_config_default(ctypes.byref(cfg))
    # lib/getEMagLs2Filters.m:35-39, dependencies/getSMAIRMatrix.m:86
    assert cfg.nfft_max_len == 2048 and cfg.f_cut_min == 1e3 and cfg.svd_regul == 0.01
    assert cfg.speed_of_sound == 343.0 and cfg.array_type == 0 and cfg.basis == 0 and cfg.precision == 0


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import emagls_b200 as em
    with pytest.raises(em.EmaglsError):
        em.Handle(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "emagls_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f


def test_mex_gateways_match_the_c_abi():
    """The MEX drop-ins cannot be built without MATLAB; compile them against a stub mex.h so that every
    call into libemagls_cuda is at least type-checked against include/emagls_cuda.h."""
    import glob
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    srcs = sorted(glob.glob(os.path.join(root, "mex", "*.c")))
    names = {os.path.basename(s)[:-2] for s in srcs}
    assert {"getLsFilters", "getMagLsFilters", "getEMagLsFilters", "getEMagLs2Filters", "getEMagLsFiltersEMAinCH",
            "getEMagLsFiltersEMAinSH", "getEMagLsFiltersFromAtf", "getSMAIRMatrix", "binauralDecode"} <= names
    for s in srcs:
        r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-Wno-unused-function",
                            "-I" + os.path.join(root, "tests", "mex_stub"), "-I" + os.path.join(root, "include"), s],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
