"""CPU check of the FP64-free phase continuation + digit split of the forward tensor-core product
(emagls_b200/csrc/phase_fixed.cuh; reference semantics: lib/getEMagLs2Filters.m:95-103, t = |H_k| y / |y|).

The header is plain C++ when compiled for the host (the MUFU.RSQ seed becomes 1 / sqrtf, optionally perturbed by the
worst-case relative error of the hardware approximation).  tests/csrc/phase_fixed_check.cpp compares
Z = rn(t 2^24 256^(T-4)) with a quad-precision evaluation over integer-valued inputs of widely varying size (zeros,
equal magnitudes, one dominant component, the largest admissible |H|) and re-assembles the balanced base-256 digits.
Bound asserted: |Z - exact| <= 0.6 units of the last digit (an exactly rounded Z has 0.5) and every digit in range.
"""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "csrc", "phase_fixed_check.cpp")


def _run(tmp_path, perturb: str, n: int):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / ("pfx_" + re.sub(r"[^0-9a-z]", "_", perturb)))
    subprocess.run([gxx, "-O2", "-std=c++17", f"-DEM_PFX_SEED_PERTURB={perturb}", "-o", exe, SRC, "-lquadmath"],
                   check=True, cwd=ROOT)
    out = subprocess.run([exe, str(n)], check=True, capture_output=True, text=True).stdout
    res = {}
    for m in re.finditer(r"T=(\d) worst ([0-9.eE+-]+) bad (\d+)", out):
        res[int(m.group(1))] = (float(m.group(2)), int(m.group(3)))
    assert set(res) == {4, 6}, out
    return res


# 2 ulp of MUFU.RSQ plus the truncation of the 24-bit operands: |eps0| <= 2^-21.5 = 3.4e-7; 4e-7 covers it
@pytest.mark.parametrize("perturb", ["0.0f", "4.0e-7f", "-4.0e-7f"])
def test_fixed_point_phase_is_within_0p6_units(tmp_path, perturb):
    res = _run(tmp_path, perturb, 300000)
    for T, (worst, bad) in res.items():
        assert bad == 0, f"T={T}: {bad} digit failures"
        assert worst <= 0.6, f"T={T}: |Z - exact| = {worst} units of the last digit"
