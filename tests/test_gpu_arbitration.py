"""High-precision arbitration on the ill-conditioned bins (SURVEY.md 7.1-1, 8-c protocol item 2; B200 only).

On bins 1..15 of em32 (sigma_min / sigma_max from 1e-13 to 1e-7) two correct FP64 implementations of
lib/getEMagLs2Filters.m:86-92 differ by up to 1e-7, so CUDA-vs-oracle parity decides nothing there.  The
fixture tests/golden/hp_goldens.npz (tests/golden/make_hp_goldens.py, oracle/hp_oracle.py) holds the rows
W(k,:) the reference's formula has in EXACT arithmetic on the oracle's FP64 inputs, and the FP64 oracle's own
distance to them.  A 1-ulp perturbation of those inputs moves the exact rows by 4e-16 (measured when the
fixture was made), so the CUDA path's own evaluation of the inputs (device Bessel / SH code) does not blur
the comparison.  Required: err_cuda <= 2 err_oracle64 on every such bin; both are printed.
"""
import os

import numpy as np
import pytest

from emagls_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hp():
    return np.load(os.path.join(ROOT, "tests", "golden", "hp_goldens.npz"))


@pytest.fixture(scope="module")
def em():
    import emagls_b200
    return emagls_b200


def _check(tag, hp, sp, capsys):
    bins, exact, err64 = hp[f"{tag}_bins"], hp[f"{tag}_exact"], hp[f"{tag}_err_oracle64"]
    rows = []
    for i, k in enumerate(bins):
        W = np.stack([sp[k, :, 0], sp[k, :, 1]])
        e = float(np.abs(W - exact[i]).max() / np.abs(exact[i]).max())
        rows.append((int(k), float(hp[f"{tag}_cond"][i]), e, float(err64[i])))
    with capsys.disabled():
        print(f"\n{tag}: bin  cond(pwGrid)  err_cuda   err_oracle64 (both against the exact rows)")
        for k, c, e, o in rows:
            print(f"{tag}: {k:3d}  {c:9.2e}  {e:9.2e}  {o:9.2e}")
    for k, c, e, o in rows:
        assert e <= 2.0 * o + 1e-13, (tag, k, e, o)     # 1e-13: both are at rounding level on the last bins
    return rows


def test_c1_ill_conditioned_bins_against_exact_rows(em, hp, grids, capsys):
    az, ze = grids["hrirGridAziRad"], grids["hrirGridZenRad"]
    hL, hR = synth.synth_hrirs(az, ze)
    h = em.Handle(0)
    _, _, sp = em.getEMagLs2Filters(hL, hR, az, ze, grids["micRadius"], grids["micGridAziRad"], grids["micGridZenRad"], 4,
                                    grids["fs"], 512, handle=h, return_spectra=True)
    rows = _check("c1", hp, sp, capsys)
    # the factored TSQR + Jacobi route keeps the relative accuracy of the graded steering matrix: it is far
    # closer to the exact rows than the LAPACK route wherever that one is limited by conditioning
    assert max(e for _, _, e, _ in rows) <= 1e-9


def test_c1_batched_orientation_matches_exact_rows(em, hp, grids, capsys):
    """The identity orientation inside a batch (the warm-started Jacobi path along the bins of a group) meets the
    same bound as the single call."""
    az, ze = grids["hrirGridAziRad"], grids["hrirGridZenRad"]
    hL, hR = synth.synth_hrirs(az, ze)
    h = em.Handle(0)
    R = np.stack([synth.rotation_yaw_pitch(20.0, -15.0), np.eye(3), synth.rotation_yaw_pitch(200.0, 35.0)])
    _, _, sp = em.getEMagLs2Filters(hL, hR, az, ze, grids["micRadius"], grids["micGridAziRad"], grids["micGridZenRad"], 4,
                                    grids["fs"], 512, rotations=R, handle=h, return_spectra=True)
    _check("c1", hp, sp[:, :, 1, :], capsys)


def test_c5_shape_ill_conditioned_bins_against_exact_rows(em, hp, grids, capsys):
    if "c5_bins" not in hp.files:
        pytest.skip("c5 rows not in the fixture")
    az, ze = grids["hrirGridAziRad"], grids["hrirGridZenRad"]
    hL, hR = synth.synth_hrirs(az, ze, fs=96000.0, taps=128, delay=40)
    maz, mze = synth.fibonacci_sphere(64)
    h = em.Handle(0)
    _, _, sp = em.getEMagLs2Filters(hL, hR, az, ze, 0.042, maz, mze, 7, 96000.0, 128, handle=h, return_spectra=True)
    _check("c5", hp, sp, capsys)
