"""Parity of the remaining filter designers against the oracle (B200 only), through the C ABI.

Same acceptance protocol as tests/test_gpu_design.py.  SH-/CH-domain variants are far better
conditioned than eMagLS2 (floor ~2e-13), so their whole-filter bound is tighter.
"""
import numpy as np
import pytest

import oracle
from emagls_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def em():
    import emagls_b200
    return emagls_b200


@pytest.fixture(scope="module")
def h(em):
    return em.Handle(0)


@pytest.fixture(scope="module")
def c1(grids):
    az, ze = grids["hrirGridAziRad"], grids["hrirGridZenRad"]
    hL, hR = synth.synth_hrirs(az, ze)
    return dict(az=az, ze=ze, hL=hL, hR=hR, r=grids["micRadius"], maz=grids["micGridAziRad"],
                mze=grids["micGridZenRad"], fs=grids["fs"])


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def bin_err(W, Wo):
    return np.abs(W - Wo).max(1) / np.abs(Wo).max(1)


# ------------------------------------------------------------------ getEMagLsFilters (SH domain)
@pytest.mark.parametrize("basis", ["real", "complex"])
def test_emagls_sh_domain_matches_oracle(em, h, c1, basis):
    args = (c1["az"], c1["ze"], c1["r"], c1["maz"], c1["mze"], 4, c1["fs"], 512)
    wL, wR, sp = em.getEMagLsFilters(c1["hL"], c1["hR"], *args, basis, handle=h, return_spectra=True)
    oL, oR, osp = oracle.getEMagLsFilters(c1["hL"], c1["hR"], *args, shDefinition=basis, return_spectra=True)
    assert wL.shape == (512, 25)
    assert wL.dtype == (np.complex128 if basis == "complex" else np.float64)
    for e, Wo in enumerate((osp["W_l"], osp["W_r"])):
        err = bin_err(sp[:, :, e], Wo)
        assert err[16:].max() <= 1e-10, err[16:].max()
        assert err[1:16].max() <= 1e-8, err[1:16].max()
        assert np.abs(sp[0, :, e] - Wo[0]).max() <= 1e-8 * np.abs(Wo[1]).max()   # DC fix in the output basis
    assert rel(wL, oL) < 1e-9 and rel(wR, oR) < 1e-9
    assert np.all(wL[0] == 0) and np.all(wL[-1] == 0)


def test_emagls_sh_domain_rotation_batch(em, h, c1):
    Rm = np.stack([np.eye(3), synth.rotation_yaw_pitch(75.0, -25.0)])
    args = (c1["az"], c1["ze"], c1["r"], c1["maz"], c1["mze"], 4, c1["fs"], 512)
    wL, wR = em.getEMagLsFilters(c1["hL"], c1["hR"], *args, rotations=Rm, handle=h)
    raz, rze = synth.rotate_grid(c1["az"], c1["ze"], Rm[1])
    oL, oR = oracle.getEMagLsFilters(c1["hL"], c1["hR"], raz, rze, *args[2:])
    assert wL.shape == (512, 25, 2)
    assert rel(wL[:, :, 1], oL) < 1e-9 and rel(wR[:, :, 1], oR) < 1e-9


# ------------------------------------------------------------------ engine routes
def test_gram_and_tsqr_routes_agree(em, h, c1, monkeypatch):
    """The Gram/Cholesky route (bins where no singular value can be clipped) and the TSQR/Jacobi
    route compute the same operator: forcing everything through TSQR changes nothing above the floor."""
    Rm = np.stack([np.eye(3), synth.rotation_yaw_pitch(10.0, 40.0)])
    args = (c1["az"], c1["ze"], c1["r"], c1["maz"], c1["mze"], 4, c1["fs"], 512)
    wL, wR, sp = em.getEMagLs2Filters(c1["hL"], c1["hR"], *args, rotations=Rm, handle=h, return_spectra=True)
    monkeypatch.setenv("EMAGLS_NO_GRAM", "1")
    xL, xR, xsp = em.getEMagLs2Filters(c1["hL"], c1["hR"], *args, rotations=Rm, handle=h, return_spectra=True)
    for o in range(2):
        for e in range(2):
            err = bin_err(sp[:, :, o, e], xsp[:, :, o, e])
            assert err[16:].max() < 1e-10, err[16:].max()
    assert rel(wL, xL) < 5e-9 and rel(wR, xR) < 5e-9


def test_orientation_chunking_is_transparent(em, h, c1, monkeypatch):
    R = synth.orientation_grid()[::700][:5]
    args = (c1["az"], c1["ze"], c1["r"], c1["maz"], c1["mze"], 4, c1["fs"], 512)
    hL2, hR2 = synth.synth_hrirs(c1["az"], c1["ze"], head_radius=0.09, ear_azi_deg=88.0, seed=5)
    HL, HR = np.stack([c1["hL"], hL2], 2), np.stack([c1["hR"], hR2], 2)
    wL, wR = em.getEMagLs2Filters(HL, HR, *args, rotations=R, handle=h)
    monkeypatch.setenv("EMAGLS_ORIENT_CHUNK", "2")
    xL, xR = em.getEMagLs2Filters(HL, HR, *args, rotations=R, handle=h)
    assert wL.shape == (512, 32, 10)
    assert rel(xL, wL) < 1e-9 and rel(xR, wR) < 1e-9
