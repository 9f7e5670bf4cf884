"""Diffuse-field covariance constraint (opt-in EXTENSION; SURVEY.md 8-f rank 4, B200 only).

The reference removed this step before the surveyed commit (CHANGELOG.md:10-18), so there is no reference code and
no usable golden (resources/*_wDC.mat were made with an HRIR set that is not available offline): "parity unpinned".
What is checked: the CUDA path against the oracle's restatement of the published formulation
(oracle.diffuseness_matrix), and the defining property itself -- the covariance of the rendered plane-wave responses
equals the covariance of the HRTF set in every bin -- against responses formed by the oracle's steering model.
"""
import math

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
def case(grids):
    az, ze = grids["hrirGridAziRad"][::3], grids["hrirGridZenRad"][::3]        # 901 directions
    hL, hR = synth.synth_hrirs(az, ze, taps=64, delay=20)
    return dict(az=az, ze=ze, hL=hL, hR=hR,
                args=(az, ze, grids["micRadius"], grids["micGridAziRad"], grids["micGridZenRad"], 4, grids["fs"], 128))


def _bins(W, Wo, lo=6):
    err = np.abs(W - Wo).max(1) / np.abs(Wo).max(1)
    return float(err[lo:].max()), float(err[:lo].max())


def test_emagls2_with_constraint_matches_oracle_and_matches_covariance(em, case):
    h = em.Handle(0)
    wL, wR, sp = em.getEMagLs2Filters(case["hL"], case["hR"], *case["args"], handle=h, return_spectra=True,
                                      applyDiffusenessConst=True)
    oL, oR, osp = oracle.getEMagLs2Filters(case["hL"], case["hR"], *case["args"], return_spectra=True,
                                           applyDiffusenessConst=True)
    for e, key in enumerate(("W_l", "W_r")):
        hi, lo = _bins(sp[:, :, e], osp[key])
        assert hi <= 1e-9 and lo <= 2e-6, (key, hi, lo)
    assert np.abs(wL - oL).max() <= 5e-8 * np.abs(oL).max() and np.abs(wR - oR).max() <= 5e-8 * np.abs(oR).max()
    assert np.all(sp[0].imag == 0)                                    # DC bin := real(constrained bin 2)
    # the constraint changes the filters (otherwise the test would not see it)
    w0L, _ = em.getEMagLs2Filters(case["hL"], case["hR"], *case["args"], handle=h)
    assert np.abs(wL - w0L).max() > 1e-4 * np.abs(w0L).max()
    # defining property on a few bins, with the oracle's steering model: cov(W pw) == cov(H)
    az, ze, r, maz, mze, order, fs, length = case["args"]
    nfft = 2 * length
    params = dict(returnRawMicSigs=True, fs=fs, irLen=nfft, oversamplingFactor=1, simulateAliasing=True,
                  radialFilter="none", smaRadius=r, smaDesignAziZenRad=np.stack([maz, mze], 1),
                  waveModel="planeWave", arrayType="rigid", shDefinition="real", shFunction=oracle.getSH, C=343.0)
    smair, _ = oracle.getSMAIRMatrix(params)
    simN = int(round(math.sqrt(smair.shape[1]))) - 1
    Yc = np.conj(oracle.getSH(simN, np.stack([az, ze], 1), "real")).T
    D = az.size
    for k in (3, 20, 77, 127):                                        # 0-based bins (not the real Nyquist bin)
        pw = smair[:, :, k] @ Yc
        Hh = np.stack([sp[k, :, 0], sp[k, :, 1]]) @ pw
        H = np.stack([osp["HL"][k], osp["HR"][k]])
        R, Rh = H @ H.conj().T / D, Hh @ Hh.conj().T / D
        assert np.abs(Rh - R).max() <= 1e-8 * np.abs(R).max(), (k, np.abs(Rh - R).max() / np.abs(R).max())


def test_emagls_sh_domain_with_constraint_matches_oracle(em, case):
    h = em.Handle(0)
    wL, wR, sp = em.getEMagLsFilters(case["hL"], case["hR"], *case["args"], handle=h, return_spectra=True,
                                     applyDiffusenessConst=True)
    oL, oR, osp = oracle.getEMagLsFilters(case["hL"], case["hR"], *case["args"], return_spectra=True,
                                          applyDiffusenessConst=True)
    assert wL.shape == (128, 25)
    for e, key in enumerate(("W_l", "W_r")):
        hi, lo = _bins(sp[:, :, e], osp[key])
        assert hi <= 1e-9 and lo <= 2e-6, (key, hi, lo)
    assert np.abs(wL - oL).max() <= 5e-8 * np.abs(oL).max()


def test_constraint_is_off_by_default_and_rejects_the_complex_basis(em, case):
    h = em.Handle(0)
    a = em.getEMagLs2Filters(case["hL"], case["hR"], *case["args"], handle=h)
    b = em.getEMagLs2Filters(case["hL"], case["hR"], *case["args"], handle=h, applyDiffusenessConst=False)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    with pytest.raises(NotImplementedError):
        em.getEMagLsFilters(case["hL"], case["hR"], *case["args"], "complex", handle=h, applyDiffusenessConst=True)


def test_constraint_in_an_orientation_batch(em, case):
    """Batched call with the constraint: page b equals the oracle on the HRIR grid rotated by R_b (as for the plain
    designer, tests/test_gpu_design.py), two HRTF sets sharing the orientation-dependent Gram matrices."""
    h = em.Handle(0)
    az, ze, r, maz, mze, order, fs, length = case["args"]
    Rm = np.stack([np.eye(3), synth.rotation_yaw_pitch(50.0, -20.0)])
    hL2 = np.stack([case["hL"], 0.7 * case["hR"]], 2)
    hR2 = np.stack([case["hR"], 1.3 * case["hL"]], 2)
    wL, wR = em.getEMagLs2Filters(hL2, hR2, *case["args"], rotations=Rm, handle=h, applyDiffusenessConst=True)
    assert wL.shape == (length, 32, 4)                                  # [len, M, sets * B], set index slowest
    for s in range(2):
        for b in range(2):
            raz, rze = synth.rotate_grid(az, ze, Rm[b])
            oL, oR = oracle.getEMagLs2Filters(hL2[:, :, s], hR2[:, :, s], raz, rze, r, maz, mze, order, fs, length,
                                              applyDiffusenessConst=True)
            page = s * 2 + b
            assert np.abs(wL[:, :, page] - oL).max() <= 5e-8 * np.abs(oL).max(), (s, b)
            assert np.abs(wR[:, :, page] - oR).max() <= 5e-8 * np.abs(oR).max(), (s, b)
