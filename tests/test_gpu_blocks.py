"""Parity of the CUDA building blocks against the oracle, through the C ABI (B200 only)."""
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


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize("N", [0, 1, 4, 19, 37])
@pytest.mark.parametrize("basis", ["real", "complex"])
def test_getsh(em, h, grids, N, basis):
    dirs = np.stack([grids["hrirGridAziRad"], grids["hrirGridZenRad_raw"]], 1)  # includes zen > pi
    dirs = np.vstack([dirs, [[0.0, 0.0], [1.0, np.pi], [-2.0, np.pi / 2]]])       # poles, equator
    assert rel(em.getSH(N, dirs, basis, handle=h), oracle.getSH(N, dirs, basis)) < 1e-13   # FP64 tolerance


@pytest.mark.parametrize("N,r,fs,K", [(19, 0.042, 48000, 513), (37, 0.042, 96000, 4097), (36, 0.08, 48000, 513)])
def test_sph_modal_coeffs_rigid(em, h, N, r, fs, K):
    kr = 2 * np.pi * np.linspace(0, fs / 2, K) / 343.0 * r
    b, bo = em.sphModalCoeffs(N, kr, "rigid", handle=h), oracle.sphModalCoeffs(N, kr, "rigid")
    assert b[0, 0] == 4 * np.pi and np.all(b[0, 1:] == 0)           # kr == 0 override
    with np.errstate(divide="ignore", invalid="ignore"):
        e = np.abs(b - bo) / np.abs(bo)
    assert np.nanmax(e) < 2e-13                                     # elementwise relative, FP64


def test_sph_modal_coeffs_open(em, h):
    kr = np.linspace(0, 18.5, 513)
    b, bo = em.sphModalCoeffs(19, kr, "open", handle=h), oracle.sphModalCoeffs(19, kr, "open")
    assert np.max(np.abs(b - bo).max(0) / np.abs(bo).max(0)) < 1e-13


def test_smair_matrix_raw(em, h, grids):
    params = dict(returnRawMicSigs=True, fs=48000, irLen=1024, oversamplingFactor=1, radialFilter="none",
                  smaRadius=0.042, smaDesignAziZenRad=np.stack([grids["micGridAziRad"], grids["micGridZenRad"]], 1))
    sm, p = em.getSMAIRMatrix(params, handle=h)
    smo, _ = oracle.getSMAIRMatrix(params)
    assert sm.shape == smo.shape == (32, 400, 513)
    assert np.all(sm[:, :, -1].imag == 0)                            # Nyquist: real(Bn)
    e = np.abs(sm - smo).max(axis=(0, 1)) / np.abs(smo).max(axis=(0, 1))
    assert e.max() < 2e-13
    assert p["order"] == 4 and p["arrayType"] == "rigid"             # defaults echoed


@pytest.mark.parametrize("raw,basis", [(False, "real"), (False, "complex"), (True, "complex")])
def test_smair_matrix_sh_domain_and_complex_basis(em, h, grids, raw, basis):
    params = dict(returnRawMicSigs=raw, order=4, fs=48000, irLen=512, oversamplingFactor=1, radialFilter="none",
                  smaRadius=0.042, shDefinition=basis,
                  smaDesignAziZenRad=np.stack([grids["micGridAziRad"], grids["micGridZenRad"]], 1))
    sm, _ = em.getSMAIRMatrix(params, handle=h)
    smo, _ = oracle.getSMAIRMatrix(params)
    assert sm.shape == smo.shape == ((32 if raw else 25), 400, 257)
    e = np.abs(sm - smo).max(axis=(0, 1)) / np.abs(smo).max(axis=(0, 1))
    assert e.max() < 1e-12


@pytest.mark.parametrize("Mc,D,grade,tol", [(32, 2702, 0, 1e-13), (8, 407, 0, 1e-13), (13, 2702, 3, 1e-11),
                                              (25, 2702, 5, 1e-9), (64, 3000, 4, 1e-10), (64, 1444, 0, 1e-13),
                                              (5, 5, 0, 1e-12), (1, 40, 0, 1e-13)])
def test_regularized_apply(em, h, Mc, D, grade, tol):
    rng = np.random.default_rng(Mc * 1000 + D)
    A = rng.standard_normal((Mc, D)) + 1j * rng.standard_normal((Mc, D))
    if grade:
        U, s, Vh = np.linalg.svd(A, full_matrices=False)
        A = (U * (s * np.logspace(0, -grade, Mc))) @ Vh
    t = rng.standard_normal((3, D)) + 1j * rng.standard_normal((3, D))
    # tolerance scales with eps * cond: both LAPACK and the TSQR/Jacobi route are backward stable
    assert rel(em.regularizedApply(A, t, 0.01, handle=h), t @ oracle.regularized_inverse(A, 0.01)) < tol
    assert rel(em.regularizedApply(A, t, 0.0, handle=h), t @ oracle.regularized_inverse(A, 0.0)) < 10 * tol


def test_regularized_apply_rank_deficient(em, h):
    rng = np.random.default_rng(5)
    A = rng.standard_normal((6, 50)) + 1j * rng.standard_normal((6, 50))
    A[3] = A[1] + A[2]            # exactly rank deficient: sigma_min is rounding noise, gets clipped
    t = rng.standard_normal((2, 50)) + 0j
    W = em.regularizedApply(A, t, 0.01, handle=h)
    assert np.all(np.isfinite(W))
    Wo = t @ oracle.regularized_inverse(A, 0.01)
    # the clipped direction carries gain 1/(c*smax); compare the well-determined part only
    assert rel(W @ A, Wo @ A) < 1e-10


def test_group_delay_matches_oracle(em):
    """a14: median(grpdelay(sum(h, 2), 1, f, fs)) (lib/getEMagLs2Filters.m:72-75) on the device against the oracle."""
    g = synth.load_grids()
    az, ze = g["hrirGridAziRad"][::7], g["hrirGridZenRad"][::7]
    for taps, delay, fs in ((128, 30, 48000.0), (96, 17, 44100.0)):
        hL, hR = synth.synth_hrirs(az, ze, taps=taps, delay=delay)
        K = taps + 1
        f = np.linspace(0.0, fs / 2.0, K)
        for h in (hL, hR):
            gd, med = em.grpdelay(h, f, fs)
            ref = oracle.grpdelay(h.sum(1), f, fs)
            assert np.abs(gd - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
            assert abs(med - np.median(ref)) <= 1e-9 * max(1.0, abs(np.median(ref)))
    # pure delay: the group delay is the delay
    x = np.zeros(64); x[11] = 1.0
    gd, med = em.grpdelay(x, np.linspace(0, 24000.0, 33), 48000.0)
    assert np.abs(gd - 11.0).max() < 1e-9 and abs(med - 11.0) < 1e-9
