"""Self-consistency tests of the oracle's building blocks (CPU only)."""
import os

import numpy as np
import pytest
from scipy.special import eval_legendre

import oracle
from emagls_b200 import synth


def test_getsh_addition_theorem(grids):
    rng = np.random.default_rng(1)
    a = np.stack([rng.uniform(-np.pi, np.pi, 40), np.arccos(rng.uniform(-1, 1, 40))], 1)
    b = np.stack([grids["hrirGridAziRad"][::7], grids["hrirGridZenRad"][::7]], 1)
    cg = np.clip(synth.unit_vectors(a[:, 0], a[:, 1]) @ synth.unit_vectors(b[:, 0], b[:, 1]).T, -1, 1)
    for basis in ("real", "complex"):
        Ya, Yb = oracle.getSH(19, a, basis), oracle.getSH(19, b, basis)
        for n in range(20):
            lhs = Ya[:, n * n:(n + 1) ** 2] @ np.conj(Yb[:, n * n:(n + 1) ** 2]).T
            rhs = (2 * n + 1) / (4 * np.pi) * eval_legendre(n, cg)
            assert np.abs(lhs - rhs).max() < 5e-13


def test_modal_coeffs_wronskian_identity():
    # rigid sphere: b_n = 4 pi i^n * (-i / x^2) / h_n^(2)'(x)
    x = np.linspace(0.03, 18.5, 300)
    b = oracle.sphModalCoeffs(19, x, "rigid")
    for n in range(20):
        ref = 4 * np.pi * (1j ** n) * (-1j / x ** 2) / oracle.dsph_hankel2(n, x)
        assert np.abs(b[:, n] - ref).max() / np.abs(ref).max() < 1e-11
    b0 = oracle.sphModalCoeffs(3, np.array([0.0]), "rigid")
    assert b0[0, 0] == 4 * np.pi and np.all(b0[0, 1:] == 0)


def test_sh_rep_to_order():
    out = oracle.sh_repToOrder(np.arange(4.0))
    assert out.tolist() == [0] + [1] * 3 + [2] * 5 + [3] * 7


def test_subsample_delay_integer_is_roll():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((64, 3))
    assert np.allclose(oracle.applySubsampleDelay(x, 5), np.roll(x, 5, axis=0), atol=1e-13)
    y = oracle.applySubsampleDelay(np.stack([x, x], 2), np.array([1.0, 2.0]).reshape(1, 1, 2))
    assert np.allclose(y[:, :, 1], np.roll(x, 2, axis=0), atol=1e-13)


def test_grpdelay_of_pure_delay():
    b = np.zeros(32)
    b[7] = 1.0
    gd = oracle.grpdelay(b, np.linspace(0, 24000, 17), 48000)
    assert np.allclose(gd, 7.0)


def test_fftfilt_and_binaural_decode_semantics():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((200, 3))
    wl, wr = rng.standard_normal((16, 3)), rng.standard_normal((16, 3))
    out = oracle.binauralDecode(x, 48000, wl, wr, 48000)
    ref = sum(np.convolve(x[:, c], wl[:, c])[:200] for c in range(3))
    assert out.shape == (200, 2) and np.allclose(out[:, 0], ref, atol=1e-12)
    outc = oracle.binauralDecode(x, 48000, wl, wr, 48000, True)
    assert outc.shape == (200 - 8 + 1, 2) and np.allclose(outc[:, 0], ref[7:], atol=1e-12)


def test_sh_rotation_matrix_matches_rotated_sh():
    rng = np.random.default_rng(4)
    E = oracle.euler2rotationMatrix(0.3, -0.7, 1.1, "zyz")
    assert np.allclose(E @ E.T, np.eye(3), atol=1e-14)
    az = rng.uniform(-np.pi, np.pi, 30)
    ze = np.arccos(rng.uniform(-1, 1, 30))
    u = synth.unit_vectors(az, ze)
    for basis in ("real", "complex"):
        Rsh = oracle.getSHrotMtx(E, 4, basis)
        assert np.allclose(Rsh @ np.conj(Rsh).T, np.eye(25), atol=1e-12)
        Y = oracle.getSH(4, np.stack([az, ze], 1), basis)
        # one of the two rotation directions must reproduce the SHs of the rotated directions
        ok = False
        for M in (E, E.T):
            a2, z2 = synth.angles_from_vectors(u @ M.T)
            Y2 = oracle.getSH(4, np.stack([a2, z2], 1), basis)
            ok = ok or np.allclose(Y @ Rsh, Y2, atol=1e-11) or np.allclose(Y @ Rsh.T, Y2, atol=1e-11)
        assert ok


def test_ch_helpers():
    az = np.linspace(0, 2 * np.pi, 13, endpoint=False)
    Yr = oracle.getCH(6, az, "real")
    assert np.allclose(Yr.T @ Yr / 13, np.eye(13), atol=1e-12)
    J = oracle.getChToShExpansionMatrix(3, "real")
    dirs = np.stack([az, np.full_like(az, np.pi / 2)], 1)
    assert np.allclose(oracle.getCH(3, az, "real") @ J.T, oracle.getSH(3, dirs, "real"), atol=1e-12)
    Jc = oracle.getChToShExpansionMatrix(3, "complex")
    assert np.allclose(oracle.getCH(3, az, "complex") @ Jc.T, oracle.getSH(3, dirs, "complex"), atol=1e-12)


def test_freq_domain_conjugates_give_real_basis_signals():
    rng = np.random.default_rng(5)
    # a real-valued SH-domain signal expressed in complex SHs obeys the symmetry -> ifft gives
    # coefficients that map back to a real signal on the sphere
    dirs = np.stack([rng.uniform(-np.pi, np.pi, 12), np.arccos(rng.uniform(-1, 1, 12))], 1)
    Yc = oracle.getSH(2, dirs, "complex")
    Yr = oracle.getSH(2, dirs, "real")
    x = rng.standard_normal((16, 9))                      # real-basis time signals
    c = np.linalg.lstsq(np.conj(Yc), Yr @ x.T, rcond=None)[0].T  # complex-basis coefficients
    X = np.fft.fft(c, axis=0)
    Xe = oracle.getShFreqDomainConjugate(X[:9])
    assert np.allclose(Xe, X, atol=1e-10)
    Xch = rng.standard_normal((9, 5)) + 1j * rng.standard_normal((9, 5))
    Xch[0] = Xch[0].real
    Xch[-1] = Xch[-1].real
    assert oracle.getChFreqDomainConjugate(Xch).shape == (16, 5)


def _small_problem(D=600, seed=7):
    rng = np.random.default_rng(seed)
    az = rng.uniform(-np.pi, np.pi, D)
    ze = np.arccos(rng.uniform(-1, 1, D))
    hL, hR = synth.synth_hrirs(az, ze, taps=32, delay=8)
    return az, ze, hL, hR


def test_real_vs_complex_basis_emagls2_agree_to_noise_floor():
    az, ze, hL, hR = _small_problem()
    g = synth.load_grids()
    args = (hL, hR, az, ze, 0.042, g["micGridAziRad"], g["micGridZenRad"], 4, 48000, 64)
    a = oracle.getEMagLs2Filters(*args, "real")[0]
    b = oracle.getEMagLs2Filters(*args, "complex")[0]
    assert np.abs(b.imag).max() < 1e-12
    assert np.abs(a - b.real).max() / np.abs(a).max() < 1e-6


def test_ema_variants_run_and_are_finite():
    az, ze, hL, hR = _small_problem(D=300)
    maz = np.linspace(0, 2 * np.pi, 9, endpoint=False)
    for basis in ("real", "complex"):
        wL, wR = oracle.getEMagLsFiltersEMAinCH(hL, hR, az, ze, 0.05, maz, 3, 16000, 32, basis)
        assert wL.shape == (32, 7) and np.all(np.isfinite(wL)) and np.abs(wL).max() > 0
        wL, wR = oracle.getEMagLsFiltersEMAinSH(hL, hR, az, ze, 0.05, maz, 3, 16000, 32, basis)
        assert wL.shape == (32, 16) and np.all(np.isfinite(wL)) and np.abs(wL).max() > 0


def test_from_atf_runs_on_reference_fixture():
    import os
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "atf_full.npz"))
    atf = d["atfIrs"].astype(float)
    ag = np.deg2rad(d["atfGridAziEleDeg"].astype(float))
    ag = np.stack([ag[:, 0], np.pi / 2 - ag[:, 1]], 1)
    g = synth.load_grids()
    az, ze = g["hrirGridAziRad"][::3], g["hrirGridZenRad"][::3]
    hL, hR = synth.synth_hrirs(az, ze)
    wL, wR, info = oracle.getEMagLsFiltersFromAtf(hL, hR, np.stack([az, ze], 1), atf, ag, 48000, 256, 1000.0,
                                                  return_spectra=True)
    assert wL.shape == (256, 8) and wL[0].max() == 0 and np.all(np.isfinite(wR))
    assert info["meanGridDevDeg"] < 5.0


def test_high_precision_oracle_agrees_with_fp64_on_a_well_conditioned_problem():
    """oracle/hp_oracle.py (exact pwGrid + Gram in integers, mpmath eigendecomposition) against the LAPACK route
    of oracle.regularized_inverse where both are accurate, including clipped singular values."""
    from oracle import hp_oracle as hp
    rng = np.random.default_rng(0)
    S, D, M = 12, 40, 6
    sm = rng.standard_normal((M, S)) + 1j * rng.standard_normal((M, S))
    sm[-1] = sm[0] + 1e-3 * (rng.standard_normal(S) + 1j * rng.standard_normal(S))   # one singular value below the 1 % clip
    Y = rng.standard_normal((S, D))
    t = rng.standard_normal((2, D)) + 1j * rng.standard_normal((2, D))
    W, sv = hp.exact_ls_rows(sm, Y, t, dps=60)
    s64 = np.linalg.svd(sm @ Y, compute_uv=False)
    assert s64[-1] < 0.01 * s64[0]
    assert np.abs(sv - s64).max() <= 1e-12 * s64[0]
    assert hp.rel_err(hp.fp64_ls_rows(sm, Y, t), W) <= 1e-10


def test_high_precision_fixture_is_consistent():
    """tests/golden/hp_goldens.npz: the stored FP64-oracle error is the distance of the stored oracle rows from
    the stored exact rows, and it grows with the condition number as SURVEY.md section 0 reports."""
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hp_goldens.npz")
    d = np.load(p)
    ex, o64, err = d["c1_exact"], d["c1_oracle64"], d["c1_err_oracle64"]
    assert ex.shape == o64.shape == (15, 2, 32)
    for i in range(15):
        assert abs(np.abs(o64[i] - ex[i]).max() / np.abs(ex[i]).max() - err[i]) <= 1e-3 * err[i] + 1e-18
    assert err[0] > 1e-8 and err[-1] < 1e-11 and d["c1_cond"][0] > 1e12


def test_diffuseness_matrix_matches_the_target_covariance():
    """EXTENSION (oracle.diffuseness_matrix): A Rhat A^H = R, and A = I when the covariances already agree."""
    rng = np.random.default_rng(5)
    for _ in range(20):
        X = rng.standard_normal((2, 7)) + 1j * rng.standard_normal((2, 7))
        Y = rng.standard_normal((2, 7)) + 1j * rng.standard_normal((2, 7))
        R, Rh = X @ X.conj().T / 7, Y @ Y.conj().T / 7
        A = oracle.diffuseness_matrix(R, Rh)
        assert np.abs(A @ Rh @ A.conj().T - R).max() <= 1e-12 * np.abs(R).max()
        assert np.abs(oracle.diffuseness_matrix(R, R) - np.eye(2)).max() <= 1e-12
