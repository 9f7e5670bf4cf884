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


# ------------------------------------------------------------------ getMagLsFilters / getLsFilters
@pytest.mark.parametrize("basis", ["real", "complex"])
def test_magls_matches_oracle(em, h, c1, basis):
    wL, wR, sp = em.getMagLsFilters(c1["hL"], c1["hR"], c1["az"], c1["ze"], 4, c1["fs"], 512, basis, handle=h,
                                    return_spectra=True)
    oL, oR, osp = oracle.getMagLsFilters(c1["hL"], c1["hR"], c1["az"], c1["ze"], 4, c1["fs"], 512, basis,
                                         return_spectra=True)
    assert wL.shape == (512, 25) and wL.dtype == (np.complex128 if basis == "complex" else np.float64)
    for e in range(2):
        assert bin_err(sp[:, :, e], osp["W"][:, :, e]).max() <= 1e-10       # every bin, DC included (LS bin)
    assert rel(wL, oL) < 1e-11 and rel(wR, oR) < 1e-11
    assert np.all(wL[0] == 0) and np.all(wL[-1] == 0)


@pytest.mark.parametrize("basis", ["real", "complex"])
def test_magls_batched_over_hrtf_sets(em, h, c1, basis):
    """Batch extension: hL, hR [samples, dirs, sets] -> [len, (N+1)^2, sets]; the one pinv(Y) serves every set
    (lib/getMagLsFilters.m:48), each page equals the single call."""
    hL2, hR2 = synth.synth_hrirs(c1["az"], c1["ze"], head_radius=0.09, ear_azi_deg=88.0, seed=5)
    HL, HR = np.stack([c1["hL"], hL2, c1["hL"]], 2), np.stack([c1["hR"], hR2, c1["hR"]], 2)
    wL, wR, sp = em.getMagLsFilters(HL, HR, c1["az"], c1["ze"], 4, c1["fs"], 512, basis, handle=h, return_spectra=True)
    assert wL.shape == (512, 25, 3) and sp.shape == (513, 25, 3, 2)
    for s_, (hl, hr) in enumerate(((c1["hL"], c1["hR"]), (hL2, hR2), (c1["hL"], c1["hR"]))):
        oL, oR, osp = oracle.getMagLsFilters(hl, hr, c1["az"], c1["ze"], 4, c1["fs"], 512, basis, return_spectra=True)
        for e in range(2):
            assert bin_err(sp[:, :, s_, e], osp["W"][:, :, e]).max() <= 1e-10
        assert rel(wL[:, :, s_], oL) < 1e-11 and rel(wR[:, :, s_], oR) < 1e-11


@pytest.mark.parametrize("basis", ["real", "complex"])
def test_ls_matches_oracle(em, h, c1, basis):
    wL, wR = em.getLsFilters(c1["hL"], c1["hR"], c1["az"], c1["ze"], 4, basis, handle=h)
    oL, oR = oracle.getLsFilters(c1["hL"], c1["hR"], c1["az"], c1["ze"], 4, basis)
    assert wL.shape == (128, 25)
    assert rel(wL, oL) < 1e-12 and rel(wR, oR) < 1e-12


def test_magls_golden_pin_through_cabi(em, h, goldens):
    """Golden pin (2) through the CUDA path: the order-4 surrogate HRIRs rebuilt from the reference's LS
    golden reproduce the LS bins of the reference's MagLS golden (7e-6 .. 5e-3), and
    the LS golden itself is reproduced from the surrogate exactly."""
    G = goldens
    az, ze = G["hrirGridAziRad"], G["hrirGridZenRad"]
    Y4 = oracle.getSH(4, np.stack([az, ze], 1), "real")
    h4L, h4R = G["real_LS_wLsL"] @ Y4.T, G["real_LS_wLsR"] @ Y4.T
    lL, lR = em.getLsFilters(h4L, h4R, az, ze, 4, handle=h)
    assert rel(lL, G["real_LS_wLsL"]) < 1e-11 and rel(lR, G["real_LS_wLsR"]) < 1e-11
    for basis in ("real", "complex"):
        wL, wR = em.getMagLsFilters(h4L, h4R, az, ze, 4, 48000, 512, basis, handle=h)
        for w, key in ((wL, "wMlsL"), (wR, "wMlsR")):
            A = np.fft.fft(w, 1024, axis=0)
            Bq = np.fft.fft(G[f"{basis}_MagLS_woDC_{key}"], 1024, axis=0)
            for lo, hi, tol in ((1, 10, 3e-5), (10, 30, 1e-3), (30, 42, 5e-2)):
                assert np.abs(A[lo:hi] - Bq[lo:hi]).max() / np.abs(Bq[lo:hi]).max() < tol


# ------------------------------------------------------------------ getEMagLsFiltersFromAtf
@pytest.fixture(scope="module")
def atf():
    import os
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "atf_full.npz"))
    ag = np.deg2rad(d["atfGridAziEleDeg"].astype(float))
    return d["atfIrs"].astype(float), np.stack([ag[:, 0], np.pi / 2 - ag[:, 1]], 1)


@pytest.mark.parametrize("step", [3, 9])   # HRIR grid larger (901 > 407) and smaller (301 < 407) than the ATF grid
def test_from_atf_matches_oracle(em, h, grids, atf, step, capsys):
    atfIrs, ag = atf[0][:, :, ::4], atf[1][::4]          # every 4th direction of the measured set: 407
    az, ze = grids["hrirGridAziRad"][::step], grids["hrirGridZenRad"][::step]
    hL, hR = synth.synth_hrirs(az, ze)
    hg = np.stack([az, ze], 1)
    wL, wR, sp, info = em.getEMagLsFiltersFromAtf(hL, hR, hg, atfIrs, ag, 48000, 256, 1000.0, handle=h,
                                                  return_spectra=True, return_info=True)
    oL, oR, osp = oracle.getEMagLsFiltersFromAtf(hL, hR, hg, atfIrs, ag, 48000, 256, 1000.0, return_spectra=True)
    assert "average grid deviation" in capsys.readouterr().out            # the reference's disp() line (:96)
    assert abs(info["meanGridDevDeg"] - osp["meanGridDevDeg"]) < 1e-9
    assert wL.shape == (256, 8)
    for e, Wo in enumerate((osp["W_l"], osp["W_r"])):
        err = bin_err(sp[:, :, e], Wo)
        assert err[1:].max() <= 1e-10, err[1:].max()                     # measured ATFs are benign (cond < 200)
    assert rel(wL, oL) < 1e-10 and rel(wR, oR) < 1e-10
    assert np.all(wL[0] == 0) and np.all(sp[0].imag == 0)


def test_from_atf_config3_full_size(em, h, grids, atf):
    """BASELINE config 3 at its real size (testEMagLsFromAtfs.m:70-73): the whole glasses-on-HATS set (1625
    directions x 8 microphones x 192 taps), the 2702-direction HRIR grid, filterLen 256, cut-on 2 kHz."""
    atfIrs, ag = atf
    assert atfIrs.shape == (192, 8, 1625)
    az, ze = grids["hrirGridAziRad"], grids["hrirGridZenRad"]
    hL, hR = synth.synth_hrirs(az, ze)
    hg = np.stack([az, ze], 1)
    wL, wR, sp = em.getEMagLsFiltersFromAtf(hL, hR, hg, atfIrs, ag, 48000, 256, 2000.0, handle=h, return_spectra=True)
    oL, oR, osp = oracle.getEMagLsFiltersFromAtf(hL, hR, hg, atfIrs, ag, 48000, 256, 2000.0, return_spectra=True)
    assert wL.shape == (256, 8)
    for e, Wo in enumerate((osp["W_l"], osp["W_r"])):
        err = bin_err(sp[:, :, e], Wo)
        assert err[1:].max() <= 1e-10, err[1:].max()
    assert rel(wL, oL) < 1e-10 and rel(wR, oR) < 1e-10


@pytest.mark.parametrize("step", [1, 9])   # ATF grid smaller (operators shared by the batch) and HRIR grid smaller
def test_from_atf_batched_over_orientations(em, h, grids, atf, step):
    """Batch extension: page b equals one reference call with the HRIR grid rotated by R_b
    (lib/getEMagLsFiltersFromAtf.m:62-95: only the nearest-neighbour matching sees the grid)."""
    atfIrs, ag = (atf[0], atf[1]) if step == 1 else (atf[0][:, :, ::4], atf[1][::4])
    az, ze = grids["hrirGridAziRad"][::step], grids["hrirGridZenRad"][::step]
    hL, hR = synth.synth_hrirs(az, ze)
    R = np.stack([np.eye(3), synth.rotation_yaw_pitch(33.3, 15.7), synth.rotation_yaw_pitch(250.9, -35.2)])
    wL, wR, sp, info = em.getEMagLsFiltersFromAtf(hL, hR, np.stack([az, ze], 1), atfIrs, ag, 48000, 256, 2000.0,
                                                  rotations=R, handle=h, return_spectra=True, return_info=True)
    assert wL.shape == (256, 8, 3) and sp.shape == (257, 8, 3, 2)
    compared = 0
    for b in range(3):
        raz, rze = synth.rotate_grid(az, ze, R[b])
        # the reference's matching takes the FIRST minimum of the distances (lib/getEMagLsFiltersFromAtf.m:82): a
        # direction with two equidistant neighbours is decided by rounding, so such an orientation proves nothing
        ua, uh = synth.unit_vectors(ag[:, 0], ag[:, 1]), synth.unit_vectors(raz, rze)
        small, large = (uh, ua) if uh.shape[0] <= ua.shape[0] else (ua, uh)
        dist = np.sort(np.sqrt(((small[:, None, :] - large[None, :, :]) ** 2).sum(2)), axis=1)
        if (dist[:, 1] - dist[:, 0]).min() < 1e-9:
            continue
        compared += 1
        oL, oR, osp = oracle.getEMagLsFiltersFromAtf(hL, hR, np.stack([raz, rze], 1), atfIrs, ag, 48000, 256, 2000.0,
                                                     return_spectra=True)
        assert abs(info["meanGridDevDeg"][b] - osp["meanGridDevDeg"]) < 1e-6
        for e, Wo in enumerate((osp["W_l"], osp["W_r"])):
            err = bin_err(sp[:, :, b, e], Wo)
            assert err[1:].max() <= 1e-10, (b, err[1:].max())
        assert rel(wL[:, :, b], oL) < 1e-10 and rel(wR[:, :, b], oR) < 1e-10
    assert compared >= 2


# ------------------------------------------------------------------ EMA designers (config 4 shape at reduced radius)
@pytest.fixture(scope="module")
def ema(grids):
    az, ze = grids["hrirGridAziRad"], grids["hrirGridZenRad"]
    hL, hR = synth.synth_hrirs(az, ze)
    maz = 2 * np.pi * np.arange(13) / 13
    return dict(az=az, ze=ze, hL=hL, hR=hR, maz=maz)


def test_ema_ch_batched_over_sets_and_orientations(em, h, ema):
    """Batch extension of getEMagLsFiltersEMAinCH: page (set, b) equals one reference call with that HRTF set and
    the HRIR grid rotated by R_b (lib/getEMagLsFiltersEMAinCH.m:57-75: the steering model sees only the angles)."""
    hL2, hR2 = synth.synth_hrirs(ema["az"], ema["ze"], head_radius=0.09, ear_azi_deg=88.0, seed=5)
    HL, HR = np.stack([ema["hL"], hL2], 2), np.stack([ema["hR"], hR2], 2)
    R = np.stack([np.eye(3), synth.rotation_yaw_pitch(40.0, 0.0), synth.rotation_yaw_pitch(200.0, 20.0)])
    args = (0.042, ema["maz"], 4, 48000, 512)
    wL, wR = em.getEMagLsFiltersEMAinCH(HL, HR, ema["az"], ema["ze"], *args, rotations=R, handle=h)
    assert wL.shape == (512, 9, 6)
    for s_, (hl, hr) in enumerate(((ema["hL"], ema["hR"]), (hL2, hR2))):
        for b in (0, 2):
            raz, rze = synth.rotate_grid(ema["az"], ema["ze"], R[b])
            oL, oR = oracle.getEMagLsFiltersEMAinCH(hl, hr, raz, rze, *args)
            assert rel(wL[:, :, s_ * 3 + b], oL) < 1e-8 and rel(wR[:, :, s_ * 3 + b], oR) < 1e-8, (s_, b)


@pytest.mark.parametrize("basis", ["real", "complex"])
@pytest.mark.parametrize("radius,order", [(0.042, 4), (0.08, 6)])
def test_ema_ch_matches_oracle(em, h, ema, basis, radius, order):
    args = (ema["az"], ema["ze"], radius, ema["maz"], order, 48000, 512)
    wL, wR, sp = em.getEMagLsFiltersEMAinCH(ema["hL"], ema["hR"], *args, basis, handle=h, return_spectra=True)
    oL, oR, osp = oracle.getEMagLsFiltersEMAinCH(ema["hL"], ema["hR"], *args, basis, return_spectra=True)
    assert wL.shape == (512, 2 * order + 1)
    assert wL.dtype == (np.complex128 if basis == "complex" else np.float64)
    for e, Wo in enumerate((osp["W_l"], osp["W_r"])):
        err = bin_err(sp[:, :, e], Wo)
        assert err[16:].max() <= 1e-10, err[16:].max()
        assert err[8:16].max() <= 1e-8, err[8:16].max()
        assert err[1:8].max() <= 2e-6, err[1:8].max()        # ill-conditioned bins: the reference's own 1-ulp floor
    assert rel(wL, oL) < 5e-9 and rel(wR, oR) < 5e-9


@pytest.mark.parametrize("basis", ["real", "complex"])
def test_ema_sh_matches_oracle(em, h, ema, basis):
    args = (ema["az"], ema["ze"], 0.042, ema["maz"], 4, 48000, 512)
    wL, wR, sp = em.getEMagLsFiltersEMAinSH(ema["hL"], ema["hR"], *args, basis, handle=h, return_spectra=True)
    oL, oR, osp = oracle.getEMagLsFiltersEMAinSH(ema["hL"], ema["hR"], *args, basis, return_spectra=True)
    assert wL.shape == (512, 25)
    assert wL.dtype == (np.complex128 if basis == "complex" else np.float64)
    for e, Wo in enumerate((osp["W_l"], osp["W_r"])):
        err = bin_err(sp[:, :, e], Wo)
        assert err[16:].max() <= 1e-10, err[16:].max()
        assert err[8:16].max() <= 1e-8, err[8:16].max()
        assert err[1:8].max() <= 2e-6, err[1:8].max()        # ill-conditioned bins: the reference's own 1-ulp floor
    assert rel(wL, oL) < 5e-9 and rel(wR, oR) < 5e-9


def test_variant_errors(em, h, c1, ema):
    with pytest.raises(em.EmaglsError, match="len too short"):
        em.getMagLsFilters(c1["hL"], c1["hR"], c1["az"], c1["ze"], 4, c1["fs"], 64, handle=h)
    with pytest.raises(em.EmaglsError, match="fewer microphones"):
        em.getEMagLsFiltersEMAinCH(ema["hL"], ema["hR"], ema["az"], ema["ze"], 0.042, ema["maz"][:5], 4, 48000, 512,
                                   handle=h)


# ------------------------------------------------------------------ custom shFunction handles
def test_custom_sh_function_is_evaluated_on_the_host_and_passed_down(em, h, c1):
    """A shFunction other than the default is evaluated on the host and reaches the device as the two basis
    matrices (emagls_design_sma_basis).  With the oracle's getSH as the handle the result is the default one
    (device getSH differs by rounding only) and equals the oracle called with the same handle
    (lib/getEMagLs2Filters.m:32,52,61; example handle at verifyEMagLs.m:356-368)."""
    calls = []

    def my_sh(order, dirs, kind):
        calls.append((order, dirs.shape[0], kind))
        return oracle.getSH(order, dirs, kind)
    args = (c1["az"], c1["ze"], c1["r"], c1["maz"], c1["mze"], 4, c1["fs"], 512)
    wL, wR, sp = em.getEMagLs2Filters(c1["hL"], c1["hR"], *args, "real", my_sh, handle=h, return_spectra=True)
    assert calls == [(19, c1["az"].size, "real"), (19, 32, "real")]
    dL, dR, dsp = em.getEMagLs2Filters(c1["hL"], c1["hR"], *args, handle=h, return_spectra=True)
    for e in range(2):
        err = bin_err(sp[:, :, e], dsp[:, :, e])
        assert err[16:].max() < 1e-10, err[16:].max()
    assert rel(wL, dL) < 5e-9 and rel(wR, dR) < 5e-9
    # batched over orientations: page 1 = the reference call with the rotated grid, handle evaluated per orientation
    Rm = np.stack([np.eye(3), synth.rotation_yaw_pitch(75.0, -25.0)])
    bL, bR = em.getEMagLs2Filters(c1["hL"], c1["hR"], *args, "real", my_sh, rotations=Rm, handle=h)
    raz, rze = synth.rotate_grid(c1["az"], c1["ze"], Rm[1])
    oL, oR = oracle.getEMagLs2Filters(c1["hL"], c1["hR"], raz, rze, *args[2:], shFunction=my_sh)
    assert rel(bL[:, :, 0], wL) < 1e-12 and rel(bL[:, :, 1], oL) < 5e-9 and rel(bR[:, :, 1], oR) < 5e-9
    # SH-domain variant
    sL, sR = em.getEMagLsFilters(c1["hL"], c1["hR"], *args, "real", my_sh, handle=h)
    oL, oR = oracle.getEMagLsFilters(c1["hL"], c1["hR"], *args, shFunction=my_sh)
    assert sL.shape == (512, 25) and rel(sL, oL) < 1e-9 and rel(sR, oR) < 1e-9
    # a different (scaled) basis really changes the design input: the handle is not ignored
    xL, _ = em.getEMagLs2Filters(c1["hL"], c1["hR"], *args, "real", lambda n, d, k: 2.0 * oracle.getSH(n, d, k), handle=h)
    assert rel(xL, wL) > 1e-3
    with pytest.raises(NotImplementedError):
        em.getEMagLs2Filters(c1["hL"], c1["hR"], *args, "complex", my_sh, handle=h)
