"""Parity of the SURVEY.md section 8(f) rows (callers either side of the hot path) against the oracle,
through the C ABI (B200 only): radial filters, SH/CH encoding and rotation of the recording, the
remaining lib/ designers.  All FP64; tolerances are written at each assert."""
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


def radial_params(grids, kind, order=4, irLen=512, osf=1, **kw):
    p = dict(order=order, fs=grids["fs"], irLen=irLen, oversamplingFactor=osf, radialFilter=kind,
             smaRadius=grids["micRadius"], waveModel="planeWave", arrayType="rigid", nfft=osf * irLen,
             smaDesignAziZenRad=np.stack([grids["micGridAziRad"], grids["micGridZenRad"]], 1))
    p.update(kw)
    return p


# ------------------------------------------------------------------ getRadialFilter
@pytest.mark.parametrize("kind,kw", [("tikhonov", {}), ("tikhonov", dict(regulConst=1e-4)),
                                     ("softlimit", dict(noiseGainDb=20)), ("full", {}), ("none", {})])
@pytest.mark.parametrize("order,irLen,osf", [(4, 512, 1), (7, 256, 2), (0, 64, 2)])
def test_radial_filter_matches_oracle(em, h, grids, kind, kw, order, irLen, osf):
    p = radial_params(grids, kind, order, irLen, osf, **kw)
    r = em.getRadialFilter(p, handle=h)
    ro = oracle.getRadialFilter(p)
    assert r.shape == ro.shape == (osf * irLen // 2 + 1, order + 1)
    fin = np.isfinite(ro)
    assert np.array_equal(np.isnan(r.real) | np.isnan(r.imag), np.isnan(ro.real) | np.isnan(ro.imag)) or kind == "full"
    # column-wise: orders differ by many decades at low kr
    for n in range(order + 1):
        m = fin[:, n] & np.isfinite(r[:, n])
        assert m.sum() >= ro.shape[0] - 1
        assert np.abs(r[m, n] - ro[m, n]).max() <= 1e-11 * np.abs(ro[m, n]).max(), (kind, n)
    if kind != "none":
        assert np.all(r[-1].imag == 0) and np.all(r[-1].real >= 0)   # abs() at Nyquist, getRadialFilter.m:68-70


def test_radial_filter_errors(em, h, grids):
    with pytest.raises(ValueError, match="Unkown radialFilter"):
        em.getRadialFilter(radial_params(grids, "regul"), handle=h)             # getRadialFilter.m:65
    with pytest.raises(NotImplementedError, match="pointSource"):
        em.getRadialFilter(radial_params(grids, "tikhonov", waveModel="pointSource"), handle=h)


# ------------------------------------------------------------------ applyRadialFilter
@pytest.mark.parametrize("n,kind,order,nfft", [(6000, "tikhonov", 4, 512), (300, "tikhonov", 2, 512),
                                                (40000, "softlimit", 3, 256), (5000, "tikhonov", 7, 1024)])
def test_apply_radial_filter_matches_oracle(em, h, grids, n, kind, order, nfft):
    p = radial_params(grids, kind, order, nfft, 1, noiseGainDb=15)
    x = np.random.default_rng(n).standard_normal((n, (order + 1) ** 2))
    y = em.applyRadialFilter(x, p, handle=h)
    yo = oracle.applyRadialFilter(x, p)
    assert y.shape == yo.shape == (max(n, nfft) - nfft // 2, (order + 1) ** 2)
    assert rel(y, yo) < 1e-11


def test_apply_radial_filter_chunked_workspace(em, h, grids, monkeypatch):
    """The per-channel FIR processes (channel, block) pairs in chunks when the workspace budget is small:
    same result with a 1 MB budget (many chunks, ragged last chunk) as with the default single chunk."""
    p = radial_params(grids, "tikhonov", 2, 256, 1)
    x = np.random.default_rng(11).standard_normal((50001, 9))    # 9 channels x 28 blocks = 252 = 7 x 32 + 28
    y1 = em.applyRadialFilter(x, p, handle=h)
    monkeypatch.setenv("EMAGLS_FIR_WS_MB", "1")
    y2 = em.applyRadialFilter(x, p, handle=h)
    assert np.array_equal(y1, y2)
    assert rel(y2, oracle.applyRadialFilter(x, p)) < 1e-11


def test_apply_radial_filter_linearity_full_size(em, h, grids):
    """Size-independent property at the length of the shipped recording: linear, and an impulse
    returns the (delay-compensated) radial-filter IR of its order."""
    p = radial_params(grids, "tikhonov", 4, 512, 1)
    n = 360290
    rng = np.random.default_rng(5)
    a, b = rng.standard_normal((n, 25)), rng.standard_normal((n, 25))
    ya, yb = em.applyRadialFilter(a, p, handle=h), em.applyRadialFilter(b, p, handle=h)
    yab = em.applyRadialFilter(2.0 * a - 3.0 * b, p, handle=h)
    assert rel(yab, 2.0 * ya - 3.0 * yb) < 1e-12
    imp = np.zeros((2000, 25))
    imp[300, :] = 1.0
    yi = em.applyRadialFilter(imp, p, handle=h)
    ir = oracle.frontend_oracle.radialFilterIr(p)
    for c in (0, 3, 8, 24):
        n_ord = int(np.sqrt(c))
        assert np.abs(yi[300 - 256:300 + 256, c] - ir[:, n_ord]).max() <= 1e-12 * np.abs(ir[:, n_ord]).max()


# ------------------------------------------------------------------ getSMAIRMatrix with a radial filter
@pytest.mark.parametrize("kind", ["tikhonov", "softlimit"])
def test_smair_matrix_with_radial_filter(em, h, grids, kind):
    p = radial_params(grids, kind, 4, 128, 2, noiseGainDb=20, returnRawMicSigs=False)
    m, _ = em.getSMAIRMatrix(p, handle=h)
    mo, _ = oracle.getSMAIRMatrix(p)
    assert m.shape == mo.shape
    fin = np.isfinite(mo)
    assert np.array_equal(fin, np.isfinite(m))
    # per bin and output order (the filtered rows span many decades)
    for k in range(0, mo.shape[2], 7):
        for n in range(5):
            sl = (slice(n * n, (n + 1) ** 2), slice(None), k)
            if fin[sl].all() and np.abs(mo[sl]).max() > 0:
                assert np.abs(m[sl] - mo[sl]).max() <= 1e-10 * np.abs(mo[sl]).max(), (k, n)


# ------------------------------------------------------------------ SH / CH encoding, rotation
@pytest.mark.parametrize("basis", ["real", "complex"])
def test_sh_encode_matches_oracle(em, h, grids, basis):
    x = np.random.default_rng(1).standard_normal((70001, 32))
    s = em.encodeSH(x, grids["micGridAziRad"], grids["micGridZenRad"], 4, basis, handle=h)
    so = oracle.encodeSH(x, grids["micGridAziRad"], grids["micGridZenRad"], 4, basis)
    assert s.shape == so.shape == (70001, 25) and s.dtype == so.dtype
    assert rel(s, so) < 1e-12


@pytest.mark.parametrize("basis", ["real", "complex"])
def test_ch_encode_matches_oracle(em, h, basis):
    azi = 2 * np.pi * np.arange(13) / 13
    x = np.random.default_rng(2).standard_normal((9000, 13))
    s = em.encodeCH(x, azi, 6, basis, handle=h)
    so = oracle.encodeCH(x, azi, 6, basis)
    assert s.shape == so.shape == (9000, 13)
    assert rel(s, so) < 1e-12


def test_sh_encode_inverts_a_band_limited_field(em, h, grids):
    """Encoding the order-4 field sampled at the 32 microphones returns its SH coefficients."""
    c = np.random.default_rng(3).standard_normal((500, 25))
    Y = oracle.getSH(4, np.stack([grids["micGridAziRad"], grids["micGridZenRad"]], 1), "real")
    s = em.encodeSH(c @ Y.T, grids["micGridAziRad"], grids["micGridZenRad"], 4, handle=h)
    assert rel(s, c) < 1e-12


@pytest.mark.parametrize("order,ypr", [(4, (0.7, 0.0, 0.0)), (3, (-1.1, 0.4, 0.2)), (7, (2.5, -0.3, 1.0)), (0, (1.0, 0, 0))])
def test_rotate_sh_matches_oracle(em, h, order, ypr):
    x = np.random.default_rng(4).standard_normal((33333, (order + 1) ** 2))
    y = em.rotateSH(x, *ypr, handle=h)
    yo = oracle.rotateSH(x, *ypr)
    assert rel(y, yo) < 1e-12
    # rotations preserve the energy of every order
    for n in range(order + 1):
        sl = slice(n * n, (n + 1) ** 2)
        assert abs(np.sum(y[:, sl] ** 2) / np.sum(x[:, sl] ** 2) - 1) < 1e-12


def test_rotate_sh_moves_a_plane_wave(em, h):
    Y = oracle.getSH(5, np.array([[0.3, 1.1]]), "real")
    Yr = em.rotateSH(np.repeat(Y, 4, 0), 0.5, handle=h)
    assert np.abs(Yr[0] - oracle.getSH(5, np.array([[0.8, 1.1]]), "real")[0]).max() < 1e-13


def test_binaural_decode_with_horizontal_rotation(em, h):
    """binauralDecode(..., horRotAngleRad) (dependencies/binauralDecode.m:26-30): rotation of the SH-domain input
    followed by the decode, against the oracle's composition of the same two reference steps."""
    rng = np.random.default_rng(8)
    x = rng.standard_normal((9000, 25))
    wL, wR = rng.standard_normal((512, 25)), rng.standard_normal((512, 25))
    y = em.binauralDecode(x, 48000, wL, wR, 48000, True, None, None, 0.6, handle=h)
    yo = oracle.binauralDecode(x, 48000, wL, wR, 48000, True, None, None, 0.6)
    assert y.shape == yo.shape and rel(y, yo) < 1e-9
    y0 = em.binauralDecode(x, 48000, wL, wR, 48000, True, handle=h)
    assert rel(y, y0) > 1e-2                       # the rotation does something
    assert np.array_equal(em.binauralDecode(x, 48000, wL, wR, 48000, True, None, None, 0.0, handle=h), y0)


def test_full_ls_render_chain(em, h, grids):
    """verifyEMagLs.m:235-257 end to end on the device: encode -> radial filter -> binauralDecode."""
    az, ze = grids["hrirGridAziRad"][::3], grids["hrirGridZenRad"][::3]
    hL, hR = synth.synth_hrirs(az, ze)
    wL, wR = em.getMagLsFilters(hL, hR, az, ze, 4, grids["fs"], 512, handle=h)
    oL, oR = oracle.getMagLsFilters(hL, hR, az, ze, 4, grids["fs"], 512)
    x = np.random.default_rng(6).standard_normal((20000, 32))
    p = radial_params(grids, "tikhonov", 4, 512, 1)
    sh = em.applyRadialFilter(em.encodeSH(x, grids["micGridAziRad"], grids["micGridZenRad"], 4, handle=h), p, handle=h)
    y = em.binauralDecode(sh, grids["fs"], wL, wR, grids["fs"], handle=h)
    sho = oracle.applyRadialFilter(oracle.encodeSH(x, grids["micGridAziRad"], grids["micGridZenRad"], 4), p)
    yo = oracle.binauralDecode(sho, grids["fs"], oL, oR, grids["fs"])
    assert y.shape == yo.shape
    assert rel(y, yo) < 1e-9   # north_star: rendered binaural signals within 1e-9 relative


# ------------------------------------------------------------------ remaining lib/ designers
@pytest.fixture(scope="module")
def hor():
    az = np.linspace(0, 2 * np.pi, 180, endpoint=False)
    hL, hR = synth.synth_hrirs(az, np.full(az.size, np.pi / 2))
    return az, hL, hR


@pytest.mark.parametrize("basis", ["real", "complex"])
@pytest.mark.parametrize("order,length", [(4, 512), (6, 256)])
def test_magls_2d_matches_oracle(em, h, hor, basis, order, length):
    az, hL, hR = hor
    wL, wR, sp = em.getMagLsFilters2D(hL, hR, az, order, 48000, length, basis, handle=h, return_spectra=True)
    oL, oR, osp = oracle.getMagLsFilters2D(hL, hR, az, order, 48000, length, basis, return_spectra=True)
    assert wL.shape == (length, 2 * order + 1) and wL.dtype == oL.dtype
    for e in range(2):
        err = np.abs(sp[:, :, e] - osp["W"][:, :, e]).max(1) / np.abs(osp["W"][:, :, e]).max(1)
        assert err.max() <= 1e-10
    assert rel(wL, oL) < 1e-11 and rel(wR, oR) < 1e-11


@pytest.mark.parametrize("r,order,fs,length", [(0.042, 4, 48000, 512), (0.042, 1, 48000, 128), (0.08, 6, 48000, 1024),
                                               (0.042, 7, 96000, 512)])
def test_spherical_head_filter_matches_oracle(em, h, r, order, fs, length):
    w, W = em.getMagLsSphericalHeadFilter(r, order, fs, length, handle=h)
    wo, Wo = oracle.getMagLsSphericalHeadFilter(r, order, fs, length)
    assert w.shape == wo.shape and W.shape == Wo.shape
    assert rel(W, Wo.real) < 1e-12
    assert rel(w, wo) < 1e-11
    assert w[0] == 0 and w[-1] == 0


@pytest.mark.parametrize("basis", ["real", "complex"])
def test_array_diffuse_filter_matches_oracle(em, h, grids, basis):
    w = em.getMagLsArrayDiffuseFilter(grids["micRadius"], grids["micGridAziRad"], grids["micGridZenRad"], 4,
                                      grids["fs"], 512, basis, handle=h)
    wo = oracle.getMagLsArrayDiffuseFilter(grids["micRadius"], grids["micGridAziRad"], grids["micGridZenRad"], 4,
                                           grids["fs"], 512, basis)
    assert rel(w, wo) < 1e-10


def test_frontend_errors(em, h, grids):
    with pytest.raises(em.EmaglsError):
        em.getMagLsSphericalHeadFilter(0.042, 30, 48000, 512, handle=h)     # order above the simulation order
    with pytest.raises(em.EmaglsError):
        em.encodeSH(np.zeros((10, 8)), np.zeros(8), np.ones(8), 4, handle=h)  # fewer mics than harmonics
    with pytest.raises(ValueError):
        em.applyRadialFilter(np.zeros((10, 24)), radial_params(grids, "tikhonov"), handle=h)
