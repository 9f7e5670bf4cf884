"""CPU checks of the oracle's restatement of the SURVEY.md section 8(f) rows (no golden exists for
them: "parity unpinned"; these tests pin internal consistency and the documented conventions)."""
import numpy as np

import oracle
from oracle import frontend_oracle as fo


def _params(grids, kind="tikhonov", order=4, nfft=512, **kw):
    p = dict(order=order, fs=48000, irLen=nfft, oversamplingFactor=1, radialFilter=kind, smaRadius=0.042,
             waveModel="planeWave", arrayType="rigid", nfft=nfft,
             smaDesignAziZenRad=np.stack([grids["micGridAziRad"], grids["micGridZenRad"]], 1))
    p.update(kw)
    return p


def test_radial_filter_kinds(grids):
    p = _params(grids)
    bn = oracle.sphModalCoeffs(4, 2 * np.pi * np.linspace(0, 24000, 257) / 343 * 0.042)
    r = oracle.getRadialFilter(p)
    ref = np.conj(bn) / (np.abs(bn) ** 2 + 1e-2)
    assert np.allclose(r[:-1], ref[:-1], rtol=1e-14, atol=0)
    assert np.allclose(r[-1], np.abs(ref[-1]), rtol=1e-14)              # getRadialFilter.m:68-70
    full = oracle.getRadialFilter(_params(grids, "full"))
    assert np.allclose(full[1:-1] * bn[1:-1], 1, rtol=1e-12)
    soft = oracle.getRadialFilter(_params(grids, "softlimit", noiseGainDb=20))
    assert np.nanmax(np.abs(soft)) <= 10.0 * (1 + 1e-12)                # limited to 20 dB
    assert np.all(oracle.getRadialFilter(_params(grids, "none")) == 1)
    # getRadialFilter.m defaults: tikhonov, oversamplingFactor 2, irLen 256
    d = oracle.getRadialFilter(dict(order=2, fs=48000, smaRadius=0.042))
    assert d.shape == (257, 3)


def test_apply_radial_filter_is_the_documented_convolution(grids):
    p = _params(grids, order=2, nfft=128)
    x = np.random.default_rng(0).standard_normal((700, 9))
    y = oracle.applyRadialFilter(x, p)
    ir = fo.radialFilterIr(p)
    assert ir.shape == (128, 3) and ir[0].max() == 0 == ir[-1].max()    # 5 % fades start/end at zero
    for c in (0, 2, 8):
        full = np.convolve(x[:, c], ir[:, int(np.sqrt(c))])[:700]
        assert np.allclose(y[:, c], full[64:], rtol=0, atol=1e-12 * np.abs(full).max())
    # short signals are zero-padded to nfft (applyRadialFilter.m:24-27)
    assert oracle.applyRadialFilter(x[:50], p).shape == (64, 9)


def test_smair_radial_branch_doubles_the_nyquist_factor(grids):
    p = _params(grids, order=1, nfft=64, oversamplingFactor=1, irLen=64, returnRawMicSigs=False)
    m, _ = oracle.getSMAIRMatrix(p)
    m0, _ = oracle.getSMAIRMatrix(dict(p, radialFilter="none"))
    rad = oracle.sh_repToOrder(oracle.getRadialFilter(p).T)[:4]
    assert np.allclose(m[:, :, 5], rad[:, None, 5] * m0[:, :, 5])
    assert np.allclose(m[:, :, -1], rad[:, None, -1].real ** 2 * m0[:, :, -1])   # getSMAIRMatrix.m:134-137


def test_encode_and_rotate(grids):
    mics = np.stack([grids["micGridAziRad"], grids["micGridZenRad"]], 1)
    c = np.random.default_rng(1).standard_normal((50, 25))
    for basis in ("real", "complex"):
        Y = oracle.getSH(4, mics, basis)
        s = oracle.encodeSH(c @ Y.T, mics[:, 0], mics[:, 1], 4, basis)
        assert np.allclose(s, c, atol=1e-12)
    azi = 2 * np.pi * np.arange(13) / 13
    cc = np.random.default_rng(2).standard_normal((20, 13))
    assert np.allclose(oracle.encodeCH(cc @ oracle.getCH(6, azi).T, azi, 6), cc, atol=1e-12)
    # a yaw of +psi moves a source from azimuth a to a + psi; composition of yaws adds
    Y = oracle.getSH(4, np.array([[0.3, 1.1]]))
    assert np.allclose(oracle.rotateSH(Y, 0.5), oracle.getSH(4, np.array([[0.8, 1.1]])), atol=1e-14)
    x = np.random.default_rng(3).standard_normal((10, 25))
    assert np.allclose(oracle.rotateSH(oracle.rotateSH(x, 0.4), 0.9), oracle.rotateSH(x, 1.3), atol=1e-13)


def test_magls_2d_ls_bins_and_window(grids):
    from emagls_b200 import synth
    az = np.linspace(0, 2 * np.pi, 90, endpoint=False)
    hL, hR = synth.synth_hrirs(az, np.full(90, np.pi / 2), taps=64, delay=20)
    wL, wR, sp = oracle.getMagLsFilters2D(hL, hR, az, 3, 48000, 128, return_spectra=True)
    assert wL.shape == (128, 7) and np.all(wL[0] == 0) and np.all(wL[-1] == 0)
    # below k_cut the solution is the plain LS fit of the delay-compensated HRTFs
    Yp = np.linalg.pinv(oracle.getCH(3, az).T)
    h = np.zeros((256, 90))
    h[:64] = hL
    H = np.fft.fft(oracle.applySubsampleDelay(h, -sp["grpD"][0]), axis=0)
    k = sp["k_cut"] - 2
    assert np.allclose(sp["W"][k, :, 0], H[k] @ Yp, rtol=1e-10, atol=1e-12)
    wLc, _ = oracle.getMagLsFilters2D(hL, hR, az, 3, 48000, 128, "complex")
    assert np.iscomplexobj(wLc)
    # the complex-CH filters are the unitary image of the real ones: same response towards every direction
    assert np.allclose(wLc @ np.conj(oracle.getCH(3, az, "complex")).T, wL @ oracle.getCH(3, az).T, atol=1e-10)


def test_diffuse_field_filters(grids):
    w, W = oracle.getMagLsSphericalHeadFilter(0.042, 4, 48000, 512)
    assert w.shape == (512,) and W.shape == (1024,) and w[0] == 0 == w[-1]
    assert abs(W[0] - 1) < 1e-12 and np.all(W[:513] <= 1 + 1e-12)       # lo_df / hi_df: order 4 holds less energy
    assert np.argmax(np.abs(w)) == 256                                   # linear-phase-like: peak at len/2
    wa = oracle.getMagLsArrayDiffuseFilter(0.042, grids["micGridAziRad"], grids["micGridZenRad"], 4, 48000, 512)
    wc = oracle.getMagLsArrayDiffuseFilter(0.042, grids["micGridAziRad"], grids["micGridZenRad"], 4, 48000, 512,
                                           "complex")
    # sum_s b_n(s) (Y_hi^H Y_lo)[s][c] runs over whole orders, which is not invariant under the change of
    # basis: the reference's result depends on shDefinition (measured 4 % here)
    assert 1e-3 < np.abs(wa - wc).max() / np.abs(wa).max() < 0.2
    assert np.argmax(np.abs(wa)) == 256
