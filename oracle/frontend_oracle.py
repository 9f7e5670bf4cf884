"""FP64 NumPy restatement of the SURVEY.md section 8(f) "next" rows: the callers either side of the
hot path (radial filters, SH encoding / rotation of the recording, the remaining ``lib/`` designers).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Every function cites the reference file:line it
follows.  Parity status: no reference test or golden pins these functions ("parity unpinned"); the
pieces they are assembled from (``sphModalCoeffs``, ``getSH``, ``getCH``, ``applySubsampleDelay``,
``getFadeWindow``, ``fftfilt``) are the pinned ones of ``emagls_oracle``.
"""
from __future__ import annotations

import math

import numpy as np

from .emagls_oracle import (DEFAULTS, applySubsampleDelay, fftfilt, getCH, getChFreqDomainConjugate,
                            getFadeWindow, getSH, getSHrotMtx, euler2rotationMatrix, grpdelay, sh_repToOrder,
                            sphModalCoeffs)

__all__ = ["getRadialFilter", "applyRadialFilter", "getMagLsFilters2D", "getMagLsSphericalHeadFilter",
           "getMagLsArrayDiffuseFilter", "encodeSH", "encodeCH", "rotateSH"]


def getRadialFilter(params: dict):
    """dependencies/getRadialFilter.m:1-71 -> ``radFilts [nfft/2+1, order+1]`` (complex)."""
    p = dict(params)
    p.setdefault("radialFilter", "tikhonov")
    p.setdefault("waveModel", "planeWave")
    p.setdefault("oversamplingFactor", 2)
    p.setdefault("irLen", 256)
    p.setdefault("dirCoeff", 0)
    C = 343.0
    nfft = int(p["oversamplingFactor"] * p["irLen"])
    f = np.linspace(0, p["fs"] / 2, nfft // 2 + 1)
    N = int(p["order"])
    kind = str(p["radialFilter"]).lower()
    if kind == "none":
        return np.ones((nfft // 2 + 1, N + 1))
    if str(p["waveModel"]).lower() == "pointsource":
        raise NotImplementedError('WaveModel parameter "pointSource" not yet implemented.')  # :37-40
    bn = sphModalCoeffs(N, 2 * np.pi * f / C * p["smaRadius"], p.get("arrayType", "rigid"), p["dirCoeff"])
    bn = bn[:, :N + 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        if kind == "tikhonov":
            regul = p.get("regulConst", 1e-2)
            rad = np.conj(bn) / (np.conj(bn) * bn + regul)
        elif kind == "softlimit":
            g = 10.0 ** (p["noiseGainDb"] / 20.0)
            rad = 2 * g / np.pi * np.abs(bn) / bn * np.arctan(np.pi / (2 * g * np.abs(bn)))
        elif kind == "full":
            rad = 1.0 / bn
        else:
            raise ValueError(f'Unkown radialFilter parameter "{p["radialFilter"]}".')  # :65
    rad = np.asarray(rad, dtype=complex)
    if nfft % 2 == 0:
        rad[-1, :] = np.abs(rad[-1, :])  # :68-70
    return rad


def radialFilterIr(params: dict):
    """dependencies/applyRadialFilter.m:9-22: the causal, faded radial-filter IRs ``[nfft, order+1]``."""
    rad = getRadialFilter(params)
    rad = np.where(np.isnan(rad), 0, rad)
    nfft = int(params["nfft"])
    ir = np.fft.ifft(np.vstack([rad, np.conj(rad[-2:0:-1, :])]), axis=0)
    # MATLAB's ifft returns a real array for a conjugate-symmetric spectrum; the delay operator then
    # leaves imaginary parts at round-off level, which binauralDecode.m:59-63 discards at the end
    ir = applySubsampleDelay(ir.real, nfft / 2)
    return ir * getFadeWindow(nfft, 0.05)[:, None]


def applyRadialFilter(inSig, params: dict):
    """dependencies/applyRadialFilter.m:1-33."""
    nfft = int(params["nfft"])
    ir = radialFilterIr(params)
    x = np.asarray(inSig, dtype=float)
    if x.shape[0] < nfft:
        x = np.vstack([x, np.zeros((nfft - x.shape[0], x.shape[1]))])  # :24-27
    irs = sh_repToOrder(ir.T).T
    y = fftfilt(irs, x)
    return y[nfft // 2:, :]  # :31


def getMagLsFilters2D(hLHor, hRHor, horHrirGridAziRad, order, fs, length, chDefinition="real",
                      return_spectra=False, **kw):
    """lib/getMagLsFilters2D.m:1-98."""
    nfft_max = kw.get("NFFT_MAX_LEN", DEFAULTS["NFFT_MAX_LEN"])
    f_cut_min = kw.get("F_CUT_MIN_FREQ", DEFAULTS["F_CUT_MIN_FREQ"])
    hL = np.asarray(hLHor, dtype=float)
    hR = np.asarray(hRHor, dtype=float)
    assert length >= hL.shape[0], "HRIR len too short"
    nfft = min(nfft_max, 2 * length)
    f = np.linspace(0, fs / 2, nfft // 2 + 1)
    K = f.size
    k_cut = int(math.ceil(max(f_cut_min, 500 * order) / f[1]))
    Y_conj = np.conj(getCH(order, horHrirGridAziRad, chDefinition)).T  # [2N+1, D]
    Y_pinv = np.linalg.pinv(Y_conj)
    grpD = np.array([np.median(grpdelay(hL.sum(1), f, fs)), np.median(grpdelay(hR.sum(1), f, fs))])
    h = np.zeros((nfft, hL.shape[1], 2))
    h[:hL.shape[0], :, 0] = hL
    h[:hR.shape[0], :, 1] = hR
    h = applySubsampleDelay(h, -grpD.reshape(1, 1, 2))
    w_LS = np.einsum("tde,dh->the", h, Y_pinv)
    H = np.fft.fft(h, nfft, axis=0)
    W = np.fft.fft(w_LS, nfft, axis=0).astype(complex)
    for k in range(k_cut, K + 1):
        i = k - 1
        for e in range(2):
            phi = np.angle(W[i - 1, :, e] @ Y_conj)
            t = np.abs(H[i, :, e]) * np.exp(1j * phi)
            if k == K:
                t = t.real
            W[i, :, e] = t @ Y_pinv
    is_real = np.isrealobj(Y_conj)
    out = []
    for e in range(2):
        We = W[:K, :, e]
        Wf = np.vstack([We, np.conj(We[-2:0:-1, :])]) if is_real else getChFreqDomainConjugate(We)
        out.append(np.fft.ifft(Wf, axis=0))
    n_shift = nfft // 2
    wL = applySubsampleDelay(out[0], n_shift)
    wR = applySubsampleDelay(out[1], n_shift + (grpD[1] - grpD[0]))
    lo = n_shift - length // 2
    win = getFadeWindow(length)[:, None]
    wL = wL[lo:lo + length] * win
    wR = wR[lo:lo + length] * win
    if is_real:
        wL, wR = wL.real.copy(), wR.real.copy()
    if return_spectra:
        return wL, wR, dict(W=W[:K], grpD=grpD, k_cut=k_cut, nfft=nfft)
    return wL, wR


def _df_response(bn_rep):
    """``rms(abs(x), 2) * sqrt(size(x, 2)) / (4*pi)`` (lib/getMagLsSphericalHeadFilter.m:42-43)."""
    a = np.abs(bn_rep)
    return np.sqrt(np.mean(a * a, axis=1)) * math.sqrt(bn_rep.shape[1]) / (4 * np.pi)


def _zero_phase_tail(Wpos, nfft, length):
    """lib/getMagLsSphericalHeadFilter.m:51-66: extend, ifft, shift by nfft/2, crop, fade."""
    K = nfft // 2 + 1
    Wf = np.concatenate([Wpos[:K], np.conj(Wpos[K - 2:0:-1])])
    w = np.fft.ifft(Wf).real
    n_shift = nfft // 2
    w = applySubsampleDelay(w[:, None], n_shift)[:, 0]
    w = w[n_shift - length // 2:n_shift + length // 2]
    return w * getFadeWindow(length)


def getMagLsSphericalHeadFilter(micRadius, order, fs, length, **kw):
    """lib/getMagLsSphericalHeadFilter.m:1-68 -> ``(wShf [len], W_Shf [nfft])``."""
    nfft = min(kw.get("NFFT_MAX_LEN", DEFAULTS["NFFT_MAX_LEN"]), 2 * length)
    f = np.linspace(0, fs / 2, nfft // 2 + 1)
    kr = 2 * np.pi * f / 343.0 * micRadius
    simN = int(math.ceil(fs * np.pi * micRadius / 343.0))
    bn_hi = sphModalCoeffs(simN, kr, "rigid", 0)
    bn_lo = bn_hi[:, :order + 1]
    hi_df = _df_response(sh_repToOrder(bn_hi.T).T)
    lo_df = _df_response(sh_repToOrder(bn_lo.T).T)
    W = 1.0 / (hi_df / lo_df)
    K = nfft // 2 + 1
    Wfull = np.concatenate([W[:K], np.conj(W[K - 2:0:-1])])
    return _zero_phase_tail(W, nfft, length), Wfull


def getMagLsArrayDiffuseFilter(micRadius, micGridAziRad, micGridZenRad, order, fs, length,
                               shDefinition="real", shFunction=None, **kw):
    """lib/getMagLsArrayDiffuseFilter.m:1-92 -> ``wAdf [len]``."""
    shFunction = shFunction or getSH
    nfft = min(kw.get("NFFT_MAX_LEN", DEFAULTS["NFFT_MAX_LEN"]), 2 * length)
    f = np.linspace(0, fs / 2, nfft // 2 + 1)
    kr = 2 * np.pi * f / 343.0 * micRadius
    simN = int(math.ceil(fs * np.pi * micRadius / 343.0))
    bn_hi = sh_repToOrder(sphModalCoeffs(simN, kr, "rigid", 0).T).T            # [K, S]
    mics = np.stack([np.asarray(micGridAziRad, float).ravel(), np.asarray(micGridZenRad, float).ravel()], 1)
    Y_hi_conj = np.conj(shFunction(simN, mics, shDefinition)).T                 # [S, M]
    bn_lo_dir = bn_hi @ Y_hi_conj                                               # [K, M]
    Y_lo = shFunction(order, mics, shDefinition)                                # [M, nsh]
    bn_lo = bn_lo_dir @ Y_lo
    hi_df = _df_response(bn_hi)
    lo_df = _df_response(bn_lo)
    lo_df = lo_df / lo_df[0]
    W_alias = hi_df / lo_df
    _, W_shf = getMagLsSphericalHeadFilter(micRadius, order, fs, length, **kw)
    W_adf = W_shf[:W_alias.shape[0]] * W_alias
    return _zero_phase_tail(W_adf, nfft, length)


def encodeSH(sig, micGridAziRad, micGridZenRad, order, shDefinition="real", shFunction=None):
    """verifyEMagLs.m:235-236 / testEMagLs.m:98-99: ``E = getSH(N, mics, def).'; sh = sig * pinv(E)``."""
    shFunction = shFunction or getSH
    mics = np.stack([np.asarray(micGridAziRad, float).ravel(), np.asarray(micGridZenRad, float).ravel()], 1)
    E = shFunction(order, mics, shDefinition).T
    return np.asarray(sig) @ np.linalg.pinv(E)


def encodeCH(sig, micGridAziRad, order, chDefinition="real"):
    """testEMagLs.m:99-102: ``EncEma = getCH(N, azi, def); ch = sig * pinv(EncEma.')``."""
    E = getCH(order, micGridAziRad, chDefinition).T
    return np.asarray(sig) @ np.linalg.pinv(E)


def rotateSH(sig, yawRad, pitchRad=0.0, rollRad=0.0):
    """dependencies/binauralDecode.m:26-30 calls ``rotateHOA_N3D(in, yaw_deg, 0, 0)`` of
    polarch/Higher-Order-Ambisonics, which the reference does NOT vendor.  Its published body is
    ``Rzyx = euler2rotationMatrix(-yaw, -pitch, roll, 'zyx'); Rshd = getSHrotMtx(Rzyx, N, 'real');
    out = in * Rshd.'`` -- restated here with the reference's own (vendored) ``euler2rotationMatrix``
    and ``getSHrotMtx``."""
    sig = np.asarray(sig, dtype=float)
    N = int(round(math.sqrt(sig.shape[1]))) - 1
    R = euler2rotationMatrix(-yawRad, -pitchRad, rollRad, "zyx")
    return sig @ getSHrotMtx(R, N, "real").T
