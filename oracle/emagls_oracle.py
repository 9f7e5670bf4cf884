"""FP64 NumPy/SciPy restatement of the eMagLS reference hot path.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Every function cites the
reference file:line (relative to the reference tree) it follows.  MathWorks
built-ins that are not in the reference tree are replaced by the stand-ins
listed in SURVEY.md section 8(c): LAPACK ``zgesdd`` via ``numpy.linalg.svd``,
AMOS Bessel via ``scipy.special.jv/yv``, ``scipy.special.lpmv`` for
``legendre``, pocketfft for ``fft/ifft``, and explicit formulas for
``grpdelay``, ``hann`` and ``fftfilt``.

Array conventions mirror MATLAB: impulse responses are ``[samples, dirs]``,
filters are ``[taps, channels]``, grids are 1-D arrays in radians
(azimuth, zenith).
"""
from __future__ import annotations

import math

import numpy as np
from scipy import signal as _sps
from scipy import special as _sp

__all__ = [
    "sph_besselj", "sph_bessely", "sph_hankel2", "dsph_besselj", "dsph_bessely",
    "dsph_hankel2", "sphModalCoeffs", "getSH", "sh_repToOrder", "getSMAIRMatrix",
    "grpdelay", "applySubsampleDelay", "getFadeWindow", "hann",
    "getShFreqDomainConjugate", "getChFreqDomainConjugate",
    "getCH", "getNnm", "getChToShExpansionMatrix", "euler2rotationMatrix",
    "getSHrotMtx", "complex2realSHMtx",
    "getLsFilters", "getMagLsFilters", "getEMagLsFilters", "getEMagLs2Filters",
    "getEMagLsFiltersEMAinCH", "getEMagLsFiltersEMAinSH", "getEMagLsFiltersFromAtf",
    "binauralDecode", "fftfilt", "regularized_inverse", "diffuseness_matrix", "matlab_round",
    "DEFAULTS",
]

# Constants that sit at the top of every reference function
# (lib/getEMagLs2Filters.m:35-39).  They are parameters here so the stress
# configuration (4096 taps) can lift NFFT_MAX_LEN.
DEFAULTS = dict(NFFT_MAX_LEN=2048, F_CUT_MIN_FREQ=1e3, SVD_REGUL_CONST=0.01, C=343.0)


def matlab_round(x: float) -> int:
    """MATLAB ``round`` (half away from zero) for the positive values used here."""
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


# --------------------------------------------------------------------------
# Array-Response-Simulator: spherical Bessel family and modal coefficients
# --------------------------------------------------------------------------
def sph_besselj(n, x):
    """dependencies/Array-Response-Simulator/sph_besselj.m:11-19."""
    x = np.asarray(x, dtype=float)
    with np.errstate(divide="ignore", invalid="ignore"):
        j = np.sqrt(np.pi / (2 * x)) * _sp.jv(n + 0.5, x)
    j = np.where(x == 0, 1.0 if n == 0 else 0.0, j)
    return j


def sph_bessely(n, x):
    """dependencies/Array-Response-Simulator/sph_bessely.m:11."""
    x = np.asarray(x, dtype=float)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.sqrt(np.pi / (2 * x)) * _sp.yv(n + 0.5, x)


def sph_hankel2(n, x):
    """dependencies/Array-Response-Simulator/sph_hankel2.m:11."""
    with np.errstate(invalid="ignore"):
        return sph_besselj(n, x) - 1j * sph_bessely(n, x)


def dsph_besselj(n, x):
    """dependencies/Array-Response-Simulator/dsph_besselj.m:11."""
    return 1.0 / (2 * n + 1) * (n * sph_besselj(n - 1, x) - (n + 1) * sph_besselj(n + 1, x))


def dsph_bessely(n, x):
    """dependencies/Array-Response-Simulator/dsph_bessely.m:11."""
    with np.errstate(invalid="ignore"):
        return 1.0 / (2 * n + 1) * (n * sph_bessely(n - 1, x) - (n + 1) * sph_bessely(n + 1, x))


def dsph_hankel2(n, x):
    """dependencies/Array-Response-Simulator/dsph_hankel2.m:11."""
    with np.errstate(invalid="ignore"):
        return dsph_besselj(n, x) - 1j * dsph_bessely(n, x)


def sphModalCoeffs(N, kr, arrayType="rigid", dirCoeff=0.0):
    """dependencies/Array-Response-Simulator/sphModalCoeffs.m:25-59.

    Returns ``b_N`` of shape ``[len(kr), N+1]`` (complex).
    """
    kr = np.asarray(kr, dtype=float).ravel()
    b_N = np.zeros((kr.size, N + 1), dtype=complex)
    for n in range(N + 1):
        if arrayType == "open":
            b_N[:, n] = 4 * np.pi * (1j ** n) * sph_besselj(n, kr)
        elif arrayType == "rigid":
            jn = sph_besselj(n, kr)
            jnprime = dsph_besselj(n, kr)
            hn = sph_hankel2(n, kr)
            hnprime = dsph_hankel2(n, kr)
            with np.errstate(divide="ignore", invalid="ignore"):
                temp = 4 * np.pi * (1j ** n) * (jn - (jnprime / hnprime) * hn)
            temp = np.where(kr == 0, 4 * np.pi if n == 0 else 0.0, temp)
            b_N[:, n] = temp
        elif arrayType == "directional":
            jn = sph_besselj(n, kr)
            jnprime = dsph_besselj(n, kr)
            b_N[:, n] = 4 * np.pi * (1j ** n) * (dirCoeff * jn - 1j * (1 - dirCoeff) * jnprime)
        else:
            raise ValueError("Wrong array type")
    b_N[np.isnan(b_N)] = 0  # sphModalCoeffs.m:59
    return b_N


# --------------------------------------------------------------------------
# Spherical-Harmonic-Transform: getSH
# --------------------------------------------------------------------------
def _legendre(n, x):
    """MATLAB ``legendre(n, x)``: rows m = 0..n, Condon-Shortley phase included."""
    m = np.arange(n + 1)[:, None]
    return _sp.lpmv(m, n, np.asarray(x, dtype=float)[None, :])


def getSH(N, dirs, basisType="real"):
    """dependencies/Spherical-Harmonic-Transform/getSH.m:17-89.

    ``dirs`` is ``[D, 2]`` = (azimuth, zenith) in radians; returns ``[D, (N+1)^2]``
    in ACN order.
    """
    dirs = np.asarray(dirs, dtype=float).reshape(-1, 2)
    azi, zen = dirs[:, 0], dirs[:, 1]
    Ndirs = dirs.shape[0]
    is_complex = basisType == "complex"
    if not is_complex and basisType != "real":
        raise ValueError("basisType must be 'real' or 'complex'")
    Y = np.zeros(((N + 1) ** 2, Ndirs), dtype=complex if is_complex else float)
    idx = 0
    cz = np.cos(zen)
    for n in range(N + 1):
        m = np.arange(n + 1)
        Lnm = _legendre(n, cz)  # (n+1) x D
        norm = np.array([math.sqrt((2 * n + 1) * math.factorial(n - mm)
                                   / (4 * math.pi * math.factorial(n + mm))) for mm in m])
        if is_complex:
            Ypos = norm[:, None] * Lnm * np.exp(1j * m[:, None] * azi[None, :])
            if n != 0:
                condon = ((-1.0) ** m[:0:-1])[:, None]
                Yneg = condon * np.conj(Ypos[:0:-1, :])
                Ynm = np.vstack([Yneg, Ypos])
            else:
                Ynm = Ypos
        else:
            Lr = Lnm
            Nn = norm[:, None] * np.ones((1, Ndirs))
            if n != 0:
                condon = ((-1.0) ** np.concatenate([m[:0:-1], m]))[:, None]
                Lr = condon * np.vstack([Lnm[:0:-1, :], Lnm])
                Nn = np.vstack([Nn[:0:-1, :], Nn])
            CosSin = np.zeros((2 * n + 1, Ndirs))
            CosSin[n, :] = 1.0
            if n != 0:
                CosSin[m[1:] + n, :] = math.sqrt(2) * np.cos(m[1:, None] * azi[None, :])
                CosSin[-m[:0:-1] + n, :] = math.sqrt(2) * np.sin(m[:0:-1, None] * azi[None, :])
            Ynm = Nn * Lr * CosSin
        Y[idx:idx + 2 * n + 1, :] = Ynm
        idx += 2 * n + 1
    return Y.T.copy()


def sh_repToOrder(x):
    """dependencies/sh_repToOrder.m:15-20: (N+1) x c  ->  (N+1)^2 x c."""
    x = np.asarray(x)
    vec = x.ndim == 1
    if vec:
        x = x[:, None]
    n = x.shape[0] - 1
    reps = np.repeat(np.arange(n + 1), 2 * np.arange(n + 1) + 1)
    out = x[reps, :]
    return out[:, 0] if vec else out


# --------------------------------------------------------------------------
# getSMAIRMatrix
# --------------------------------------------------------------------------
def getSMAIRMatrix(params: dict):
    """dependencies/getSMAIRMatrix.m:1-141.

    ``params`` keys follow the reference struct fields.  Returns
    ``(smairMat [rows, S, K] complex, params_with_defaults)``.
    """
    p = dict(params)
    if "smaDesignAziZenRad" not in p:
        raise ValueError("default mic layout needs des.3.32.7.txt, which the reference "
                         "tree does not ship (getSMAIRMatrix.m:30-36)")
    p.setdefault("order", 4)
    p.setdefault("fs", 48000)
    p.setdefault("smaRadius", 0.042)
    p.setdefault("arrayType", "rigid")
    p.setdefault("radialFilter", "regul")
    p.setdefault("sourceDist", 2)
    p.setdefault("dirCoeff", 0)
    p.setdefault("waveModel", "planeWave")
    p.setdefault("noiseGainDb", 20)
    p.setdefault("zStyleMaxRe", 1)
    p.setdefault("sourcePosCart", np.array([p["sourceDist"], 0, 0], dtype=float))
    p.setdefault("oversamplingFactor", 4)
    p.setdefault("irLen", 2048)
    p.setdefault("returnRawMicSigs", False)
    p.setdefault("shDefinition", "real")
    p.setdefault("shFunction", getSH)
    C = p.get("C", DEFAULTS["C"])

    nfft = int(p["oversamplingFactor"] * p["irLen"])
    assert nfft % 2 == 0  # getSMAIRMatrix.m:89
    f = np.linspace(0, p["fs"] / 2, nfft // 2 + 1)
    p["sourceDist"] = float(np.linalg.norm(p["sourcePosCart"]))

    simulationOrder = max(int(p["order"]), int(math.ceil(p["fs"] * math.pi * p["smaRadius"] / C)))
    numShsOut = (int(p["order"]) + 1) ** 2
    K = f.size
    mics = np.asarray(p["smaDesignAziZenRad"], dtype=float)
    Y_Hi = p["shFunction"](simulationOrder, mics, p["shDefinition"])
    Y_Lo_pinv = np.linalg.pinv(Y_Hi[:, :numShsOut])

    # getSMAIRMatrix.m:107-108 -- note the minus sign
    bnAll = -sphModalCoeffs(simulationOrder, 2 * np.pi * f / C * p["smaRadius"],
                            p["arrayType"], p["dirCoeff"]).T  # (simN+1) x K

    bn_rep = sh_repToOrder(bnAll)  # S x K
    bn_rep[:, -1] = bn_rep[:, -1].real  # Nyquist: real(Bn), getSMAIRMatrix.m:115-117
    pMics = Y_Hi[:, :, None] * bn_rep[None, :, :]  # Y_Hi * diag(.) per bin

    if p["returnRawMicSigs"]:
        return pMics, p
    smairMat = np.einsum("om,msk->osk", Y_Lo_pinv, pMics)
    if str(p["radialFilter"]).lower() != "none":
        # getSMAIRMatrix.m:129-139.  At the Nyquist bin the reference multiplies by BnTi and then
        # once more by real(BnTi) (the result of the first product is overwritten in place).
        from .frontend_oracle import getRadialFilter
        rad = sh_repToOrder(getRadialFilter(p).T)[:numShsOut, :]  # numShsOut x K
        smairMat = rad[:, None, :] * smairMat
        smairMat[:, :, -1] = rad[:, None, -1].real * smairMat[:, :, -1]
    return smairMat, p


# --------------------------------------------------------------------------
# DSP helpers
# --------------------------------------------------------------------------
def grpdelay(b, f, fs):
    """MathWorks ``grpdelay(b, 1, f, fs)`` for an FIR ``b`` at frequencies ``f`` (Hz).

    Re(sum n b_n z^-n / sum b_n z^-n), with bins where ``abs(den) < 10*eps``
    set to zero (call sites: lib/getEMagLs2Filters.m:74-75).
    """
    b = np.asarray(b, dtype=float).ravel()
    w = 2 * np.pi * np.asarray(f, dtype=float) / fs
    n = np.arange(b.size)
    E = np.exp(-1j * w[:, None] * n[None, :])
    num = E @ (b * n)
    den = E @ b
    bad = np.abs(den) < 10 * np.finfo(float).eps
    num = np.where(bad, 0, num)
    den = np.where(bad, 1, den)
    return (num / den).real


def applySubsampleDelay(sig, delay_samples):
    """dependencies/applySubsampleDelay.m:10-18.

    ``sig`` is ``[n, ...]``; ``delay_samples`` is a scalar or broadcastable to the
    trailing dims (the reference passes 1x1x2 for the two ears).
    """
    sig = np.asarray(sig)
    n = sig.shape[0]
    omega = np.linspace(0, 0.5, n // 2 + 1).reshape((-1,) + (1,) * (sig.ndim - 1))
    d = np.asarray(delay_samples, dtype=float)
    if d.ndim:  # MATLAB broadcasting is left-aligned: pad trailing singleton dims
        d = d.reshape(d.shape + (1,) * (sig.ndim - d.ndim))
    exp_omega = np.exp(-1j * 2 * np.pi * omega * d)
    exp_omega[-1, ...] = exp_omega[-1, ...].real
    exp_omega = np.concatenate([exp_omega, np.conj(exp_omega[-2:0:-1, ...])], axis=0)
    Sig = np.fft.fft(sig, axis=0) * exp_omega
    out = np.fft.ifft(Sig, axis=0)
    if np.isrealobj(sig):
        out = out.real  # MATLAB's ifft returns real for conjugate-symmetric input
    return out


def hann(N):
    """Signal Processing Toolbox ``hann(N)`` (symmetric, zero end points)."""
    if N == 1:
        return np.ones(1)
    n = np.arange(N)
    return 0.5 * (1 - np.cos(2 * np.pi * n / (N - 1)))


def getFadeWindow(irLen, relFadeLen=0.15):
    """dependencies/getFadeWindow.m:9-16."""
    n_in = matlab_round(relFadeLen * irLen)
    n_out = matlab_round(relFadeLen * irLen)
    hin = hann(2 * n_in)
    hout = hann(2 * n_out)
    return np.concatenate([hin[:n_in], np.ones(irLen - (n_in + n_out)), hout[n_out:]])


def getShFreqDomainConjugate(y):
    """dependencies/sh-symmetries/lib/getShFreqDomainConjugate.m:12-27."""
    y = np.asarray(y)
    numShs = y.shape[1]
    shOrder = int(round(math.sqrt(numShs))) - 1
    yConjFlip = np.conj(y[-2:0:-1, :])
    yNeg = np.zeros_like(yConjFlip)
    yNeg[:, 0] = yConjFlip[:, 0]
    for nn in range(1, shOrder + 1):
        for mm in range(-nn, nn + 1):
            yNeg[:, nn * nn + nn + mm] = (-1.0) ** mm * yConjFlip[:, nn * nn + nn - mm]
    return np.vstack([y, yNeg])


def getChFreqDomainConjugate(y):
    """dependencies/sh-symmetries/lib/getChFreqDomainConjugate.m:11-23."""
    y = np.asarray(y)
    numChs = y.shape[1]
    chOrder = (numChs - 1) // 2
    yConjFlip = np.conj(y[-2:0:-1, :])
    yNeg = np.zeros_like(yConjFlip)
    yNeg[:, 0] = yConjFlip[:, 0]
    for ii in range(1, chOrder + 1):
        yNeg[:, 2 * ii - 1] = yConjFlip[:, 2 * ii]
        yNeg[:, 2 * ii] = yConjFlip[:, 2 * ii - 1]
    return np.vstack([y, yNeg])


# --------------------------------------------------------------------------
# circular harmonics / EMA helpers
# --------------------------------------------------------------------------
def getCH(N, aziRad, basisType="real"):
    """dependencies/getCH.m:17-28: [D, 2N+1], order [0,-1,+1,-2,+2,...]."""
    azi = np.asarray(aziRad, dtype=float).ravel()
    if basisType not in ("real", "complex"):
        raise ValueError("basisType must be 'real' or 'complex'")
    Y = np.zeros((azi.size, 2 * N + 1), dtype=complex if basisType == "complex" else float)
    Y[:, 0] = 1
    for nn in range(1, N + 1):
        if basisType == "real":
            Y[:, 2 * nn - 1] = math.sqrt(2) * np.sin(nn * azi)
            Y[:, 2 * nn] = math.sqrt(2) * np.cos(nn * azi)
        else:
            Y[:, 2 * nn - 1] = np.exp(-1j * nn * azi)
            Y[:, 2 * nn] = np.exp(1j * nn * azi)
    return Y


def getNnm(N, zenRad, harmonicsDef="real"):
    """dependencies/getNnm.m:13-30."""
    Nnm = np.zeros((N + 1) ** 2)
    cz = np.array([math.cos(zenRad)])
    for nn in range(N + 1):
        Pn = _legendre(nn, cz)[:, 0]
        for mm in range(-nn, nn + 1):
            am = abs(mm)
            if harmonicsDef == "complex":
                if mm < 0:
                    Pnm = (-1.0) ** am * math.factorial(nn - am) / math.factorial(nn + am) * Pn[am]
                else:
                    Pnm = Pn[mm]
                Nnm[nn * nn + nn + mm] = math.sqrt(((2 * nn + 1) * math.factorial(nn - mm))
                                                   / ((4 * math.pi) * math.factorial(nn + mm))) * Pnm
            elif harmonicsDef == "real":
                Nnm[nn * nn + nn + mm] = (-1.0) ** mm * math.sqrt(
                    ((2 * nn + 1) * math.factorial(nn - am))
                    / ((4 * math.pi) * math.factorial(nn + am))) * Pn[am]
            else:
                raise ValueError("harmonicsDef must be 'real' or 'complex'")
    return Nnm


def getChToShExpansionMatrix(order, harmonicsDef="real"):
    """dependencies/getChToShExpansionMatrix.m:11-18."""
    J = np.zeros(((order + 1) ** 2, 2 * order + 1))
    Nnm = getNnm(order, math.pi / 2, harmonicsDef)
    for n in range(order + 1):
        for m in range(-n, n + 1):
            acn = n * n + n + m
            J[acn, 2 * abs(m) - (1 if m < 0 else 0)] = Nnm[acn]
    return J


def euler2rotationMatrix(alpha, beta, gamma, convention="zyz"):
    """dependencies/Spherical-Harmonic-Transform/euler2rotationMatrix.m:19-50."""
    def Rx(t):
        return np.array([[1, 0, 0], [0, math.cos(t), math.sin(t)], [0, -math.sin(t), math.cos(t)]])

    def Ry(t):
        return np.array([[math.cos(t), 0, -math.sin(t)], [0, 1, 0], [math.sin(t), 0, math.cos(t)]])

    def Rz(t):
        return np.array([[math.cos(t), math.sin(t), 0], [-math.sin(t), math.cos(t), 0], [0, 0, 1]])

    sel = {"x": Rx, "y": Ry, "z": Rz}
    R1 = sel[convention[0]](alpha)
    R2 = sel[convention[1]](beta)
    R3 = sel[convention[2]](gamma)
    return R3 @ R2 @ R1


def complex2realSHMtx(N):
    """dependencies/Spherical-Harmonic-Transform/complex2realSHMtx.m:25-47."""
    T = np.zeros(((N + 1) ** 2, (N + 1) ** 2), dtype=complex)
    T[0, 0] = 1
    idx = 1
    for n in range(1, N + 1):
        m = np.arange(1, n + 1)
        diagT = np.concatenate([1j * np.ones(n), [math.sqrt(2) / 2], (-1.0) ** m]) / math.sqrt(2)
        adiagT = np.concatenate([-1j * (-1.0) ** m[::-1], [math.sqrt(2) / 2], np.ones(n)]) / math.sqrt(2)
        tempT = np.diag(diagT) + np.fliplr(np.diag(adiagT))
        T[idx:idx + 2 * n + 1, idx:idx + 2 * n + 1] = tempT
        idx += 2 * n + 1
    return T


def getSHrotMtx(Rxyz, L, basisType="real"):
    """dependencies/Spherical-Harmonic-Transform/getSHrotMtx.m:59-188
    (Ivanic-Ruedenberg band recursion)."""
    Rxyz = np.asarray(Rxyz, dtype=float)
    R = np.zeros(((L + 1) ** 2, (L + 1) ** 2))
    R[0, 0] = 1
    if L == 0:
        return R.astype(complex) if basisType == "complex" else R
    # first band, index = m + 1 for m in (-1, 0, 1) <-> (y, z, x)
    R_1 = np.array([
        [Rxyz[1, 1], Rxyz[1, 2], Rxyz[1, 0]],
        [Rxyz[2, 1], Rxyz[2, 2], Rxyz[2, 0]],
        [Rxyz[0, 1], Rxyz[0, 2], Rxyz[0, 0]],
    ])
    R[1:4, 1:4] = R_1
    R_lm1 = R_1

    def P(i, l, a, b):
        ri1, rim1, ri0 = R_1[i + 1, 2], R_1[i + 1, 0], R_1[i + 1, 1]
        if b == -l:
            return ri1 * R_lm1[a + l - 1, 0] + rim1 * R_lm1[a + l - 1, 2 * l - 2]
        if b == l:
            return ri1 * R_lm1[a + l - 1, 2 * l - 2] - rim1 * R_lm1[a + l - 1, 0]
        return ri0 * R_lm1[a + l - 1, b + l - 1]

    def U(l, m, n):
        return P(0, l, m, n)

    def V(l, m, n):
        if m == 0:
            return P(1, l, 1, n) + P(-1, l, -1, n)
        if m > 0:
            d = 1.0 if m == 1 else 0.0
            return P(1, l, m - 1, n) * math.sqrt(1 + d) - P(-1, l, -m + 1, n) * (1 - d)
        d = 1.0 if m == -1 else 0.0
        return P(1, l, m + 1, n) * (1 - d) + P(-1, l, -m - 1, n) * math.sqrt(1 + d)

    def Wf(l, m, n):
        if m > 0:
            return P(1, l, m + 1, n) + P(-1, l, -m - 1, n)
        return P(1, l, m - 1, n) - P(-1, l, -m + 1, n)

    band = 4
    for l in range(2, L + 1):
        R_l = np.zeros((2 * l + 1, 2 * l + 1))
        for m in range(-l, l + 1):
            for n in range(-l, l + 1):
                d = 1.0 if m == 0 else 0.0
                denom = (2 * l) * (2 * l - 1) if abs(n) == l else (l * l - n * n)
                u = math.sqrt((l * l - m * m) / denom)
                v = math.sqrt((1 + d) * (l + abs(m) - 1) * (l + abs(m)) / denom) * (1 - 2 * d) * 0.5
                w = math.sqrt((l - abs(m) - 1) * (l - abs(m)) / denom) * (1 - d) * (-0.5)
                if u != 0:
                    u = u * U(l, m, n)
                if v != 0:
                    v = v * V(l, m, n)
                if w != 0:
                    w = w * Wf(l, m, n)
                R_l[m + l, n + l] = u + v + w
        R[band:band + 2 * l + 1, band:band + 2 * l + 1] = R_l
        R_lm1 = R_l
        band += 2 * l + 1
    if basisType == "complex":
        W = complex2realSHMtx(L)
        R = W.T @ R @ np.conj(W)
    return R


# --------------------------------------------------------------------------
# the shared hot loop
# --------------------------------------------------------------------------
def regularized_inverse(pwGrid, svd_regul=DEFAULTS["SVD_REGUL_CONST"]):
    """lib/getEMagLs2Filters.m:87-89.

    ``[U,s,V] = svd(pwGrid.','econ','vector'); s = 1./max(s, c*max(s));
    Y_reg_inv = conj(U) * (s .* V.')`` -> ``[D, M]``.
    """
    U, s, Vh = np.linalg.svd(pwGrid.T, full_matrices=False)
    s = 1.0 / np.maximum(s, svd_regul * s.max())
    return np.conj(U) @ (s[:, None] * np.conj(Vh))


def diffuseness_matrix(R, Rhat):
    """2 x 2 mixing matrix A with A Rhat A^H = R that changes the rendered ear signals least (EXTENSION).

    The diffuse-field covariance constraint ("wDC") was removed from the reference before the surveyed commit
    (CHANGELOG.md:10-18; only its outputs resources/*_wDC.mat remain), so there is no reference code to follow and
    this part is "parity unpinned".  It restates the published formulation the CHANGELOG names (Zaunschirm,
    Schoerkhuber, Hoeldrich 2018, eqs. 12-15): with Cholesky factors R = X^H X, Rhat = Xh^H Xh every
    A = X^H Q Xh^-H with unitary Q matches the covariance; Q = V U^H from Xh X^H = U S V^H (orthogonal Procrustes)
    minimises || (A - I) Xh^H ||_F.
    """
    X = np.linalg.cholesky(R).conj().T          # upper, R = X^H X
    Xh = np.linalg.cholesky(Rhat).conj().T
    U, _, Vh = np.linalg.svd(Xh @ X.conj().T)
    Q = Vh.conj().T @ U.conj().T
    return X.conj().T @ Q @ np.linalg.inv(Xh.conj().T)


def _magls_loop(pw_of_k, HL, HR, numPosFreqs, numCh, k_cut, svd_regul, nyquist_real=True,
                diag=None, diffuseness=False):
    """Hot loop shared by the five eMagLS variants (lib/getEMagLs2Filters.m:85-106).

    ``pw_of_k(k)`` returns the ``[channels, D]`` steering matrix of MATLAB bin
    index ``k`` (1-based).  ``k_cut`` is the reference's 1-based index.
    """
    W_l = np.zeros((numPosFreqs, numCh), dtype=complex)
    W_r = np.zeros((numPosFreqs, numCh), dtype=complex)
    for k in range(2, numPosFreqs + 1):
        pw = pw_of_k(k)
        if diag is not None:
            s = np.linalg.svd(pw, compute_uv=False)
            diag.setdefault("cond", {})[k] = s[0] / s[-1] if s[-1] > 0 else np.inf
        Yri = regularized_inverse(pw, svd_regul)
        i = k - 1
        if k < k_cut:
            W_l[i] = HL[i] @ Yri
            W_r[i] = HR[i] @ Yri
        else:
            phi_l = np.angle(W_l[i - 1] @ pw)
            phi_r = np.angle(W_r[i - 1] @ pw)
            tl = np.abs(HL[i]) * np.exp(1j * phi_l)
            tr = np.abs(HR[i]) * np.exp(1j * phi_r)
            if k == numPosFreqs and nyquist_real:
                tl = tl.real
                tr = tr.real
            W_l[i] = tl @ Yri
            W_r[i] = tr @ Yri
    if diffuseness:
        # EXTENSION (see diffuseness_matrix): after the loop and before the DC bin is set, as the CHANGELOG describes
        # the removed code (CHANGELOG.md:7,20-22).  Target covariance from the HRTF set, rendered covariance from the
        # plane-wave responses of the array through the filters; at a real Nyquist bin the real parts, so that the
        # filters stay real.
        for k in range(2, numPosFreqs + 1):
            i = k - 1
            pw = pw_of_k(k)
            D = pw.shape[1]
            H = np.stack([HL[i], HR[i]])
            Wk = np.stack([W_l[i], W_r[i]])
            Hhat = Wk @ pw
            R = H @ H.conj().T / D
            Rhat = Hhat @ Hhat.conj().T / D
            if k == numPosFreqs and nyquist_real:
                R, Rhat = R.real.astype(complex), Rhat.real.astype(complex)
            A = diffuseness_matrix(R, Rhat)
            Wk = A @ Wk
            W_l[i], W_r[i] = Wk[0], Wk[1]
    # DC fix, lib/getEMagLs2Filters.m:109-110
    W_l[0] = W_l[1].real
    W_r[0] = W_r[1].real
    return W_l, W_r


def _prep_hrirs(hL, hR, nfft, f, fs):
    """lib/getEMagLs2Filters.m:72-81: zero-pad, remove group delay, FFT."""
    T = hL.shape[0]
    hLp = np.zeros((nfft, hL.shape[1]))
    hRp = np.zeros((nfft, hR.shape[1]))
    hLp[:T] = hL
    hRp[:T] = hR
    grpDL = float(np.median(grpdelay(hLp.sum(axis=1), f, fs)))
    grpDR = float(np.median(grpdelay(hRp.sum(axis=1), f, fs)))
    hLp = applySubsampleDelay(hLp, -grpDL)
    hRp = applySubsampleDelay(hRp, -grpDR)
    return np.fft.fft(hLp, axis=0), np.fft.fft(hRp, axis=0), grpDL, grpDR


def _tail(W_l, W_r, numPosFreqs, nfft, length, grpDL, grpDR, extend):
    """lib/getEMagLs2Filters.m:113-135: spectrum extension, ifft, shift, crop, fade."""
    W_l = extend(W_l[:numPosFreqs])
    W_r = extend(W_r[:numPosFreqs])
    wL = np.fft.ifft(W_l, axis=0)
    wR = np.fft.ifft(W_r, axis=0)
    n_shift = nfft // 2
    wL = applySubsampleDelay(wL, n_shift)
    wR = applySubsampleDelay(wR, n_shift + grpDR - grpDL)
    lo = n_shift - length // 2
    wL = wL[lo:lo + length]
    wR = wR[lo:lo + length]
    win = getFadeWindow(length)[:, None]
    return wL * win, wR * win


def _real_extend(W):
    return np.vstack([W, np.conj(W[-2:0:-1])])


def _finish_real(wL, wR, is_real_basis):
    if is_real_basis:
        # the reference asserts isreal(ifft(...)); numerically imag is ~1e-17
        scale = max(np.abs(wL).max(), np.abs(wR).max(), 1e-300)
        assert np.abs(wL.imag).max() <= 1e-9 * scale, "Resulting decoding filters are not real valued."
        assert np.abs(wR.imag).max() <= 1e-9 * scale, "Resulting decoding filters are not real valued."
        return wL.real.copy(), wR.real.copy()
    return wL, wR


def _freq_setup(fs, length, order, cfg):
    nfft = min(cfg["NFFT_MAX_LEN"], 2 * length)
    f = np.linspace(0, fs / 2, nfft // 2 + 1)
    f_cut = max(cfg["F_CUT_MIN_FREQ"], 500 * order)
    k_cut = int(math.ceil(f_cut / f[1]))
    return nfft, f, f.size, k_cut


def _cfg(kw):
    cfg = dict(DEFAULTS)
    cfg.update({k: v for k, v in kw.items() if k in DEFAULTS})
    return cfg


def getEMagLs2Filters(hL, hR, hrirGridAziRad, hrirGridZenRad, micRadius, micGridAziRad,
                      micGridZenRad, order, fs, length, shDefinition="real", shFunction=None,
                      return_spectra=False, **kw):
    """lib/getEMagLs2Filters.m:1-137 -> (wMlsL, wMlsR), each [length, numMics]."""
    cfg = _cfg(kw)
    shFunction = shFunction or getSH
    hL = np.asarray(hL, dtype=float)
    hR = np.asarray(hR, dtype=float)
    assert length >= hL.shape[0], "len too short"
    nfft, f, K, k_cut = _freq_setup(fs, length, order, cfg)
    az = np.asarray(hrirGridAziRad, float).ravel()
    ze = np.asarray(hrirGridZenRad, float).ravel()
    maz = np.asarray(micGridAziRad, float).ravel()
    mze = np.asarray(micGridZenRad, float).ravel()
    params = dict(returnRawMicSigs=True, fs=fs, irLen=nfft, oversamplingFactor=1,
                  simulateAliasing=True, radialFilter="none", smaRadius=micRadius,
                  smaDesignAziZenRad=np.stack([maz, mze], 1), waveModel="planeWave",
                  arrayType="rigid", shDefinition=shDefinition, shFunction=shFunction, C=cfg["C"])
    smairMat, _ = getSMAIRMatrix(params)
    simN = int(round(math.sqrt(smairMat.shape[1]))) - 1
    numMics = maz.size
    Y_conj = np.conj(shFunction(simN, np.stack([az, ze], 1), shDefinition)).T  # S x D
    HL, HR, grpDL, grpDR = _prep_hrirs(hL, hR, nfft, f, fs)
    diag = kw.get("diag")
    W_l, W_r = _magls_loop(lambda k: smairMat[:, :, k - 1] @ Y_conj, HL, HR, K, numMics,
                           k_cut, cfg["SVD_REGUL_CONST"], diag=diag,
                           diffuseness=bool(kw.get("applyDiffusenessConst", False)))
    wL, wR = _tail(W_l, W_r, K, nfft, length, grpDL, grpDR, _real_extend)
    wL, wR = _finish_real(wL, wR, np.isrealobj(Y_conj))
    if return_spectra:
        return wL, wR, dict(W_l=W_l, W_r=W_r, grpDL=grpDL, grpDR=grpDR, k_cut=k_cut, nfft=nfft,
                            HL=HL[:K], HR=HR[:K])
    return wL, wR


def getEMagLsFilters(hL, hR, hrirGridAziRad, hrirGridZenRad, micRadius, micGridAziRad,
                     micGridZenRad, order, fs, length, shDefinition="real", shFunction=None,
                     return_spectra=False, **kw):
    """lib/getEMagLsFilters.m:1-144 -> (wMlsL, wMlsR), each [length, (order+1)^2]."""
    cfg = _cfg(kw)
    shFunction = shFunction or getSH
    hL = np.asarray(hL, dtype=float)
    hR = np.asarray(hR, dtype=float)
    assert length >= hL.shape[0], "len too short"
    nfft, f, K, k_cut = _freq_setup(fs, length, order, cfg)
    az = np.asarray(hrirGridAziRad, float).ravel()
    ze = np.asarray(hrirGridZenRad, float).ravel()
    maz = np.asarray(micGridAziRad, float).ravel()
    mze = np.asarray(micGridZenRad, float).ravel()
    params = dict(order=order, fs=fs, irLen=nfft, oversamplingFactor=1, simulateAliasing=True,
                  radialFilter="none", smaRadius=micRadius,
                  smaDesignAziZenRad=np.stack([maz, mze], 1), waveModel="planeWave",
                  arrayType="rigid", shDefinition=shDefinition, shFunction=shFunction, C=cfg["C"])
    smairMat, _ = getSMAIRMatrix(params)
    simN = int(round(math.sqrt(smairMat.shape[1]))) - 1
    numHarm = (order + 1) ** 2
    Y_Hi_conj = np.conj(shFunction(simN, np.stack([az, ze], 1), shDefinition)).T
    HL, HR, grpDL, grpDR = _prep_hrirs(hL, hR, nfft, f, fs)
    W_l, W_r = _magls_loop(lambda k: smairMat[:, :, k - 1] @ Y_Hi_conj, HL, HR, K, numHarm,
                           k_cut, cfg["SVD_REGUL_CONST"],
                           diffuseness=bool(kw.get("applyDiffusenessConst", False)))
    is_real = np.isrealobj(Y_Hi_conj)
    extend = _real_extend if is_real else getShFreqDomainConjugate
    wL, wR = _tail(W_l, W_r, K, nfft, length, grpDL, grpDR, extend)
    wL, wR = _finish_real(wL, wR, is_real)
    if return_spectra:
        return wL, wR, dict(W_l=W_l, W_r=W_r, grpDL=grpDL, grpDR=grpDR, k_cut=k_cut, nfft=nfft)
    return wL, wR


def getEMagLsFiltersEMAinCH(hL, hR, hrirGridAziRad, hrirGridZenRad, micRadius, micGridAziRad,
                            order, fs, length, shDefinition="real", shFunction=None,
                            chFunction=None, return_spectra=False, **kw):
    """lib/getEMagLsFiltersEMAinCH.m:1-150 -> filters [length, 2*order+1]."""
    cfg = _cfg(kw)
    shFunction = shFunction or getSH
    chFunction = chFunction or getCH
    hL = np.asarray(hL, dtype=float)
    hR = np.asarray(hR, dtype=float)
    assert length >= hL.shape[0], "len too short"
    nfft, f, K, k_cut = _freq_setup(fs, length, order, cfg)
    az = np.asarray(hrirGridAziRad, float).ravel()
    ze = np.asarray(hrirGridZenRad, float).ravel()
    maz = np.asarray(micGridAziRad, float).ravel()
    params = dict(returnRawMicSigs=True, order=order, fs=fs, irLen=nfft, oversamplingFactor=1,
                  radialFilter="none", smaRadius=micRadius,
                  smaDesignAziZenRad=np.stack([maz, np.full_like(maz, np.pi / 2)], 1),
                  waveModel="planeWave", arrayType="rigid", shDefinition=shDefinition,
                  shFunction=shFunction, C=cfg["C"])
    smairMat, _ = getSMAIRMatrix(params)
    simN = int(round(math.sqrt(smairMat.shape[1]))) - 1
    numHarm = 2 * order + 1
    Y_hor_conj = np.conj(shFunction(simN, np.stack([az, ze], 1), shDefinition)).T
    Y_CH_Mic_pinv = np.linalg.pinv(chFunction(order, maz, shDefinition))
    # lib/getEMagLsFiltersEMAinCH.m:74-75 (two pagemtimes)
    smairMat_CH = np.matmul(smairMat.transpose(2, 0, 1), Y_hor_conj)   # [K, M, D] (batched BLAS)
    smairMat_CH = np.matmul(Y_CH_Mic_pinv, smairMat_CH)                # [K, C, D]
    HL, HR, grpDL, grpDR = _prep_hrirs(hL, hR, nfft, f, fs)
    W_l, W_r = _magls_loop(lambda k: smairMat_CH[k - 1], HL, HR, K, numHarm, k_cut,
                           cfg["SVD_REGUL_CONST"])
    is_real = np.isrealobj(Y_hor_conj)
    extend = _real_extend if is_real else getChFreqDomainConjugate
    wL, wR = _tail(W_l, W_r, K, nfft, length, grpDL, grpDR, extend)
    wL, wR = _finish_real(wL, wR, is_real)
    if return_spectra:
        return wL, wR, dict(W_l=W_l, W_r=W_r, grpDL=grpDL, grpDR=grpDR, k_cut=k_cut, nfft=nfft)
    return wL, wR


def getEMagLsFiltersEMAinSH(hL, hR, hrirGridAziRad, hrirGridZenRad, micRadius, micGridAziRad,
                            order, fs, length, shDefinition="real", shFunction=None,
                            chFunction=None, return_spectra=False, **kw):
    """lib/getEMagLsFiltersEMAinSH.m:1-180 -> filters [length, (order+1)^2]."""
    cfg = _cfg(kw)
    shFunction = shFunction or getSH
    chFunction = chFunction or getCH
    hL = np.asarray(hL, dtype=float)
    hR = np.asarray(hR, dtype=float)
    assert length >= hL.shape[0], "len too short"
    nfft, f, K, k_cut = _freq_setup(fs, length, order, cfg)
    az = np.asarray(hrirGridAziRad, float).ravel()
    ze = np.asarray(hrirGridZenRad, float).ravel()
    maz = np.asarray(micGridAziRad, float).ravel()
    params = dict(returnRawMicSigs=True, order=order, fs=fs, irLen=nfft, oversamplingFactor=1,
                  radialFilter="none", smaRadius=micRadius,
                  smaDesignAziZenRad=np.stack([maz, np.full_like(maz, np.pi / 2)], 1),
                  waveModel="planeWave", arrayType="rigid", shDefinition=shDefinition,
                  shFunction=shFunction, C=cfg["C"])
    emaIrMat, _ = getSMAIRMatrix(params)
    simN = int(round(math.sqrt(emaIrMat.shape[1]))) - 1
    # lib/getEMagLsFiltersEMAinSH.m:68-69: directions projected onto the equator
    Y_hor_conj = np.conj(shFunction(simN, np.stack([az, np.full_like(az, np.pi / 2)], 1),
                                    shDefinition)).T
    emaIrDir = np.matmul(emaIrMat.transpose(2, 0, 1), Y_hor_conj)  # K x M x D
    numHarm = (order + 1) ** 2
    D = hL.shape[1]
    YCh = chFunction(order, maz, shDefinition)
    J = getChToShExpansionMatrix(order, shDefinition)
    dec = np.linalg.pinv(YCh.T) @ J.T  # M x numHarm   (:81-83)
    emaIrDir_sh = np.matmul(dec.T, emaIrDir)  # K x H x D
    for d in range(D):  # :86-101
        if ze[d] != np.pi / 2:
            E = euler2rotationMatrix(-az[d], ze[d] - np.pi / 2, az[d], "zyz")
            Rm = getSHrotMtx(E, order, shDefinition)
            emaIrDir_sh[:, :, d] = emaIrDir_sh[:, :, d] @ Rm
    HL, HR, grpDL, grpDR = _prep_hrirs(hL, hR, nfft, f, fs)
    W_l, W_r = _magls_loop(lambda k: emaIrDir_sh[k - 1], HL, HR, K, numHarm, k_cut,
                           cfg["SVD_REGUL_CONST"])
    is_real = np.isrealobj(Y_hor_conj)
    extend = _real_extend if is_real else getShFreqDomainConjugate
    wL, wR = _tail(W_l, W_r, K, nfft, length, grpDL, grpDR, extend)
    wL, wR = _finish_real(wL, wR, is_real)
    if return_spectra:
        return wL, wR, dict(W_l=W_l, W_r=W_r, grpDL=grpDL, grpDR=grpDR, k_cut=k_cut, nfft=nfft,
                            pwGridAll=emaIrDir_sh)
    return wL, wR


def _sph2cart_unit(azi, zen):
    ele = np.pi / 2 - zen
    return np.stack([np.cos(ele) * np.cos(azi), np.cos(ele) * np.sin(azi), np.sin(ele)], 1)


def getEMagLsFiltersFromAtf(hL, hR, hrirGridAziZenRad, atfIrs, atfGridAziZenRad, fs, filterLen,
                            fTrans, return_spectra=False, **kw):
    """lib/getEMagLsFiltersFromAtf.m:1-152 -> filters [filterLen, numMics].

    ``atfIrs`` is ``[T, M, Datf]``.
    """
    cfg = _cfg(kw)
    hL = np.asarray(hL, dtype=float)
    hR = np.asarray(hR, dtype=float)
    atfIrs = np.asarray(atfIrs, dtype=float)
    hg = np.asarray(hrirGridAziZenRad, float).reshape(-1, 2)
    ag = np.asarray(atfGridAziZenRad, float).reshape(-1, 2)
    assert filterLen >= hL.shape[0], "len too short"
    nfft = min(cfg["NFFT_MAX_LEN"], 2 * filterLen)
    f = np.linspace(0, fs / 2, nfft // 2 + 1)
    K = f.size
    kTrans = int(math.ceil(fTrans / f[1]))
    numMics = atfIrs.shape[1]
    hLp = np.zeros((nfft, hL.shape[1]))
    hRp = np.zeros((nfft, hR.shape[1]))
    hLp[:hL.shape[0]] = hL
    hRp[:hR.shape[0]] = hR
    grpDL = float(np.median(grpdelay(hLp.sum(1), f, fs)))
    grpDR = float(np.median(grpdelay(hRp.sum(1), f, fs)))
    hLp = np.roll(hLp, -matlab_round(grpDL), axis=0)  # :48-49 integer shift
    hRp = np.roll(hRp, -matlab_round(grpDR), axis=0)
    HL = np.fft.fft(hLp, nfft, axis=0)
    HR = np.fft.fft(hRp, nfft, axis=0)
    atfs = np.fft.fft(atfIrs, nfft, axis=0)  # nfft x M x Datf
    hc = _sph2cart_unit(hg[:, 0], hg[:, 1])
    ac = _sph2cart_unit(ag[:, 0], ag[:, 1])
    sizes = [hL.shape[1], atfIrs.shape[2]]
    numDirections = min(sizes)
    hrtf_smaller = sizes[0] <= sizes[1]  # MATLAB min() picks the first on a tie (:62)
    if hrtf_smaller:
        dirCart, toMatch = hc, ac
        HLm, HRm = HL[:K], HR[:K]
        atfsM = np.zeros((K, numMics, numDirections), dtype=complex)
    else:
        dirCart, toMatch = ac, hc
        HLm = np.zeros((K, numDirections), dtype=complex)
        HRm = np.zeros((K, numDirections), dtype=complex)
        atfsM = atfs[:K]
    dev = np.zeros(numDirections)
    for ii in range(numDirections):
        dist = np.sqrt(((toMatch - dirCart[ii]) ** 2).sum(1))
        ci = int(np.argmin(dist))
        dev[ii] = math.degrees(math.acos(min(1.0, max(-1.0, float(dirCart[ii] @ toMatch[ci])))))
        if hrtf_smaller:
            atfsM[:, :, ii] = atfs[:K, :, ci]
        else:
            HLm[:, ii] = HL[:K, ci]
            HRm[:, ii] = HR[:K, ci]
    W_l, W_r = _magls_loop(lambda k: atfsM[k - 1], HLm, HRm, K, numMics, kTrans,
                           cfg["SVD_REGUL_CONST"], nyquist_real=(nfft % 2 == 0))
    W_lf = _real_extend(W_l)
    W_rf = _real_extend(W_r)
    wL = np.fft.ifft(W_lf, axis=0)
    wR = np.fft.ifft(W_rf, axis=0)
    n_shift = matlab_round(nfft / 2)
    wL = np.roll(wL, n_shift, axis=0)
    wR = np.roll(wR, n_shift, axis=0)
    lo = n_shift - filterLen // 2
    wL = wL[lo:lo + filterLen]
    wR = wR[lo:lo + filterLen]
    win = getFadeWindow(filterLen)[:, None]  # inline copy at :145-149 is identical
    wL, wR = wL * win, wR * win
    # atfs of real IRs give a conjugate-symmetric spectrum; MATLAB's ifft returns real
    wL, wR = wL.real.copy(), wR.real.copy()
    if return_spectra:
        return wL, wR, dict(W_l=W_l, W_r=W_r, grpDL=grpDL, grpDR=grpDR, kTrans=kTrans, nfft=nfft,
                            meanGridDevDeg=float(dev.mean()))
    return wL, wR


# --------------------------------------------------------------------------
# LS / MagLS in the SH domain
# --------------------------------------------------------------------------
def getLsFilters(hL, hR, hrirGridAziRad, hrirGridZenRad, order, shDefinition="real",
                 shFunction=None):
    """lib/getLsFilters.m:27-34."""
    shFunction = shFunction or getSH
    az = np.asarray(hrirGridAziRad, float).ravel()
    ze = np.asarray(hrirGridZenRad, float).ravel()
    Y_conj = np.conj(shFunction(order, np.stack([az, ze], 1), shDefinition)).T
    Y_pinv = np.linalg.pinv(Y_conj)
    return np.asarray(hL) @ Y_pinv, np.asarray(hR) @ Y_pinv


def getMagLsFilters(hL, hR, hrirGridAziRad, hrirGridZenRad, order, fs, length,
                    shDefinition="real", shFunction=None, return_spectra=False, **kw):
    """lib/getMagLsFilters.m:1-98."""
    cfg = _cfg(kw)
    shFunction = shFunction or getSH
    hL = np.asarray(hL, dtype=float)
    hR = np.asarray(hR, dtype=float)
    assert length >= hL.shape[0], "HRIR len too short"
    nfft, f, K, k_cut = _freq_setup(fs, length, order, cfg)
    az = np.asarray(hrirGridAziRad, float).ravel()
    ze = np.asarray(hrirGridZenRad, float).ravel()
    Y_conj = np.conj(shFunction(order, np.stack([az, ze], 1), shDefinition)).T  # H x D
    Y_pinv = np.linalg.pinv(Y_conj)  # D x H
    # :52-56 -- note: grpdelay is taken on the *unpadded* HRIR sums here
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
        Wf = _real_extend(We) if is_real else getShFreqDomainConjugate(We)
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


# --------------------------------------------------------------------------
# render
# --------------------------------------------------------------------------
def fftfilt(b, x):
    """Signal Processing Toolbox ``fftfilt(b, x)``: first ``len(x)`` samples of the
    linear convolution of the columns of ``x`` with ``b``."""
    x = np.asarray(x)
    b = np.asarray(b)
    if x.ndim == 1:
        return _sps.fftconvolve(x, b)[: x.shape[0]]
    return _sps.fftconvolve(x, b[:, None] if b.ndim == 1 else b, axes=0)[: x.shape[0]]


def binauralDecode(inp, inFs, decodingFilterLeft, decodingFilterRight, decodingFilterFs,
                   compensateDelay=False, sig=None, signalFs=None, horRotAngleRad=None):
    """dependencies/binauralDecode.m:1-65."""
    inp = np.asarray(inp)
    wL = np.asarray(decodingFilterLeft)
    wR = np.asarray(decodingFilterRight)
    if sig is not None and signalFs is not None and signalFs != inFs:
        raise NotImplementedError("resample path (binauralDecode.m:12-16) is not on the hot path")
    if decodingFilterFs != inFs:
        raise NotImplementedError("resample path (binauralDecode.m:18-23) is not on the hot path")
    if horRotAngleRad is not None and horRotAngleRad != 0:
        # binauralDecode.m:26-30: in = rotateHOA_N3D(in, rad2deg(horRotAngleRad), 0, 0); the callee is not
        # vendored by the reference -- its published body is restated in frontend_oracle.rotateSH
        from .frontend_oracle import rotateSH
        inp = rotateSH(inp, float(horRotAngleRad))
    n = inp.shape[0]
    cplx = np.iscomplexobj(inp) or np.iscomplexobj(wL) or np.iscomplexobj(wR)
    left = np.zeros(n, dtype=complex if cplx else float)
    right = np.zeros(n, dtype=complex if cplx else float)
    for ii in range(inp.shape[1]):
        left = left + fftfilt(wL[:, ii], inp[:, ii])
        right = right + fftfilt(wR[:, ii], inp[:, ii])
    if sig is not None:
        s = np.asarray(sig)
        s0 = s[:, 0] if s.ndim > 1 else s
        left = fftfilt(left, s0)
        right = fftfilt(right, s0)
    out = np.stack([left, right], 1)
    if compensateDelay:
        dl = wL.shape[0] // 2
        out = out[dl - 1:, :]
    if np.iscomplexobj(out):
        out = out.real
    return out
