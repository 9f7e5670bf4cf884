"""High-precision arbitration oracle for the per-bin regularised inverse (TEST INFRASTRUCTURE ONLY).

SURVEY.md 7.1-1 / 8-c protocol item 2: on the ill-conditioned bins (sigma_min/sigma_max down to 1e-13 for
em32) two correct FP64 implementations of

    [U,s,V] = svd(pwGrid.','econ','vector'); s = 1./max(s, c*max(s)); Y_reg_inv = conj(U)*(s.*V.');
    W(k,:)  = H(k,:) * Y_reg_inv                                    (lib/getEMagLs2Filters.m:86-92)

differ by far more than 1e-10, so parity between them decides nothing.  This module computes the result
the reference's formula has in exact arithmetic on given FP64 inputs:

  * ``pwGrid = smairMat(:,:,k) * Y_conj`` (lib/getEMagLs2Filters.m:86) is formed EXACTLY: every FP64 entry
    is an integer times a power of two, the products are accumulated in Python integers;
  * the Gram matrix ``G = pwGrid.' ^H pwGrid.'`` is exact as well (squaring the condition number is harmless
    without rounding);
  * its Hermitian eigendecomposition G = V diag(s^2) V^H is computed by mpmath at ``dps`` decimal digits
    (default 120: cond(G) = 1e26 for em32 bin 1 leaves > 90 digits);
  * ``conj(U) diag(g) V^T = conj(A) conj(V) diag(g / s) V^T`` with A = pwGrid.', so the filter row is
    ``W = (H conj(A)) conj(V) diag(1 / (s max(s, c s_max))) V^T``, evaluated in mpmath and rounded once.

Only tests/ and tests/golden/make_hp_goldens.py import this module (never the product path).
"""
from __future__ import annotations

import math

import numpy as np


def _to_int_matrix(x: np.ndarray):
    """FP64 array -> (object array of Python ints, e) with x == ints * 2**e exactly."""
    x = np.asarray(x, dtype=np.float64)
    m, ex = np.frexp(x)                       # x = m * 2**ex, 0.5 <= |m| < 1
    mi = np.ldexp(m, 53).astype(np.int64)     # exact: 53-bit mantissas
    ex = ex.astype(np.int64) - 53
    nz = mi != 0
    e_min = int(ex[nz].min()) if nz.any() else 0
    out = np.zeros(x.shape, dtype=object)
    flat_m, flat_e, flat_o = mi.ravel(), ex.ravel(), out.ravel()
    for i in range(flat_m.size):
        v = int(flat_m[i])
        flat_o[i] = (v << int(flat_e[i] - e_min)) if v else 0
    return out, e_min


def exact_pw_transpose(smair_k: np.ndarray, Y_conj: np.ndarray):
    """A = (smair_k @ Y_conj).T exactly: returns (Are, Aim, e) integer [D x M] with A = (Are + i Aim) 2**e."""
    Sr, es_r = _to_int_matrix(np.ascontiguousarray(smair_k.real))
    Si, es_i = _to_int_matrix(np.ascontiguousarray(smair_k.imag))
    es = min(es_r, es_i)
    Sr = Sr * (1 << (es_r - es)) if es_r > es else Sr
    Si = Si * (1 << (es_i - es)) if es_i > es else Si
    Y, ey = _to_int_matrix(Y_conj)
    Yt = Y.T                                   # [D x S]
    return Yt.dot(Sr.T), Yt.dot(Si.T), es + ey


def exact_ls_rows(smair_k, Y_conj, targets, svd_regul=0.01, dps=120):
    """Exact-arithmetic ``targets @ Y_reg_inv`` (rows of W for one bin), rounded to complex128.

    smair_k [M x S] complex128, Y_conj [S x D] float64, targets [n x D] complex128.
    Also returns the singular values (float64) of pwGrid.
    """
    import mpmath as mp

    Are, Aim, _ea = exact_pw_transpose(np.asarray(smair_k, complex), np.asarray(Y_conj, float))
    D, M = Are.shape
    # exact Gram G = A^H A (integers; the common factor 2**(2 ea) cancels in the final formula except for
    # one power, restored below)
    Gre = Are.T.dot(Are) + Aim.T.dot(Aim)
    Gim = Are.T.dot(Aim) - Aim.T.dot(Are)
    # exact z = targets * conj(A)  ([n x M]); targets are FP64 as well
    Tr, et_r = _to_int_matrix(np.ascontiguousarray(np.asarray(targets).real))
    Ti, et_i = _to_int_matrix(np.ascontiguousarray(np.asarray(targets).imag))
    et = min(et_r, et_i)
    Tr = Tr * (1 << (et_r - et)) if et_r > et else Tr
    Ti = Ti * (1 << (et_i - et)) if et_i > et else Ti
    Zre = Tr.dot(Are) + Ti.dot(Aim)            # (tr + i ti)(ar - i ai)
    Zim = Ti.dot(Are) - Tr.dot(Aim)
    with mp.workdps(dps):
        # scale the integer Gram matrix to O(1) by its largest diagonal entry (exact power of two)
        gmax = max(int(Gre[i, i]) for i in range(M))
        sh = gmax.bit_length()
        G = mp.matrix(M, M)
        for i in range(M):
            for j in range(M):
                G[i, j] = mp.mpc(mp.ldexp(mp.mpf(int(Gre[i, j])), -sh), mp.ldexp(mp.mpf(int(Gim[i, j])), -sh))
        lam, V = mp.eigh(G)                    # G = V diag(lam) V^H, ascending
        s = [mp.sqrt(l) if l > 0 else mp.mpf(0) for l in lam]
        smax = max(s)
        gain = [1 / (si * max(si, svd_regul * smax)) if si > 0 else mp.mpf(0) for si in s]
        # Phi = conj(V) diag(gain) V^T  (in units of 2**(-sh) 2**(-2 ea));  W = z Phi with z in units 2**(et + ea)
        n = Zre.shape[0]
        W = np.zeros((n, M), dtype=complex)
        for r in range(n):
            z = [mp.mpc(mp.mpf(int(Zre[r, j])), mp.mpf(int(Zim[r, j]))) for j in range(M)]
            # y_c = sum_j z_j conj(V[j, c])
            y = [mp.fsum(z[j] * mp.conj(V[j, c]) for j in range(M)) * gain[c] for c in range(M)]
            for m_ in range(M):
                w = mp.fsum(y[c] * V[m_, c] for c in range(M))
                # units: z 2**(et+ea), Phi 2**(-sh - 2 ea)  ->  2**(et - ea - sh)
                w = mp.ldexp(w.real, et - _ea - sh) + 1j * mp.ldexp(w.imag, et - _ea - sh)
                W[r, m_] = complex(w)
        sv = np.array([float(mp.ldexp(si, (sh + 2 * _ea) // 2) * (mp.sqrt(2) if (sh + 2 * _ea) % 2 else 1)) for si in s])[::-1]
    return W, sv


def fp64_ls_rows(smair_k, Y_conj, targets, svd_regul=0.01):
    """The FP64 oracle's value of the same quantity (LAPACK SVD route of oracle.regularized_inverse)."""
    from .emagls_oracle import regularized_inverse
    pw = np.asarray(smair_k) @ np.asarray(Y_conj)
    return np.asarray(targets) @ regularized_inverse(pw, svd_regul)


def rel_err(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))
