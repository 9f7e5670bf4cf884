"""Offline (CPU, NumPy) study of the one-sided Jacobi iteration of factor_kernel on the R factors of the em32
steering matrices: sweeps and accuracy of the clipped projector Pb against a LAPACK SVD, for
  * the stop rule ("a sweep whose largest cosine^2 stayed below thr ends the iteration"),
  * QR / QRCP preconditioning (R -> qr(R^H)),
  * incrementally updated column norms (de Rijk) refreshed once per sweep.
Row-cyclic pair order (the kernel uses a round-robin order; the sweep counts agree to +-1).
usage: PYTHONPATH=. python tools/proto_jacobi.py > profiles/r01_jacobi_offline_study.txt"""
import numpy as np
import scipy.linalg as sla

import oracle
from emagls_b200 import synth

g = synth.load_grids()
az, ze, maz, mze = g["hrirGridAziRad"], g["hrirGridZenRad"], g["micGridAziRad"], g["micGridZenRad"]
fs, K = 48000, 513
f = np.linspace(0, fs / 2, K)
Ymic = oracle.getSH(19, np.stack([maz, mze], 1), "real")
Yc = oracle.getSH(19, np.stack([az, ze], 1), "real").T


def jacobi(X0, bigthr, incremental_norms=False):
    X = X0.copy()
    n = X.shape[1]
    J = np.eye(n, dtype=complex)
    tol2 = (np.finfo(float).eps * np.sqrt(n)) ** 2
    for sw in range(1, 41):
        big = False
        nrm = np.einsum("ij,ij->j", X.conj(), X).real if incremental_norms else None
        for p in range(n - 1):
            for q in range(p + 1, n):
                xp, xq = X[:, p].copy(), X[:, q].copy()
                if incremental_norms:
                    a, b = nrm[p], nrm[q]
                else:
                    a, b = np.vdot(xp, xp).real, np.vdot(xq, xq).real
                gpq = np.vdot(xp, xq)
                gg = abs(gpq) ** 2
                if gg > tol2 * a * b and gg > 0:
                    if gg > bigthr * a * b:
                        big = True
                    d = b - a
                    root = np.sqrt(d * d + 4 * gg)
                    tw = np.copysign(2.0 / (abs(d) + root), d) if d != 0 else 2.0 / root
                    cs = 1 / np.sqrt(1 + tw * tw * gg)
                    sph = gpq * cs * tw
                    X[:, p], X[:, q] = cs * xp - np.conj(sph) * xq, cs * xq + sph * xp
                    jp, jq = J[:, p].copy(), J[:, q].copy()
                    J[:, p], J[:, q] = cs * jp - np.conj(sph) * jq, cs * jq + sph * jp
                    if incremental_norms:
                        nrm[p], nrm[q] = a - tw * gg, b + tw * gg
        if not big:
            return X, J, sw
    return X, J, 40


def projector(X, J, c=0.01):
    s = np.linalg.norm(X, axis=0)
    return np.conj(J) @ np.diag(1 / (s * np.maximum(s, c * s.max()))) @ X.T


print("bin  cond(R)    | stop rule thr (cos^2): sweeps / rel. error of Pb          | preconditioned (thr 1e-18)  | de Rijk norms (thr 1e-14)")
for k in (2, 10, 30, 60, 85):
    bn = -oracle.sphModalCoeffs(19, np.array([2 * np.pi * f[k] / 343.0 * 0.042]))[0]
    pw = (Ymic * oracle.sh_repToOrder(bn[:, None])[:, 0][None, :]) @ Yc
    R = np.linalg.qr(pw.T, mode="r")
    U, s, Vh = np.linalg.svd(R)
    Pref = np.conj(U @ np.diag(1 / np.maximum(s, 0.01 * s[0])) @ Vh)
    err = lambda P: np.abs(P - Pref).max() / np.abs(Pref).max()      # noqa: E731
    cols = []
    for thr in (1e-18, 1e-14, 1e-10, 1e-6):
        X, J, sw = jacobi(R.conj().T, thr)
        cols.append(f"{thr:.0e}: {sw} / {err(projector(X, J)):.1e}")
    _, R2, _ = sla.qr(R, pivoting=True)
    sw_qrcp = jacobi(R2.conj().T, 1e-18)[2]
    sw_qr = jacobi(np.linalg.qr(R.conj().T, mode="r").conj().T, 1e-18)[2]
    X, J, sw_dr = jacobi(R.conj().T, 1e-14, incremental_norms=True)
    print(f"{k:3d}  {s[0] / s[-1]:.1e}    | " + "   ".join(cols) + f" | QRCP {sw_qrcp}, QR(R^H) {sw_qr} sweeps | {sw_dr} / {err(projector(X, J)):.1e}")
