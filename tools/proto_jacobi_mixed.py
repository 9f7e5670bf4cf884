"""Offline (CPU, NumPy) study for the next step on svdclip_kernel (tsqr_kernels.cu): mixed-precision one-sided Jacobi.
For a few em32 bins (R factors of the steering matrices, lib/getEMagLs2Filters.m:86-89): sweeps of the FP64 iteration
from a cold start and from the previous bin's J (what the kernel does today), against FP32 sweeps from that warm start
followed by one or two Newton-Schulz steps on the accumulated rotation (its unitarity is only 1e-6 after FP32 sweeps), X = R^H V
in FP64 and FP64 sweeps to convergence; relative error of the clipped projector Pb against a LAPACK SVD in each case.
Row-cyclic pair order (the kernel uses a round-robin order; the sweep counts agree to +-1).
usage: PYTHONPATH=. python tools/proto_jacobi_mixed.py > profiles/r02_jacobi_mixed_offline_study.txt"""
import numpy as np

import oracle
from emagls_b200 import synth
g = synth.load_grids()
az, ze, maz, mze = g["hrirGridAziRad"], g["hrirGridZenRad"], g["micGridAziRad"], g["micGridZenRad"]
fs, K = 48000, 513
f = np.linspace(0, fs / 2, K)
Ymic = oracle.getSH(19, np.stack([maz, mze], 1), "real")
Yc = oracle.getSH(19, np.stack([az, ze], 1), "real").T

def jacobi(X0, J0, bigthr, ctype, rtype, maxsw=40):
    X = X0.astype(ctype).copy(); J = J0.astype(ctype).copy()
    n = X.shape[1]
    eps = np.finfo(rtype).eps
    tol2 = rtype((eps * np.sqrt(n)) ** 2)
    for sw in range(1, maxsw + 1):
        big = False
        for p in range(n - 1):
            for q in range(p + 1, n):
                xp, xq = X[:, p].copy(), X[:, q].copy()
                a, b = rtype(np.vdot(xp, xp).real), rtype(np.vdot(xq, xq).real)
                gpq = ctype(np.vdot(xp, xq))
                gg = rtype(abs(gpq) ** 2)
                if gg > tol2 * a * b and gg > 0:
                    if gg > rtype(bigthr) * a * b: big = True
                    d = b - a
                    root = np.sqrt(d * d + 4 * gg)
                    tw = rtype(np.copysign(2.0 / (abs(d) + root), d) if d != 0 else 2.0 / root)
                    cs = rtype(1 / np.sqrt(1 + tw * tw * gg))
                    sph = ctype(gpq * cs * tw)
                    X[:, p], X[:, q] = cs * xp - np.conj(sph) * xq, cs * xq + sph * xp
                    jp, jq = J[:, p].copy(), J[:, q].copy()
                    J[:, p], J[:, q] = cs * jp - np.conj(sph) * jq, cs * jq + sph * jp
        if not big: return X, J, sw
    return X, J, maxsw

def projector(X, J, c=0.01):
    s = np.linalg.norm(X, axis=0)
    return np.conj(J) @ np.diag(1 / (s * np.maximum(s, c * s.max()))) @ X.T

def Rk(k):
    bn = -oracle.sphModalCoeffs(19, np.array([2 * np.pi * f[k] / 343.0 * 0.042]))[0]
    pw = (Ymic * oracle.sh_repToOrder(bn[:, None])[:, 0][None, :]) @ Yc
    return np.linalg.qr(pw.T, mode="r")

I = np.eye(32, dtype=complex)
for k0 in (4, 12, 30, 60, 84):
    R0 = Rk(k0 - 1)
    _, Jprev, _ = jacobi(R0.conj().T, I, 1e-14, np.complex128, np.float64)
    R = Rk(k0)
    U, s, Vh = np.linalg.svd(R)
    Pref = np.conj(U @ np.diag(1 / np.maximum(s, 0.01 * s[0])) @ Vh)
    err = lambda P: np.abs(P - Pref).max() / np.abs(Pref).max()
    Xc, Jc, swc = jacobi(R.conj().T, I, 1e-14, np.complex128, np.float64)
    Xw, Jw, sww = jacobi(R.conj().T @ Jprev, Jprev, 1e-14, np.complex128, np.float64)
    res = []
    for ns_steps in (1, 2):
        X32, J32, sw32 = jacobi(R.conj().T @ Jprev, Jprev, 1e-8, np.complex64, np.float32, maxsw=12)
        V = J32.astype(np.complex128)
        for _ in range(ns_steps):
            V = V @ (1.5 * I - 0.5 * (V.conj().T @ V))   # Newton-Schulz step towards the unitary polar factor
        orth = np.abs(V.conj().T @ V - I).max()
        Xm, Jm, swm = jacobi(R.conj().T @ V, V, 1e-14, np.complex128, np.float64)
        res.append(f"{ns_steps} Newton-Schulz step(s): fp32 sweeps {sw32}, orth {orth:.0e}, fp64 sweeps {swm}, err {err(projector(Xm, Jm)):.1e}")
    print(f"bin {k0} cond {s[0]/s[-1]:.1e}: cold {swc} (err {err(projector(Xc,Jc)):.1e}), warm {sww} (err {err(projector(Xw,Jw)):.1e}) | " + " | ".join(res), flush=True)
