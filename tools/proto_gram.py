"""Development prototype (NumPy) of the hybrid Gram/TSQR solver (v2 of the CUDA hot loop).

Well-conditioned bins (cond_F(G) <= thr):  G = A^H A = Ym (conj(b) Gh b) Ym^T  (Gh = Y_h^T Y_h),
    W = ((t Y_h) * conj(b)) Ym^T G^-T      -- no clipping possible (rigorous), normal equations
other bins: TSQR + Jacobi operator (tools/proto_f2.py).
Forward for every bin: y = Y_h (b * (Ym^T w)).
"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import oracle
from emagls_b200 import synth
from tools.proto_f2 import small_operator


def design(hL, hR, az, ze, mic_xyz, radius, order, fs, length, Rm, thr, regul=0.01, c=343.0):
    nfft = min(2048, 2 * length); K = nfft // 2 + 1
    f = np.linspace(0, fs / 2, K)
    k_cut = int(np.ceil(max(1e3, 500 * order) / f[1]))
    simN = max(order, int(np.ceil(fs * np.pi * radius / c)))
    Yh = oracle.getSH(simN, np.stack([az, ze], 1), 'real')
    Q, R = np.linalg.qr(Yh)
    Gh = Yh.T @ Yh
    bn = -oracle.sphModalCoeffs(simN, 2 * np.pi * f / c * radius, 'rigid').T
    brep = oracle.sh_repToOrder(bn); brep[:, -1] = brep[:, -1].real
    HL, HR, gL, gR = oracle.emagls_oracle._prep_hrirs(hL, hR, nfft, f, fs)
    H = np.stack([HL[:K], HR[:K]], 0)
    TQ = H @ Q
    TY = H @ Yh
    absH = np.abs(H)
    maz, mze = synth.angles_from_vectors(mic_xyz @ Rm)
    Ym = oracle.getSH(simN, np.stack([maz, mze], 1), 'real')
    M = Ym.shape[0]
    W = np.zeros((2, K, M), dtype=complex)
    mode = np.zeros(K, int); condF = np.zeros(K)
    for k in range(2, K + 1):
        i = k - 1
        b = brep[:, i]
        Bm = b[:, None] * Ym.T                     # S x M  (diag(b) Ym^T)
        G = Bm.conj().T @ (Gh @ Bm)                # M x M hermitian
        G = 0.5 * (G + G.conj().T)
        gram = False
        try:
            L = np.linalg.cholesky(G)
            Li = np.linalg.inv(L)
            Ginv = Li.conj().T @ Li
            condF[i] = np.linalg.norm(G) * np.linalg.norm(Ginv)
            gram = condF[i] <= thr
        except np.linalg.LinAlgError:
            condF[i] = np.inf
        mode[i] = gram
        if not gram:
            C = R @ Bm
            Qc, Pf, Pb, sweeps, s = small_operator(C, regul)
        for e in range(2):
            if k < k_cut:
                z = TY[e, i]; tq = TQ[e, i]
            else:
                u = b * (Ym.T @ W[e, i - 1])
                y = Yh @ u
                ay = np.abs(y)
                t = absH[e, i] * np.where(ay > 0, y / np.where(ay > 0, ay, 1), 1.0)
                if k == K:
                    t = t.real
                z = t @ Yh
                tq = t @ Q
            if gram:
                v = (z * np.conj(b)) @ Ym.T        # 1 x M
                W[e, i] = v @ Ginv.T
            else:
                W[e, i] = (tq @ np.conj(Qc)) @ Pb
    W[:, 0] = W[:, 1].real
    return W, mode, condF


if __name__ == '__main__':
    g = synth.load_grids()
    az, ze = g['hrirGridAziRad'], g['hrirGridZenRad']
    hL, hR = synth.synth_hrirs(az, ze)
    mic_xyz = synth.unit_vectors(g['micGridAziRad'], g['micGridZenRad'])
    Rm = synth.rotation_yaw_pitch(33.0, 15.0)
    raz, rze = synth.rotate_grid(az, ze, Rm)
    wL, wR, sp = oracle.getEMagLs2Filters(hL, hR, raz, rze, g['micRadius'], g['micGridAziRad'], g['micGridZenRad'], 4, g['fs'], 512, return_spectra=True)
    for thr in (1e2, 1e3, 1e4, 1e5, 1e6):
        W, mode, condF = design(hL, hR, az, ze, mic_xyz, g['micRadius'], 4, g['fs'], 512, Rm, thr)
        out = []
        for e, Wo in enumerate((sp['W_l'], sp['W_r'])):
            err = np.abs(W[e] - Wo).max(1) / np.abs(Wo).max(1)
            gi = np.where(mode == 1)[0]
            out.append('ear%d gram-bins max err %.2e, all bins>=16 max %.2e' % (e, err[gi].max() if gi.size else 0, err[16:].max()))
        print('thr %.0e: gram bins %d/%d (first %d) | %s' % (thr, mode.sum(), mode.size - 1, np.where(mode == 1)[0].min() if mode.sum() else -1, ' | '.join(out)))
