"""Development prototype (NumPy) of the factored eMagLS2 solver that the CUDA kernels implement.

A_k^T = Y_h diag(b_k) Ym^T = Q (R diag(b_k) Ym^T) = Q C_k,  C_k = Q_C R_C,
R_C^H J = X (one-sided Jacobi) -> R_C = J S Vt^H ; Z = conj(Q_C) conj(J) diag(s'/s) X^T
Used to validate the algorithm against the oracle before/while writing CUDA.
"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import oracle
from emagls_b200 import synth


def rr_rounds(n):
    arr = list(range(n)); rounds = []
    for r in range(n - 1):
        rounds.append(([arr[i] for i in range(n // 2)], [arr[n - 1 - i] for i in range(n // 2)]))
        arr = [arr[0]] + [arr[-1]] + arr[1:-1]
    return rounds


_RR = {}


def jacobi_right(X, tol=None, max_sweeps=30):
    """One-sided Jacobi (round-robin parallel ordering) on the columns of X: X J = Xf."""
    n = X.shape[1]
    if n not in _RR:
        _RR[n] = [(np.array(p), np.array(q)) for p, q in rr_rounds(n)]
    X = X.copy(); J = np.eye(n, dtype=complex)
    eps = np.finfo(float).eps
    tol = tol or eps * np.sqrt(n)
    for sweep in range(max_sweeps):
        rot = 0
        for p, q in _RR[n]:
            Xp, Xq = X[:, p], X[:, q]
            a = (np.abs(Xp) ** 2).sum(0); b = (np.abs(Xq) ** 2).sum(0)
            g = (np.conj(Xp) * Xq).sum(0)
            ag = np.abs(g)
            act = (ag > tol * np.sqrt(a * b)) & (ag > 0)
            if not act.any():
                continue
            rot += int(act.sum())
            ags = np.where(act, ag, 1.0)
            ph = np.where(act, g / ags, 1.0)
            zeta = (b - a) / (2 * ags)
            t = np.where(zeta >= 0, 1.0, -1.0) / (np.abs(zeta) + np.sqrt(1 + zeta * zeta))
            c = 1 / np.sqrt(1 + t * t); s = c * t
            c = np.where(act, c, 1.0); s = np.where(act, s, 0.0)
            X[:, p] = c * Xp - s * np.conj(ph) * Xq
            X[:, q] = s * ph * Xp + c * Xq
            Jp, Jq = J[:, p], J[:, q]
            J[:, p], J[:, q] = c * Jp - s * np.conj(ph) * Jq, s * ph * Jp + c * Jq
        if rot == 0:
            break
    return X, J, sweep + 1


def small_operator(C, regul):
    """C (S x M complex) -> Q_C (S x M), Pf = R_C^T (M x M), Pb (M x M), sweeps."""
    Qc, Rc = np.linalg.qr(C)
    X, J, sweeps = jacobi_right(Rc.conj().T)
    s = np.linalg.norm(X, axis=0)
    smax = s.max()
    g = np.where(s > 0, 1.0 / (s * np.maximum(s, regul * smax)), 0.0)   # s'/s
    Pb = np.conj(J) @ (g[:, None] * X.T)
    return Qc, Rc.T.copy(), Pb, sweeps, s


def design(hL, hR, az, ze, mic_xyz, radius, order, fs, length, rotations, regul=0.01, c=343.0,
           nfft_max=2048, f_cut_min=1e3):
    nfft = min(nfft_max, 2 * length); K = nfft // 2 + 1
    f = np.linspace(0, fs / 2, K)
    k_cut = int(np.ceil(max(f_cut_min, 500 * order) / f[1]))
    simN = max(order, int(np.ceil(fs * np.pi * radius / c)))
    S = (simN + 1) ** 2
    Yh = oracle.getSH(simN, np.stack([az, ze], 1), 'real')
    Q, R = np.linalg.qr(Yh)
    bn = -oracle.sphModalCoeffs(simN, 2 * np.pi * f / c * radius, 'rigid').T
    brep = oracle.sh_repToOrder(bn); brep[:, -1] = brep[:, -1].real
    HL, HR, gL, gR = oracle.emagls_oracle._prep_hrirs(hL, hR, nfft, f, fs)
    H = np.stack([HL[:K], HR[:K]], 0)           # 2 x K x D
    T = H @ Q                                     # 2 x K x S
    absH = np.abs(H)
    B = rotations.shape[0]
    M = mic_xyz.shape[0]
    W = np.zeros((B, 2, K, M), dtype=complex)
    stats = []
    for o in range(B):
        maz, mze = synth.angles_from_vectors(mic_xyz @ rotations[o])   # R^T applied to mics
        Ym = oracle.getSH(simN, np.stack([maz, mze], 1), 'real')        # M x S
        for k in range(2, K + 1):
            i = k - 1
            C = R @ (brep[:, i, None] * Ym.T)      # S x M
            Qc, Pf, Pb, sweeps, s = small_operator(C, regul)
            stats.append(sweeps)
            for e in range(2):
                if k < k_cut:
                    tq = T[e, i]
                else:
                    cv = (W[o, e, i - 1] @ Pf) @ Qc.T        # 1 x S
                    y = cv @ Q.T                              # 1 x D
                    ay = np.abs(y)
                    t = absH[e, i] * np.where(ay > 0, y / np.where(ay > 0, ay, 1), 1.0)
                    if k == K:
                        t = t.real
                    tq = t @ Q
                W[o, e, i] = (tq @ np.conj(Qc)) @ Pb
        W[o, :, 0] = W[o, :, 1].real
    return W, dict(gL=gL, gR=gR, k_cut=k_cut, sweeps=np.array(stats))


if __name__ == '__main__':
    g = synth.load_grids()
    az, ze = g['hrirGridAziRad'], g['hrirGridZenRad']
    hL, hR = synth.synth_hrirs(az, ze)
    mic_xyz = synth.unit_vectors(g['micGridAziRad'], g['micGridZenRad'])
    Rm = synth.rotation_yaw_pitch(33.0, 15.0)
    raz, rze = synth.rotate_grid(az, ze, Rm)
    t = time.time()
    wL, wR, sp = oracle.getEMagLs2Filters(hL, hR, raz, rze, g['micRadius'], g['micGridAziRad'], g['micGridZenRad'], 4, g['fs'], 512, return_spectra=True)
    print('oracle', time.time() - t)
    t = time.time()
    W, st = design(hL, hR, az, ze, mic_xyz, g['micRadius'], 4, g['fs'], 512, Rm[None])
    print('proto', time.time() - t, 'sweeps max/mean', st['sweeps'].max(), st['sweeps'].mean())
    for e, Wo in enumerate((sp['W_l'], sp['W_r'])):
        err = np.abs(W[0, e] - Wo).max(1) / np.abs(Wo).max(1)
        print('ear', e, 'per-bin rel err: bins<16 max %.2e, 16..42 max %.2e, >=43 max %.2e' % (err[1:16].max(), err[16:42].max(), err[42:].max()))
        print('   first bins', ' '.join('%.1e' % x for x in err[1:12]))
    np.savez('/tmp/t/proto_W.npz', W=W)
