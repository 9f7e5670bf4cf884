"""NumPy mirrors of three device algorithms, statement by statement, so that their arithmetic is pinned on
the CPU as well (the CUDA kernels themselves are tested against the oracle in the `-m gpu` suite):

* balanced base-256 slicing of `oz::slice_digits` (emagls_b200/csrc/ozaki.cuh) and the exactness of the
  sliced product with the i + j < T truncation,
* the Hermitian sweep inversion of `gram_sweep_kernel` (emagls_b200/csrc/gram_kernels.cu),
* the radix-2/4 Stockham FFT with real pack/unpack of `fused_render_kernel` (emagls_b200/csrc/render.cu).
"""
import numpy as np
import pytest


def slice_digits(a, T):
    """oz::slice_digits<T>: |a| <= 64 -> digits q[T] with a = sum q_s 256^-s + r."""
    nh = T - 3
    bias_hi = {3: 0x808080, 2: 0x8080, 1: 0x80}[nh]
    ah = a * float(1 << (8 * (nh - 1)))
    hi = np.rint(ah).astype(np.int64)
    lo = np.rint((ah - hi) * 16777216.0).astype(np.int64)
    yl = lo + 0x808080
    hi = hi + (yl >> 24)
    zl = yl ^ 0x808080
    sb = lambda v: ((v & 0xFF) + 128) % 256 - 128          # noqa: E731  (int)(signed char)
    q = np.zeros((T,) + a.shape, dtype=np.int64)
    q[T - 1], q[T - 2], q[T - 3] = sb(zl), sb(zl >> 8), sb(zl >> 16)
    zh = (hi + bias_hi) ^ bias_hi
    for j in range(nh):
        q[nh - 1 - j] = sb(zh >> (8 * j))
    return q


@pytest.mark.parametrize("T", [4, 5, 6])
def test_balanced_base256_digits(T):
    rng = np.random.default_rng(T)
    a = np.concatenate([rng.uniform(-64, 64, 100000),
                        [64.0, -64.0, 0.0, np.nextafter(64.0, 0), -np.nextafter(64.0, 0), 0.4999999, 2.0 ** -30,
                         0.498046875 + 1e-9, 63.5, -63.5]])
    q = slice_digits(a, T)
    assert q.min() >= -128 and q.max() <= 127 and np.abs(q[0]).max() <= 64      # int8 operands
    rec = sum(q[s].astype(float) * 256.0 ** -s for s in range(T))
    assert np.abs(rec - a).max() <= 0.5 * 256.0 ** -(T - 1) * (1 + 1e-12)        # |r| <= half the last digit


def test_sliced_product_error_and_int32_range():
    """C = sA sB sum_{i+j<T} 256^-(i+j) qA_i . qB_j: exact integer partial sums, 46 bits relative to
    max|a| max|b| K for T = 6, and every diagonal stays inside int32 for the contraction lengths in use."""
    rng = np.random.default_rng(0)
    K, T = 2720, 6
    a, b = rng.standard_normal(K), rng.standard_normal(K)
    ea, eb = np.frexp(np.abs(a).max())[1], np.frexp(np.abs(b).max())[1]
    qa, qb = slice_digits(a * 2.0 ** (6 - ea), T), slice_digits(b * 2.0 ** (6 - eb), T)
    acc = [sum(int(qa[i] @ qb[d - i]) for i in range(d + 1)) for d in range(T)]
    assert max(abs(v) for v in acc) < 2 ** 31
    assert T * 2 ** 14 * K < 2 ** 31                                             # oz::contraction_fits
    c = 2.0 ** (ea - 6) * 2.0 ** (eb - 6) * sum(acc[d] * 256.0 ** -d for d in range(T))
    assert abs(c - a @ b) <= 2.0 ** -44 * np.abs(a).max() * np.abs(b).max() * K


@pytest.mark.parametrize("M,cond", [(32, 4e3), (25, 50.0), (7, 1e2)])
def test_hermitian_sweep_inversion(M, cond):
    """gram_sweep_kernel: A <- sweep_k(A) for k = 0..M-1 gives -G^-1; the pivots are the Cholesky pivots."""
    rng = np.random.default_rng(M)
    U, _ = np.linalg.qr(rng.standard_normal((M, M)) + 1j * rng.standard_normal((M, M)))
    lam = np.logspace(0, -np.log10(cond), M)
    G = (U * lam) @ U.conj().T
    G = (G + G.conj().T) / 2
    A = G.copy()
    piv = []
    for k in range(M):
        col = A[:, k].copy()                      # published pivot column
        d = col[k].real
        piv.append(d)
        inv = 1.0 / d
        for i in range(M):                        # lane i
            if i == k:
                A[k, :] = A[k, :] * inv
                A[k, k] = -inv
            else:
                f = A[i, k] * inv
                for j in range(M):
                    if j != k:
                        A[i, j] -= f * np.conj(col[j])
                A[i, k] = f
    Ginv = np.linalg.inv(G)
    assert np.abs(-A - Ginv).max() <= 1e-12 * cond * np.abs(Ginv).max()
    L = np.linalg.cholesky(G)
    assert np.allclose(piv, np.abs(np.diag(L)) ** 2, rtol=1e-9)


def stockham(a, WM, inverse):
    """stockham_fft<INV, 1>: leading radix-2 pass when log2(M) is odd, then radix-4 passes."""
    M = a.size
    logM = int(round(np.log2(M)))
    src, dst = a.copy(), np.empty_like(a)
    p = 1
    if logM & 1:
        j = np.arange(M // 2)
        dst[2 * j], dst[2 * j + 1] = src[j] + src[j + M // 2], src[j] - src[j + M // 2]
        src, dst = dst, src
        p = 2
    T = M // 4
    while p < M:
        t = np.arange(T)
        k = t & (p - 1)
        tw = M // (4 * p)
        w1, w2 = WM[k * tw], WM[2 * k * tw]
        if inverse:
            w1, w2 = np.conj(w1), np.conj(w2)
        w3 = w1 * w2
        u0, u1, u2, u3 = src[t], src[t + T] * w1, src[t + 2 * T] * w2, src[t + 3 * T] * w3
        v0, v1, v2, v3 = u0 + u2, u0 - u2, u1 + u3, (u1 - u3) * (1j if inverse else -1j)
        j = ((t - k) << 2) + k
        dst[j], dst[j + p], dst[j + 2 * p], dst[j + 3 * p] = v0 + v2, v1 + v3, v0 - v2, v1 - v3
        src, dst = dst, src
        p <<= 2
    return src


@pytest.mark.parametrize("N", [8, 64, 2048, 4096])
def test_stockham_real_fft_roundtrip(N):
    rng = np.random.default_rng(N)
    M = N // 2
    WM = np.exp(-2j * np.pi * np.arange(max(M // 2, 1)) / M)
    WN = np.exp(-2j * np.pi * np.arange(M + 1) / N)
    x = rng.standard_normal(N)
    Z = stockham(x[0::2] + 1j * x[1::2], WM, False)
    k = np.arange(M + 1)
    zk, zm = Z[k % M], np.conj(Z[(M - k) % M])
    X = 0.5 * (zk + zm) - 0.5j * WN[k] * (zk - zm)                               # rfft unpack
    assert np.abs(X - np.fft.rfft(x)).max() <= 1e-12 * np.abs(X).max()
    k = np.arange(M)
    yk, ym = X[k], np.conj(X[M - k])
    zy = stockham((yk + ym) + 1j * np.conj(WN[k]) * (yk - ym), WM, True) / N     # irfft pack
    y = np.empty(N)
    y[0::2], y[1::2] = zy.real, zy.imag
    assert np.abs(y - x).max() <= 1e-12 * np.abs(x).max()
