"""Deterministic synthetic inputs for tests and benchmarks.

Input generation only -- nothing here is on the filter-design path.  The HRIR
set is a rigid-sphere head model (the HRIR set the reference downloads at run
time is not available offline), the grids are the 2702-direction grid stored in
the reference's golden files and the em32 layout of verifyEMagLs.m:28-31.
"""
from __future__ import annotations

import math
import os

import numpy as np
from scipy import special as _sp

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "grids.npz")


def load_grids():
    """2702-direction HRIR grid + em32 microphone layout (radians, zenith convention)."""
    d = np.load(_DATA)
    g = {k: (d[k].copy() if d[k].ndim else float(d[k])) for k in d.files}
    # The golden files store float32-rounded angles; one zenith is float32(pi) > pi, i.e. a
    # direction "past the south pole".  Clip so the grid is a set of genuine directions that
    # can be rotated as unit vectors (the raw values stay available as *_raw).
    g["hrirGridZenRad_raw"] = g["hrirGridZenRad"].copy()
    g["hrirGridZenRad"] = np.clip(g["hrirGridZenRad"], 0.0, np.pi)
    return g


def unit_vectors(azi, zen):
    azi = np.asarray(azi, float)
    zen = np.asarray(zen, float)
    return np.stack([np.sin(zen) * np.cos(azi), np.sin(zen) * np.sin(azi), np.cos(zen)], -1)


def angles_from_vectors(v):
    v = np.asarray(v, float)
    azi = np.arctan2(v[..., 1], v[..., 0])
    zen = np.arctan2(np.hypot(v[..., 0], v[..., 1]), v[..., 2])  # well conditioned at the poles
    return azi, zen


def rotation_yaw_pitch(yaw_deg, pitch_deg):
    """R = Rz(yaw) * Ry(pitch); world direction of grid direction u is R u."""
    a, b = math.radians(yaw_deg), math.radians(pitch_deg)
    Rz = np.array([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1]])
    Ry = np.array([[math.cos(b), 0, math.sin(b)], [0, 1, 0], [-math.sin(b), 0, math.cos(b)]])
    return Rz @ Ry


def orientation_grid(n_yaw=360, pitches=(-45, -35, -25, -15, -5, 5, 15, 25, 35, 45)):
    """The 3600-orientation head-tracking grid of BASELINE config 2 as [B,3,3] rotations."""
    R = [rotation_yaw_pitch(y * 360.0 / n_yaw, p) for p in pitches for y in range(n_yaw)]
    return np.stack(R, 0)


def rotate_grid(azi, zen, R):
    """Angles of R u for every grid direction u (what one reference call per orientation gets)."""
    v = unit_vectors(azi, zen) @ np.asarray(R, float).T
    return angles_from_vectors(v)


def fibonacci_sphere(n):
    """n-point Fibonacci sphere (azi, zen) -- the 64-microphone layout of BASELINE config 5."""
    i = np.arange(n) + 0.5
    zen = np.arccos(1.0 - 2.0 * i / n)
    azi = np.mod(i * math.pi * (3.0 - math.sqrt(5.0)), 2 * math.pi)
    return azi, zen


def _rigid_bn(N, x):
    """4 pi i^n (j_n - j_n'/h_n' h_n), h_n = j_n - i y_n, for x > 0: [len(x), N+1]."""
    n = np.arange(N + 1)[None, :]
    x = np.asarray(x, float)[:, None]
    jn = _sp.spherical_jn(n, x)
    yn = _sp.spherical_yn(n, x)
    djn = _sp.spherical_jn(n, x, derivative=True)
    dyn = _sp.spherical_yn(n, x, derivative=True)
    hn, dhn = jn - 1j * yn, djn - 1j * dyn
    return 4 * np.pi * (1j ** n) * (jn - djn / dhn * hn)


def synth_hrirs(azi, zen, fs=48000.0, taps=128, head_radius=0.0875, ear_azi_deg=90.0,
                delay=40, noise_db=-80.0, seed=20261017, order=40, c=343.0):
    """Rigid-sphere HRIR pair on the given grid: (hL, hR), each [taps, D] float64."""
    rng = np.random.default_rng(seed)
    u = unit_vectors(azi, zen)
    f = np.arange(taps // 2 + 1) * fs / taps
    ka = 2 * np.pi * f[1:] / c * head_radius
    bn = np.conj(_rigid_bn(order, ka))  # [F-1, order+1]
    out = []
    for sgn in (+1.0, -1.0):
        ea = math.radians(sgn * ear_azi_deg)
        ear = np.array([math.cos(ea), math.sin(ea), 0.0])
        cosg = np.clip(u @ ear, -1.0, 1.0)
        P = np.stack([_sp.eval_legendre(n, cosg) for n in range(order + 1)], 0)  # [order+1, D]
        wts = (2 * np.arange(order + 1) + 1) / (4 * np.pi)
        H = np.ones((f.size, u.shape[0]), dtype=complex)
        H[1:] = (bn * wts[None, :]) @ P
        H *= np.exp(-2j * np.pi * f[:, None] / fs * delay)
        H[-1] = H[-1].real
        h = np.fft.irfft(H, n=taps, axis=0)
        h = h + rng.standard_normal(h.shape) * (np.abs(h).max() * 10 ** (noise_db / 20))
        out.append(np.ascontiguousarray(h))
    return out[0], out[1]
