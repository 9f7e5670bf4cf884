"""Bring-up diagnostics on a GPU box: each CUDA building block against the oracle (prints errors)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import emagls_b200 as em
from emagls_b200 import synth

def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

which = sys.argv[1:] or ["sh", "bn", "reg", "smair", "design1", "designrot", "render", "perf"]
g = synth.load_grids()
az, ze = g["hrirGridAziRad"], g["hrirGridZenRad"]
maz, mze, r, fs = g["micGridAziRad"], g["micGridZenRad"], g["micRadius"], g["fs"]
h = em.Handle(0)

if "sh" in which:
    for N in (4, 19, 37):
        dirs = np.stack([az, ze], 1)
        for basis in ("real", "complex"):
            t = time.time(); Y = em.getSH(N, dirs, basis, handle=h); dt = time.time() - t
            Yo = oracle.getSH(N, dirs, basis)
            print(f"getSH N={N} {basis}: rel err {rel(Y, Yo):.2e}  ({dt*1e3:.1f} ms)")
    dirs = np.stack([g["micGridAziRad"], g["micGridZenRad"]], 1)
    print("getSH mics N=19:", rel(em.getSH(19, dirs, "real", handle=h), oracle.getSH(19, dirs, "real")))
    # raw grid with zen = float32(pi) > pi
    dirs = np.stack([az, g["hrirGridZenRad_raw"]], 1)
    print("getSH raw grid N=19:", rel(em.getSH(19, dirs, "real", handle=h), oracle.getSH(19, dirs, "real")))

if "bn" in which:
    for N, rr, fss, K in ((19, 0.042, 48000, 513), (37, 0.042, 96000, 4097), (36, 0.08, 48000, 513)):
        f = np.linspace(0, fss / 2, K)
        kr = 2 * np.pi * f / 343.0 * rr
        for typ in ("rigid", "open"):
            b = em.sphModalCoeffs(N, kr, typ, handle=h)
            bo = oracle.sphModalCoeffs(N, kr, typ)
            with np.errstate(divide="ignore", invalid="ignore"):
                e = np.abs(b - bo) / np.abs(bo)
            e[~np.isfinite(e)] = 0
            print(f"sphModalCoeffs N={N} r={rr} {typ}: max elementwise rel err {e.max():.2e} at {np.unravel_index(e.argmax(), e.shape)}; colmax-normalised {np.max(np.abs(b-bo).max(0)/np.abs(bo).max(0)):.2e}")

if "reg" in which:
    rng = np.random.default_rng(0)
    for (Mc, D, grade) in ((32, 2702, 0), (32, 2702, 8), (8, 407, 0), (13, 2702, 3), (25, 2702, 5), (64, 3000, 4), (64, 1444, 0)):
        A = rng.standard_normal((Mc, D)) + 1j * rng.standard_normal((Mc, D))
        if grade:
            # graded singular spectrum
            U, s, Vh = np.linalg.svd(A, full_matrices=False)
            s = s * np.logspace(0, -grade, Mc)
            A = (U * s) @ Vh
        t = rng.standard_normal((3, D)) + 1j * rng.standard_normal((3, D))
        Wo = t @ oracle.regularized_inverse(A, 0.01)
        W = em.regularizedApply(A, t, 0.01, handle=h)
        Wo0 = t @ oracle.regularized_inverse(A, 0.0)
        W0 = em.regularizedApply(A, t, 0.0, handle=h)
        print(f"regularizedApply Mc={Mc} D={D} grade=1e-{grade}: rel err {rel(W, Wo):.2e}; pinv (regul=0) {rel(W0, Wo0):.2e}")

if "smair" in which:
    params = dict(returnRawMicSigs=True, fs=fs, irLen=1024, oversamplingFactor=1, radialFilter="none", smaRadius=r,
                  smaDesignAziZenRad=np.stack([maz, mze], 1))
    t = time.time(); sm, p = em.getSMAIRMatrix(params, handle=h); dt = time.time() - t
    smo, _ = oracle.getSMAIRMatrix(params)
    e = np.abs(sm - smo).max(axis=(0, 1)) / np.abs(smo).max(axis=(0, 1))
    print(f"getSMAIRMatrix raw {sm.shape}: per-bin max rel err {e.max():.2e} ({dt*1e3:.0f} ms)")

hL, hR = synth.synth_hrirs(az, ze)

def bin_report(tag, W, Wo):
    err = np.abs(W - Wo).max(1) / np.abs(Wo).max(1)
    print(f"  {tag}: bins 1..15 max {err[1:16].max():.2e} | 16..41 max {err[16:42].max():.2e} | >=42 max {err[42:].max():.2e} | DC {err[0]:.2e}")
    print("     first bins:", " ".join(f"{x:.1e}" for x in err[1:10]))

if "design1" in which:
    t = time.time()
    wL, wR, sp = em.getEMagLs2Filters(hL, hR, az, ze, r, maz, mze, 4, fs, 512, handle=h, return_spectra=True)
    dt = time.time() - t
    t = time.time()
    oL, oR, osp = oracle.getEMagLs2Filters(hL, hR, az, ze, r, maz, mze, 4, fs, 512, return_spectra=True)
    dto = time.time() - t
    print(f"design1 (single reference call): cuda {dt:.2f}s oracle {dto:.2f}s")
    bin_report("L", sp[:, :, 0].T.T, osp["W_l"])
    bin_report("R", sp[:, :, 1], osp["W_r"])
    print(f"  filters: rel err L {rel(wL, oL):.2e} R {rel(wR, oR):.2e}; end taps {wL[0,0]:.1e} {wL[-1,0]:.1e}")
    t = time.time()
    wL2, wR2 = em.getEMagLs2Filters(hL, hR, az, ze, r, maz, mze, 4, fs, 512, handle=h)
    print(f"  second call {time.time()-t:.3f}s, repeatable: {np.array_equal(wL, wL2)}")

if "designrot" in which:
    Rm = np.stack([synth.rotation_yaw_pitch(33.0, 15.0), synth.rotation_yaw_pitch(-120.0, -35.0), np.eye(3)])
    t = time.time()
    wL, wR, sp = em.getEMagLs2Filters(hL, hR, az, ze, r, maz, mze, 4, fs, 512, rotations=Rm, handle=h, return_spectra=True)
    print(f"designrot B=3: cuda {time.time()-t:.2f}s")
    for o in range(2):
        raz, rze = synth.rotate_grid(az, ze, Rm[o])
        oL, oR, osp = oracle.getEMagLs2Filters(hL, hR, raz, rze, r, maz, mze, 4, fs, 512, return_spectra=True)
        bin_report(f"o={o} L", sp[:, :, o, 0], osp["W_l"])
        bin_report(f"o={o} R", sp[:, :, o, 1], osp["W_r"])
        print(f"  o={o} filters rel err L {rel(wL[:, :, o], oL):.2e} R {rel(wR[:, :, o], oR):.2e}")

if "render" in which:
    rng = np.random.default_rng(3)
    for (n, ch, ln, comp) in ((5000, 32, 512, False), (70000, 32, 512, True), (1000, 25, 512, False), (300, 4, 64, True), (360290, 32, 512, False)):
        x = rng.standard_normal((n, ch)); wl = rng.standard_normal((ln, ch)); wr = rng.standard_normal((ln, ch))
        t = time.time(); y = em.binauralDecode(x, 48000, wl, wr, 48000, comp, handle=h); dt = time.time() - t
        yo = oracle.binauralDecode(x, 48000, wl, wr, 48000, comp)
        print(f"binauralDecode n={n} ch={ch} len={ln} comp={comp}: shape {y.shape} rel err {rel(y, yo):.2e} ({dt*1e3:.0f} ms)")

if "perf" in which:
    import ctypes as C
    for B in (64, 512):
        Rm = synth.orientation_grid()[:B]
        t = time.time()
        wL, wR = em.getEMagLs2Filters(hL, hR, az, ze, r, maz, mze, 4, fs, 512, rotations=Rm, handle=h)
        dt = time.time() - t
        t = time.time()
        wL, wR = em.getEMagLs2Filters(hL, hR, az, ze, r, maz, mze, 4, fs, 512, rotations=Rm, handle=h)
        dt2 = time.time() - t
        print(f"perf B={B}: first {dt:.2f}s second {dt2:.2f}s -> {B/dt2:.1f} sets/s; launches so far {h.launches}")
print("done")

if "prof" in which:
    for B in (1, 512, 3600):
        Rm = synth.orientation_grid()[:B]
        wL, wR = em.getEMagLs2Filters(hL, hR, az, ze, r, maz, mze, 4, fs, 512, rotations=Rm, handle=h)
        h.profile(True); h.profile_read()
        l0 = h.launches
        t = time.time()
        wL, wR = em.getEMagLs2Filters(hL, hR, az, ze, r, maz, mze, 4, fs, 512, rotations=Rm, handle=h)
        dt = time.time() - t
        pr = h.profile_read(); h.profile(False)
        tot = sum(v["ms"] for v in pr.values())
        print(f"prof B={B}: wall {dt:.3f}s ({B/dt:.1f} sets/s), launches {h.launches-l0}, event-sum {tot:.0f} ms")
        for k, v in pr.items():
            if v["n"]:
                print(f"    {k:12s} {v['ms']:9.1f} ms  n={v['n']:5d}  avg {v['ms']/v['n']*1e3:9.1f} us")
