"""Generate tests/golden/hp_goldens.npz: exact-arithmetic filter rows W(k,:) of the ill-conditioned LS bins
(oracle/hp_oracle.py: exact pwGrid and Gram matrix in integers, mpmath eigendecomposition at 120 digits) for

  c1  BASELINE config 1 (em32, N = 4, 512 taps, 48 kHz), bins 1..15 (0-based; 47 Hz .. 703 Hz)
  c5  BASELINE config 5 shape at reduced length (64-mic Fibonacci sphere, N = 7, 96 kHz, 128 taps), bins 1..7

on the synthetic problems of tests/conftest.py / tests/test_gpu_design.py, next to the FP64 oracle's own rows
(LAPACK SVD route) and its error against the exact rows.  SURVEY.md 8-c item 2: the CUDA path must be no
further from the exact rows than twice the FP64 oracle's distance (tests/test_gpu_arbitration.py).

usage: PYTHONPATH=. python tests/golden/make_hp_goldens.py [c1] [c5]      (about 10 + 30 minutes of CPU)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle                      # noqa: E402
from oracle import hp_oracle as hp  # noqa: E402
from emagls_b200 import synth      # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "hp_goldens.npz")


def run(tag, hL, hR, az, ze, r, maz, mze, order, fs, length, simN, bins):
    wL, wR, sp = oracle.getEMagLs2Filters(hL, hR, az, ze, r, maz, mze, order, fs, length, return_spectra=True)
    params = dict(returnRawMicSigs=True, fs=fs, irLen=sp["nfft"], oversamplingFactor=1, simulateAliasing=True,
                  radialFilter="none", smaRadius=r, smaDesignAziZenRad=np.stack([maz, mze], 1), waveModel="planeWave",
                  arrayType="rigid", shDefinition="real", shFunction=oracle.getSH, C=343.0)
    smair, _ = oracle.getSMAIRMatrix(params)
    assert smair.shape[1] == (simN + 1) ** 2
    Yc = np.conj(oracle.getSH(simN, np.stack([az, ze], 1), "real")).T
    M = maz.size
    exact = np.zeros((len(bins), 2, M), complex)
    o64 = np.zeros((len(bins), 2, M), complex)
    err = np.zeros(len(bins))
    cond = np.zeros(len(bins))
    for i, k in enumerate(bins):
        assert k + 1 < sp["k_cut"], "LS bins only"
        t0 = time.time()
        T = np.stack([sp["HL"][k], sp["HR"][k]])
        exact[i], sv = hp.exact_ls_rows(smair[:, :, k], Yc, T)
        o64[i] = np.stack([sp["W_l"][k], sp["W_r"][k]])
        err[i] = hp.rel_err(o64[i], exact[i])
        cond[i] = sv[0] / sv[-1]
        print(f"{tag} bin {k}: cond {cond[i]:.2e} err_oracle64 {err[i]:.2e} ({time.time() - t0:.0f} s)", flush=True)
    return {f"{tag}_bins": np.array(bins), f"{tag}_exact": exact, f"{tag}_oracle64": o64, f"{tag}_err_oracle64": err,
            f"{tag}_cond": cond}


def main():
    which = sys.argv[1:] or ["c1", "c5"]
    data = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    g = synth.load_grids()
    az, ze = g["hrirGridAziRad"], g["hrirGridZenRad"]
    if "c1" in which:
        hL, hR = synth.synth_hrirs(az, ze)
        data.update(run("c1", hL, hR, az, ze, g["micRadius"], g["micGridAziRad"], g["micGridZenRad"], 4, g["fs"], 512, 19,
                        list(range(1, 16))))
        np.savez_compressed(OUT, **data)
    if "c5" in which:
        hL, hR = synth.synth_hrirs(az, ze, fs=96000.0, taps=128, delay=40)
        maz, mze = synth.fibonacci_sphere(64)
        data.update(run("c5", hL, hR, az, ze, 0.042, maz, mze, 7, 96000.0, 128, 37, list(range(1, 8))))
        np.savez_compressed(OUT, **data)


if __name__ == "__main__":
    main()
