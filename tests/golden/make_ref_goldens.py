"""Pack the reference's own golden outputs (resources/*.mat) into small fixtures.

Run in the build container only (needs /root/reference):
    python tests/golden/make_ref_goldens.py
Writes
    tests/golden/ref_goldens.npz        -- the woDC + LS golden filters (real & complex basis)
    emagls_b200/data/grids.npz          -- 2702-direction HRIR grid + em32 layout (input data)
    tests/golden/atf_full.npz           -- glasses-on-HATS ATF set, float32 (input data; tests decimate it where needed)
The em32 layout is the two degree lists of verifyEMagLs.m:30-31 (r = 0.042 m, :29).
"""
import os
import numpy as np
import scipy.io as sio

REF = "/root/reference/resources"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def main():
    out = {}
    pre = "HRIR_L2702_512samples_32channels_sh4_"
    for basis in ("real", "complex"):
        for meth, keys in (("LS", ("wLsL", "wLsR")), ("MagLS_woDC", ("wMlsL", "wMlsR")),
                           ("eMagLS_woDC", ("wEMlsL", "wEMlsR")),
                           ("eMagLS2_woDC", ("wEMls2L", "wEMls2R"))):
            d = sio.loadmat(os.path.join(REF, f"{pre}{basis}_{meth}.mat"))
            for k in keys:
                out[f"{basis}_{meth}_{k}"] = d[k]
    d = sio.loadmat(os.path.join(REF, f"{pre}real_eMagLS2_woDC.mat"))
    grid = dict(hrirGridAziRad=d["hrirGridAziRad"].ravel(), hrirGridZenRad=d["hrirGridZenRad"].ravel(),
                micGridAziRad=d["micGridAziRad"].ravel(), micGridZenRad=d["micGridZenRad"].ravel(),
                micRadius=float(d["micRadius"].ravel()[0]), fs=float(d["fs"].ravel()[0]))
    # cross-check the stored em32 layout against verifyEMagLs.m:30-31
    azi_deg = [0, 32, 0, 328, 0, 45, 69, 45, 0, 315, 291, 315, 91, 90, 90, 89, 180, 212, 180, 148,
               180, 225, 249, 225, 180, 135, 111, 135, 269, 270, 270, 271]
    zen_deg = [69, 90, 111, 90, 32, 55, 90, 125, 148, 125, 90, 55, 21, 58, 121, 159, 69, 90, 111,
               90, 32, 55, 90, 125, 148, 125, 90, 55, 21, 58, 122, 159]
    assert np.allclose(np.deg2rad(azi_deg), grid["micGridAziRad"])
    assert np.allclose(np.deg2rad(zen_deg), grid["micGridZenRad"])
    np.savez_compressed(os.path.join(HERE, "ref_goldens.npz"), **out, **grid)
    np.savez_compressed(os.path.join(ROOT, "emagls_b200", "data", "grids.npz"), **grid)

    a = sio.loadmat(os.path.join(REF, "glasses_on_HATS_ATFs_sphere.mat"))
    # the whole measured set (1625 directions x 8 microphones x 192 taps, BASELINE config 3), rounded to float32:
    # it is INPUT data of the parity tests (the oracle and the CUDA path get the same rounded array)
    np.savez_compressed(os.path.join(HERE, "atf_full.npz"),
                        atfIrs=a["atfIrs"].astype(np.float32),
                        atfGridAziEleDeg=a["atfGridAziEleDeg"].astype(np.int16),
                        fs=float(a["fs"].ravel()[0]))
    for f in ("ref_goldens.npz", "atf_full.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
