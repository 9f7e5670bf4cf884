"""Sweep the overlap-save geometry of the render (FFT size, L2 working set) on a device-resident
10-minute 32-channel signal.  usage (GPU box): python tools/gpu_render_sweep.py [seconds]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emagls_b200 as em  # noqa: E402

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
n, ch, ln = int(secs * 48000), 32, 512
dev = torch.device("cuda", 0)
h = em.Handle(0)
stream = torch.cuda.ExternalStream(h.stream, device=dev)
x = torch.randn((ch, n), dtype=torch.float64, device=dev)
y = torch.empty((2, n), dtype=torch.float64, device=dev)
rng = np.random.default_rng(0)
wl = torch.from_numpy(rng.standard_normal((ch, ln))).to(dev)
wr = torch.from_numpy(rng.standard_normal((ch, ln))).to(dev)
yref = None


def run(reps=3):
    def step():
        h.check(h.lib.emagls_binaural_decode_dev(h.ptr, x.data_ptr(), n, ch, wl.data_ptr(), wr.data_ptr(), ln, 0,
                                                 y.data_ptr()))
    step(); step()
    torch.cuda.synchronize()
    h.profile(True); h.profile_read()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        step()
    b.record(stream)
    torch.cuda.synchronize()
    p = h.profile_read(); h.profile(False)
    return a.elapsed_time(b) / reps, {k: round(v["ms"] / reps, 3) for k, v in p.items() if v["n"]}


for N in (1024, 2048, 4096, 8192):
    for ws, direct in ((0, 1), (0, 0), (2048, 1), (512, 1)):
        os.environ["EMAGLS_RENDER_FFT"] = str(N)
        os.environ["EMAGLS_RENDER_WS_MB"] = str(ws)
        os.environ["EMAGLS_RENDER_DIRECT"] = str(direct)
        ms, cls = run()
        if yref is None:
            yref = y.clone()
            err = 0.0
        else:
            err = float((y - yref).abs().max() / yref.abs().max())
        print(f"N={N:6d} direct={direct} ws={ws:7d} MB  {ms:8.3f} ms  {n / ms / 1e3:9.1f} Msamples/s  dev-vs-first {err:.1e}  {cls}", flush=True)
