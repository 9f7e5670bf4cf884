#!/usr/bin/env python
"""BASELINE config 5 (stress): synthetic 64-microphone Fibonacci-sphere array, SH order 7, 4096-tap filters at
96 kHz (NFFT_MAX_LEN lifted to 8192, K = 4097 bins: SURVEY.md H5), a batch of (HRTF set x orientation) filter sets
sharded over the ranks.  The full bank of the config (100k sets x 4.19 MB = 419 GB) does not fit one GPU, so the
banks are NOT gathered: every rank streams its shard chunk by chunk to pinned host memory (DESIGN.md
section 6) while the next chunk is designed.

    python tools/bench_c5.py [--sets-per-gpu 256] [--chunk 128] [--steps 1]
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_c5.py      (one rank per GPU)

Prints one JSON line (rank 0): filter sets/s over all ranks, D2H inside the timed region.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets-per-gpu", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=128, help="orientations per design call (one bank chunk)")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--len", type=int, default=4096)
    args = ap.parse_args()
    import torch
    import emagls_b200 as em
    from emagls_b200 import dist as emdist, synth
    rank, local_rank, world = emdist.init("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    h = em.Handle(local_rank)
    cfg = h.default_config()
    cfg.nfft_max_len = 2 * args.len
    g = synth.load_grids()
    az, ze = g["hrirGridAziRad"], g["hrirGridZenRad"]
    rng = np.random.default_rng(20261017 + rank)
    a = 0.0875 if rank == 0 else float(rng.uniform(0.075, 0.10))
    hL, hR = synth.synth_hrirs(az, ze, fs=96000.0, taps=256, delay=60, head_radius=a, seed=20261017 + rank)
    maz, mze = synth.fibonacci_sphere(64)
    Rg = synth.orientation_grid()
    n_local = args.sets_per_gpu
    R = np.ascontiguousarray(np.concatenate([Rg] * (n_local // len(Rg) + 1))[:n_local].reshape(-1, 9))
    M, LEN, T, D = 64, args.len, hL.shape[0], az.size
    stream = torch.cuda.ExternalStream(h.stream, device=dev)
    copy_stream = torch.cuda.Stream(device=dev)

    def dt64(x):
        return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to(dev)
    d_hL, d_hR, d_az, d_ze, d_maz, d_mze, d_R = (dt64(hL.T), dt64(hR.T), dt64(az), dt64(ze), dt64(maz), dt64(mze), dt64(R))
    chunk = min(args.chunk, n_local)
    banks = [[torch.empty((chunk, M, LEN), dtype=torch.float64, device=dev) for _ in range(2)] for _ in range(2)]
    host = [torch.empty((n_local, M, LEN), dtype=torch.float64).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def step():
        for ci, o0 in enumerate(range(0, n_local, chunk)):
            nb = min(chunk, n_local - o0)
            buf = ci & 1
            stream.wait_event(freed[buf])                    # the D2H of the chunk that used this buffer is over
            rc = h.lib.emagls_design_emagls2_dev(
                h.ptr, C.byref(cfg), d_hL.data_ptr(), d_hR.data_ptr(), T, D, d_az.data_ptr(), d_ze.data_ptr(), 0.042,
                d_maz.data_ptr(), d_mze.data_ptr(), M, 7, 96000.0, LEN, 1, nb, d_R[o0:].data_ptr(),
                banks[buf][0].data_ptr(), banks[buf][1].data_ptr(), None)
            h.check(rc)
            done[buf].record(stream)
            with torch.cuda.stream(copy_stream):             # stream the chunk out while the next one is designed
                copy_stream.wait_event(done[buf])
                for e in range(2):
                    host[e][o0:o0 + nb].copy_(banks[buf][e][:nb], non_blocking=True)
                freed[buf].record(copy_stream)
        copy_stream.synchronize()
        stream.synchronize()

    for f in freed:
        f.record(copy_stream)
    step()                                                   # warm-up (plans, memory pool)
    h.profile(True); h.profile_read(); h.stats_read()
    emdist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    dt = emdist.max_over_ranks(time.perf_counter() - t0, dev)
    prof = h.profile_read(); h.profile(False)
    st = h.stats_read()
    finite = bool(torch.isfinite(host[0]).all() and torch.isfinite(host[1]).all())
    zero_ends = bool((host[0][:, :, 0] == 0).all() and (host[0][:, :, -1] == 0).all())
    if rank == 0:
        tot = sum(v["ms"] for v in prof.values()) or 1.0
        line = {"metric": "emagls2_filter_sets_per_sec", "value": world * n_local * args.steps / dt, "unit": "filter sets/s",
                "n_gpus": world, "steps": args.steps, "ms_per_step": dt / args.steps * 1e3, "scaling": "weak", "dtype": "f64",
                "config": {"workload": "BASELINE config 5 (stress): 64-mic Fibonacci sphere r = 4.2 cm, SH order 7, 96 kHz, "
                                       f"{LEN} taps (nfft {2 * LEN}), 2702-direction HRIR grid, 256-tap synthetic HRIRs",
                           "sets_per_gpu_per_step": n_local, "orientations_per_design_call": chunk,
                           "banks": "left sharded; each rank streams its chunks to pinned host memory (D2H overlapped)"},
                "d2h_bytes_per_step": world * 2 * n_local * M * LEN * 8,
                "time_to_100k_sets_s": 100000.0 / (world * n_local * args.steps / dt),
                "class_time_share": {k: round(v["ms"] / tot, 4) for k, v in prof.items() if v["n"]},
                "stats": st, "checks": {"finite": finite, "zero_end_taps": zero_ends}}
        print(json.dumps(line), flush=True)
    emdist.barrier()


if __name__ == "__main__":
    main()
