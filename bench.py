#!/usr/bin/env python
"""Benchmark of the eMagLS hot path on B200: eMagLS2 filter sets / s (em32, N=4, 512 taps).

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --gpus N ...          # the reference algorithm on host cores

A "step" designs one batch of head-orientation filter sets (BASELINE config 2: the 3600-orientation
head-tracking grid, one HRTF set per rank -> weak scaling, no data-path collective).  Rank 0 prints
ONE JSON line.  `value` is device-resident throughput (inputs already in HBM, CUDA events on the
library's stream, max over ranks); `e2e` goes through the public host API (pinned host buffers,
H2D + D2H inside the timed region, plus the NCCL gather of the banks onto rank 0 for N > 1).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "emagls2_filter_sets_per_sec"
UNIT = "filter sets/s"
ORDER, LEN = 4, 512
# direct formulation, per filter set (512 bins): 22.3 GFLOP
ALGO_GFLOP_PER_SET = 22.3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--orient", type=int, default=3600, help="orientations per rank per step")
    ap.add_argument("--render-seconds", type=float, default=600.0,
                    help="length of the rendered 32-channel signal (SURVEY.md 8-d: 10 min)")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def problem(rank: int, n_orient: int):
    from emagls_b200 import synth
    g = synth.load_grids()
    az, ze = g["hrirGridAziRad"], g["hrirGridZenRad"]
    # one HRTF set per rank (seed = 20261017 + set index, randomised head)
    rng = np.random.default_rng(20261017 + rank)
    a = 0.0875 if rank == 0 else float(rng.uniform(0.075, 0.10))
    ear = 90.0 if rank == 0 else float(rng.uniform(85.0, 100.0))
    hL, hR = synth.synth_hrirs(az, ze, head_radius=a, ear_azi_deg=ear, seed=20261017 + rank)
    R = synth.orientation_grid()
    if n_orient <= R.shape[0]:
        R = R[:: max(1, R.shape[0] // n_orient)][:n_orient]
    else:
        R = np.concatenate([R] * (n_orient // R.shape[0] + 1))[:n_orient]
    return dict(g=g, az=az, ze=ze, hL=hL, hR=hR, R=np.ascontiguousarray(R), r=g["micRadius"],
                maz=g["micGridAziRad"], mze=g["micGridZenRad"], fs=g["fs"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_port_round(pr, first, workers):
    """One round of the NumPy restatement on the host cores: `workers` filter sets (orientations
    first..first+workers-1) designed concurrently, one thread each with single-threaded BLAS/LAPACK
    inside (the per-bin 2702 x 32 SVDs do not scale inside one call; NumPy releases the GIL in
    BLAS/LAPACK).  Overrides the OMP_NUM_THREADS=1 torchrun exports.  Returns seconds."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    from threadpoolctl import threadpool_limits
    from emagls_b200 import synth

    def one(i):
        raz, rze = synth.rotate_grid(pr["az"], pr["ze"], pr["R"][i % len(pr["R"])])
        oracle.getEMagLs2Filters(pr["hL"], pr["hR"], raz, rze, pr["r"], pr["maz"], pr["mze"], ORDER, pr["fs"], LEN)

    t0 = time.perf_counter()
    with threadpool_limits(limits=1):
        with ThreadPoolExecutor(max_workers=workers) as ex:
            list(ex.map(one, range(first, first + workers)))
    return time.perf_counter() - t0


def cpu_workers():
    return max(1, min(os.cpu_count() or 1, 32))


def run_reference(args, rank, world):
    """The reference's own algorithm on the box's host cores.  The reference is MATLAB, which this
    image cannot run, so this is the NumPy restatement (oracle/), labelled kind = "port"."""
    if rank != 0:
        return
    pr = problem(0, args.orient)
    workers = cpu_workers()
    sample = (f"{workers} filter sets (orientations) of the 3600-orientation batch per step, one host thread each "
              f"({os.cpu_count()} logical cores)")
    for i in range(args.warmup):
        cpu_port_round(pr, i * workers, workers)
    dt = 0.0
    for i in range(args.steps):
        dt += cpu_port_round(pr, (args.warmup + i) * workers, workers)
    v = args.steps * workers / dt
    ncores = workers
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "eMagLS2 em32 order 4, 512 taps, 2702-dir grid; bounded sample of the "
                                   "3600-orientation batch", "sample_sets_per_step": workers},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": ncores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import ctypes as C
    import torch
    import emagls_b200 as em
    from emagls_b200 import dist as emdist

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    emdist.init("nccl")
    dev = torch.device("cuda", local_rank)
    h = em.Handle(local_rank)
    cfg = h.default_config()
    stream = torch.cuda.ExternalStream(h.stream, device=dev)
    pr = problem(rank, args.orient)
    B, M, D, T = args.orient, pr["maz"].size, pr["az"].size, pr["hL"].shape[0]

    def dt64(x):
        return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to(dev)

    # ---------------- device-resident inputs ([T, D] column-major == [D, T] row-major)
    d_hL, d_hR = dt64(pr["hL"].T), dt64(pr["hR"].T)
    d_az, d_ze, d_maz, d_mze, d_R = dt64(pr["az"]), dt64(pr["ze"]), dt64(pr["maz"]), dt64(pr["mze"]), dt64(pr["R"])
    d_wL = torch.empty((B, M, LEN), dtype=torch.float64, device=dev)   # [len, M, B] column-major
    d_wR = torch.empty_like(d_wL)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_dev():
        rc = h.lib.emagls_design_emagls2_dev(
            h.ptr, C.byref(cfg), d_hL.data_ptr(), d_hR.data_ptr(), T, D, d_az.data_ptr(), d_ze.data_ptr(),
            float(pr["r"]), d_maz.data_ptr(), d_mze.data_ptr(), M, ORDER, float(pr["fs"]), LEN, 1, B,
            d_R.data_ptr(), d_wL.data_ptr(), d_wR.data_ptr(), None)
        h.check(rc)

    def l2_flush():
        with torch.cuda.stream(stream):
            flush.fill_(1)

    for _ in range(max(args.warmup, 3)):
        l2_flush()
        step_dev()
    torch.cuda.synchronize()
    h.profile(True)
    h.profile_read()
    launches0 = h.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    emdist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local_rank) as clk:
        for i in range(args.steps):
            l2_flush()
            ev[i][0].record(stream)
            step_dev()
            ev[i][1].record(stream)
        torch.cuda.synchronize()
    emdist.barrier()
    ms_local = sum(a.elapsed_time(b) for a, b in ev)
    ms = emdist.max_over_ranks(ms_local, dev)
    prof = h.profile_read()
    h.profile(False)
    launches = h.launches - launches0
    value = world * B * args.steps / (ms * 1e-3)

    # ---------------- end to end through the public API (pinned host buffers)
    def pinned(shape):
        t = torch.empty(shape[::-1], dtype=torch.float64, pin_memory=True)
        return t, t.numpy().T  # Fortran-ordered view of the pinned block

    p_hL, hLv = pinned((T, D))
    p_hR, hRv = pinned((T, D))
    hLv[...] = pr["hL"]
    hRv[...] = pr["hR"]
    p_wL, wLv = pinned((LEN, M, B))
    p_wR, wRv = pinned((LEN, M, B))
    h2d = 2 * T * D * 8 + 2 * D * 8 + 2 * M * 8 + B * 9 * 8
    d2h = 2 * LEN * M * B * 8
    bank_total = None

    def step_e2e():
        nonlocal bank_total
        em.getEMagLs2Filters(hLv, hRv, pr["az"], pr["ze"], pr["r"], pr["maz"], pr["mze"], ORDER, pr["fs"], LEN,
                             rotations=pr["R"], handle=h, out=(wLv, wRv))
        if world > 1:  # NCCL gather of the finished banks onto rank 0 (the only collective)
            bank = torch.stack([d_wL, d_wR], 1)          # device copy of the last device-resident banks
            bank_total = emdist.gather_banks(bank, world * B, dst=0)
            torch.cuda.synchronize()

    step_e2e()
    emdist.barrier()
    torch.cuda.synchronize()
    e2e_steps = max(1, min(args.steps, 2))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = emdist.max_over_ranks(time.perf_counter() - t0, dev)
    e2e_value = world * B * e2e_steps / e2e_s

    # ---------------- roofline of the dominant tensor kernel
    # The two direction-grid contractions run on the int8 tensor cores (tcgen05.mma kind::i8, FP64
    # accuracy through T = 6 error-free base-256 slices: T (T + 1) / 2 = 21 int8 GEMMs per product).  The roofline
    # counts the int8 operations the kernel EXECUTES against the int8 tensor peak; the FP64-equivalent
    # rate and the DGEMM (DMMA) peak of this GPU are reported next to it.
    S = (19 + 1) ** 2
    use_oz = os.environ.get("EMAGLS_GEMM", "") != "dmma"
    oz_T = int(os.environ.get("EMAGLS_OZAKI_SLICES", "6"))
    pairs = oz_T * (oz_T + 1) // 2
    KpS, KpD = (S + 31) // 32 * 32, (D + 31) // 32 * 32
    dgemm_peak = i8_peak = None
    i8_src = None
    if rank == 0:
        def best_ms(fn, n=5):
            fn()
            best = 1e9
            for _ in range(n):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            return best
        a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        dgemm_peak = 2 * 8192 ** 3 / (best_ms(lambda: torch.matmul(a, b)) * 1e-3) / 1e12
        del a, b
        try:   # library int8 GEMM (cuBLASLt) as the measured int8 tensor peak
            ai = torch.randint(-64, 64, (8192, 8192), dtype=torch.int8, device=dev)
            bi = torch.randint(-64, 64, (8192, 8192), dtype=torch.int8, device=dev).t()
            i8_peak = 2 * 8192 ** 3 / (best_ms(lambda: torch._int_mm(ai, bi)) * 1e-3) / 1e12
            i8_src = "live torch._int_mm (cuBLASLt int8) 8192^3, best of 5"
            del ai, bi
        except Exception as ex:  # noqa: BLE001
            i8_peak = None
            i8_src = f"torch._int_mm unavailable ({type(ex).__name__})"
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            alt = 2.0 * float(mp["bf16_tflops"])
            if i8_peak is None or i8_peak < alt:
                i8_peak, i8_src = alt, ("2 x MEASURED_PEAKS.json bf16_tflops (burst): tcgen05 kind::i8 runs at twice "
                                        "the bf16 rate; " + (i8_src or ""))
        except Exception:
            if i8_peak is None:
                i8_peak, i8_src = 4500.0, "nominal dense int8 (B200_PROFILING.md fallback)"
    fp64_flops = {"gemm_fwd": 2.0 * D * 4 * B * S, "gemm_bwd": 2.0 * 4 * B * S * D}
    i8_ops = {"gemm_fwd": 2.0 * D * 4 * B * KpS * pairs, "gemm_bwd": 2.0 * 4 * B * S * KpD * pairs}
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    shares = {k: round(v["ms"] / tot_ms, 4) for k, v in prof.items() if v["n"]}
    dom = max(("gemm_fwd", "gemm_bwd"), key=lambda k: prof[k]["ms"])
    avg_ms = prof[dom]["ms"] / max(prof[dom]["n"], 1)
    fp64_equiv = fp64_flops[dom] / (avg_ms * 1e-3) / 1e12
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_oz_traffic.json" if use_oz else "r01_gemm_traffic.json")))
        want = ("EpiPhase" if dom == "gemm_fwd" else "EpiStore")
        if B == 3600:
            traffic = next(x["traffic_bytes"] for x in tj["launches"] if want in x["kernel"])
            traffic_src = ("profiles/" + ("r01_oz_traffic.json" if use_oz else "r01_gemm_traffic.json") +
                           " (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full launch)")
    except Exception:
        pass
    # The largest single class of the step is the per-(orientation, bin) TSQR + Jacobi kernel of the clipped
    # bins (FP64 on the CUDA cores, neither an HBM nor a tensor-core roofline): reported beside the GEMM.
    # Algorithmic flops per problem: 8 (S M^2 - M^3 / 3) for the QR, ~1800 flops per rotated column pair
    # x M (M - 1) / 2 pairs x 8 sweeps (9 were measured with the earlier, stricter stop rule; the current rule
    # saves one on most bins), 8 M^3 for the projector.
    fac = None
    if prof.get("factor", {}).get("n"):
        fl = 8.0 * (S * M * M - M ** 3 / 3.0) + 8 * 1800.0 * M * (M - 1) / 2 + 8.0 * M ** 3
        fac_ms = prof["factor"]["ms"] / args.steps
        fac = {"kernel": "factor_kernel (TSQR + one-sided Jacobi, FP64 CUDA cores)", "ms_per_step": fac_ms,
               "share": shares.get("factor"), "algorithmic_mflop_per_problem": fl / 1e6,
               "note": "problems per step = clipped bins x orientations (em32 @48 kHz: 87 x orientations); "
                       "peak = the measured DGEMM figure (the FP64 pipe of this part)"}
        if M == 32 and abs(pr["fs"] - 48000.0) < 1:
            fac["achieved_tflops"] = fl * 87 * B / (fac_ms * 1e-3) / 1e12
            fac["frac_of_dgemm_peak"] = (fac["achieved_tflops"] / dgemm_peak) if dgemm_peak else None
    if use_oz:
        achieved = i8_ops[dom] / (avg_ms * 1e-3) / 1e12
        roofline = {"kernel": f"ozaki_gemm_kernel<{oz_T}> ({dom}: tcgen05.mma kind::i8, TMEM accumulators, TMA operands)",
                    "bound": "tensor", "achieved": achieved, "peak": i8_peak, "unit": "TOP/s (int8, executed)",
                    "frac": (achieved / i8_peak) if i8_peak else None, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": i8_src, "int8_ops_per_launch": i8_ops[dom], "slice_pairs": pairs,
                    "fp64_equivalent_tflops": fp64_equiv, "fp64_flops_per_launch": fp64_flops[dom],
                    "dgemm_peak_tflops": dgemm_peak, "fp64_equivalent_vs_dgemm_peak": (fp64_equiv / dgemm_peak) if dgemm_peak else None,
                    "dgemm_peak_source": "live cuBLAS DGEMM 8192^3, best of 5 (MEASURED_PEAKS.json has no FP64 entry)",
                    "avg_launch_ms": avg_ms, "class_time_share": shares,
                    "algorithmic_tflops_whole_job": value / world * ALGO_GFLOP_PER_SET / 1e3}
    else:
        roofline = {"kernel": f"gemm_f64_kernel ({dom}, DMMA.8x8x4)", "bound": "tensor", "achieved": fp64_equiv,
                    "peak": dgemm_peak, "unit": "TFLOP/s", "frac": (fp64_equiv / dgemm_peak) if dgemm_peak else None,
                    "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": "live cuBLAS DGEMM 8192^3, best of 5 (MEASURED_PEAKS.json has no FP64 entry)",
                    "flops_per_launch": fp64_flops[dom], "avg_launch_ms": avg_ms, "class_time_share": shares,
                    "algorithmic_tflops_whole_job": value / world * ALGO_GFLOP_PER_SET / 1e3}

    roofline["largest_class"] = fac

    # ---------------- render (secondary metric: Msamples/s of the 32 -> 2 channel, 512-tap FIR)
    render = None
    if not args.no_render and rank == 0:
        n = int(args.render_seconds * 48000)
        x = torch.randn((32, n), dtype=torch.float64, device=dev)           # [n, 32] column-major
        y = torch.empty((2, n), dtype=torch.float64, device=dev)
        wl, wr = d_wL[0].contiguous(), d_wR[0].contiguous()                 # [M, LEN] == [len, M] col-major

        def rstep():
            h.check(h.lib.emagls_binaural_decode_dev(h.ptr, x.data_ptr(), n, 32, wl.data_ptr(), wr.data_ptr(), LEN, 0,
                                                     y.data_ptr()))
        for _ in range(2):
            rstep()
        torch.cuda.synchronize()
        h.profile(True)
        h.profile_read()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record(stream)
        for _ in range(3):
            rstep()
        r1.record(stream)
        torch.cuda.synchronize()
        rp = h.profile_read()
        h.profile(False)
        rms = r0.elapsed_time(r1) / 3
        hbm = None
        try:
            hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
        except Exception:
            pass
        hbm = hbm or 6650.0
        algo_bytes = n * 272.0                                              # 272 B / frame
        render = {"metric": "render_msamples_per_sec", "value": n / (rms * 1e-3) / 1e6, "unit": "Msamples/s",
                  "frames": n, "channels": 32, "taps": LEN, "ms": rms,
                  "roofline": {"bound": "hbm", "achieved": algo_bytes / (rms * 1e-3) / 1e9, "peak": hbm,
                               "unit": "GB/s", "frac": algo_bytes / (rms * 1e-3) / 1e9 / hbm,
                               "class_ms": {k: v["ms"] / 3 for k, v in rp.items() if v["n"]}}}
        del x, y

    # ---------------- CPU baseline (the reference algorithm restated, on this box's host cores)
    cpu = None
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        workers = cpu_workers()
        dtc = cpu_port_round(pr, 0, workers)
        cpu = {"value": workers / dtc, "unit": UNIT, "cores": workers, "kind": "port",
               "sample": f"{workers} of the {B} filter sets of one step, designed concurrently with one host thread each "
                         f"(NumPy restatement of the MATLAB reference; {os.cpu_count()} logical cores)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "BASELINE config 2: eMagLS2, em32 (32 mics, r=4.2 cm), SH order 4, 512 taps, "
                                       "2702-direction HRIR grid @48 kHz, head-orientation batch",
                           "orientations_per_gpu_per_step": B, "hrtf_sets_per_gpu": 1, "parallelism": f"shard{world}",
                           "l2": "flushed between steps (256 MiB write); per-step working set is several GB"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "includes_nccl_gather": world > 1},
                "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roofline,
                "cpu_baseline": cpu, "render": render}
        print(json.dumps(line), flush=True)
    emdist.barrier()


if __name__ == "__main__":
    main()
