"""World-size-2 gloo test of the sharding + gather plumbing used for N > 1 (CPU)."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

from emagls_b200 import dist as emdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 450, 3600, 100001):
        for world in (1, 2, 3, 8):
            spans = [emdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


WORKER = textwrap.dedent("""
    import os, sys, torch
    sys.path.insert(0, %r)
    from emagls_b200 import dist as emdist
    rank, local_rank, world = emdist.init("gloo")
    n_total = 7                      # ragged: 4 + 3
    lo, hi = emdist.shard_range(n_total, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float64).reshape(-1, 1, 1) * torch.ones(1, 3, 2, dtype=torch.float64)
    out = emdist.gather_banks(local, n_total, dst=0)
    # gather straight into a preallocated bank; rank 0's own shard is a view of it (no staging copies)
    total = torch.full((n_total, 3, 2), -1.0, dtype=torch.float64) if rank == 0 else None
    mine = local
    if rank == 0:
        total[lo:hi] = local
        mine = total[lo:hi]
    out2 = emdist.gather_into(mine, total, n_total, dst=0)
    m = emdist.max_over_ranks(float(rank + 1))
    assert m == float(world)
    emdist.barrier()
    if rank == 0:
        assert out.shape == (n_total, 3, 2)
        assert torch.equal(out[:, 0, 0], torch.arange(n_total, dtype=torch.float64))
        assert out2 is total and torch.equal(out2, out)
        print("GATHER_OK")
    else:
        assert out is None and out2 is None
""")


def test_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0]


def test_render_time_block_sharding_with_halo_reproduces_the_whole_render():
    """Shard the render by time blocks with a (taps - 1)-frame input halo: concatenating the
    blocks equals the unsharded oracle render, for ragged splits and blocks shorter than the filter."""
    import numpy as np
    import oracle
    rng = np.random.default_rng(3)
    for n, taps, world in ((1000, 64, 3), (517, 128, 8), (90, 64, 4), (5, 8, 8)):
        x = rng.standard_normal((n, 4))
        wl, wr = rng.standard_normal((taps, 4)), rng.standard_normal((taps, 4))
        ref = oracle.binauralDecode(x, 48000, wl, wr, 48000)
        parts = []
        for r in range(world):
            lo, hi, halo = emdist.render_shard(n, taps, r, world)
            assert halo == min(taps - 1, lo) and 0 <= lo <= hi <= n
            if hi > lo:
                parts.append(oracle.binauralDecode(x[lo - halo:hi], 48000, wl, wr, 48000)[halo:])
        got = np.concatenate(parts, 0)
        assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
