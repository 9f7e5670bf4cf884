import sys, os, time, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle, emagls_b200 as em
rng = np.random.default_rng(3)
h = em.Handle(0)
ch, ln = 32, 512
wl = rng.standard_normal((ln, ch)); wr = rng.standard_normal((ln, ch))
for chunk, n in ((128, 512*300), (512, 512*512), (512, 512*600), (64, 512*200), (2048, 512*3000)):
    os.environ["EMAGLS_RENDER_CHUNK"] = str(chunk)
    x = rng.standard_normal((n, ch))
    y = em.binauralDecode(x, 48000, wl, wr, 48000, False, handle=h)
    t = time.time(); y = em.binauralDecode(x, 48000, wl, wr, 48000, False, handle=h); dt = time.time() - t
    yo = oracle.binauralDecode(x, 48000, wl, wr, 48000, False)
    e = np.abs(y - yo).max(1)
    bad = np.nonzero(e > 1e-9 * np.abs(yo).max())[0]
    print(f"chunk={chunk} n={n}: rel err {e.max()/np.abs(yo).max():.2e}; bad samples {bad.size} first {bad[:3]} last {bad[-3:]}  ({dt*1e3:.0f} ms e2e host)")
