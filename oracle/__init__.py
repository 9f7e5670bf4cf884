"""CPU oracle for the eMagLS hot path (TEST INFRASTRUCTURE ONLY).

This package is a FP64 NumPy/SciPy restatement of the reference MATLAB
algorithm (thomasdeppisch/eMagLS @ d204b49).  It is the *checker* for the CUDA
path; it is never imported by the product package ``emagls_b200``.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.

Parity status: the reference is pure MATLAB, which cannot run in this image
(no MATLAB, no Octave).  The restatement is pinned against the reference's own
golden ``.mat`` outputs as far as they can pin anything without the HRIR set
the reference downloads at run time (see tests/test_oracle_goldens.py and
DESIGN.md "Oracle pinning"): window/shift conventions exactly, MagLS LS-bins
to 7e-6, eMagLS/eMagLS2 conventions to a few percent.  The end-to-end hot loop
itself is therefore "parity unpinned" by any runnable in-tree reference test.
"""
from .emagls_oracle import *  # noqa: F401,F403
from .frontend_oracle import *  # noqa: F401,F403,E402  
