"""Jacobi sweep / route statistics of the factor kernel (EMAGLS_DEBUG_INFO) on a small orientation batch."""
import os, sys
os.environ["EMAGLS_DEBUG_INFO"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emagls_b200 as em
from emagls_b200 import synth
g = synth.load_grids()
az, ze = g["hrirGridAziRad"], g["hrirGridZenRad"]
hL, hR = synth.synth_hrirs(az, ze)
Rm = synth.orientation_grid()[:: 3600 // 64][:64]
h = em.Handle(0)
em.getEMagLs2Filters(hL, hR, az, ze, g["micRadius"], g["micGridAziRad"], g["micGridZenRad"], 4, g["fs"], 512, rotations=Rm, handle=h)
