import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emagls_b200 as em
from emagls_b200 import synth
g = synth.load_grids()
az, ze = g["hrirGridAziRad"], g["hrirGridZenRad"]
hL, hR = synth.synth_hrirs(az, ze)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
Rm = synth.orientation_grid()[:B]
h = em.Handle(0)
wL, wR = em.getEMagLs2Filters(hL, hR, az, ze, g["micRadius"], g["micGridAziRad"], g["micGridZenRad"], 4, g["fs"], 512, rotations=Rm, handle=h)
print("ok", wL.shape)
