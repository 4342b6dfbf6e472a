import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'rgbd-pl-slam_b200'))
import numpy as np
import plslam_b200 as pl
from plslam_b200.synth import synth_frame
from oracle import bindings as ob
img = synth_frame(0)
ex = pl.ORBextractor()
k,d = ex(img)
o = ob.OrbOracle(); ok,od = o.extract(img)
for l in range(8):
    g = ex.candidates(0,l); c = o.candidates(l)
    print(l, len(g), len(c), np.array_equal(g,c))
g = ex.candidates(0,0); print(g[:12]); print(o.candidates(0)[:12])
