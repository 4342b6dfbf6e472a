import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'rgbd-pl-slam_b200'))
import numpy as np
import plslam_b200 as pl
from plslam_b200.synth import synth_frame
img = synth_frame(0)
ex = pl.ORBextractor()
k,d = ex(img)
print(len(k))
