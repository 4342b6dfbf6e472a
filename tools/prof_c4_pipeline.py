"""Throughput of the pipelined front-end with the C4 extras switched on one at a time (8 batches in flight)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'rgbd-pl-slam_b200'))
import numpy as np, torch
import plslam_b200 as pl
from plslam_b200.synth import synth_pair, synth_vocabulary_arrays
B, depth, steps = 256, 8, 16
base = [im for s in range(8) for im in synth_pair(s)]
imgs = torch.from_numpy(np.stack(base * (B // 16))).cuda()
voc = pl.ORBVocabulary.from_arrays(10, 6, *synth_vocabulary_arrays(10, 6, 0))
fe = pl.Frontend(1000, 1.2, 8, 20, 7, 40, depth=depth)
outs = [fe.alloc(B, device="cuda") for _ in range(depth)]
streams = [torch.cuda.Stream() for _ in range(depth)]
fvs = [voc.featvec_batch_device(o["descriptors"], o["kp_counts"], 4) for o in outs]
bows = [pl.bow_pairs_device(o["keypoints"], o["descriptors"], o["kp_counts"], fv) for o, fv in zip(outs, fvs)]
torch.cuda.synchronize()
def run(n, featvec, bow):
    main = torch.cuda.current_stream()
    for s in streams: s.wait_stream(main)
    for k in range(n):
        i = k % depth
        fe.process_device(imgs, outs[i], True, stream=streams[i])
        if featvec: voc.featvec_batch_device(outs[i]["descriptors"], outs[i]["kp_counts"], 4, out=fvs[i], stream=streams[i])
        if bow: pl.bow_pairs_device(outs[i]["keypoints"], outs[i]["descriptors"], outs[i]["kp_counts"], fvs[i], out=bows[i], stream=streams[i])
    for s in streams: main.wait_stream(s)
for featvec, bow in ((0, 0), (1, 0), (0, 1), (1, 1)):
    run(depth, featvec, bow); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(steps, featvec, bow); e1.record(); torch.cuda.synchronize()
    print("featvec=%d bow=%d: %.2f ms/step" % (featvec, bow, e0.elapsed_time(e1) / steps), flush=True)
