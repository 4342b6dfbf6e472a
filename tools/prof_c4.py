"""Device times of the C4 extras (ComputeBoW batch + SearchByBoW pairs) on a 256-frame batch."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'rgbd-pl-slam_b200'))
import numpy as np, torch
import plslam_b200 as pl
from plslam_b200.synth import synth_pair, synth_vocabulary_arrays
B = 256
base = [im for s in range(8) for im in synth_pair(s)]
imgs = torch.from_numpy(np.stack(base * (B // 16))).cuda()
voc = pl.ORBVocabulary.from_arrays(10, 6, *synth_vocabulary_arrays(10, 6, 0))
ex = pl.ORBextractor()
d_kps, d_desc, d_cnt = ex.extract_batch_device(imgs)
fv = voc.featvec_batch_device(d_desc, d_cnt, 4)
res = pl.bow_pairs_device(d_kps, d_desc, d_cnt, fv)
torch.cuda.synchronize()
def timeit(fn, n=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("featvec (descent + CSR): %.3f ms" % timeit(lambda: voc.featvec_batch_device(d_desc, d_cnt, 4, out=fv)))
print("SearchByBoW pairs:       %.3f ms" % timeit(lambda: pl.bow_pairs_device(d_kps, d_desc, d_cnt, fv, out=res)))
print("matches per pair:", res["nmatches"][:8].tolist(), "nodes per frame:", fv["fv_count"][:4].tolist())
