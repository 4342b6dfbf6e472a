"""Boundary fixtures for ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) (ORBmatcher.h:78, @0x80d00).

The reference binary computes the projection with a float division (vdivss @0x81c92), fused multiply-adds (vfmadd213ss
@0x81cba / @0x81cd9, vfnmadd132ss @0x81eb5) and cv::gemm's small-matrix path (float sums).  An evaluation in double
differs from it in the last bit of u, v or ur for a fraction of the points; the decision only differs when a current
key point sits within one ulp of the search window's edge.  This script builds exactly such cases: for points whose
u (or ur) differs between the two evaluations it plants a current key point with the point's own descriptor at a
position that is inside the window under one evaluation and outside under the other, runs the REFERENCE'S OWN CODE
(tests/golden/reference_code.py) on them and stores inputs and outputs in projection_boundary.npz.

    python tests/golden/make_projection_boundary.py        (needs /root/reference; run in the build container)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "rgbd-pl-slam_b200"), os.path.join(ROOT, "tests"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

f32 = np.float32


def fma32(a, b, c):
    return f32(np.float64(a) * np.float64(b) + np.float64(c))


def project_binary(Rc, X, cam):
    """u, v, invzc the way the binary computes them."""
    pc = []
    for r in range(3):
        p0, p1, p2 = f32(Rc[r, 0] * X[0]), f32(Rc[r, 1] * X[1]), f32(Rc[r, 2] * X[2])
        s = f32(f32(p0 + p1) + p2)
        pc.append(f32(np.float64(s) + np.float64(Rc[r, 3])))
    invz = f32(f32(1.0) / pc[2])
    return fma32(f32(pc[0] * cam[0]), invz, cam[2]), fma32(f32(pc[1] * cam[1]), invz, cam[3]), invz


def project_double(Rc, X, cam):
    """the same quantities with double accumulators, a double division and unfused operations (what round 1 computed)."""
    pc = []
    for r in range(3):
        s = sum(np.float64(Rc[r, k]) * np.float64(X[k]) for k in range(3))
        pc.append(f32(s + np.float64(Rc[r, 3])))
    invz = f32(1.0 / np.float64(pc[2]))
    return f32(f32(f32(cam[0] * pc[0]) * invz) + cam[2]), f32(f32(f32(cam[1] * pc[1]) * invz) + cam[3]), invz


def plant(u_a, u_b, radius):
    """x with |x - u_a| < radius and |x - u_b| >= radius (float32 arithmetic), or None."""
    for sign in (1.0, -1.0):
        x = f32(u_a + f32(sign) * radius)
        for _ in range(6):
            x = np.nextafter(x, f32(u_a), dtype=np.float32)
            ina = abs(f32(x - u_a)) < radius
            inb = abs(f32(x - u_b)) < radius
            if ina and not inb:
                return x
    return None


def build_case(oo, seed, motion, th, mono, rng):
    from matchdata import frame_grid, projection_case
    from plslam_b200.synth import synth_pair
    a, b = synth_pair(seed)
    (ka, da), (kb, db) = oo.extract(a), oo.extract(b)
    sf = oo.tables()["scale"]
    last, cur, cam, sf2, tc, tl = projection_case(ka, da, kb, db, sf, seed=seed, motion=motion)
    n1, n2 = len(last["desc"]), len(cur["desc"])
    free = list(rng.permutation(n2))
    planted = 0
    for i in range(n1):
        if not last["valid"][i] or not free:
            continue
        ua, va, ia = project_binary(tc, last["xyz"][i], cam)
        ub, vb, ib = project_double(tc, last["xyz"][i], cam)
        if not (ia >= 0 and cam[6] + 20 < ua < cam[7] - 20 and cam[8] + 20 < va < cam[9] - 20):
            continue
        radius = f32(f32(th) * sf2[last["octave"][i]])
        j = None
        if ua != ub:  # u differs in the last bit: plant on the window edge in x, alternately inside under either evaluation
            x = plant(ua, ub, radius) if planted % 2 == 0 else plant(ub, ua, radius)
            if x is not None:
                j = free.pop()
                cur["xy"][j] = (x, va)
                cur["uright"][j] = -1.0
        elif not mono:
            ura, urb = fma32(-ia, cam[4], ua), f32(ua - f32(cam[4] * ib))
            if ura != urb:  # ur differs: plant the stereo coordinate on the edge |ur - uR| > radius
                for sign in (1.0, -1.0):
                    r2 = f32(ura + f32(sign) * radius)
                    for _ in range(6):
                        ina, inb = not (abs(f32(ura - r2)) > radius), not (abs(f32(urb - r2)) > radius)
                        if ina != inb and r2 > 0:
                            break
                        r2 = np.nextafter(r2, f32(ura + f32(sign) * 4 * radius), dtype=np.float32)
                    else:
                        continue
                    j = free.pop()
                    cur["xy"][j] = (ua, va)
                    cur["uright"][j] = r2
                    break
        if j is None:
            continue
        cur["octave"][j] = last["octave"][i]
        cur["desc"][j] = last["desc"][i]
        cur["angle"][j] = last["angle"][i]
        cur["taken"][j] = 0
        planted += 1
    gs, gi, _ = frame_grid(cur["xy"], 640, 480)
    cur["grid_start"], cur["grid_items"] = gs, gi
    return last, cur, cam, sf2, tc, tl, planted


def main():
    from oracle import bindings as ob
    from reference_code import RefLibrary
    ob.build()
    oo = ob.OrbOracle()
    R = RefLibrary()
    out, k = {}, 0
    rng = np.random.default_rng(7)
    for seed in (3, 4):
        for motion in (0.02, 0.3, -0.3):
            for th, mono in ((7.0, 0), (15.0, 1)):
                last, cur, cam, sf, tc, tl, planted = build_case(oo, seed, motion, th, mono, rng)
                m, n = R.search_by_projection(last, cur, cam, sf, tc, tl, th, bool(mono), True)
                mo, no = ob.search_by_projection(last, cur, cam, sf, tc, tl, th, bool(mono), True)
                same = no == n and np.array_equal(mo, m)
                print("case %d seed %d motion %+.2f th %g mono %d: %d planted, reference %d matches, oracle %s" %
                      (k, seed, motion, th, mono, planted, n, "identical" if same else "DIFFERS"))
                assert same
                out["c%d_args" % k] = np.array([seed, motion, th, mono, planted], np.float64)
                for key, v in last.items():
                    out["c%d_last_%s" % (k, key)] = v
                for key, v in cur.items():
                    out["c%d_cur_%s" % (k, key)] = v
                out["c%d_cam" % k], out["c%d_sf" % k], out["c%d_tc" % k], out["c%d_tl" % k] = cam, sf, tc, tl
                out["c%d_match" % k], out["c%d_n" % k] = m, np.array(n)
                k += 1
    out["n"] = np.array(k)
    np.savez_compressed(os.path.join(HERE, "projection_boundary.npz"), **out)
    print("wrote projection_boundary.npz: %d cases" % k)


if __name__ == "__main__":
    main()
