"""Host logic of bench.py's end-to-end leg: how the steps of a run are cut into plslam_frontend_submit_host_wave calls."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench


def test_auto_ramp_only_for_single_pass_runs():
    assert bench.wave_ramp_sizes("auto", 20, 20) == [4, 16]      # the driver's --steps 20: slots capped at 20
    assert bench.wave_ramp_sizes("auto", 64, 32) == []           # longer runs keep equal waves (the phase)
    assert bench.wave_ramp_sizes("auto", 3, 3) == []             # too few slots to split
    assert bench.wave_ramp_sizes("2,4,14", 64, 32) == [2, 4, 14]
    assert bench.wave_ramp_sizes("", 20, 20) == []


def test_wave_plan_covers_every_step_once():
    for n, depth, wave, ramp in [(20, 20, 10, [4, 16]), (64, 32, 16, []), (40, 20, 10, [4, 16]), (5, 20, 10, [4, 16]),
                                 (7, 3, 1, []), (33, 32, 16, [40]), (1, 1, 1, [])]:
        plan = bench.wave_plan(n, depth, wave, ramp)
        assert sum(plan) == n and all(1 <= m <= depth for m in plan), (n, depth, wave, ramp, plan)
    assert bench.wave_plan(20, 20, 10, [4, 16]) == [4, 16]
    assert bench.wave_plan(64, 32, 16, []) == [16, 16, 16, 16]
    assert bench.wave_plan(40, 20, 10, [4, 16]) == [4, 16, 10, 10]   # warm-up of two passes: the ramp, then equal waves
