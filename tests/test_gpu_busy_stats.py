"""GPU: rpe_scorer_busy_stats — device time of the scorer launches inside a pipelined region as the union of their CUDA-event
intervals. With one context (nothing overlaps) it must agree with the plain per-launch event pairs; with several contexts it
must stay below the wall time of the region and above what the launches cost alone."""
import ctypes as C
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _busy(rpe, reset):
    s, c = C.c_double(0), C.c_longlong(0)
    assert rpe.lib.rpe_scorer_busy_stats(0, C.byref(s), C.byref(c), 1 if reset else 0) == 0
    return s.value, c.value


def test_union_of_launch_intervals(rpe):
    n, H = 307200, 1024
    q, t = rpe.sim_pose(11)
    Q, P, _ = rpe.sim_3d_3d(12, q, t, n, noise=0.1, outlier_ratio=0.5)
    S = rpe.sample_table(1, n, 3, H)
    with rpe.Context(0) as a:
        a.enable_stage_timing(2)
        a.upload(xc=P, xw=Q)
        a.ransac("shinji", S, thr3d=0.25, confidence=0.9999, want_mask=False)  # warm-up
        _busy(rpe, True)
        sm, ct = C.c_double(0), C.c_longlong(0)
        rpe.lib.rpe_scorer_time_stats(a._h, C.byref(sm), C.byref(ct), 1)
        for _ in range(8):
            a.ransac("shinji", S, thr3d=0.25, confidence=0.9999, want_mask=False)
        rpe.lib.rpe_scorer_time_stats(a._h, C.byref(sm), C.byref(ct), 1)
        busy, cnt = _busy(rpe, True)
        assert cnt == 8 and ct.value == 8
        # one context, blocking calls: intervals are disjoint, the union is the sum of the launches' own event pairs
        # (the clock's events sit inside the per-launch pair: a few microseconds less per launch)
        assert 0.8 * sm.value < busy <= sm.value * 1.001
        assert 0.15 < busy / cnt < 0.30
    # several contexts, frames in flight together: the union is at most the wall time and at least one launch per frame's worth
    ctxs = [rpe.Context(0) for _ in range(4)]
    try:
        for c in ctxs:
            c.enable_stage_timing(2)
            c.upload(xc=P, xw=Q)
        _busy(rpe, True)
        t0 = time.perf_counter()
        for rep in range(6):
            for c in ctxs:
                c.ransac_async("shinji", S, thr3d=0.25, confidence=0.9999)
        for c in ctxs:
            c.sync()
        wall_ms = (time.perf_counter() - t0) * 1e3
        busy, cnt = _busy(rpe, True)
        assert cnt == 24
        assert busy <= wall_ms
        assert busy / cnt > 0.15
    finally:
        for c in ctxs:
            c.close()
    assert _busy(rpe, False) == (0.0, 0)
    assert rpe.lib.rpe_scorer_busy_stats(64, None, None, 0) != 0
