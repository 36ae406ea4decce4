#!/usr/bin/env python
"""Wall time of the binary64 path (rpe_upload_f64 + rpe_ransac_f64) at config-#3 size and on a dense frame."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rgbd_pose_estimation_b200 as rpe  # noqa: E402

F = 585.0
th = dict(thr3d=0.2, cos_thr2d=float(np.cos(np.arctan(8.0 / F))), cos_thrN=float(np.cos(0.1)))
ctx = rpe.Context(0)
for n, H in [(50000, 1024), (307200, 1024)]:
    q, t = rpe.sim_pose(1)
    d = rpe.sim_2d_3d_nl(2, q, t, n, n2d=1.0, or2d=0.3, n3d=0.05, or3d=0.3, nnl=float(np.deg2rad(2.0)), ornl=0.3)
    arrs = {}
    for k in ("bv", "xc", "nc", "xw", "nw"):
        a = d[k].astype(np.float64)
        if k in ("bv", "nc", "nw"):
            a = a / np.linalg.norm(a, axis=1, keepdims=True)
        arrs[k] = a
    for name in ("shinji", "kneip", "nl_shinji_kneip"):
        m = rpe.METHODS[name]
        S = rpe.sample_table(1, n, rpe.method_sample_size(m), H)
        ctx.upload_f64(**arrs)
        ts = []
        for i in range(4):
            ctx.sync()
            t0 = time.perf_counter()
            r = ctx.ransac_f64(name, S, confidence=0.99, want_mask=False, **th)
            ts.append((time.perf_counter() - t0) * 1e3)
        slots = H * rpe.method_slots(m)
        print(f"n={n} {name}: {min(ts[1:]):.3f} ms per rpe_ransac_f64 ({slots} slots, {slots * n / (min(ts[1:]) * 1e-3) / 1e9:.1f} G slot-corr/s), "
              f"max_votes {r['max_votes']} iter {r['iter_final']}")
