#!/usr/bin/env python
"""One-off: a 4 000 003-correspondence frame (48 MB per array, ragged tail, > 2^31 bytes of evaluations) against the
oracle — index arithmetic at sizes far beyond the benchmark frame."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rgbd_pose_estimation_b200 as rpe  # noqa: E402
import orc  # noqa: E402

orc.set_math_mode(orc.DET)
n, H = 4000003, 96
q, t = rpe.sim_pose(1)
for name, method in (("shinji", 0), ("nl_shinji_kneip", 5)):
    d = rpe.sim_2d_3d_nl(2, q, t, n)
    arrs = {k: d[k] for k in ("bv", "xc", "nc", "xw", "nw")} if method else {"xc": d["xc"], "xw": d["xw"]}
    S = rpe.sample_table(3, n, 4 if method else 3, H)
    th = dict(thr3d=0.2, cos_thr=float(np.cos(np.arctan(np.float32(8.0) / np.float32(585.0)))), cos_nl=float(np.cos(np.float32(0.1))))
    ref = orc.ransac(method, S, confidence=0.99, full=True, nthreads=16, **th, **arrs)
    with rpe.Context(0) as ctx:
        ctx.upload(**arrs)
        got = ctx.ransac(name, S, thr3d=th["thr3d"], cos_thr2d=th["cos_thr"], cos_thrN=th["cos_nl"], confidence=0.99)
        slots = H * rpe.method_slots(method)
        ok = (np.array_equal(ctx.get_votes(slots), ref["votes"]) and np.array_equal(got["mask"], ref["mask"]) and
              (got["winner"], got["max_votes"], got["iter_final"]) == (ref["winner"], ref["max_votes"], ref["iter_final"]))
        print(name, "n", n, "ok", ok, "max_votes", got["max_votes"], "borderline", got["n_borderline"], "flags", got["flags"])
        assert ok
