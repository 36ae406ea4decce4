#!/usr/bin/env python
"""Stage times and evaluation rates of every estimator family on one dense frame (N = 307 200, 1 024 iterations)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rgbd_pose_estimation_b200 as rpe  # noqa: E402

N, H, F = 307200, 1024, 585.0
OR = float(os.environ.get("OUTLIER", "0.3"))
q, t = rpe.sim_pose(1000)
d = rpe.sim_2d_3d_nl(1001, q, t, N, n2d=1.0, or2d=OR, n3d=0.05, or3d=OR, nnl=float(np.deg2rad(2.0)), ornl=OR)
arrs = {k: d[k] for k in ("bv", "xc", "nc", "xw", "nw")}
th = dict(thr3d=0.2, cos_thr2d=float(np.cos(np.arctan(np.float32(8.0) / np.float32(F)))), cos_thrN=float(np.cos(np.float32(0.1))))
FLOP = {"shinji": 26, "kneip": 30, "shinji_kneip": 38, "nl_kneip": 50, "nl_shinji": 46, "nl_shinji_kneip": 58}
ctx = rpe.Context(0)
ctx.enable_stage_timing(1)
out = {}
ONLY = [x for x in os.environ.get("METHODS", "").split(",") if x]   # e.g. METHODS=kneip,nl_shinji_kneip
for name, m in rpe.METHODS.items():
    if ONLY and name not in ONLY:
        continue
    S = rpe.sample_table(1, N, rpe.method_sample_size(m), H)
    ms = []
    for i in range(5):
        ctx.upload(**arrs)
        r = ctx.ransac(name, S, confidence=0.99, want_mask=False, **th)
        ms.append(ctx.last_stage_ms())
    st = {k: float(np.median([x[k] for x in ms[1:]])) for k in ms[0]}
    slots = r["n_slots"]
    evals = slots * N
    out[name] = {"slots": slots, "stages_ms": {k: round(v, 4) for k, v in st.items()},
                 "score_ms": st["score"], "score_fast_ms": st["score_fast"], "total_ms": st["total"],
                 "G_evals_per_s": evals / (st["score_fast"] * 1e-3) / 1e9 if st["score_fast"] > 0 else None,
                 "n_borderline": r["n_borderline"], "flags": r["flags"], "max_votes": r["max_votes"]}
    print(name, json.dumps(out[name]))
