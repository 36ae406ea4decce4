#!/usr/bin/env python
"""Run a few config-#4 frames through the C-ABI on one context (for ncu / stage timing)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rgbd_pose_estimation_b200 as rpe  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=6)
ap.add_argument("--n", type=int, default=307200)
ap.add_argument("--hyp", type=int, default=1024)
ap.add_argument("--packed", type=int, default=1)
ap.add_argument("--gn-iters", type=int, default=3)
ap.add_argument("--variant", type=int, default=14)
ap.add_argument("--nosync", type=int, default=0)
args = ap.parse_args()

rpe.lib.rpe_debug_set_packed(args.packed)
rpe.lib.rpe_debug_set_score_variant(args.variant)
rpe.lib.rpe_debug_set_nosync(args.nosync)
q, t = rpe.sim_pose(1000)
Q, P, _ = rpe.sim_3d_3d(1001, q, t, args.n, noise=0.1, outlier_ratio=0.5)
S = rpe.sample_table(1, args.n, 3, args.hyp)
ctx = rpe.Context(0)
ctx.enable_stage_timing(True)
out = []
for i in range(args.frames):
    ctx.upload(xc=P, xw=Q)
    r = ctx.ransac("shinji", S, thr3d=0.25, confidence=0.9999, want_mask=False)
    st = ctx.last_stage_ms()
    k = ctx.refit("kabsch_inliers")
    g = ctx.refit("gn", max_iters=args.gn_iters)
    st["gn"] = ctx.last_stage_ms()["gn"]
    out.append(st)
print(json.dumps({"stage_ms_last": out[-1], "stage_ms_median": {k: float(np.median([o[k] for o in out[1:]])) for k in out[0]},
                  "max_votes": r["max_votes"], "iter_final": r["iter_final"], "n_borderline": r["n_borderline"],
                  "gn_evals": g["refit_evals"], "gn_cost": g["refit_cost"]}))
