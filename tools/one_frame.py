#!/usr/bin/env python
"""A few blocking config-#4 frames (RANSAC + Kabsch + LM), for ncu captures of the small kernels of a frame."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rgbd_pose_estimation_b200 as rpe  # noqa: E402

n, H = 307200, 1024
q, t = rpe.sim_pose(1000)
Q, P, _ = rpe.sim_3d_3d(1001, q, t, n, noise=0.1, outlier_ratio=0.5)
S = rpe.sample_table(1, n, 3, H)
with rpe.Context(0) as ctx:
    for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
        ctx.upload(xc=P, xw=Q)
        r = ctx.ransac("shinji", S, thr3d=0.25, confidence=0.9999)
        ctx.refit("kabsch_inliers")
        g = ctx.refit("gn", max_iters=3)
    print(r["max_votes"], r["iter_final"], g["refit_ok"])
