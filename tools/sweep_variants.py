#!/usr/bin/env python
"""Time the 3-D scorer variants (rpe_debug_set_score_variant) on config-#4 frames, one context, scorer alone
(CUDA events around the kernel: stage 'score_fast'). Votes of every variant are compared with variant 14's."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rgbd_pose_estimation_b200 as rpe  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--variants", type=str, default="14,20,21,22,23,24,25,26,27")
ap.add_argument("--frames", type=int, default=10)
ap.add_argument("--rounds", type=int, default=2)
ap.add_argument("--n", type=int, default=307200)
ap.add_argument("--hyp", type=int, default=1024)
args = ap.parse_args()

q, t = rpe.sim_pose(1000)
Q, P, _ = rpe.sim_3d_3d(1001, q, t, args.n, noise=0.1, outlier_ratio=0.5)
S = rpe.sample_table(1, args.n, 3, args.hyp)
ctx = rpe.Context(0)
ctx.enable_stage_timing(True)
ref_votes = None
res = {}
for rnd in range(args.rounds):
    for v in [int(x) for x in args.variants.split(",")]:
        rpe.lib.rpe_debug_set_score_variant(v)
        ms = []
        for i in range(args.frames):
            ctx.upload(xc=P, xw=Q)
            r = ctx.ransac("shinji", S, thr3d=0.25, confidence=0.9999, want_mask=False)
            st = ctx.last_stage_ms()
            ms.append(st.get("score_fast", st["score"]))
        votes = ctx.get_votes(args.hyp)
        if ref_votes is None:
            ref_votes = votes.copy()
        same = bool(np.array_equal(votes, ref_votes))
        res.setdefault(v, []).append((float(np.median(ms[2:])), float(np.min(ms[2:])), same, r["n_borderline"]))
rpe.lib.rpe_debug_set_score_variant(14)
for v, rr in res.items():
    print(json.dumps({"variant": v, "median_ms": [round(x[0], 5) for x in rr], "min_ms": [round(x[1], 5) for x in rr],
                      "votes_same": all(x[2] for x in rr), "borderline": rr[0][3]}))
