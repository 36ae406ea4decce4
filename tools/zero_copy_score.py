#!/usr/bin/env python
"""What does the 3-D scorer do when its bulk-TMA loads read the frame straight from page-locked HOST memory
(rpe_upload_device with the pinned arrays' addresses: under UVA they are valid device pointers)? Scorer alone, by variant."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rgbd_pose_estimation_b200 as rpe  # noqa: E402

n, H = 307200, 1024
q, t = rpe.sim_pose(1000)
Q, P, _ = rpe.sim_3d_3d(1001, q, t, n, noise=0.1, outlier_ratio=0.5)
hq, hp = rpe.pinned_empty((n, 3)), rpe.pinned_empty((n, 3))
hq[:], hp[:] = Q, P
S = rpe.sample_table(1, n, 3, H)
ctx = rpe.Context(0)
ctx.enable_stage_timing(True)
out = {}
for v in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "14,24,1,7,16").split(",")]:
    rpe.lib.rpe_debug_set_score_variant(v)
    row = {}
    for where in ("device", "host"):
        ms = []
        for i in range(8):
            if where == "device":
                ctx.upload(xc=P, xw=Q)
            else:
                ctx.upload_device(n, xc=hp.ctypes.data, xw=hq.ctypes.data)
            r = ctx.ransac("shinji", S, thr3d=0.25, confidence=0.9999, want_mask=False)
            st = ctx.last_stage_ms()
            ms.append((st["score_fast"], st["generate"], st["mask_refit"]))
        row[where] = {"score_fast_ms": round(float(np.median([m[0] for m in ms[2:]])), 4),
                      "generate_ms": round(float(np.median([m[1] for m in ms[2:]])), 4),
                      "mask_ms": round(float(np.median([m[2] for m in ms[2:]])), 4), "votes": r["max_votes"]}
    out[v] = row
rpe.lib.rpe_debug_set_score_variant(14)
print(json.dumps(out, indent=1))
