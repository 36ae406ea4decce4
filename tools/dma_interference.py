#!/usr/bin/env python
"""Does host-to-device DMA traffic slow the scorer down?  Scores device-resident config-#4 frames (scorer alone, CUDA
events around the launch) with and without a background stream of 7.4 MB page-locked uploads into an unrelated buffer."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rgbd_pose_estimation_b200 as rpe  # noqa: E402

n, H = int(os.environ.get("DMA_N", "307200")), 1024
q, t = rpe.sim_pose(1000)
Q, P, _ = rpe.sim_3d_3d(1001, q, t, n, noise=0.1, outlier_ratio=0.5)
S = rpe.sample_table(1, n, 3, H)
ctx = rpe.Context(0)
ctx.enable_stage_timing(True)
ctx.upload(xc=P, xw=Q)


def frames(k):
    ms = []
    for _ in range(k):
        ctx.ransac("shinji", S, thr3d=0.25, confidence=0.9999, want_mask=False)
        ms.append(ctx.last_stage_ms()["score_fast"])
    return float(np.median(ms[2:])), float(np.min(ms[2:])), float(np.max(ms[2:]))


out = {"quiet": frames(40)}
host = torch.empty(2 * n * 3, dtype=torch.float32).pin_memory()
dev = torch.empty_like(host, device="cuda")
side = torch.cuda.Stream()
for direction in ("h2d", "d2h"):
    with torch.cuda.stream(side):
        for _ in range(4000):
            if direction == "h2d":
                dev.copy_(host, non_blocking=True)
            else:
                host.copy_(dev, non_blocking=True)
    out["during_" + direction] = frames(40)
    out["copies_still_running_" + direction] = not side.query()
    side.synchronize()
out["quiet_again"] = frames(40)
print(json.dumps(out))
