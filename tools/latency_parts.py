#!/usr/bin/env python
"""Blocking frame (rpe_set_upload_overlap(4), page-locked arrays) with the pipeline cut short at different points."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rgbd_pose_estimation_b200 as rpe  # noqa: E402

n, H = 307200, 1024
q, t = rpe.sim_pose(1)
frames = []
for i in range(4):
    Q, P, _ = rpe.sim_3d_3d(2 + i, q, t, n, noise=0.1, outlier_ratio=0.5)
    hq, hp = rpe.pinned_empty((n, 3)), rpe.pinned_empty((n, 3))
    hq[:], hp[:] = Q, P
    frames.append((hq, hp))
tab = rpe.pinned_empty((H, 4), np.int32)
tab[:] = rpe.sample_table(1, n, 3, H)
mask = rpe.pinned_empty((2, n), np.int16)
chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
out = {}
with rpe.Context(0) as c:
    c.set_upload_overlap(chunks)
    for name in ("ransac_nomask", "ransac_mask", "ransac_mask_kabsch", "full"):
        ms = []
        for i in range(14):
            hq, hp = frames[i % 4]
            t0 = time.perf_counter()
            c.upload_async(xc=hp, xw=hq)
            c.ransac_async("shinji", tab, thr3d=0.25, confidence=0.9999, mask=None if name == "ransac_nomask" else mask)
            if name in ("ransac_mask_kabsch", "full"):
                c.refit_async("kabsch_inliers")
            if name == "full":
                c.refit_async("gn", max_iters=3)
            c.sync()
            ms.append((time.perf_counter() - t0) * 1e3)
        out[name] = round(float(np.median(ms[3:])), 4)
print(json.dumps({"chunks": chunks, "median_ms": out}))
