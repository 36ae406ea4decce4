#!/usr/bin/env python
"""Host-side timeline of one blocking frame (page-locked host arrays, rpe_set_upload_overlap(k)): when each C-ABI call returns."""
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
with rpe.Context(0) as c:
    c.set_upload_overlap(chunks)
    rows = []
    for i in range(14):
        hq, hp = frames[i % 4]
        t0 = time.perf_counter()
        c.upload_async(xc=hp, xw=hq)
        t1 = time.perf_counter()
        r = c.ransac_async("shinji", tab, thr3d=0.25, confidence=0.9999, mask=mask)
        t2 = time.perf_counter()
        c.refit_async("kabsch_inliers")
        t3 = time.perf_counter()
        c.refit_async("gn", max_iters=3)
        t4 = time.perf_counter()
        c.sync()
        t5 = time.perf_counter()
        rows.append([(x - t0) * 1e6 for x in (t1, t2, t3, t4, t5)])
    med = np.median(np.array(rows[3:]), axis=0)
    print(json.dumps({"chunks": chunks, "us_after_call": dict(zip(["upload_async", "ransac_async", "refit_kabsch", "refit_gn", "sync"], [round(float(x), 1) for x in med]))}))
