#!/usr/bin/env python
"""RPE_OVERLAP_TRACE=1 python tools/overlap_trace.py [chunks] — device timeline of blocking frames with rpe_set_upload_overlap."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rgbd_pose_estimation_b200 as rpe  # noqa: E402

chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n, H = 307200, 1024
q, t = rpe.sim_pose(1)
Q, P, _ = rpe.sim_3d_3d(2, q, t, n)
hq, hp = rpe.pinned_empty((n, 3)), rpe.pinned_empty((n, 3))
hq[:], hp[:] = Q, P
tab = rpe.pinned_empty((H, 4), np.int32)
tab[:] = rpe.sample_table(1, n, 3, H)
mask = rpe.pinned_empty((2, n), np.int16)
with rpe.Context(0) as c:
    c.set_upload_overlap(chunks)
    for i in range(6):
        t0 = time.perf_counter()
        c.upload_async(xc=hp, xw=hq)
        c.ransac_async("shinji", tab, thr3d=0.25, confidence=0.9999, mask=mask)
        c.refit_async("kabsch_inliers")
        c.refit_async("gn", max_iters=3)
        c.sync()
        print("frame", i, "host ms", round((time.perf_counter() - t0) * 1e3, 3))
