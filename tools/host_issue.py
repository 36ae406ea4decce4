#!/usr/bin/env python
"""Host-side cost of enqueueing one config-#4 frame (no GPU wait): wall time of the issuing loop for a burst that fits
the result queues, stage timing on/off."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rgbd_pose_estimation_b200 as rpe  # noqa: E402

N, H, NCTX, BURST = 307200, 1024, int(os.environ.get("NCTX", "4")), 64
dev = torch.device("cuda:0")
q, t = rpe.sim_pose(1)
Q, P, _ = rpe.sim_3d_3d(2, q, t, N, noise=0.1, outlier_ratio=0.5)
dQ, dP = torch.from_numpy(Q).to(dev), torch.from_numpy(P).to(dev)
dS = torch.from_numpy(rpe.sample_table(1, N, 3, H)).to(dev)
streams = [torch.cuda.Stream(device=dev) for _ in range(NCTX)]
ctxs = [rpe.Context(0, stream=s.cuda_stream) for s in streams]


hQ, hP = rpe.pinned_empty((N, 3), np.float32), rpe.pinned_empty((N, 3), np.float32)
hQ[:], hP[:] = Q, P
hS = rpe.pinned_empty((H, 4), np.int32)
hS[:] = rpe.sample_table(1, N, 3, H)
hM = [rpe.pinned_empty((2, N), np.int16) for _ in ctxs]
HOST = len(sys.argv) > 1 and sys.argv[1] == "host"


def frame(c):
    if HOST:
        c.upload_async(xc=hP, xw=hQ)
        c.ransac_async("shinji", hS, thr3d=0.25, confidence=0.9999, mask=hM[ctxs.index(c)])
    else:
        c.upload_device(N, xc=dP.data_ptr(), xw=dQ.data_ptr())
        c.ransac_async("shinji", dS.data_ptr(), H=H, thr3d=0.25, confidence=0.9999)
    c.refit_async("kabsch_inliers")
    c.refit_async("gn", max_iters=3)


for timing in (True, False):
    for c in ctxs:
        c.enable_stage_timing(timing)
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(BURST):
            frame(ctxs[k % NCTX])
        t1 = time.perf_counter()
        for c in ctxs:
            c.sync()
            c._keep = []
        t2 = time.perf_counter()
        print(f"timing={timing} issue {1e3 * (t1 - t0) / BURST:.4f} ms/frame, drained after {1e3 * (t2 - t0) / BURST:.4f} ms/frame")
