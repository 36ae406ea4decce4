#!/usr/bin/env python
"""How long does the generator take when its sample points live in page-locked HOST memory (idle PCIe bus)?"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rgbd_pose_estimation_b200 as rpe  # noqa: E402

n, H = 307200, 1024
q, t = rpe.sim_pose(1)
Q, P, _ = rpe.sim_3d_3d(2, q, t, n)
hq, hp = rpe.pinned_empty((n, 3)), rpe.pinned_empty((n, 3))
hq[:], hp[:] = Q, P
with rpe.Context(0) as c:
    c.enable_stage_timing(1)
    for name, kw in (("host arrays", dict(xc=hp.ctypes.data, xw=hq.ctypes.data)),):
        c.upload_device(n, **kw)
        for i in range(5):
            S = rpe.sample_table(10 + i, n, 3, H)
            t0 = time.perf_counter()
            c.generate("shinji", S)
            c.sync()
            dt = (time.perf_counter() - t0) * 1e3
            hy, va = c.get_hypotheses(H)
            print(name, "generate wall ms", round(dt, 3), "valid", int(va.sum()))
    c.upload(xc=P, xw=Q)
    for i in range(3):
        S = rpe.sample_table(10 + i, n, 3, H)
        t0 = time.perf_counter()
        c.generate("shinji", S)
        c.sync()
        print("device arrays generate wall ms", round((time.perf_counter() - t0) * 1e3, 3))
