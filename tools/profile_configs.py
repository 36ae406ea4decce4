#!/usr/bin/env python
"""Blocking latency of BASELINE.json configs #1-#3 through the C-ABI (host buffers in, pose + mask out) next to the
single-threaded CPU oracle on the same inputs (the reference is single-threaded)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rgbd_pose_estimation_b200 as rpe  # noqa: E402
import orc  # noqa: E402

F = 585.0
cos_thr = float(np.cos(np.arctan(np.float32(8.0) / np.float32(F))))
cos_nl = float(np.cos(np.float32(0.1)))
ctx = rpe.Context(0)
ctx.set_first_pass_iters(int(os.environ.get("FIRST_PASS", "1024")))


def timed(fn, reps=20):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


q, t = rpe.sim_pose(11)
# config 1: 1000 3-D/3-D correspondences, 50 % outliers, Iter0 = 100 000, conf 0.9999, shinji_ransac2 + shinji_ls1
Q, P, _ = rpe.sim_3d_3d(12, q, t, 1000, noise=0.1, outlier_ratio=0.5)
S = rpe.sample_table(1, 1000, 3, 100000)


def cfg1():
    ctx.upload(xc=P, xw=Q)
    r = ctx.ransac("shinji", S, thr3d=0.25, confidence=0.9999)
    ctx.refit("kabsch_inliers")
    return r


g = timed(cfg1)
t0 = time.perf_counter()
ref = orc.ransac(0, S, thr3d=0.25, confidence=0.9999, full=False, xc=P, xw=Q, want_arrays=False)
c = (time.perf_counter() - t0) * 1e3
print(f"config 1 (N=1000, Iter0=100000): GPU {g:.3f} ms per call, CPU oracle {c:.2f} ms (iterations run {ref['iters_run']})")

# config 2: 10 000 2-D/3-D correspondences, 70 % outliers, kneip_ransac + LM
Q, U, Pgt, W = rpe.sim_2d_3d(13, q, t, 10000, noise_px=1.0, outlier_ratio=0.7)
S = rpe.sample_table(2, 10000, 4, 100000)


def cfg2():
    ctx.upload(bv=U, xw=Q)
    r = ctx.ransac("kneip", S, cos_thr2d=cos_thr, confidence=0.99)
    ctx.refit("gn", max_iters=8)
    return r


g = timed(cfg2)
t0 = time.perf_counter()
ref = orc.ransac(1, S, cos_thr=cos_thr, confidence=0.99, full=False, bv=U, xw=Q, want_arrays=False)
c = (time.perf_counter() - t0) * 1e3
print(f"config 2 (N=10000, 70 % outliers, kneip + LM): GPU {g:.3f} ms per call, CPU oracle RANSAC {c:.2f} ms "
      f"(iterations run {ref['iters_run']})")

# config 3: 50 000 correspondences, 2-D + 3-D + normals, nl_shinji_kneip_ransac + nl_shinji_kneip_ls
d = rpe.sim_2d_3d_nl(14, q, t, 50000, n2d=1.0, or2d=0.3, n3d=0.05, or3d=0.3, nnl=float(np.deg2rad(2.0)), ornl=0.3)
arrs = {k: d[k] for k in ("bv", "xc", "nc", "xw", "nw")}
S = rpe.sample_table(3, 50000, 4, 1024)


def cfg3():
    ctx.upload(**arrs)
    r = ctx.ransac("nl_shinji_kneip", S, thr3d=0.2, cos_thr2d=cos_thr, cos_thrN=cos_nl, confidence=0.99)
    ctx.refit("nl_sk_ls")
    return r


g = timed(cfg3)
t0 = time.perf_counter()
ref = orc.ransac(5, S, thr3d=0.2, cos_thr=cos_thr, cos_nl=cos_nl, confidence=0.99, full=False, want_arrays=False, **arrs)
c = (time.perf_counter() - t0) * 1e3
print(f"config 3 (N=50000, three modalities, Iter0=1024): GPU {g:.3f} ms per call (all 1024 iterations scored), CPU oracle "
      f"{c:.2f} ms (stops after {ref['iters_run']} iterations)")
