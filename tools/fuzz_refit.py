#!/usr/bin/env python
"""Randomised sweep of the refits against the CPU oracle (tolerances: north_star's 1e-6 rad, 1e-6 x scene scale): Kabsch over the
inliers / over all points, LM with random modality weights and iteration caps (statistics-based and per-row paths),
nl_shinji_kneip_ls with and without dynamic weights; explicit random masks through rpe_set_mask."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rgbd_pose_estimation_b200 as rpe  # noqa: E402
import orc  # noqa: E402


def angle(qa, qb):
    a = np.asarray(qa, np.float64) / np.linalg.norm(np.asarray(qa, np.float64))
    b = np.asarray(qb, np.float64) / np.linalg.norm(np.asarray(qb, np.float64))
    av, aw, bv, bw = a[:3], a[3], -b[:3], b[3]
    w = aw * bw - np.dot(av, bv)
    v = aw * bv + bw * av + np.cross(av, bv)
    return 2.0 * np.arctan2(np.linalg.norm(v), abs(w))


TOL_ANG = float(os.environ.get('RPE_FUZZ_TOL_ANG', 1e-6))      # rad
TOL_T = float(os.environ.get('RPE_FUZZ_TOL_T', 1e-6))          # x scene scale
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
orc.set_math_mode(orc.DET)
ctx = rpe.Context(0)
bad = 0
worst = {"kabsch": 0.0, "gn": 0.0, "nlsk": 0.0}
for i in range(cases):
    rng = np.random.default_rng(seed0 + i)
    n = int(rng.choice([200, 1000, 5000, 20000]))
    ors = rng.uniform(0.0, 0.5, 3)
    q, t = rpe.sim_pose(7000 + seed0 + i)
    d = rpe.sim_2d_3d_nl(9000 + seed0 + i, q, t, n, n2d=float(rng.uniform(0.2, 2.0)), or2d=float(ors[0]),
                         n3d=float(rng.uniform(0.01, 0.1)), or3d=float(ors[1]), nnl=float(np.deg2rad(rng.uniform(0.5, 3.0))),
                         ornl=float(ors[2]))
    arrs = {k: d[k] for k in ("bv", "xc", "nc", "xw", "nw")}
    scale = float(np.abs(arrs["xc"]).max())
    S = rpe.sample_table(seed0 + i, n, 4, 128)
    th = dict(thr3d=0.2, cos_thr2d=float(np.cos(np.arctan(np.float32(8.0) / np.float32(585.0)))), cos_thrN=float(np.cos(np.float32(0.1))))
    ctx.upload(**arrs)
    got = ctx.ransac("nl_shinji_kneip", S, confidence=0.99, **th)
    mask = got["mask"]
    if rng.random() < 0.5:  # explicit mask: thin the inlier columns at random
        mask = mask.copy()
        for c in range(3):
            mask[c, rng.random(n) < rng.uniform(0.0, 0.6)] = 0
        ctx.set_mask(mask)
    if min(int(mask[0].sum()), int(mask[1].sum()), int(mask[2].sum())) < 30:
        continue
    msgs = []
    # Kabsch over the 3-D inliers / over everything
    ctx.set_pose(got["q"], got["t"])
    fit = ctx.refit("kabsch_inliers")
    rq, rt, _ = orc.shinji_ls(arrs["xc"], arrs["xw"], mask[1], dt=np.float64)
    e = max(angle(fit["q"], rq) / TOL_ANG, np.abs(fit["t"] - rt).max() / (TOL_T * scale))
    worst["kabsch"] = max(worst["kabsch"], e)
    if e > 1:
        msgs.append(f"kabsch_inliers {e:.2f}")
    # LM with random weights
    w = [float(rng.choice([0.0, 0.5, 1.0, 2.0])) for _ in range(3)]
    if w[1] == 0.0 and w[0] == 0.0:
        w[1] = 1.0  # normals alone leave the translation free
    iters = int(rng.integers(1, 9))
    ctx.set_pose(got["q"], got["t"])
    fit = ctx.refit("gn", weights=w, max_iters=iters)
    tq, tt, info = orc.refine_gn(got["q"], got["t"], mask, w=tuple(w), max_iters=iters, **arrs)
    e = max(angle(fit["q"], tq) / TOL_ANG, np.abs(fit["t"].astype(np.float64) - tt.astype(np.float64)).max() / (TOL_T * scale))
    worst["gn"] = max(worst["gn"], e)
    if e > 1:
        msgs.append(f"gn w={w} iters={iters} {e:.2f} evals gpu {fit['refit_evals']} ref {info['evals']}")
    # the reference's multi-modal refinement
    ctx.set_pose(got["q"], got["t"])
    use_w = rng.random() < 0.5
    W = d["weights"] if use_w else None
    fit = ctx.refit("nl_sk_ls", weights=W)
    rq, rt = orc.nl_shinji_kneip_ls(got["q"], got["t"], mask, got["max_votes"], weights3=None if W is None else W.astype(np.float64),
                                    dt=np.float64, **arrs)
    if np.isfinite(rq).all() and fit["refit_ok"] == 1:
        e = max(angle(fit["q"], rq) / TOL_ANG, np.abs(fit["t"].astype(np.float64) - rt).max() / (TOL_T * scale))
        worst["nlsk"] = max(worst["nlsk"], e)
        if e > 1:
            msgs.append(f"nl_sk_ls weights={use_w} {e:.2f}")
    elif np.isfinite(rq).all() != (fit["refit_ok"] == 1):
        msgs.append(f"nl_sk_ls validity differs: oracle finite {np.isfinite(rq).all()} gpu ok {fit['refit_ok']}")
    if msgs:
        bad += 1
        print(f"MISMATCH case {seed0 + i}: n {n} inliers {mask.sum(axis=1)} :: " + "; ".join(msgs))
print(f"fuzz_refit: {cases} cases, {bad} outside tolerance; worst (in units of the tolerance) {worst}")
sys.exit(1 if bad else 0)
