#!/usr/bin/env python
"""Randomised parity sweep: GPU (binary32 path, and binary64 path for some cases) against the CPU oracle in DET mode over
random sizes, iteration counts, outlier ratios, noise levels, thresholds, NaN camera points and all seven families."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rgbd_pose_estimation_b200 as rpe  # noqa: E402
import orc  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
orc.set_math_mode(orc.DET)
ctx = rpe.Context(0)
ctx.peer_import(0, 1, ctx.peer_export())  # a world of one rank: rpe_ransac_sharded must equal rpe_ransac
names = {v: k for k, v in rpe.METHODS.items()}
bad = 0
for i in range(cases):
    rng = np.random.default_rng(seed0 + i)
    method = int(rng.integers(0, 7))
    n = int(rng.choice([3, 4, 7, 16, 33, 100, 257, 1000, 2049, 5000]))
    if method != 0:
        n = max(n, 4)
    H = int(rng.choice([1, 2, 31, 64, 100, 257, 600]))
    ors = rng.uniform(0.0, 0.8, 3)
    n2d, n3d, nnl = float(rng.uniform(0.1, 3.0)), float(rng.uniform(0.005, 0.2)), float(np.deg2rad(rng.uniform(0.2, 5.0)))
    thr3d = float(np.float32(rng.uniform(0.02, 0.8)))
    if rng.random() < 0.08:  # thresholds at the rounding level of the coordinates, and absurdly large ones
        thr3d = float(np.float32(rng.choice([1e-7, 1e-6, 3e-5, 50.0, 1e4])))
    cos_thr = float(np.cos(np.arctan(np.float32(rng.uniform(0.5, 20.0)) / np.float32(585.0))))
    cos_nl = float(np.cos(np.float32(rng.uniform(0.02, 0.5))))
    if rng.random() < 0.08:
        cos_thr = float(np.cos(np.arctan(np.float32(rng.choice([0.01, 300.0, 1e5])) / np.float32(585.0))))
    if rng.random() < 0.08:
        cos_nl = float(np.cos(np.float32(rng.choice([1e-4, 1.2, 1.5707]))))
    conf = float(rng.choice([0.9, 0.99, 0.9999]))
    q, t = rpe.sim_pose(1000 + seed0 + i)
    d = rpe.sim_2d_3d_nl(5000 + seed0 + i, q, t, n, n2d=n2d, or2d=float(ors[0]), n3d=n3d, or3d=float(ors[1]), nnl=nnl,
                         ornl=float(ors[2]))
    arrs = {k: d[k] for k in ("bv", "xc", "nc", "xw", "nw")}
    if rng.random() < 0.3 and n > 8:
        arrs["xc"][rng.integers(0, n, max(1, n // 10))] = np.nan
    degenerate = rng.random()
    if degenerate < 0.15 and n > 8:      # duplicated correspondences: rank-deficient samples
        src = rng.integers(0, n, n // 3)
        dst = rng.integers(0, n, n // 3)
        for k in arrs:
            arrs[k][dst] = arrs[k][src]
    elif degenerate < 0.25:              # collinear world points in a block
        k0 = int(rng.integers(0, max(1, n - 3)))
        arrs["xw"][k0:k0 + 3] = arrs["xw"][k0] * np.array([[1.0], [1.5], [2.0]], np.float32)
    elif degenerate < 0.35:              # large scene scale
        sc = np.float32(rng.choice([50.0, 1000.0]))
        arrs["xw"] *= sc
        arrs["xc"] *= sc
        thr3d = float(np.float32(thr3d) * sc)
    elif degenerate < 0.40:              # infinities
        arrs["xc"][int(rng.integers(0, n))] = np.inf
        arrs["xw"][int(rng.integers(0, n)), 1] = -np.inf
    if method == 0:
        arrs = {"xc": arrs["xc"], "xw": arrs["xw"]}
    m = 3 if method == 0 else 4
    S = np.stack([rng.choice(n, m, replace=False) for _ in range(H)]).astype(np.int32)
    if m == 3:
        S = np.concatenate([S, -np.ones((H, 1), np.int32)], axis=1)
    f64 = rng.random() < 0.25
    first_pass = int(rng.choice([1024, 16, 100]))
    ctx.set_first_pass_iters(first_pass)
    kw = dict(thr3d=thr3d, cos_thr=cos_thr, cos_nl=cos_nl, confidence=conf, full=True)
    slots = H * rpe.method_slots(method)
    if f64:
        a64 = {k: v.astype(np.float64) for k, v in arrs.items()}
        ref = orc.ransac(method, S, dt=np.float64, **kw, **a64)
        ctx.upload_f64(**a64)
        got = ctx.ransac_f64(names[method], S, thr3d=thr3d, cos_thr2d=cos_thr, cos_thrN=cos_nl, confidence=conf)
        pose_ok = np.array_equal(got["qd"].view(np.uint64), ref["q"].view(np.uint64)) or got["winner"] < 0 or (
            np.isnan(got["qd"]).all() and np.isnan(ref["q"]).all())  # NaN payloads differ between host and device
    else:
        ref = orc.ransac(method, S, **kw, **arrs)
        ctx.upload(**arrs)
        gk = dict(thr3d=thr3d, cos_thr2d=cos_thr, cos_thrN=cos_nl, confidence=conf)
        mode = int(rng.integers(0, 4))
        if mode == 1:    # rows on demand, pass by pass
            got = ctx.ransac_stream(names[method], lambda first, count: S[first:first + count], H, **gk)
        elif mode == 2 and H <= first_pass:  # the one-call sharded frame (one rank), mask through the stage API afterwards
            r = ctx.ransac_sharded(names[method], S, **gk)
            got = ctx.finish(names[method], H, **gk)
            assert (r["winner"], r["max_votes"], r["iter_final"]) == (got["winner"], got["max_votes"], got["iter_final"])
        elif mode == 3 and H <= first_pass:  # stage API, the slot range scored in random pieces
            ns = ctx.generate(names[method], S)
            cuts = sorted(set([0, ns] + [int(c) for c in rng.integers(0, ns + 1, int(rng.integers(0, 4)))]))
            for b, e in zip(cuts[:-1], cuts[1:]):
                ctx.score(names[method], b, e, thr3d=thr3d, cos_thr2d=cos_thr, cos_thrN=cos_nl)
            got = ctx.finish(names[method], H, **gk)
        else:
            got = ctx.ransac(names[method], S, **gk)
        pose_ok = np.array_equal(got["q"].view(np.uint32), ref["q"].view(np.uint32)) or got["winner"] < 0 or (
            np.isnan(got["q"]).all() and np.isnan(ref["q"]).all())
    # several device passes: the vote table on the device holds the last pass only
    votes = ctx.get_votes(slots) if H <= first_pass else ref["votes"]
    ok = (np.array_equal(votes, ref["votes"]) and (got["winner"], got["max_votes"], got["iter_final"]) ==
          (ref["winner"], ref["max_votes"], ref["iter_final"]) and np.array_equal(got["mask"], ref["mask"]) and pose_ok)
    if not ok:
        bad += 1
        print(f"  mask_equal {np.array_equal(got['mask'], ref['mask'])} pose_ok {pose_ok} mask sums gpu {got['mask'].sum(axis=1)} "
              f"ref {ref['mask'].sum(axis=1)} q gpu {got['qd'] if f64 else got['q']} ref {ref['q']} t gpu {got['t']} ref {ref['t']}")
        print(f"MISMATCH case {seed0 + i}: method {method} n {n} H {H} f64 {f64} votes_equal {np.array_equal(votes, ref['votes'])} "
              f"gpu {(got['winner'], got['max_votes'], got['iter_final'])} ref {(ref['winner'], ref['max_votes'], ref['iter_final'])}")
print(f"fuzz: {cases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
