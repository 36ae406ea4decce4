#!/usr/bin/env python
"""Generates tests/golden/ref_shim_stale_golden.json: runs of the REFERENCE'S OWN nl_shinji_ransac / nl_shinji_kneip_ransac
(oracle/_ref/libref_shim.so, see make_ref_shim_golden.py) on frames where a third of the camera points have no depth
(all-NaN). On such frames the reference's nl_2p reads STALE sample columns (AbsoluteOrientationNormal.hpp:48-75,
299-315); the product reproduces that only with rpe_set_stale_sample_columns(ctx, 1). `differs_from_default` marks the
cases where the reference's accepted result is NOT what the default (NaN-propagating) behaviour gives — the cases that
prove the option does something. Runs only where /root/reference exists."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rgbd_pose_estimation_b200 as rpe  # noqa: E402
from tests import orc, refshim  # noqa: E402

assert refshim.available(), "needs /root/reference (or a prebuilt oracle/_ref/libref_shim.so)"
F = 585.0
out = {"focal": F, "cases": [], "dropped": []}
for dtname, dt in (("f32", np.float32), ("f64", np.float64)):
    for method in (4, 5):
        if dt == np.float64 and method == 5:
            continue  # P3P in binary64 trips Sophus' 1e-10 ENSURE on float-precision unit vectors (see make_ref_shim_golden.py)
        for seed in (1, 2, 8, 43, 77, 91):
            n, H = 700, 150
            case = {"dtype": dtname, "method": method, "pose_seed": 6000 + seed, "data_seed": 6100 + seed, "sample_seed": seed,
                    "n": n, "H": H, "outliers": 0.7, "nan_every": 3, "thr3d": 0.2, "thr2d_px": 8.0, "thrN_rad": 0.1,
                    "confidence": 0.99}
            q, t = rpe.sim_pose(case["pose_seed"])
            g = rpe.sim_2d_3d_nl(case["data_seed"], q, t, n, or2d=0.7, or3d=0.7, ornl=0.7)
            arrs = {k: np.ascontiguousarray(g[k]).astype(dt) for k in ("bv", "xc", "nc", "xw", "nw")}
            arrs["xc"][::3] = np.nan
            ct, cn = refshim.cos_thr(8.0, F, dt), refshim.cos_nl(0.1, dt)
            case["cos_thr"], case["cos_nl"] = float(ct), float(cn)
            r = refshim.ransac(method, seed, H, thr3d=0.2, thr2d=8.0, focal=F, thrN=0.1, confidence=0.99, dt=dt, **arrs)
            S = rpe.sample_table(seed, n, 4, H)
            orc.set_math_mode(orc.DET)
            orc.set_stale_sample_buffers(True)
            o = orc.ransac(method, S, thr3d=0.2, cos_thr=ct, cos_nl=cn, confidence=0.99, full=True, dt=dt, **arrs)
            orc.set_stale_sample_buffers(False)
            o0 = orc.ransac(method, S, thr3d=0.2, cos_thr=ct, cos_nl=cn, confidence=0.99, full=True, dt=dt, **arrs)
            orc.set_math_mode(orc.LIBM)
            cols = o["mask"].shape[0]
            same = (r["ensure_failures"] == 0 and o["max_votes"] == r["max_votes"] and o["iter_final"] == r["iter_final"]
                    and np.array_equal(o["mask"], r["mask"][:cols]))
            if not same:
                out["dropped"].append({k2: case[k2] for k2 in ("dtype", "method", "sample_seed")})
                continue
            case["differs_from_default"] = bool(o0["max_votes"] != r["max_votes"] or not np.array_equal(o0["mask"], r["mask"][:cols]))
            case["pose_bits_exact"] = bool(np.array_equal(o["q"], r["q"]) and np.array_equal(o["t"], r["t"]))
            bits = np.uint32 if dt == np.float32 else np.uint64
            case["expect"] = {"max_votes": r["max_votes"], "iter_final": r["iter_final"],
                              "mask_sums": [int(v) for v in r["mask"][:cols].sum(axis=1)],
                              "mask_sha1": hashlib.sha1(np.ascontiguousarray(r["mask"][:cols]).tobytes()).hexdigest(),
                              "q_bits": [str(v) for v in r["q"].view(bits).tolist()],
                              "t_bits": [str(v) for v in r["t"].view(bits).tolist()]}
            out["cases"].append(case)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shim_stale_golden.json"), "w"), indent=1)
print("wrote", len(out["cases"]), "cases,", len(out["dropped"]), "dropped,",
      sum(c["differs_from_default"] for c in out["cases"]), "differ from the default behaviour")
