#!/usr/bin/env python
"""Generates tests/golden/oracle_golden.json from the CPU oracle (regression pins; the reference has no
golden vectors of its own and cannot be run here — see oracle/README.md)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rgbd_pose_estimation_b200 as rpe  # noqa: E402
from tests import orc  # noqa: E402

out = {"rand_seed1_first8": [int(v) for v in orc.rand_seq(1, 8)],
       "sample_table_seed1_n1000_m3": orc.sample_table(1, 1000, 3, 4).tolist(), "ransac": []}
orc.set_math_mode(orc.DET)
cos_thr = float(np.cos(np.arctan(8.0 / 585.0)))
cos_nl = float(np.cos(0.1))
for method in range(7):
    case = {"method": method, "pose_seed": 100 + method, "data_seed": 200 + method, "sample_seed": 1 + method, "n": 1200,
            "H": 300, "thr3d": 0.2, "cos_thr": cos_thr, "cos_nl": cos_nl, "confidence": 0.99}
    q, t = rpe.sim_pose(case["pose_seed"])
    d = rpe.sim_2d_3d_nl(case["data_seed"], q, t, case["n"])
    arrs = {k: d[k] for k in ("bv", "xc", "nc", "xw", "nw")}
    S = rpe.sample_table(case["sample_seed"], case["n"], 3 if method == 0 else 4, case["H"])
    r = orc.ransac(method, S, thr3d=case["thr3d"], cos_thr=cos_thr, cos_nl=cos_nl, confidence=case["confidence"],
                   full=True, **arrs)
    case["expect"] = [r["winner"], r["max_votes"], r["iter_final"], int(r["votes"].astype(np.int64).sum())]
    case["expect_mask_sums"] = [int(v) for v in r["mask"].sum(axis=1)]
    case["q_bits"] = np.array(r["q"], np.float32).view(np.uint32).tolist()
    out["ransac"].append(case)
# Tp = double: the same generator output widened to binary64 (no renormalisation, so the input bits are reproducible
# anywhere); families without P3P, whose double instantiation rejects bearings that are unit only to float precision.
out["ransac_f64"] = []
for method in (0, 4):
    case = {"method": method, "pose_seed": 300 + method, "data_seed": 400 + method, "sample_seed": 11 + method, "n": 1500,
            "H": 300, "thr3d": 0.2, "cos_thr": cos_thr, "cos_nl": cos_nl, "confidence": 0.99}
    q, t = rpe.sim_pose(case["pose_seed"])
    d = rpe.sim_2d_3d_nl(case["data_seed"], q, t, case["n"])
    arrs = {k: d[k].astype(np.float64) for k in ("bv", "xc", "nc", "xw", "nw")}
    S = rpe.sample_table(case["sample_seed"], case["n"], 3 if method == 0 else 4, case["H"])
    r = orc.ransac(method, S, thr3d=case["thr3d"], cos_thr=cos_thr, cos_nl=cos_nl, confidence=case["confidence"],
                   full=True, dt=np.float64, **arrs)
    case["expect"] = [r["winner"], r["max_votes"], r["iter_final"], int(r["votes"].astype(np.int64).sum())]
    case["expect_mask_sums"] = [int(v) for v in r["mask"].sum(axis=1)]
    case["q_bits"] = [str(v) for v in np.array(r["q"], np.float64).view(np.uint64).tolist()]
    out["ransac_f64"].append(case)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.json"), "w"), indent=1)
print("wrote", len(out["ransac"]), "cases")
