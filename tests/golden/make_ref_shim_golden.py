#!/usr/bin/env python
"""Generates tests/golden/ref_shim_golden.json from the REFERENCE'S OWN pose headers (oracle/_ref/libref_shim.so:
/root/reference/pose/*.hpp compiled unmodified against the Eigen / Sophus API stand-ins of oracle/ref_shim/, its
RandomElements / ProsacSampler drawing from ::rand() after srand(sample_seed)). Runs only where /root/reference exists.

Every case is also run through the oracle in DET math mode (the arithmetic the GPU uses for acos / sincos / log / cbrt);
a case is kept only if that agrees with the reference run in every vote (mask bit for bit, counts, final Iter), so the
GPU test that reads this file (tests/test_gpu_ref_golden.py) can ask for exact equality there; `pose_bits_exact` says
whether the accepted hypothesis agrees bit for bit as well (else 2e-5 rad / 2e-5 x scale). Dropped cases are listed in the file.
Inputs are regenerated from seeds by the library's host-side Simulator, thresholds are stored as the bits the
reference's own expressions produce (P3P.hpp:323, AbsoluteOrientationNormal.hpp:223)."""
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
    for method in range(7):
        for k, (n, H, outl, nan_every) in enumerate(((1500, 256, 0.4, 0), (2600, 320, 0.65, 9))):
            sampler = 1 if method == 6 else 0
            if dt == np.float64 and (sampler or method in (1, 2, 3, 5)):
                # binary64 cases: families without P3P only — the generator's unit vectors are unit to float precision,
                # which Sophus' 1e-10 ENSURE rejects for P3P rotations, and renormalising in numpy would make the
                # input bits depend on the numpy build
                continue
            case = {"dtype": dtname, "method": method, "sampler": sampler, "pose_seed": 7000 + 10 * method + k,
                    "data_seed": 8000 + 10 * method + k, "sample_seed": 21 + method + 5 * k, "n": n, "H": H, "outliers": outl,
                    "nan_every": nan_every, "thr3d": 0.2, "thr2d_px": 8.0, "thrN_rad": 0.1, "confidence": 0.99}
            q, t = rpe.sim_pose(case["pose_seed"])
            d = rpe.sim_2d_3d_nl(case["data_seed"], q, t, n, or2d=outl, or3d=outl, ornl=outl)
            arrs = {key: np.ascontiguousarray(d[key]).astype(dt) for key in ("bv", "xc", "nc", "xw", "nw")}
            if nan_every:
                arrs["xc"][::nan_every] = np.nan
            w = np.ascontiguousarray(d["weights"]).astype(dt)
            ct, cn = refshim.cos_thr(case["thr2d_px"], F, dt), refshim.cos_nl(case["thrN_rad"], dt)
            case["cos_thr"], case["cos_nl"] = float(ct), float(cn)
            m = 3 if method == 0 else 4
            r = refshim.ransac(method, case["sample_seed"], H, sampler=sampler, thr3d=case["thr3d"], thr2d=case["thr2d_px"],
                               focal=F, thrN=case["thrN_rad"], confidence=case["confidence"], weights3=w if sampler else None,
                               dt=dt, **arrs)
            S = (rpe.prosac_table(case["sample_seed"], n, m, H, w[0]) if sampler else rpe.sample_table(case["sample_seed"], n, m, H))
            orc.set_math_mode(orc.DET)
            o = orc.ransac(method, S, thr3d=case["thr3d"], cos_thr=ct, cos_nl=cn, confidence=case["confidence"], full=True,
                           dt=dt, **arrs)
            orc.set_math_mode(orc.LIBM)
            cols = o["mask"].shape[0]
            same = (r["ensure_failures"] == 0 and o["max_votes"] == r["max_votes"] and o["iter_final"] == r["iter_final"]
                    and np.array_equal(o["mask"], r["mask"][:cols]))
            if not same:
                out["dropped"].append({k2: case[k2] for k2 in ("dtype", "method", "pose_seed")})
                continue
            # P3P / nl_2p hypotheses go through pow / cbrt / acos / sincos: libm (the reference) and the deterministic
            # helpers (the GPU) may differ in the last bits of the pose while every vote agrees
            case["pose_bits_exact"] = bool(np.array_equal(o["q"], r["q"]) and np.array_equal(o["t"], r["t"]))
            bits = np.uint32 if dt == np.float32 else np.uint64
            case["expect"] = {"max_votes": r["max_votes"], "iter_final": r["iter_final"],
                              "mask_sums": [int(v) for v in r["mask"][:cols].sum(axis=1)],
                              "mask_sha1": hashlib.sha1(np.ascontiguousarray(r["mask"][:cols]).tobytes()).hexdigest(),
                              "q_bits": [str(v) for v in r["q"].view(bits).tolist()],
                              "t_bits": [str(v) for v in r["t"].view(bits).tolist()]}
            out["cases"].append(case)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shim_golden.json"), "w"), indent=1)
print("wrote", len(out["cases"]), "cases,", len(out["dropped"]), "dropped")
