#!/usr/bin/env python
"""Generates tests/golden/dual_driver_golden.json: the output of tests/cpp/dual_driver.cpp built against the REFERENCE'S
OWN headers (/root/reference/pose + the Eigen / Sophus stand-ins of oracle/ref_shim), on inputs any box can regenerate
from seeds. The GPU box runs the same source built against the drop-in headers and compares
(tests/test_gpu_dual_driver.py). Runs only where /root/reference exists."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rgbd_pose_estimation_b200 as rpe  # noqa: E402
from tests import dual_driver_common as dd  # noqa: E402

with tempfile.TemporaryDirectory() as tmp:
    exe, inp = os.path.join(tmp, "driver_ref"), os.path.join(tmp, "in.bin")
    dd.build_reference(exe)
    dd.write_input(rpe, inp)
    got = dd.parse(subprocess.run([exe, inp], capture_output=True, text=True, check=True).stdout)
out = {"n": dd.N, "pose_seed": dd.POSE_SEED, "data_seed": dd.DATA_SEED, "outliers": dd.OUTLIERS, "cases": {}}
for name, g in got.items():
    out["cases"][name] = {"max_votes": g["max_votes"], "iter": g["iter"], "mask_hash": g["mask_hash"], "n_idx": g["n_idx"],
                          "q_hex": [float(v).hex() for v in g["q"]], "t_hex": [float(v).hex() for v in g["t"]]}
json.dump(out, open(dd.GOLDEN, "w"), indent=1)
print("wrote", len(out["cases"]), "cases")
