"""Shared by tests/test_dual_driver.py (CPU), tests/test_gpu_dual_driver.py (GPU) and
tests/golden/make_dual_driver_golden.py: inputs, builds and output parsing of tests/cpp/dual_driver.cpp."""
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "dual_driver.cpp")
SHIM = os.path.join(ROOT, "oracle", "ref_shim")
REFERENCE = "/root/reference/pose"
GOLDEN = os.path.join(ROOT, "tests", "golden", "dual_driver_golden.json")
N, POSE_SEED, DATA_SEED = 2000, 501, 502
OUTLIERS = dict(or2d=0.35, or3d=0.45, ornl=0.25)


def inputs(rpe):
    q, t = rpe.sim_pose(POSE_SEED)
    d = rpe.sim_2d_3d_nl(DATA_SEED, q, t, N, **OUTLIERS)
    return q, t, d


def write_input(rpe, path):
    q, t, d = inputs(rpe)
    with open(path, "wb") as f:
        f.write(np.int32(N).tobytes())
        for k in ("bv", "xc", "nc", "xw", "nw"):
            f.write(np.ascontiguousarray(d[k], np.float32).tobytes())
        f.write(np.ascontiguousarray(d["weights"], np.float32).tobytes())
    return q, t, d


def build_reference(exe):
    """(A) the driver against the reference's own headers + the Eigen / Sophus stand-ins."""
    subprocess.run(["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-fno-fast-math", "-DNDEBUG", "-w", "-I", SHIM, "-I", REFERENCE,
                    "-o", exe, SRC], check=True)


def build_dropin(exe):
    """(B) the same source against include/rpe + librpe_b200.so (Eigen stand-in only for the matrix type)."""
    libdir = os.path.join(ROOT, "rgbd_pose_estimation_b200")
    subprocess.run(["g++", "-std=c++14", "-O2", "-I", SHIM, "-I", os.path.join(ROOT, "include", "rpe"), "-I",
                    os.path.join(ROOT, "include"), "-o", exe, SRC, "-L", libdir, "-lrpe_b200", "-Wl,-rpath," + libdir], check=True)


def parse(stdout):
    out = {}
    for line in stdout.splitlines():
        if not line.startswith("{"):
            continue
        r = json.loads(line)
        r["q"] = np.array([float.fromhex(v) for v in r["q"]], np.float32)
        r["t"] = np.array([float.fromhex(v) for v in r["t"]], np.float32)
        out[r["case"]] = r
    return out


def angle(qa, qb):
    a = np.asarray(qa, np.float64) / np.linalg.norm(np.asarray(qa, np.float64))
    b = np.asarray(qb, np.float64) / np.linalg.norm(np.asarray(qb, np.float64))
    return 2.0 * np.arccos(min(1.0, abs(float(np.dot(a, b)))))


# which of the three index lists (2-D, 3-D, normal) an estimator refreshes; the others keep whatever an earlier call on the
# same adapter left there (the reference only calls the cvtInlier of the classes it knows about)
FRESH_LISTS = {"shinji_prosac": [1], "shinji_ransac2": [1], "kneip_prosac": [0], "kneip_ransac": [0],
               "shinji_kneip_ransac": [0, 1], "nl_kneip_ransac": [0, 2], "nl_shinji_ransac": [1, 2],
               "nl_shinji_kneip_ransac": [0, 1, 2]}
