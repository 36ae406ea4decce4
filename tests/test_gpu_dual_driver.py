"""GPU part of the one-source / two-builds check: tests/cpp/dual_driver.cpp — written in the reference's vocabulary only
(Eigen matrices, Sophus::SO3, the reference's adapter / estimator / refit signatures, ::rand() sampling) — built here
against the drop-in headers + librpe_b200.so and run on the GPU, against tests/golden/dual_driver_golden.json: the output
of the SAME source built against the reference's own headers (made where /root/reference exists by
tests/golden/make_dual_driver_golden.py; checked against the oracle by tests/test_dual_driver.py).

Bars: accepted votes, final Iter, the hash of every inlier flag and the refreshed index-list lengths identical; the
accepted hypothesis within 2e-5 rad / 2e-5 x scale (bit-identical unless libm and the deterministic cbrt differ by an ulp
inside P3P); closed-form refits within 2e-6 rad / 2e-6 x scale of the reference's binary32 result, nl_shinji_kneip_ls
within 1e-5 rad / 1e-5 x scale.

Written when the round's GPU minutes were spent: non-strict xfail until it has been seen green on a B200 once."""
import json
import subprocess

import numpy as np
import pytest

from tests import dual_driver_common as dd

SCALE = 10.0  # metres


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="not yet run on a GPU box (added after the round's GPU budget was spent)")
def test_dropin_build_reproduces_the_reference_build(tmp_path, rpe):
    gold = json.load(open(dd.GOLDEN))
    assert (gold["n"], gold["pose_seed"], gold["data_seed"]) == (dd.N, dd.POSE_SEED, dd.DATA_SEED)
    exe, inp = str(tmp_path / "driver_b200"), str(tmp_path / "in.bin")
    dd.build_dropin(exe)
    dd.write_input(rpe, inp)
    p = subprocess.run([exe, inp], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    got = dd.parse(p.stdout)
    assert sorted(got) == sorted(gold["cases"])
    for name, c in gold["cases"].items():
        g = got[name]
        qe = np.array([float.fromhex(v) for v in c["q_hex"]], np.float32)
        te = np.array([float.fromhex(v) for v in c["t_hex"]], np.float32)
        dq, dt = dd.angle(g["q"], qe), float(np.linalg.norm(g["t"].astype(np.float64) - te.astype(np.float64)))
        if name in dd.FRESH_LISTS:  # a RANSAC / PROSAC run
            assert (g["max_votes"], g["iter"], g["mask_hash"]) == (c["max_votes"], c["iter"], c["mask_hash"]), name
            for lst in dd.FRESH_LISTS[name]:
                assert g["n_idx"][lst] == c["n_idx"][lst], (name, lst)
            assert dq <= 2e-5 and dt <= 2e-5 * SCALE, (name, dq, dt)
        elif "nl_shinji_kneip_ls" in name:
            assert dq <= 1e-5 and dt <= 1e-5 * SCALE, (name, dq, dt)
        else:
            assert dq <= 2e-6 and dt <= 2e-6 * SCALE, (name, dq, dt)
