"""Opt-in reproduction of the reference's stale sample columns (rpe_set_stale_sample_columns): on frames with invalid
depth the reference's nl_2p pairs the current world sample with camera-side columns of an EARLIER sample
(/root/reference/pose/AbsoluteOrientationNormal.hpp:48-75, 299-315). Golden vectors: tests/golden/
ref_shim_stale_golden.json, produced by the reference's own sources (make_ref_shim_stale_golden.py)."""
import json
import os

import numpy as np
import pytest

from tests.test_gpu_ref_golden import _check

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_shim_stale_golden.json")


def _inputs(rpe, case):
    dt = np.float32 if case["dtype"] == "f32" else np.float64
    q, t = rpe.sim_pose(case["pose_seed"])
    o = case["outliers"]
    d = rpe.sim_2d_3d_nl(case["data_seed"], q, t, case["n"], or2d=o, or3d=o, ornl=o)
    arrs = {k: np.ascontiguousarray(d[k]).astype(dt) for k in ("bv", "xc", "nc", "xw", "nw")}
    arrs["xc"][::case["nan_every"]] = np.nan
    return dt, arrs, rpe.sample_table(case["sample_seed"], case["n"], 4, case["H"])


def test_golden_file_is_sane():
    g = json.load(open(GOLDEN))
    assert len(g["cases"]) >= 12 and sum(c["differs_from_default"] for c in g["cases"]) >= 2
    assert {c["method"] for c in g["cases"]} == {4, 5}


def test_oracle_stale_model_reproduces_reference_golden(orc, rpe):
    g = json.load(open(GOLDEN))
    orc.set_math_mode(orc.DET)
    orc.set_stale_sample_buffers(True)
    try:
        for case in g["cases"]:
            dt, arrs, S = _inputs(rpe, case)
            r = orc.ransac(case["method"], S, thr3d=case["thr3d"], cos_thr=case["cos_thr"], cos_nl=case["cos_nl"],
                           confidence=case["confidence"], full=True, dt=dt, **arrs)
            _check(case, r["max_votes"], r["iter_final"], r["mask"], r["q"], r["t"])
    finally:
        orc.set_stale_sample_buffers(False)
        orc.set_math_mode(orc.LIBM)


@pytest.mark.gpu
def test_gpu_with_the_option_reproduces_reference_golden(rpe, orc):
    g = json.load(open(GOLDEN))
    orc.set_math_mode(orc.DET)
    seen_difference = 0
    with rpe.Context(0) as ctx:
        for first_pass in (1024, 40):  # 40: the frame takes several device passes, the stale state is carried across them
            ctx.set_first_pass_iters(first_pass)
            for case in g["cases"]:
                dt, arrs, S = _inputs(rpe, case)
                kw = dict(thr3d=case["thr3d"], cos_thr2d=case["cos_thr"], cos_thrN=case["cos_nl"], confidence=case["confidence"])

                def run():
                    if case["dtype"] == "f32":
                        ctx.upload(**arrs)
                        r = ctx.ransac(case["method"], S, **kw)
                        return r, r["q"], r["t"]
                    ctx.upload_f64(**arrs)
                    r = ctx.ransac_f64(case["method"], S, **kw)
                    return r, r["qd"], r["td"]
                ctx.set_stale_sample_columns(True)
                r, q, t = run()
                _check(case, r["max_votes"], r["iter_final"], r["mask"], q, t)
                if first_pass == 1024:
                    # the whole vote table against the oracle's model of the reference (every nl_2p slot, stale or not)
                    slots = case["H"] * rpe.method_slots(case["method"])
                    orc.set_stale_sample_buffers(True)
                    try:
                        ref = orc.ransac(case["method"], S, thr3d=case["thr3d"], cos_thr=case["cos_thr"], cos_nl=case["cos_nl"],
                                         confidence=case["confidence"], full=True, dt=dt, **arrs)
                    finally:
                        orc.set_stale_sample_buffers(False)
                    assert np.array_equal(ctx.get_votes(slots), ref["votes"])
                # default behaviour: NaN goes into nl_2p; differs from the reference exactly where the golden file says so
                ctx.set_stale_sample_columns(False)
                r0, _, _ = run()
                differs = r0["max_votes"] != case["expect"]["max_votes"] or \
                    [int(v) for v in r0["mask"].sum(axis=1)] != case["expect"]["mask_sums"]
                assert differs == case["differs_from_default"] or not case["differs_from_default"]
                seen_difference += int(differs)
    assert seen_difference >= 2


@pytest.mark.gpu
def test_option_changes_nothing_without_invalid_points(rpe):
    n, H = 4000, 200
    q, t = rpe.sim_pose(11)
    d = rpe.sim_2d_3d_nl(12, q, t, n)
    arrs = {k: d[k] for k in ("bv", "xc", "nc", "xw", "nw")}
    S = rpe.sample_table(3, n, 4, H)
    kw = dict(thr3d=0.2, cos_thr2d=float(np.cos(np.arctan(8 / 585.0))), cos_thrN=float(np.cos(0.1)), confidence=0.99)
    with rpe.Context(0) as ctx:
        ctx.upload(**arrs)
        a = ctx.ransac("nl_shinji_kneip", S, **kw)
        va = ctx.get_votes(3 * H).copy()
        ctx.set_stale_sample_columns(True)
        b = ctx.ransac("nl_shinji_kneip", S, **kw)
        assert np.array_equal(va, ctx.get_votes(3 * H)) and np.array_equal(a["mask"], b["mask"])
        assert np.array_equal(a["q"].view(np.uint32), b["q"].view(np.uint32))
