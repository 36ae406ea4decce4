"""GPU parity for the 2-D / hybrid / normal-aware estimator families through the C-ABI.

Reference paths: kneip_ransac (P3P.hpp:320-392), kneip_prosac's scoring form (:439-453),
shinji_kneip_ransac (AbsoluteOrientation.hpp:367-438), nl_kneip_ransac / nl_shinji_ransac /
nl_shinji_kneip_ransac (AbsoluteOrientationNormal.hpp:215-445), on Simulator.hpp:316-367 inputs.
Bars: generated hypotheses (Kneip P3P + Ferrari quartic, shinji, nl_2p) BIT-IDENTICAL to the oracle in DET
math mode; vote tables, winner, final Iter, masks and per-column counts identical; LM refinement within
1e-6 rad / 1e-6 x scale of its CPU twin.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NAMES = {1: "kneip", 2: "shinji_kneip", 3: "nl_kneip", 4: "nl_shinji", 5: "nl_shinji_kneip", 6: "kneip_quat"}
F = 585.0


def _data(rpe, seed, n, **kw):
    q, t = rpe.sim_pose(seed)
    d = rpe.sim_2d_3d_nl(seed + 1, q, t, n, **kw)
    return q, t, {k: d[k] for k in ("bv", "xc", "nc", "xw", "nw")}, d["weights"]


def _thr(thr2d_px=8.0, thrn=0.1, thr3d=0.2):
    return dict(thr3d=thr3d, cos_thr=float(np.cos(np.arctan(np.float32(thr2d_px) / np.float32(F)))),
                cos_nl=float(np.cos(np.float32(thrn))))


def _angle(qa, qb):
    a = np.asarray(qa, np.float64) / np.linalg.norm(np.asarray(qa, np.float64))
    b = np.asarray(qb, np.float64) / np.linalg.norm(np.asarray(qb, np.float64))
    av, aw, bv, bw = a[:3], a[3], -b[:3], b[3]
    w = aw * bw - np.dot(av, bv)
    v = aw * bv + bw * av + np.cross(av, bv)
    return 2.0 * np.arctan2(np.linalg.norm(v), abs(w))


@pytest.mark.parametrize("method", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("n,H,seed", [(800, 512, 3), (5001, 300, 5)])
def test_generation_bit_exact(rpe, orc, gpu_ctx, method, n, H, seed):
    orc.set_math_mode(orc.DET)
    q, t, arrs, _ = _data(rpe, seed + 10 * method, n)
    S = rpe.sample_table(seed, n, 4, H)
    th = _thr()
    ref = orc.ransac(method, S, confidence=0.99, full=True, **th, **arrs)
    gpu_ctx.upload(**arrs)
    ns = gpu_ctx.generate(method, S)
    hyps, valid = gpu_ctx.get_hypotheses(ns)
    assert np.array_equal(valid, (ref["votes"] >= 0).astype(np.int32))
    sel = valid == 1
    assert sel.sum() > H // 2
    same = hyps[sel].view(np.uint32) == ref["hyps"][sel].view(np.uint32)
    assert same.all(), f"{(~same.all(axis=1)).sum()} of {sel.sum()} hypotheses differ in some bit"


@pytest.mark.parametrize("raw", [1, 0])  # 1: the scorer streams the caller's arrays (bulk TMA + transpose); 0: packed copy
@pytest.mark.parametrize("method", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("n,H,seed,kw", [
    (800, 512, 3, {}),
    (5001, 300, 5, dict(n2d=2.0, or2d=0.5, n3d=0.1, or3d=0.5, nnl=float(np.deg2rad(4.0)), ornl=0.5)),
    (64, 100, 7, dict(or2d=0.0, or3d=0.0, ornl=0.0)),
    (40003, 200, 9, dict(or2d=0.3, or3d=0.3, ornl=0.3, nan_rows=True)),  # several stages per CTA, ragged tail, invalid depth
])
def test_ransac_identical_to_oracle(rpe, orc, gpu_ctx, method, n, H, seed, kw, raw):
    orc.set_math_mode(orc.DET)
    kw = dict(kw)
    nan_rows = kw.pop("nan_rows", False)
    q, t, arrs, _ = _data(rpe, seed + 10 * method, n, **kw)
    if nan_rows:  # all-NaN camera points: no 3-D and no normal vote, the 2-D vote still counts (Appendix A of SURVEY.md)
        bad = np.random.default_rng(seed).choice(n, n // 20, replace=False)
        bad = bad[bad > 64]  # keep the first rows valid so that most samples stay usable
        arrs["xc"] = arrs["xc"].copy()
        arrs["xc"][bad] = np.nan
        arrs["xc"][n - 1] = np.nan
    rpe.lib.rpe_debug_set_raw_tiles(raw)
    S = rpe.sample_table(seed, n, 4, H)
    th = _thr()
    ref = orc.ransac(method, S, confidence=0.99, full=True, **th, **arrs)
    gpu_ctx.upload(**arrs)
    got = gpu_ctx.ransac(method, S, thr3d=th["thr3d"], cos_thr2d=th["cos_thr"], cos_thrN=th["cos_nl"], confidence=0.99)
    votes = gpu_ctx.get_votes(ref["votes"].shape[0])
    assert np.array_equal(votes, ref["votes"]), f"votes differ at {np.nonzero(votes != ref['votes'])[0][:10]}"
    for k in ("winner", "max_votes", "iter_final"):
        assert got[k] == ref[k], k
    assert np.array_equal(got["q"].view(np.uint32), ref["q"].view(np.uint32))
    assert np.array_equal(got["t"].view(np.uint32), ref["t"].view(np.uint32))
    assert np.array_equal(got["mask"], ref["mask"])
    assert got["n_inliers"][:got["mask"].shape[0]] == [int(v) for v in ref["mask"].sum(axis=1)]
    assert sum(got["n_inliers"]) == ref["max_votes"]


def test_config2_pnp_10k_70pct_outliers(rpe, orc, gpu_ctx):
    """BASELINE.json config #2: P3P RANSAC on 10k 2-D/3-D correspondences, 70 % outliers, then LM refinement."""
    orc.set_math_mode(orc.DET)
    n, H = 10000, 1024
    q, t = rpe.sim_pose(101)
    Q, U, P, _ = rpe.sim_2d_3d(102, q, t, n, noise_px=1.0, outlier_ratio=0.7)
    S = rpe.sample_table(1, n, 4, H)
    th = _thr(thr2d_px=8.0)
    ref = orc.ransac(1, S, confidence=0.99, full=True, cos_thr=th["cos_thr"], bv=U, xw=Q)
    gpu_ctx.upload(bv=U, xw=Q)
    got = gpu_ctx.ransac("kneip", S, cos_thr2d=th["cos_thr"], confidence=0.99)
    assert np.array_equal(gpu_ctx.get_votes(H), ref["votes"])
    assert got["winner"] == ref["winner"] and got["iter_final"] == ref["iter_final"] and got["max_votes"] == ref["max_votes"]
    assert np.array_equal(got["mask"], ref["mask"])
    assert 0.25 * n < got["max_votes"] < 0.35 * n
    twin_q, twin_t, info = orc.refine_gn(got["q"], got["t"], got["mask"], max_iters=8, bv=U, xw=Q)
    fit = gpu_ctx.refit("gn", max_iters=8)
    assert fit["refit_ok"] == 1
    assert _angle(fit["q"], twin_q) < 1e-6
    assert np.abs(fit["t"].astype(np.float64) - twin_t.astype(np.float64)).max() < 1e-6 * 10.0
    assert _angle(fit["q"], q) < _angle(got["q"], q) + 1e-4  # refinement does not move away from ground truth
    assert _angle(fit["q"], q) < 5e-3


def test_config3_normal_ao_50k_int32_indices(rpe, orc, gpu_ctx):
    """BASELINE.json config #3: 50 000 correspondences with 2-D/3-D/normal residuals. The reference's `short` index
    loops break above 32 767 (PnPPoseAdapter.hpp:232); masks here cover all 50 000 rows."""
    orc.set_math_mode(orc.DET)
    n, H = 50000, 256
    q, t, arrs, _ = _data(rpe, 201, n, n2d=1.0, or2d=0.3, n3d=0.05, or3d=0.3, nnl=float(np.deg2rad(2.0)), ornl=0.3)
    S = rpe.sample_table(1, n, 4, H)
    th = _thr()
    ref = orc.ransac(5, S, confidence=0.99, full=True, **th, **arrs)
    gpu_ctx.upload(**arrs)
    got = gpu_ctx.ransac("nl_shinji_kneip", S, thr3d=th["thr3d"], cos_thr2d=th["cos_thr"], cos_thrN=th["cos_nl"],
                         confidence=0.99)
    assert np.array_equal(gpu_ctx.get_votes(3 * H), ref["votes"])
    assert got["winner"] == ref["winner"] and got["iter_final"] == ref["iter_final"]
    assert np.array_equal(got["mask"], ref["mask"])
    assert got["mask"][:, 40000:].sum() > 1000  # rows beyond the reference's short range carry inliers
    twin_q, twin_t, info = orc.refine_gn(got["q"], got["t"], got["mask"], max_iters=8, **arrs)
    fit = gpu_ctx.refit("gn", max_iters=8)
    assert _angle(fit["q"], twin_q) < 1e-6
    assert np.abs(fit["t"].astype(np.float64) - twin_t.astype(np.float64)).max() < 1e-6 * 10.0
    assert _angle(fit["q"], q) < 2e-3


def test_prosac_tables_feed_the_same_kernels(rpe, orc, gpu_ctx):
    """kneip_prosac / shinji_kneip_prosac differ from the RANSAC variants only by the sample table
    (Utility.hpp:161-250) — and, for kneip_prosac, by the quaternion-form rotation in scoring (P3P.hpp:442)."""
    orc.set_math_mode(orc.DET)
    n, H = 2000, 400
    q, t, arrs, W = _data(rpe, 301, n)
    th = _thr()
    S = rpe.prosac_table(1, n, 4, H, W[0])
    S = np.minimum(S, n - 1)  # the reference's sampler can emit index n (Utility.hpp:238): clamp for the test
    for method in (6, 2):
        ref = orc.ransac(method, S, confidence=0.99, full=True, **th, **arrs)
        gpu_ctx.upload(**arrs)
        got = gpu_ctx.ransac(method, S, thr3d=th["thr3d"], cos_thr2d=th["cos_thr"], cos_thrN=th["cos_nl"], confidence=0.99)
        assert np.array_equal(gpu_ctx.get_votes(ref["votes"].shape[0]), ref["votes"])
        assert got["winner"] == ref["winner"] and got["iter_final"] == ref["iter_final"]
        assert np.array_equal(got["mask"], ref["mask"])


def test_missing_modality_is_an_error(rpe, gpu_ctx):
    q, t = rpe.sim_pose(1)
    Q, P, _ = rpe.sim_3d_3d(2, q, t, 100)
    gpu_ctx.upload(xc=P, xw=Q)
    S = rpe.sample_table(1, 100, 4, 10)
    with pytest.raises(rpe.RpeError):
        gpu_ctx.ransac("kneip", S, cos_thr2d=0.999)
    with pytest.raises(rpe.RpeError):
        gpu_ctx.ransac("nl_shinji", S, thr3d=0.1, cos_thrN=0.99)


@pytest.mark.parametrize("method", [1, 2, 3, 4, 5, 6])
def test_fast_tiled_scorer_equals_exact_kernel(rpe, orc, gpu_ctx, method):
    """The FFMA2 tiled scorer (+ guard band + exact fix-up) and the exact-order kernel give the same vote table,
    and the fast path really is the one that ran (borderline evaluations were queued)."""
    n, H = 20000, 512
    q, t, arrs, _ = _data(rpe, 400 + method, n, n2d=2.0, or2d=0.4, n3d=0.08, or3d=0.4, nnl=float(np.deg2rad(3.0)), ornl=0.4)
    S = rpe.sample_table(3, n, 4, H)
    th = _thr()
    gpu_ctx.upload(**arrs)
    rpe.lib.rpe_debug_force_exact_multi(1)
    a = gpu_ctx.ransac(method, S, thr3d=th["thr3d"], cos_thr2d=th["cos_thr"], cos_thrN=th["cos_nl"], confidence=0.99)
    va = gpu_ctx.get_votes(a["n_slots"])
    rpe.lib.rpe_debug_force_exact_multi(0)
    b = gpu_ctx.ransac(method, S, thr3d=th["thr3d"], cos_thr2d=th["cos_thr"], cos_thrN=th["cos_nl"], confidence=0.99)
    vb = gpu_ctx.get_votes(b["n_slots"])
    assert a["n_borderline"] == 0 and b["flags"] == 0
    assert np.array_equal(va, vb)
    assert np.array_equal(a["mask"], b["mask"]) and a["winner"] == b["winner"] and a["iter_final"] == b["iter_final"]


def test_borderline_2d_and_normal_residuals_go_to_the_exact_path(rpe, orc, gpu_ctx):
    """Adversarial data: bearing vectors at exactly the threshold angle and normals at exactly the threshold angle
    (+- a few ulps) from the ground-truth pose, scored with hypotheses that reproduce that pose."""
    orc.set_math_mode(orc.DET)
    n = 4096
    q, t = rpe.sim_pose(71)
    d = rpe.sim_2d_3d_nl(72, q, t, n, n2d=0.0, or2d=0.0, n3d=0.0, or3d=0.0, nnl=0.0, ornl=0.0)
    rng = np.random.default_rng(5)
    th = _thr()
    ang2 = np.arccos(np.float64(th["cos_thr"]))
    angn = np.arccos(np.float64(th["cos_nl"]))

    def tilt(v, ang):
        v = v.astype(np.float64)
        r = rng.normal(size=v.shape)
        r -= np.sum(r * v, axis=1, keepdims=True) * v
        r /= np.linalg.norm(r, axis=1, keepdims=True)
        a = ang * (1.0 + rng.integers(-4, 5, size=(v.shape[0], 1)) * 2e-8)
        return (np.cos(a) * v + np.sin(a) * r).astype(np.float32)

    bv = tilt(d["bv"], ang2)
    nc = tilt(d["nc"], angn)
    bv[:16], nc[:16] = d["bv"][:16], d["nc"][:16]  # clean rows for the minimal samples
    arrs = dict(bv=bv, xc=d["xc"], nc=nc, xw=d["xw"], nw=d["nw"])
    S = np.tile(np.array([[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11], [12, 13, 14, 15]], np.int32), (8, 1))
    for method in (5, 3, 1):
        ref = orc.ransac(method, S, confidence=0.99, full=True, **th, **arrs)
        gpu_ctx.upload(**arrs)
        got = gpu_ctx.ransac(method, S, thr3d=th["thr3d"], cos_thr2d=th["cos_thr"], cos_thrN=th["cos_nl"], confidence=0.99)
        votes = gpu_ctx.get_votes(ref["votes"].shape[0])
        assert got["n_borderline"] > 1000, method
        assert np.array_equal(votes, ref["votes"]), method
        assert np.array_equal(got["mask"], ref["mask"]), method
    orc.set_math_mode(orc.LIBM)


def test_gn_from_statistics_3d_and_normal_rows(rpe, orc, gpu_ctx):
    """Without 2-D rows the LM loop runs from the sufficient statistics (one launch, no pass over the data):
    (a) statistics left by the mask kernel, (b) recomputed after rpe_set_mask, both against the per-row twin."""
    orc.set_math_mode(orc.DET)
    n, H = 20000, 128
    q, t, arrs, _ = _data(rpe, 211, n, n2d=1.0, or2d=0.3, n3d=0.05, or3d=0.3, nnl=float(np.deg2rad(2.0)), ornl=0.3)
    S = rpe.sample_table(2, n, 4, H)
    th = _thr()
    gpu_ctx.upload(**arrs)
    got = gpu_ctx.ransac("nl_shinji_kneip", S, thr3d=th["thr3d"], cos_thr2d=th["cos_thr"], cos_thrN=th["cos_nl"],
                         confidence=0.99)
    for w in [(0.0, 1.0, 1.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0), (0.0, 0.5, 2.0)]:
        twin_q, twin_t, info = orc.refine_gn(got["q"], got["t"], got["mask"], w=w, max_iters=8, **arrs)
        gpu_ctx.set_pose(got["q"], got["t"])
        fit = gpu_ctx.refit("gn", weights=w, max_iters=8)
        # normal rows alone leave the translation unobservable: both sides stop after the first evaluation
        assert fit["refit_ok"] == 1 and abs(fit["refit_evals"] - info["evals"]) <= 1
        assert _angle(fit["q"], twin_q) < 1e-6, w
        assert np.abs(fit["t"].astype(np.float64) - twin_t.astype(np.float64)).max() < 1e-6 * 10.0, w
        assert abs(fit["refit_cost"] - info["cost"]) <= 1e-5 * max(1.0, abs(info["cost"])), w
    # (b) an explicit mask invalidates the cached statistics
    rng = np.random.default_rng(5)
    mask = got["mask"].copy()
    mask[1, rng.random(n) < 0.3] = 0
    mask[2, rng.random(n) < 0.3] = 0
    gpu_ctx.set_mask(mask)
    gpu_ctx.set_pose(got["q"], got["t"])
    w = (0.0, 1.0, 1.0)
    twin_q, twin_t, info = orc.refine_gn(got["q"], got["t"], mask, w=w, max_iters=8, **arrs)
    fit = gpu_ctx.refit("gn", weights=w, max_iters=8)
    assert _angle(fit["q"], twin_q) < 1e-6
    assert np.abs(fit["t"].astype(np.float64) - twin_t.astype(np.float64)).max() < 1e-6 * 10.0
    fit2 = gpu_ctx.refit("gn", weights=w, max_iters=8)  # second call: statistics now cached, starts from the refined pose
    assert _angle(fit2["q"], twin_q) < 1e-6


def test_many_borderline_2d_evaluations_stay_exact(rpe, orc):
    """Low outlier ratio + pixel-level 2-D threshold: about a percent of a good hypothesis' evaluations fall inside the
    rigorous guard band. They are queued on the spot (segmented worklist), the list grows after an overflow, and the
    votes stay bit-identical to the CPU path."""
    orc.set_math_mode(orc.DET)
    n, H = 40000, 256
    q, t, arrs, _ = _data(rpe, 401, n, n2d=1.0, or2d=0.1, n3d=0.05, or3d=0.1, nnl=float(np.deg2rad(2.0)), ornl=0.1)
    S = rpe.sample_table(4, n, 4, H)
    th = _thr()
    ref = orc.ransac(5, S, confidence=0.99, full=True, nthreads=8, **th, **arrs)
    with rpe.Context(0) as ctx:
        ctx.upload(**arrs)
        for rep in range(3):
            got = ctx.ransac("nl_shinji_kneip", S, thr3d=th["thr3d"], cos_thr2d=th["cos_thr"], cos_thrN=th["cos_nl"],
                             confidence=0.99)
            assert np.array_equal(ctx.get_votes(3 * H), ref["votes"]), rep
            assert got["winner"] == ref["winner"] and got["iter_final"] == ref["iter_final"]
            assert np.array_equal(got["mask"], ref["mask"])
        assert got["n_borderline"] > 50000
        assert got["flags"] == 0  # by now the worklist holds them all


def test_gpu_against_committed_golden_vectors(rpe, gpu_ctx):
    """The CUDA path against tests/golden/oracle_golden.json directly (no oracle at run time): all seven families in
    binary32, two families in binary64."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.json")))
    for case in g["ransac"]:
        q, t = rpe.sim_pose(case["pose_seed"])
        d = rpe.sim_2d_3d_nl(case["data_seed"], q, t, case["n"])
        arrs = {k: d[k] for k in ("bv", "xc", "nc", "xw", "nw")}
        S = rpe.sample_table(case["sample_seed"], case["n"], 3 if case["method"] == 0 else 4, case["H"])
        gpu_ctx.upload(**arrs)
        r = gpu_ctx.ransac(case["method"], S, thr3d=case["thr3d"], cos_thr2d=case["cos_thr"], cos_thrN=case["cos_nl"],
                           confidence=case["confidence"])
        slots = case["H"] * rpe.method_slots(case["method"])
        assert [r["winner"], r["max_votes"], r["iter_final"]] == case["expect"][:3], case["method"]
        assert int(gpu_ctx.get_votes(slots).astype(np.int64).sum()) == case["expect"][3]
        assert [int(v) for v in r["mask"].sum(axis=1)] == case["expect_mask_sums"]
        assert r["q"].view(np.uint32).tolist() == case["q_bits"]
    for case in g["ransac_f64"]:
        q, t = rpe.sim_pose(case["pose_seed"])
        d = rpe.sim_2d_3d_nl(case["data_seed"], q, t, case["n"])
        arrs = {k: d[k].astype(np.float64) for k in ("bv", "xc", "nc", "xw", "nw")}
        S = rpe.sample_table(case["sample_seed"], case["n"], 3 if case["method"] == 0 else 4, case["H"])
        gpu_ctx.upload_f64(**arrs)
        r = gpu_ctx.ransac_f64(case["method"], S, thr3d=case["thr3d"], cos_thr2d=case["cos_thr"], cos_thrN=case["cos_nl"],
                               confidence=case["confidence"])
        slots = case["H"] * rpe.method_slots(case["method"])
        assert [r["winner"], r["max_votes"], r["iter_final"]] == case["expect"][:3], case["method"]
        assert int(gpu_ctx.get_votes(slots).astype(np.int64).sum()) == case["expect"][3]
        assert [int(v) for v in r["mask"].sum(axis=1)] == case["expect_mask_sums"]
        assert [str(v) for v in r["qd"].view(np.uint64).tolist()] == case["q_bits"]


def test_randomised_parity_sweep(rpe):
    """tools/fuzz_parity.py: random sizes (3 .. 5000), iteration counts (1 .. 600), outlier ratios, noise, thresholds,
    NaN camera points, all seven families, a quarter of the cases on the binary64 path — votes, winner, Iter, masks and
    the accepted pose against the CPU oracle. (10 000 cases were run once when this test was written: 0 mismatches after
    the no-winner mask column it found had been fixed.)"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_parity.py"), "250", "31337"], capture_output=True,
                       text=True, timeout=600)
    assert p.returncode == 0 and "250 cases, 0 mismatches" in p.stdout, p.stdout[-3000:] + p.stderr[-2000:]


def test_no_accepted_hypothesis_leaves_all_flags_set(rpe, orc, gpu_ctx):
    """Every sample invalid (a NaN camera point in each) -> nothing is accepted; the adapters' flags keep their initial
    setOnes() state in EVERY column, also the 2-D column a 3-D-only family never writes (found by the sweep above)."""
    orc.set_math_mode(orc.DET)
    n, H = 100, 4
    q, t = rpe.sim_pose(5)
    Q, P, _ = rpe.sim_3d_3d(6, q, t, n)
    P[:8] = np.nan
    S = np.array([[0, 1, 2, -1], [3, 4, 20, -1], [5, 30, 6, -1], [40, 7, 1, -1]], np.int32)
    ref = orc.ransac(0, S, thr3d=0.25, confidence=0.99, full=True, xc=P, xw=Q)
    gpu_ctx.upload(xc=P, xw=Q)
    got = gpu_ctx.ransac("shinji", S, thr3d=0.25, confidence=0.99)
    assert got["winner"] == ref["winner"] == -1 and got["iter_final"] == ref["iter_final"] == H
    assert np.array_equal(got["mask"], ref["mask"]) and int(got["mask"].sum()) == 2 * n


def test_randomised_refit_sweep(rpe):
    """tools/fuzz_refit.py: Kabsch over inliers, LM with random modality weights / iteration caps (both the
    statistics-based and the per-row kernels) and nl_shinji_kneip_ls with and without dynamic weights, on RANSAC masks and
    on explicit random masks, against the CPU oracle within north_star's 1e-6 rad / 1e-6 x scene scale for ALL of them
    (400 cases on a B200 in round 2: worst 0.10 of that tolerance; 3 000 cases in round 1, which exposed the per-row LM
    kernel forming its rows in binary32 while its twin uses binary64.)"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_refit.py"), "120", "4242"], capture_output=True,
                       text=True, timeout=600)
    assert p.returncode == 0 and "120 cases, 0 outside tolerance" in p.stdout, p.stdout[-3000:] + p.stderr[-2000:]
