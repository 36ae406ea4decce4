"""GPU (B200): the binary64 RANSAC path (rpe_upload_f64 / rpe_ransac_f64) against the oracle instantiated for double,
DET math mode — the reference's TestMain.cpp runs its estimators as <double> (e.g. TestMain.cpp:186,210)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F = 585.0


def _data64(rpe, seed, n, **kw):
    """Simulator output widened to binary64, directions renormalised in binary64 (a double Simulator's bearings and
    normals are unit vectors to double precision; P3P relies on that), plus sub-float noise so the data is not
    representable in binary32."""
    q, t = rpe.sim_pose(seed)
    d = rpe.sim_2d_3d_nl(seed + 1, q, t, n, **kw)
    rng = np.random.default_rng(seed)
    out = {}
    for k in ("bv", "xc", "nc", "xw", "nw"):
        a = d[k].astype(np.float64)
        a = a * (1.0 + 1e-9 * rng.standard_normal(a.shape))
        if k in ("bv", "nc", "nw"):
            a = a / np.linalg.norm(a, axis=1, keepdims=True)
        out[k] = np.ascontiguousarray(a)
    return q, t, out


def _thr64(thr2d_px=8.0, thrn=0.1, thr3d=0.2):
    return dict(thr3d=thr3d, cos_thr=float(np.cos(np.arctan(thr2d_px / F))), cos_nl=float(np.cos(thrn)))


@pytest.mark.parametrize("exact_only", [0, 1])
@pytest.mark.parametrize("method,name", [(0, "shinji"), (1, "kneip"), (2, "shinji_kneip"), (3, "nl_kneip"),
                                         (4, "nl_shinji"), (5, "nl_shinji_kneip"), (6, "kneip_quat")])
def test_f64_path_bit_identical_to_double_oracle(rpe, orc, gpu_ctx, method, name, exact_only):
    """exact_only = 0: binary32 tiled prefilter + binary64 evaluation of the borderline evaluations (the default);
    exact_only = 1: every evaluation in binary64. Both must reproduce the double CPU path bit for bit."""
    orc.set_math_mode(orc.DET)
    rpe.lib.rpe_debug_f64_exact_only(exact_only)
    n, H = 3000, 300
    q, t, arrs = _data64(rpe, 500 + method, n, n2d=1.0, or2d=0.3, n3d=0.05, or3d=0.3, nnl=float(np.deg2rad(2.0)), ornl=0.3)
    S = rpe.sample_table(9, n, 3 if method == 0 else 4, H)
    th = _thr64()
    ref = orc.ransac(method, S, confidence=0.99, full=True, dt=np.float64, **th, **arrs)
    gpu_ctx.upload_f64(**arrs)
    got = gpu_ctx.ransac_f64(name, S, thr3d=th["thr3d"], cos_thr2d=th["cos_thr"], cos_thrN=th["cos_nl"], confidence=0.99)
    slots = H * rpe.method_slots(method)
    hyps, valid = gpu_ctx.get_hypotheses_f64(slots)
    assert np.array_equal(valid, (ref["votes"] >= 0).astype(np.int32))
    sel = valid == 1
    assert np.array_equal(hyps[sel].view(np.uint64), ref["hyps"][sel].view(np.uint64)), "hypotheses differ in some bit"
    assert np.array_equal(gpu_ctx.get_votes(slots), ref["votes"])
    assert got["flags"] & 2
    assert (got["winner"], got["max_votes"], got["iter_final"]) == (ref["winner"], ref["max_votes"], ref["iter_final"])
    assert np.array_equal(got["qd"].view(np.uint64), ref["q"].view(np.uint64))
    assert np.array_equal(got["td"].view(np.uint64), ref["t"].view(np.uint64))
    assert np.array_equal(got["mask"], ref["mask"])
    rpe.lib.rpe_debug_f64_exact_only(0)
    if not exact_only:
        assert got["n_borderline"] > 0 and not (got["flags"] & 1)
    # the float path on the same (rounded) data need not agree evaluation by evaluation — that is why this path exists
    orc.set_math_mode(orc.LIBM)


def test_f64_long_iter_runs_in_passes_and_refits(rpe, orc, gpu_ctx):
    orc.set_math_mode(orc.DET)
    n, H = 2000, 20000
    q, t, arrs = _data64(rpe, 601, n, n2d=1.0, or2d=0.6, n3d=0.05, or3d=0.6, nnl=float(np.deg2rad(2.0)), ornl=0.6)
    S = rpe.sample_table(3, n, 4, H)
    th = _thr64()
    ref = orc.ransac(5, S, confidence=0.999, full=False, dt=np.float64, want_arrays=False, **th, **arrs)
    gpu_ctx.upload_f64(**arrs)
    got = gpu_ctx.ransac_f64("nl_shinji_kneip", S, thr3d=th["thr3d"], cos_thr2d=th["cos_thr"], cos_thrN=th["cos_nl"],
                             confidence=0.999)
    assert (got["winner"], got["max_votes"], got["iter_final"]) == (ref["winner"], ref["max_votes"], ref["iter_final"])
    assert np.array_equal(got["mask"], ref["mask"])
    assert np.array_equal(got["qd"].view(np.uint64), ref["q"].view(np.uint64))
    # refits after a binary64 RANSAC: statistics from the float copies, binary64 accumulation
    ls_q, ls_t, ok = orc.shinji_ls(arrs["xc"], arrs["xw"], ref["mask"][1], dt=np.float64)
    fit = gpu_ctx.refit("kabsch_inliers")
    assert fit["refit_ok"] == 1
    a = np.asarray(fit["q"], np.float64)
    b = np.asarray(ls_q, np.float64)
    ang = 2.0 * np.arccos(min(1.0, abs(float(np.dot(a / np.linalg.norm(a), b / np.linalg.norm(b))))))
    assert ang < 2e-6 and np.abs(fit["t"].astype(np.float64) - ls_t).max() < 1e-5
    # a float upload afterwards puts the context back on the binary32 path
    gpu_ctx.upload(**{k: v.astype(np.float32) for k, v in arrs.items()})
    got32 = gpu_ctx.ransac("nl_shinji_kneip", S[:256], thr3d=th["thr3d"], cos_thr2d=th["cos_thr"], cos_thrN=th["cos_nl"],
                           confidence=0.999)
    assert not (got32["flags"] & 2)
    orc.set_math_mode(orc.LIBM)


def test_f64_prefilter_at_threshold_adversarial(rpe, orc, gpu_ctx):
    """Half of the camera points fit the ground-truth pose exactly (so hypotheses drawn from them reproduce it to
    ~1e-15), the other half sit within 1e-12 .. 1e-6 (relative) of the 3-D threshold sphere of that pose: the binary32
    prefilter must hand every such evaluation to the binary64 evaluation."""
    orc.set_math_mode(orc.DET)
    n, H = 4000, 64
    q, t, arrs = _data64(rpe, 777, n, n2d=1.0, or2d=0.0, n3d=0.0, or3d=0.0, nnl=0.0, ornl=0.0)
    th = _thr64()
    # exact pose in binary64 -> move x_c radially so that |x_c - (R x_w + t)| = thr3d (1 + eps_i)
    from scipy.spatial.transform import Rotation as Rot
    R = Rot.from_quat(np.asarray(q, np.float64)).as_matrix()
    y = arrs["xw"] @ R.T + np.asarray(t, np.float64)
    rng = np.random.default_rng(1)
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    eps = rng.choice([1e-12, -1e-12, 1e-9, -1e-9, 3e-8, -3e-8, 1e-6, -1e-6], size=n)
    on_sphere = np.arange(n) % 2 == 1
    arrs["xc"] = np.where(on_sphere[:, None], y + d * (th["thr3d"] * (1.0 + eps))[:, None], y)
    S = rpe.sample_table(3, n, 3, H)
    ref = orc.ransac(0, S, confidence=0.99, full=True, dt=np.float64, thr3d=th["thr3d"], xc=arrs["xc"], xw=arrs["xw"])
    gpu_ctx.upload_f64(xc=arrs["xc"], xw=arrs["xw"])
    got = gpu_ctx.ransac_f64("shinji", S, thr3d=th["thr3d"], confidence=0.99)
    assert np.array_equal(gpu_ctx.get_votes(H), ref["votes"])
    assert np.array_equal(got["mask"], ref["mask"])
    assert got["n_borderline"] > 1000  # several hypotheses are the exact pose: all their sphere points are borderline
    assert got["max_votes"] >= n // 2
    orc.set_math_mode(orc.LIBM)
