"""GPU: device-side Simulator: distributions of Simulator.hpp:158-367 generated in HBM."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _R(q):
    x, y, z, w = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def test_device_3d_3d_statistics_and_ransac(rpe, gpu_ctx):
    n = 100000
    q, t = rpe.sim_pose(7)
    gpu_ctx.sim_3d_3d_device(11, q, t, n, noise=0.1, outlier_ratio=0.5)
    d = gpu_ctx.download(("xc", "xw"))
    P, Q = d["xc"].astype(np.float64), d["xw"].astype(np.float64)
    f = 585.0
    assert (np.abs(P[:, 0] / P[:, 2]) < 320 / f).all() and (np.abs(P[:, 1] / P[:, 2]) < 240 / f).all()
    assert P[:, 2].min() >= 0.4 and P[:, 2].max() <= 8.0
    e = P - (Q @ _R(q).T + t)
    inl = np.linalg.norm(e, axis=1) < 0.6
    assert abs(inl.mean() - 0.5) < 0.01               # exactly n/2 outlier positions (+ the odd far-out inlier)
    assert abs(e[inl].std(axis=0).mean() - 0.1) < 0.005 and np.abs(e[inl].mean(axis=0)).max() < 0.003
    # depth is uniform in [0.4, 8] before frustum rejection only in z: check z is not degenerate
    assert 3.5 < P[:, 2].mean() < 6.5
    S = rpe.sample_table(1, n, 3, 512)
    r = gpu_ctx.ransac("shinji", S, thr3d=0.25, confidence=0.9999, want_mask=False)
    fit = gpu_ctx.refit("kabsch_inliers")
    assert 0.4 * n < r["max_votes"] < 0.5 * n
    Re = _R(fit["q"])
    ang = np.arccos(np.clip((np.trace(Re @ _R(q).T) - 1) / 2, -1, 1))
    assert ang < 0.03 and np.abs(fit["t"] - t).max() < 0.3
    # different seeds give different frames, same seed the same frame
    gpu_ctx.sim_3d_3d_device(11, q, t, n)
    assert np.array_equal(gpu_ctx.download(("xc",))["xc"], d["xc"])
    gpu_ctx.sim_3d_3d_device(12, q, t, n)
    assert not np.array_equal(gpu_ctx.download(("xc",))["xc"], d["xc"])


def test_device_multimodal_statistics(rpe, gpu_ctx):
    n = 60000
    q, t = rpe.sim_pose(9)
    gpu_ctx.sim_2d_3d_nl_device(21, q, t, n, n2d=1.0, or2d=0.3, n3d=0.05, or3d=0.2, nnl=np.deg2rad(2.0), ornl=0.1)
    d = gpu_ctx.download(("bv", "xc", "nc", "xw", "nw"))
    R = _R(q)
    y = d["xw"].astype(np.float64) @ R.T + t
    assert np.abs(np.linalg.norm(d["bv"], axis=1) - 1).max() < 1e-5
    assert np.abs(np.linalg.norm(d["nc"], axis=1) - 1).max() < 1e-5 and np.abs(np.linalg.norm(d["nw"], axis=1) - 1).max() < 1e-5
    assert (d["nc"][:, 2] <= 1e-6).all()
    cos2 = np.sum(y / np.linalg.norm(y, axis=1, keepdims=True) * d["bv"], axis=1)
    in2 = cos2 > np.cos(np.arctan(8.0 / 585.0))
    in3 = np.linalg.norm(d["xc"] - y, axis=1) < 0.3
    inn = np.sum((d["nw"].astype(np.float64) @ R.T) * d["nc"], axis=1) > np.cos(0.15)
    assert abs(in2.mean() - 0.7) < 0.02 and abs(in3.mean() - 0.8) < 0.02 and abs(inn.mean() - 0.9) < 0.03
    S = rpe.sample_table(1, n, 4, 256)
    r = gpu_ctx.ransac("nl_shinji_kneip", S, thr3d=0.2, cos_thr2d=float(np.cos(np.arctan(8 / 585.0))),
                       cos_thrN=float(np.cos(0.1)), confidence=0.99, want_mask=False)
    Re = _R(r["q"])
    ang = np.arccos(np.clip((np.trace(Re @ R.T) - 1) / 2, -1, 1))
    assert ang < 0.05 and r["max_votes"] > 1.8 * n


def test_device_kinect_noise_model(rpe, gpu_ctx):
    """rpe_sim_kinect_2d_3d_nl_device: camera points carry the Kinect axial (quadratic in depth) and lateral noise."""
    n, f = 200000, 585.0
    q, t = rpe.sim_pose(31)
    gpu_ctx.sim_kinect_2d_3d_nl_device(32, q, t, n, n2d=1.0, or2d=0.0, or3d=0.2, nnl=float(np.deg2rad(2.0)), ornl=0.0)
    d = gpu_ctx.download(("xc", "xw", "nc", "nw"))
    P = d["xc"].astype(np.float64)
    Pgt = d["xw"].astype(np.float64) @ _R(q).T + np.asarray(t, np.float64)
    e = P - Pgt
    inl = np.abs(e).max(axis=1) < 0.5
    assert abs((~inl).mean() - 0.2) < 0.01
    z = Pgt[:, 2]
    ngt = d["nw"].astype(np.float64) @ _R(q).T            # true camera-frame normals (noise-free up to 2 deg)
    theta = np.arccos(np.clip(-ngt[:, 2], -1, 1))
    frontal = inl & (theta < np.deg2rad(55.0))
    base = 0.0012 + 0.0019 * (z - 0.4) ** 2

    def robust_sigma(v):
        return 1.4826 * np.median(np.abs(v))

    for lo, hi, tol in [(3.0, 4.0, 0.06), (6.5, 7.5, 0.04)]:
        sel = frontal & (z > lo) & (z < hi)
        assert abs(robust_sigma(e[sel, 2] / base[sel]) - 1) < tol
    sel = frontal & (z > 3.0) & (z < 4.0)
    lat = (0.8 + 0.035 * theta / (np.pi / 2 - theta)) * z / f
    assert abs(robust_sigma(e[sel, 0] / lat[sel]) - 1) < 0.06 and abs(robust_sigma(e[sel, 1] / lat[sel]) - 1) < 0.06
    # grazing surfaces are noisier in depth than the quadratic term alone
    graz = inl & (theta > np.deg2rad(75.0)) & (z > 3.0) & (z < 4.0)
    assert robust_sigma(e[graz, 2] / base[graz]) > 1.05


def test_frame_replacement_leaves_binary64_mode(rpe, orc):
    """Regression (round-1 driver run): a context that was in binary64 mode (rpe_upload_f64, small n) must score the NEW
    frame after rpe_sim_3d_3d_device / rpe_upload / rpe_upload_device — not the stale binary64 arrays, and never read
    past them."""
    rng = np.random.default_rng(5)
    with rpe.Context(0) as ctx:
        n_small = 2600
        q, t = rpe.sim_pose(3)
        Q, P, _ = rpe.sim_3d_3d(4, q, t, n_small)
        ctx.upload_f64(xc=P.astype(np.float64), xw=Q.astype(np.float64))
        r64 = ctx.ransac_f64("shinji", rpe.sample_table(1, n_small, 3, 64), thr3d=0.25, confidence=0.99)
        assert r64["flags"] & 2
        # 1. device-side simulator
        n = 100000
        ctx.sim_3d_3d_device(11, q, t, n, noise=0.1, outlier_ratio=0.5)
        S = rpe.sample_table(1, n, 3, 256)
        r = ctx.ransac("shinji", S, thr3d=0.25, confidence=0.9999)
        assert not (r["flags"] & 2) and 0.4 * n < r["max_votes"] < 0.5 * n
        d = ctx.download(("xc", "xw"))
        ref = orc.ransac(0, S, thr3d=0.25, confidence=0.9999, full=True, xc=d["xc"], xw=d["xw"])
        assert r["winner"] == ref["winner"] and r["max_votes"] == ref["max_votes"]
        assert np.array_equal(r["mask"], ref["mask"])
        # 2. float upload after a binary64 one
        ctx.upload_f64(xc=P.astype(np.float64), xw=Q.astype(np.float64))
        Q2, P2, _ = rpe.sim_3d_3d(6, q, t, 30000)
        ctx.upload(xc=P2, xw=Q2)
        S2 = rpe.sample_table(2, 30000, 3, 128)
        r2 = ctx.ransac("shinji", S2, thr3d=0.25, confidence=0.99)
        ref2 = orc.ransac(0, S2, thr3d=0.25, confidence=0.99, full=True, xc=P2, xw=Q2)
        assert not (r2["flags"] & 2) and r2["winner"] == ref2["winner"] and np.array_equal(r2["mask"], ref2["mask"])
        # 3. rpe_ransac_f64 on a context that left binary64 mode is refused, not served from stale arrays
        with pytest.raises(rpe.RpeError):
            ctx.ransac_f64("shinji", S2, thr3d=0.25, confidence=0.99)
        del rng
