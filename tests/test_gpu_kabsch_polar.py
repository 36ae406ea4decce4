"""GPU: the Kabsch refits (shinji_ls / shinji_ls1 / shinji_ls2, /root/reference/pose/AbsoluteOrientation.hpp:273-342) take
the rotation of a well-conditioned cross-covariance from a scaled Newton polar iteration instead of the Jacobi SVD
(pipeline.cu: polar_rotation_newton). Both routes must agree far inside the 1e-6 rad / 1e-6 x scale bar, and degenerate
covariances (planar points, reflections, three points) must still take the SVD route."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHINJI = 0


def _angle(qa, qb):
    a = np.asarray(qa, np.float64) / np.linalg.norm(np.asarray(qa, np.float64))
    b = np.asarray(qb, np.float64) / np.linalg.norm(np.asarray(qb, np.float64))
    av, aw, bv, bw = a[:3], a[3], -b[:3], b[3]
    w = aw * bw - np.dot(av, bv)
    v = aw * bv + bw * av + np.cross(av, bv)
    return 2.0 * np.arctan2(np.linalg.norm(v), abs(w))


def _both(rpe, ctx, kind, rerun):
    """The refit with the polar iteration switched off and on; `rerun` repeats the RANSAC call first (its mask kernel
    solves the inliers' Kabsch problem itself, with the switch as it stands at that moment)."""
    out = {}
    for on in (0, 1):
        assert rpe.lib.rpe_debug_set_kabsch_polar(on) == 0
        rerun()
        out[on] = ctx.refit(kind)
    rpe.lib.rpe_debug_set_kabsch_polar(1)
    return out[0], out[1]


@pytest.mark.parametrize("n,seed,outlier,noise", [(2000, 3, 0.5, 0.1), (307200, 5, 0.5, 0.1), (50000, 7, 0.8, 0.3),
                                                   (300, 9, 0.1, 0.01), (20000, 11, 0.0, 0.0)])
def test_polar_equals_svd_and_oracle(rpe, orc, gpu_ctx, n, seed, outlier, noise):
    orc.set_math_mode(orc.DET)
    q, t = rpe.sim_pose(seed)
    Q, P, _ = rpe.sim_3d_3d(seed + 1, q, t, n, noise=noise, outlier_ratio=outlier)
    S = rpe.sample_table(seed, n, 3, 256)
    gpu_ctx.upload(xc=P, xw=Q)
    got = gpu_ctx.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999)
    scale = float(np.abs(P).max())
    for kind in ("kabsch_inliers", "kabsch_all"):
        svd, pol = _both(rpe, gpu_ctx, kind, lambda: gpu_ctx.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999))
        assert svd["refit_ok"] == 1 and pol["refit_ok"] == 1
        assert _angle(svd["q"], pol["q"]) < 1e-6 and np.abs(svd["t"] - pol["t"]).max() <= 1e-6 * scale
    ls_q, ls_t, _ = orc.shinji_ls(P, Q, got["mask"][1], dt=np.float64)
    gpu_ctx.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999)
    pol = gpu_ctx.refit("kabsch_inliers")
    assert _angle(pol["q"], ls_q) < 1e-6 and np.abs(pol["t"] - ls_t).max() <= 1e-6 * scale


def test_degenerate_covariances_take_the_svd_route(rpe, gpu_ctx):
    """Planar world points (rank 2), a mirrored cloud (det < 0) and three points: identical output with the switch on and off."""
    n = 4000
    q, t = rpe.sim_pose(21)
    Q, P, _ = rpe.sim_3d_3d(22, q, t, n, noise=0.0, outlier_ratio=0.0)
    cases = []
    Qp, Pp = Q.copy(), P.copy()
    Pp[:, 2] = 1.0  # camera points in a plane: the covariance loses a rank
    cases.append((Qp, Pp))
    Qm = Q.copy()
    Qm[:, 0] = -Qm[:, 0]  # a reflection fits best: det(M) < 0
    cases.append((Qm, P.copy()))
    cases.append((Q[:3].copy(), P[:3].copy()))
    for Qc, Pc in cases:
        m = Qc.shape[0]
        gpu_ctx.upload(xc=Pc, xw=Qc)
        S = rpe.sample_table(5, m, 3, 64)
        svd, pol = _both(rpe, gpu_ctx, "kabsch_all", lambda: gpu_ctx.ransac(SHINJI, S, thr3d=1e9, confidence=0.9999))
        assert svd["refit_ok"] == pol["refit_ok"]
        assert np.array_equal(np.asarray(svd["q"]).view(np.uint32), np.asarray(pol["q"]).view(np.uint32))
        assert np.array_equal(np.asarray(svd["t"]).view(np.uint32), np.asarray(pol["t"]).view(np.uint32))
