"""GPU parity: the correspondence-stationary 3-D scorer (score_ur.cu: thread <-> pairs of correspondences, hypothesis
scalars in uniform registers, per-warp REDUX of packed sign words) against the CPU oracle and against the
hypothesis-stationary scorer it replaces for dense frames (reference loop: /root/reference/pose/AbsoluteOrientation.hpp:
135-143 = :192-200). Variant 30 forces the kernel for every frame it can take; 24 is the round-1 kernel. Bars: vote
tables, winner, final Iter and masks BIT-IDENTICAL."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHINJI = 0


def _frame(rpe, seed, n, outlier=0.5, noise=0.1):
    q, t = rpe.sim_pose(seed)
    Q, P, _ = rpe.sim_3d_3d(seed + 1, q, t, n, noise=noise, outlier_ratio=outlier)
    return Q, P


@pytest.fixture
def ur_variant(rpe):
    rpe.lib.rpe_debug_set_score_variant(30)
    yield
    rpe.lib.rpe_debug_set_score_variant(14)


@pytest.mark.parametrize("n,H,seed,outlier,noise,thr", [
    (1000, 1024, 3, 0.5, 0.1, 0.25),     # far fewer pairs than threads
    (1001, 300, 5, 0.2, 0.05, 0.1),      # odd n: the last pair is half NaN padding
    (37, 64, 7, 0.0, 0.01, 0.05),
    (20000, 1024, 9, 0.5, 0.1, 0.25),
    (4097, 513, 11, 0.7, 0.2, 0.5),
    (50003, 1000, 13, 0.5, 0.1, 0.25),   # n % 4 = 3: frame tail read element-wise
    (66000, 700, 15, 0.3, 0.02, 0.05),
])
def test_votes_winner_iter_mask_identical_to_oracle(rpe, orc, gpu_ctx, ur_variant, n, H, seed, outlier, noise, thr):
    orc.set_math_mode(orc.DET)
    Q, P = _frame(rpe, seed, n, outlier, noise)
    S = rpe.sample_table(seed, n, 3, H)
    ref = orc.ransac(SHINJI, S, thr3d=thr, confidence=0.9999, full=True, xc=P, xw=Q)
    gpu_ctx.upload(xc=P, xw=Q)
    got = gpu_ctx.ransac(SHINJI, S, thr3d=thr, confidence=0.9999)
    votes = gpu_ctx.get_votes(H)
    assert got["flags"] == 0
    assert np.array_equal(votes, ref["votes"]), f"votes differ at {np.nonzero(votes != ref['votes'])[0][:10]}"
    assert got["winner"] == ref["winner"] and got["max_votes"] == ref["max_votes"] and got["iter_final"] == ref["iter_final"]
    assert np.array_equal(got["mask"], ref["mask"])


def test_invalid_depth_and_empty_slots(rpe, orc, gpu_ctx, ur_variant):
    """NaN camera points (never inliers, never borderline) and degenerate samples (empty hypothesis slots)."""
    orc.set_math_mode(orc.DET)
    n, H = 30001, 777
    Q, P = _frame(rpe, 21, n)
    rng = np.random.default_rng(5)
    P = P.copy()
    P[rng.random(n) < 0.2] = np.nan
    S = rpe.sample_table(21, n, 3, H)
    S[5, :3] = S[5, 0]  # a degenerate sample
    ref = orc.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999, full=True, xc=P, xw=Q)
    gpu_ctx.upload(xc=P, xw=Q)
    got = gpu_ctx.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999)
    votes = gpu_ctx.get_votes(H)
    assert np.array_equal(votes, ref["votes"])
    assert got["winner"] == ref["winner"] and np.array_equal(got["mask"], ref["mask"])


def test_at_threshold_values_go_through_the_exact_path(rpe, orc, gpu_ctx, ur_variant):
    """Correspondences placed within a few ulp of the threshold sphere of the true pose: the second pass has work to do."""
    orc.set_math_mode(orc.DET)
    n, H = 40000, 512
    q, t = rpe.sim_pose(31)
    Q, P, _ = rpe.sim_3d_3d(32, q, t, n, noise=0.0, outlier_ratio=0.0)
    rng = np.random.default_rng(7)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    P = (P.astype(np.float64) + 0.25 * d * (1.0 + rng.integers(-3, 4, size=(n, 1)) * 6e-8)).astype(np.float32)
    S = rpe.sample_table(31, n, 3, H)
    ref = orc.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999, full=True, xc=P, xw=Q)
    gpu_ctx.upload(xc=P, xw=Q)
    got = gpu_ctx.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999)
    votes = gpu_ctx.get_votes(H)
    assert np.array_equal(votes, ref["votes"])
    assert np.array_equal(got["mask"], ref["mask"])


def test_dense_frame_same_votes_as_the_hypothesis_stationary_scorer(rpe, gpu_ctx):
    """Config #4 (307 200 x 1 024; 2 076 pairs per CTA column = 512 x 4 + a tail of 28): the uniform-register scorer = the
    default and the round-1 kernels, vote by vote."""
    n, H = 307200, 1024
    Q, P = _frame(rpe, 41, n)
    S = rpe.sample_table(41, n, 3, H)
    gpu_ctx.upload(xc=P, xw=Q)
    out = {}
    for v in (24, 14, 30):
        rpe.lib.rpe_debug_set_score_variant(v)
        r = gpu_ctx.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999)
        out[v] = (gpu_ctx.get_votes(H).copy(), r)
    rpe.lib.rpe_debug_set_score_variant(14)
    for v in (14, 30):
        assert np.array_equal(out[v][0], out[24][0])
        assert out[v][1]["winner"] == out[24][1]["winner"] and out[v][1]["iter_final"] == out[24][1]["iter_final"]
        assert np.array_equal(out[v][1]["mask"], out[24][1]["mask"])
