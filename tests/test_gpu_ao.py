"""GPU parity: AO (3-D/3-D) RANSAC through the C-ABI vs the CPU oracle.

Reference path: shinji_ransac2 + shinji_ls1 (AbsoluteOrientation.hpp:158-213, 298-320) on
simulate_3d_3d_correspondences inputs (Simulator.hpp:268-314). Bars: hypotheses,
votes[H], winner, final Iter, masks and counts BIT-IDENTICAL (oracle in DET math mode, see
oracle/README.md); refit rotation within 1e-6 rad and translation within 1e-6 x scene scale against the
oracle evaluated in binary64.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHINJI = 0


def _frame(rpe, seed, n, outlier=0.5, noise=0.1):
    q, t = rpe.sim_pose(seed)
    Q, P, _ = rpe.sim_3d_3d(seed + 1, q, t, n, noise=noise, outlier_ratio=outlier)
    return q, t, Q, P


def _angle(qa, qb):
    """Geodesic angle between two rotations given as (x,y,z,w) quaternions, accurate near zero."""
    a = np.asarray(qa, np.float64) / np.linalg.norm(np.asarray(qa, np.float64))
    b = np.asarray(qb, np.float64) / np.linalg.norm(np.asarray(qb, np.float64))
    # relative rotation a * conj(b)
    av, aw, bv, bw = a[:3], a[3], -b[:3], b[3]
    w = aw * bw - np.dot(av, bv)
    v = aw * bv + bw * av + np.cross(av, bv)
    return 2.0 * np.arctan2(np.linalg.norm(v), abs(w))


@pytest.mark.parametrize("n,H,seed", [(1000, 1024, 3), (1001, 300, 5), (37, 64, 7), (10000, 2048, 9)])
def test_generation_bit_exact(rpe, orc, gpu_ctx, n, H, seed):
    orc.set_math_mode(orc.DET)
    q, t, Q, P = _frame(rpe, seed, n)
    S = rpe.sample_table(seed, n, 3, H)
    ref = orc.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999, full=True, xc=P, xw=Q)
    gpu_ctx.upload(xc=P, xw=Q)
    ns = gpu_ctx.generate(SHINJI, S)
    hyps, valid = gpu_ctx.get_hypotheses(ns)
    ref_valid = (ref["votes"] >= 0).astype(np.int32)
    assert np.array_equal(valid, ref_valid)
    sel = valid == 1
    assert np.array_equal(hyps[sel].view(np.uint32), ref["hyps"][sel].view(np.uint32)), "hypotheses differ in some bit"


@pytest.mark.parametrize("packed", [1, 0])
@pytest.mark.parametrize("n,H,seed,outlier,noise,thr", [
    (1000, 1024, 3, 0.5, 0.1, 0.25),
    (1001, 300, 5, 0.2, 0.05, 0.1),
    (37, 64, 7, 0.0, 0.01, 0.05),
    (20000, 1024, 9, 0.5, 0.1, 0.25),
    (4097, 513, 11, 0.7, 0.2, 0.5),
])
def test_ransac_votes_winner_iter_mask_identical(rpe, orc, gpu_ctx, packed, n, H, seed, outlier, noise, thr):
    orc.set_math_mode(orc.DET)
    rpe.lib.rpe_debug_set_packed(packed)
    q, t, Q, P = _frame(rpe, seed, n, outlier, noise)
    S = rpe.sample_table(seed, n, 3, H)
    ref = orc.ransac(SHINJI, S, thr3d=thr, confidence=0.9999, full=True, xc=P, xw=Q)
    gpu_ctx.upload(xc=P, xw=Q)
    got = gpu_ctx.ransac(SHINJI, S, thr3d=thr, confidence=0.9999)
    votes = gpu_ctx.get_votes(H)
    rpe.lib.rpe_debug_set_packed(1)
    assert got["flags"] == 0
    assert np.array_equal(votes, ref["votes"]), f"votes differ at {np.nonzero(votes != ref['votes'])[0][:10]}"
    assert got["winner"] == ref["winner"]
    assert got["max_votes"] == ref["max_votes"]
    assert got["iter_final"] == ref["iter_final"]
    assert np.array_equal(got["q"].view(np.uint32), ref["q"].view(np.uint32))
    assert np.array_equal(got["t"].view(np.uint32), ref["t"].view(np.uint32))
    assert np.array_equal(got["mask"], ref["mask"])
    assert got["n_inliers"][1] == int(ref["mask"][1].sum()) == ref["max_votes"]


def test_early_stop_loop_equals_full_replay(rpe, orc):
    """The oracle's literal early-stopping loop and its score-everything-then-replay form agree."""
    orc.set_math_mode(orc.DET)
    for seed in range(20, 30):
        q, t = rpe.sim_pose(seed)
        Q, P, _ = rpe.sim_3d_3d(seed + 1, q, t, 500, noise=0.1, outlier_ratio=0.5)
        S = rpe.sample_table(seed, 500, 3, 2000)
        a = orc.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999, full=True, xc=P, xw=Q)
        b = orc.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999, full=False, xc=P, xw=Q)
        for k in ("winner", "max_votes", "iter_final"):
            assert a[k] == b[k]
        assert np.array_equal(a["mask"], b["mask"])


def test_kabsch_refit_within_tolerance(rpe, orc, gpu_ctx):
    orc.set_math_mode(orc.DET)
    n, H = 20000, 512
    q, t, Q, P = _frame(rpe, 31, n)
    S = rpe.sample_table(31, n, 3, H)
    gpu_ctx.upload(xc=P, xw=Q)
    got = gpu_ctx.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999)
    ref_q, ref_t, ok = orc.shinji_ls(P, Q, got["mask"][1], dt=np.float64)
    assert ok
    fit = gpu_ctx.refit("kabsch_inliers")
    assert fit["refit_ok"] == 1
    scale = float(np.abs(P).max())
    assert _angle(fit["q"], ref_q) < 1e-6  # rad
    assert np.abs(fit["t"].astype(np.float64) - ref_t).max() < 1e-6 * scale
    # shinji_ls2 (all points)
    ref_q2, ref_t2, _ = orc.shinji_ls(P, Q, None, dt=np.float64)
    fit2 = gpu_ctx.refit("kabsch_all")
    assert _angle(fit2["q"], ref_q2) < 1e-6
    assert np.abs(fit2["t"].astype(np.float64) - ref_t2).max() < 1e-6 * scale
    # and the refit is close to ground truth (Simulator's known answer)
    assert _angle(fit["q"], q) < 3e-2  # consensus-set bias at thr=2.5 sigma; the oracle shows the same


def test_gn_refinement_matches_twin_and_closed_form(rpe, orc, gpu_ctx):
    orc.set_math_mode(orc.DET)
    n, H = 5000, 256
    q, t, Q, P = _frame(rpe, 41, n)
    S = rpe.sample_table(41, n, 3, H)
    gpu_ctx.upload(xc=P, xw=Q)
    got = gpu_ctx.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999)
    twin_q, twin_t, info = orc.refine_gn(got["q"], got["t"], got["mask"], max_iters=8, xc=P, xw=Q)
    fit = gpu_ctx.refit("gn", max_iters=8)
    assert fit["refit_ok"] == 1 and fit["refit_evals"] >= 2
    scale = float(np.abs(P).max())
    assert _angle(fit["q"], twin_q) < 1e-6
    assert np.abs(fit["t"].astype(np.float64) - twin_t.astype(np.float64)).max() < 1e-6 * scale
    # same least-squares objective as Kabsch on pure 3-D data: converges to the closed form
    ls_q, ls_t, _ = orc.shinji_ls(P, Q, got["mask"][1], dt=np.float64)
    assert _angle(fit["q"], ls_q) < 1e-6
    assert np.abs(fit["t"].astype(np.float64) - ls_t).max() < 1e-5


def test_nan_points_and_invalid_samples(rpe, orc, gpu_ctx):
    """isValid semantics (AOOnlyPoseAdapter.hpp:161-166): all-NaN camera points never vote; a sample that
    hits one consumes the iteration without a hypothesis (AbsoluteOrientation.hpp:185)."""
    orc.set_math_mode(orc.DET)
    n, H = 3000, 512
    q, t, Q, P = _frame(rpe, 51, n)
    rng = np.random.default_rng(0)
    bad = rng.choice(n, n // 5, replace=False)
    P = P.copy()
    P[bad] = np.nan
    P[bad[:50], 1] = 1.0  # partially-NaN points are "valid" for isValid but can never be inliers
    S = rpe.sample_table(51, n, 3, H)
    ref = orc.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999, full=True, xc=P, xw=Q)
    gpu_ctx.upload(xc=P, xw=Q)
    got = gpu_ctx.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999)
    votes = gpu_ctx.get_votes(H)
    assert (ref["votes"] == -1).sum() > 0
    assert np.array_equal(votes, ref["votes"])
    assert got["winner"] == ref["winner"] and got["iter_final"] == ref["iter_final"]
    assert np.array_equal(got["mask"], ref["mask"])


def test_borderline_band_forces_exact_path(rpe, orc, gpu_ctx):
    """Put many residuals exactly around the threshold: the fast path must hand them to the exact fix-up."""
    orc.set_math_mode(orc.DET)
    n = 4096
    q, t = rpe.sim_pose(61)
    Q, P, _ = rpe.sim_3d_3d(62, q, t, n, noise=0.0, outlier_ratio=0.0)
    # displace camera points by exactly thr along random directions (+- a few ulps)
    rng = np.random.default_rng(1)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    thr = 0.25
    P2 = (P.astype(np.float64) + d * thr * (1.0 + rng.integers(-3, 4, size=(n, 1)) * 1e-7)).astype(np.float32)
    P2[:8] = P[:8]  # a few clean points for the sampler below
    S = np.tile(np.array([[0, 1, 2, -1], [3, 4, 5, -1], [1, 5, 7, -1]], np.int32), (11, 1))
    ref = orc.ransac(SHINJI, S, thr3d=thr, confidence=0.99, full=True, xc=P2, xw=Q)
    gpu_ctx.upload(xc=P2, xw=Q)
    got = gpu_ctx.ransac(SHINJI, S, thr3d=thr, confidence=0.99)
    votes = gpu_ctx.get_votes(S.shape[0])
    assert got["n_borderline"] > 1000
    assert np.array_equal(votes, ref["votes"])
    assert np.array_equal(got["mask"], ref["mask"])


def test_config1_iter_100000_runs_in_passes(rpe, orc, gpu_ctx):
    """BASELINE config #1 as SimpleMain.cpp:30-45 runs it: N=1000, 50 % outliers, Iter0 = 100 000, conf 0.9999.
    The library generates and scores 1024, 2048, 4096, 8192, 8192, ... iterations per device pass and stops as soon as
    the replayed adaptive bound is reached; the outcome equals the CPU loop over the same sample stream."""
    orc.set_math_mode(orc.DET)
    n, H = 1000, 100000
    for seed, outlier in [(71, 0.5), (72, 0.9)]:
        q, t, Q, P = _frame(rpe, seed, n, outlier)
        S = rpe.sample_table(1, n, 3, H)
        ref = orc.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999, full=False, xc=P, xw=Q, want_arrays=False)
        gpu_ctx.upload(xc=P, xw=Q)
        got = gpu_ctx.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999)
        assert (got["winner"], got["max_votes"], got["iter_final"]) == (ref["winner"], ref["max_votes"], ref["iter_final"])
        assert np.array_equal(got["mask"], ref["mask"])
        ends, p, tot = set(), 1024, 0
        while tot < H:
            tot = min(H, tot + p)
            ends.add(tot)
            p = min(2 * p, 8192)
        assert got["n_slots"] in ends
        assert got["n_slots"] >= min(H, ref["iter_final"])
        if outlier == 0.5:
            assert got["n_slots"] == 1024  # the bound (a few hundred) is reached inside the first pass
    # with 90 % outliers the bound stays above one pass: several passes were needed
    assert got["n_slots"] > 8192


def test_worklist_overflow_falls_back_to_exact_rescoring(rpe, orc):
    """If more evaluations are borderline than the worklist holds, the frame is rescored in exact order on the device
    (flags bit 0 is set) and the result is still identical to the CPU path."""
    import ctypes
    orc.set_math_mode(orc.DET)
    n, H = 6000, 300
    q, t, Q, P = _frame(rpe, 91, n)
    S = rpe.sample_table(91, n, 3, H)
    ref = orc.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999, full=True, xc=P, xw=Q)
    with rpe.Context(0) as ctx:
        rpe.lib.rpe_debug_set_worklist_capacity.argtypes = [ctypes.c_void_p, ctypes.c_uint]
        rpe.lib.rpe_debug_set_worklist_capacity(ctx.handle, 3)
        ctx.upload(xc=P, xw=Q)
        got = ctx.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999)
        votes = ctx.get_votes(H)
        assert got["flags"] & 1
        assert np.array_equal(votes, ref["votes"])
        assert got["winner"] == ref["winner"] and got["iter_final"] == ref["iter_final"]
        assert np.array_equal(got["mask"], ref["mask"])
        # and the next frame on the same context is clean again
        rpe.lib.rpe_debug_set_worklist_capacity(ctx.handle, 1 << 21)
        got2 = ctx.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999)
        assert got2["flags"] == 0 and np.array_equal(ctx.get_votes(H), ref["votes"])


def test_async_queue_longer_than_the_staging_ring(rpe, orc, gpu_ctx):
    """More results in flight than pinned staging slots (256): the oldest are delivered early, none is lost."""
    n, H = 600, 32
    q, t, Q, P = _frame(rpe, 77, n)
    gpu_ctx.upload(xc=P, xw=Q)
    tables = [rpe.sample_table(1000 + i, n, 3, H) for i in range(8)]
    want = [gpu_ctx.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999, want_mask=False) for S in tables]
    pend = []
    for i in range(300):
        pend.append((i % 8, gpu_ctx.ransac_async(SHINJI, tables[i % 8], thr3d=0.25, confidence=0.9999)))
        pend.append((-1, gpu_ctx.refit_async("kabsch_inliers")))
    gpu_ctx.sync()
    for k, r in pend:
        if k >= 0:
            assert r.max_votes == want[k]["max_votes"] and r.winner == want[k]["winner"]
            assert r.iter_final == want[k]["iter_final"]
        else:
            assert r.refit_ok == 1
    gpu_ctx._keep = []


def test_ransac_stream_draws_rows_pass_by_pass(rpe, orc, gpu_ctx):
    """rpe_ransac_stream asks for the rows of one pass at a time (1024, 2048, ...) and stops asking once the
    replayed adaptive bound ends the loop; same result as the whole-table call."""
    orc.set_math_mode(orc.DET)
    n, H = 1500, 100000
    q, t, Q, P = _frame(rpe, 61, n)
    S = rpe.sample_table(5, n, 3, H)
    gpu_ctx.upload(xc=P, xw=Q)
    want = gpu_ctx.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999)
    got = gpu_ctx.ransac_stream(SHINJI, lambda first, count: S[first:first + count], H, thr3d=0.25, confidence=0.9999)
    for k in ("winner", "max_votes", "iter_final", "n_slots"):
        assert got[k] == want[k], k
    assert np.array_equal(got["mask"], want["mask"])
    assert got["passes"] == [(0, 1024)]  # the bound (a few hundred iterations) is reached inside the first pass
    ref = orc.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999, full=False, xc=P, xw=Q, want_arrays=False)
    assert (got["winner"], got["iter_final"]) == (ref["winner"], ref["iter_final"])
    # iterations the reference's loop executes, as the drop-in headers reconstruct them from the result
    assert ref["iters_run"] == max(got["iter_final"], got["winner"] + 1)
    # 92 % outliers: the bound stays in the thousands, several passes are requested in order
    q, t, Q, P = _frame(rpe, 62, n, 0.92)
    gpu_ctx.upload(xc=P, xw=Q)
    got = gpu_ctx.ransac_stream(SHINJI, lambda first, count: S[first:first + count], H, thr3d=0.25, confidence=0.9999)
    ref = orc.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999, full=False, xc=P, xw=Q, want_arrays=False)
    assert (got["winner"], got["max_votes"], got["iter_final"]) == (ref["winner"], ref["max_votes"], ref["iter_final"])
    assert got["passes"][:3] == [(0, 1024), (1024, 2048), (3072, 4096)]
    assert ref["iters_run"] == max(got["iter_final"], got["winner"] + 1)
    # a shorter first pass changes the schedule, not the result
    gpu_ctx.set_first_pass_iters(128)
    got2 = gpu_ctx.ransac_stream(SHINJI, lambda first, count: S[first:first + count], H, thr3d=0.25, confidence=0.9999)
    gpu_ctx.set_first_pass_iters(1024)
    assert got2["passes"][:4] == [(0, 128), (128, 256), (384, 512), (896, 1024)]
    assert (got2["winner"], got2["max_votes"], got2["iter_final"]) == (got["winner"], got["max_votes"], got["iter_final"])
    assert np.array_equal(got2["mask"], got["mask"])


def test_c_abi_error_behaviour(rpe):
    """The reference asserts / aborts; the C-ABI returns a status, keeps a message and leaves the context usable."""
    import ctypes as C
    with rpe.Context(0) as ctx:
        S = rpe.sample_table(1, 100, 3, 8)
        res = rpe.capi._Result()
        # nothing uploaded yet
        assert rpe.lib.rpe_ransac(ctx.handle, 0, S.ctypes.data, 8, 0.25, 0.0, 0.0, 0.99, C.byref(res), None) != 0
        assert b"" != rpe.lib.rpe_last_error(ctx.handle)
        assert rpe.lib.rpe_refit(ctx.handle, 0, None, 0, C.byref(res)) != 0
        q, t, Q, P = _frame(rpe, 5, 100)
        ctx.upload(xc=P, xw=Q)
        for bad in [dict(method=99), dict(H=0), dict(samples=None)]:
            m = bad.get("method", 0)
            H = bad.get("H", 8)
            sp = None if "samples" in bad else S.ctypes.data
            assert rpe.lib.rpe_ransac(ctx.handle, m, sp, H, 0.25, 0.0, 0.0, 0.99, C.byref(res), None) != 0
        # a 2-D method without bearing vectors is a state error, not a crash
        with pytest.raises(rpe.RpeError):
            ctx.ransac("kneip", rpe.sample_table(1, 100, 4, 8), cos_thr2d=0.999)
        # GN before any mask exists
        with pytest.raises(rpe.RpeError):
            ctx.refit("gn")
        # and the context still works
        r = ctx.ransac("shinji", S, thr3d=0.25, confidence=0.99)
        assert r["winner"] >= 0 and r["max_votes"] > 10
    assert rpe.lib.rpe_ransac(None, 0, None, 0, 0.0, 0.0, 0.0, 0.0, None, None) != 0


@pytest.mark.parametrize("n", [2, 3, 5, 37, 1001, 4098, 20003])
def test_raw_array_scorer_matches_packed_and_oracle(rpe, orc, gpu_ctx, n):
    """The 3-D / 3-D scorer streams the caller's arrays through bulk TMA (no packed copy) when they are 16-byte
    aligned; ragged sizes exercise the 0..3-correspondence tail that bypasses TMA. Same votes as the packed path."""
    orc.set_math_mode(orc.DET)
    H = 200
    q, t, Q, P = _frame(rpe, 300 + n, max(n, 3))
    Q, P = Q[:n].copy(), P[:n].copy()
    if n > 10:
        P[n // 2] = np.nan            # an invalid camera point
        Q[n // 3, 1] = np.float32(3e4)  # a far-away world point widens that stage's guard band only
    rng = np.random.default_rng(n)
    S = np.stack([rng.choice(n, 3, replace=n < 3) if n >= 3 else rng.integers(0, n, 3) for _ in range(H)]).astype(np.int32)
    S = np.concatenate([S, -np.ones((H, 1), np.int32)], axis=1)
    ref = orc.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999, full=True, xc=P, xw=Q)
    out = {}
    for raw in (1, 0):
        rpe.lib.rpe_debug_set_raw_tiles(raw)
        gpu_ctx.upload(xc=P, xw=Q)
        got = gpu_ctx.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999)
        out[raw] = (gpu_ctx.get_votes(H).copy(), got)
    rpe.lib.rpe_debug_set_raw_tiles(1)
    for raw in (1, 0):
        votes, got = out[raw]
        assert np.array_equal(votes, ref["votes"]), raw
        assert (got["winner"], got["max_votes"], got["iter_final"]) == (ref["winner"], ref["max_votes"], ref["iter_final"])
        assert np.array_equal(got["mask"], ref["mask"])


def test_misaligned_device_arrays_fall_back_to_the_packed_path(rpe, orc):
    import torch
    orc.set_math_mode(orc.DET)
    n, H = 5000, 128
    q, t, Q, P = _frame(rpe, 55, n)
    S = rpe.sample_table(55, n, 3, H)
    ref = orc.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999, full=True, xc=P, xw=Q)
    dev = torch.device("cuda", 0)
    buf_w = torch.zeros(3 * n + 1, dtype=torch.float32, device=dev)
    buf_c = torch.zeros(3 * n + 1, dtype=torch.float32, device=dev)
    buf_w[1:] = torch.from_numpy(Q.reshape(-1)).to(dev)   # 4-byte offset: not 16-byte aligned
    buf_c[1:] = torch.from_numpy(P.reshape(-1)).to(dev)
    torch.cuda.synchronize()
    with rpe.Context(0) as ctx:
        ctx.upload_device(n, xc=buf_c.data_ptr() + 4, xw=buf_w.data_ptr() + 4)
        got = ctx.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999)
        assert np.array_equal(ctx.get_votes(H), ref["votes"])
        assert np.array_equal(got["mask"], ref["mask"])


def test_full_size_frame_properties(rpe, orc, gpu_ctx):
    """BASELINE config #4 at full size (307 200 correspondences x 1 024 hypotheses), against the oracle and through
    size-independent properties:
      * the tiled scorer's vote table equals the exact-order kernel's (same device, independent code path: the
        worklist is shrunk to nothing, every borderline evaluation overflows it, the frame is rescored exactly);
      * vote counts are invariant under a permutation of the correspondences (with the sample indices remapped);
      * the winner's vote count equals the number of set flags in its mask; the replayed Iter matches the host rule;
      * the CPU oracle agrees on the full vote table (all 1 024 hypotheses), the hypotheses, the winner and the mask."""
    import ctypes
    from rgbd_pose_estimation_b200 import sharding
    orc.set_math_mode(orc.DET)
    n, H = 307200, 1024
    q, t, Q, P = _frame(rpe, 4242, n)
    S = rpe.sample_table(4242, n, 3, H)
    gpu_ctx.upload(xc=P, xw=Q)
    got = gpu_ctx.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999)
    votes = gpu_ctx.get_votes(H).copy()
    assert got["flags"] == 0 and got["n_slots"] == H
    assert got["max_votes"] == int(votes.max()) == int(got["mask"][1].sum()) == got["n_inliers"][1]
    win, best, it = sharding.replay(votes, 0, n, 0.9999)
    assert (win, best, it) == (got["winner"], got["max_votes"], got["iter_final"])
    # the CPU oracle on the WHOLE frame: all 1 024 hypotheses x 307 200 correspondences (3.1e8 evaluations in the
    # reference's operation order, hypotheses sharded over the host cores: about half a second on the GPU box)
    import os
    ref = orc.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999, full=True, nthreads=os.cpu_count() or 1, xc=P, xw=Q)
    assert np.array_equal(votes, ref["votes"]), f"votes differ at {np.nonzero(votes != ref['votes'])[0][:10]}"
    assert (got["winner"], got["max_votes"], got["iter_final"]) == (ref["winner"], ref["max_votes"], ref["iter_final"])
    assert np.array_equal(got["mask"], ref["mask"])
    assert np.array_equal(got["q"].view(np.uint32), ref["q"].view(np.uint32))
    assert np.array_equal(got["t"].view(np.uint32), ref["t"].view(np.uint32))
    gpu_ctx.generate(SHINJI, S)
    hyps, valid = gpu_ctx.get_hypotheses(H)
    sel = valid == 1
    assert np.array_equal(hyps[sel].view(np.uint32), ref["hyps"][sel].view(np.uint32))
    # the exact-order kernel scores the whole frame (worklist capacity 0 forces the overflow -> exact path)
    rpe.lib.rpe_debug_set_worklist_capacity.argtypes = [ctypes.c_void_p, ctypes.c_uint]
    rpe.lib.rpe_debug_set_worklist_capacity(gpu_ctx.handle, 0)
    got_exact = gpu_ctx.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999)
    rpe.lib.rpe_debug_set_worklist_capacity(gpu_ctx.handle, 1 << 30)
    assert got_exact["flags"] & 1
    assert np.array_equal(gpu_ctx.get_votes(H), votes)
    assert np.array_equal(got_exact["mask"], got["mask"])
    # permutation invariance
    rng = np.random.default_rng(7)
    perm = rng.permutation(n)
    inv = np.empty(n, np.int64)
    inv[perm] = np.arange(n)
    S2 = S.copy()
    S2[:, :3] = inv[S[:, :3]]
    gpu_ctx.upload(xc=np.ascontiguousarray(P[perm]), xw=np.ascontiguousarray(Q[perm]))
    got2 = gpu_ctx.ransac(SHINJI, S2, thr3d=0.25, confidence=0.9999)
    assert np.array_equal(gpu_ctx.get_votes(H), votes)
    assert (got2["winner"], got2["max_votes"], got2["iter_final"]) == (got["winner"], got["max_votes"], got["iter_final"])
    assert np.array_equal(got2["mask"][1][inv], got["mask"][1])


def test_sample_rows_outside_the_frame_give_empty_slots(rpe, orc, gpu_ctx):
    """A caller-made table with an index < 0 or >= n (the reference's samplers cannot produce one): that iteration is an
    empty slot (votes -1) instead of an out-of-bounds read; the other iterations are unaffected."""
    orc.set_math_mode(orc.DET)
    n, H = 500, 64
    q, t, Q, P = _frame(rpe, 66, n)
    S = rpe.sample_table(66, n, 3, H)
    ref = orc.ransac(SHINJI, S, thr3d=0.25, confidence=0.9999, full=True, xc=P, xw=Q)
    bad = S.copy()
    bad[5, 1] = n
    bad[17, 0] = -3
    bad[40, 2] = 2 ** 30
    gpu_ctx.upload(xc=P, xw=Q)
    got = gpu_ctx.ransac(SHINJI, bad, thr3d=0.25, confidence=0.9999)
    votes = gpu_ctx.get_votes(H)
    assert list(votes[[5, 17, 40]]) == [-1, -1, -1]
    keep = np.ones(H, bool)
    keep[[5, 17, 40]] = False
    assert np.array_equal(votes[keep], ref["votes"][keep])
    assert got["winner"] >= 0
