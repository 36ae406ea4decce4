"""GPU: rpe_set_upload_overlap — the frame is uploaded in chunks from page-locked host memory while the generator reads
its sample points from the host arrays and the scorer runs chunk by chunk; with chunks = 1 nothing is uploaded at all: the
scorer streams the frame from the host arrays with bulk TMA and leaves the device copy behind (H = 2048 falls back to the
plain upload: the frame would cross the bus once per hypothesis column). Results must be what the plain path and the
CPU oracle give (reference loop: /root/reference/pose/AbsoluteOrientation.hpp:101-213)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pinned(rpe, a):
    b = rpe.pinned_empty(a.shape, a.dtype)
    b[:] = a
    return b


@pytest.mark.parametrize("n,H,chunks", [(307200, 1024, 4), (100003, 600, 3), (70000, 1024, 8), (200000, 2048, 2),
                                        (307200, 1024, 1), (100003, 600, 1), (70001, 1024, 1), (200000, 2048, 1)])
def test_overlapped_upload_gives_identical_results(rpe, orc, n, H, chunks):
    import os
    orc.set_math_mode(orc.DET)
    frames = []
    for i in range(3):
        q, t = rpe.sim_pose(500 + i)
        Q, P, _ = rpe.sim_3d_3d(600 + i, q, t, n, noise=0.1, outlier_ratio=0.5)
        if i == 1:
            P[::37] = np.nan  # invalid depth: never votes, never borderline
        frames.append((_pinned(rpe, Q), _pinned(rpe, P), _pinned(rpe, rpe.sample_table(700 + i, n, 3, H))))
    mask_a = rpe.pinned_empty((2, n), np.int16)
    with rpe.Context(0) as plain, rpe.Context(0) as over:
        over.set_upload_overlap(chunks)
        over.set_first_pass_iters(2048)
        plain.set_first_pass_iters(2048)
        for rep in range(2):          # frames back to back on one context: the next upload must wait for the previous readers
            pend = []
            for Q, P, S in frames:
                over.upload_async(xc=P, xw=Q)
                r = over.ransac_async("shinji", S, thr3d=0.25, confidence=0.9999, mask=mask_a)
                k = over.refit_async("kabsch_inliers")
                pend.append((r, k))
            over.sync()
            for (Q, P, S), (r, k) in zip(frames, pend):
                plain.upload(xc=np.asarray(P), xw=np.asarray(Q))
                w = plain.ransac("shinji", np.asarray(S), thr3d=0.25, confidence=0.9999)
                wk = plain.refit("kabsch_inliers")
                # (n_borderline is a diagnostic and depends on how the frame is cut into scorer stages: the guard band uses the
                # magnitude bound of each stage; every borderline evaluation is resolved exactly either way)
                assert (r.winner, r.max_votes, r.iter_final) == (w["winner"], w["max_votes"], w["iter_final"])
                assert np.array_equal(np.array(k.q, np.float32).view(np.uint32), wk["q"].view(np.uint32))
        # one frame in full detail: vote table and mask against the plain path and the oracle
        Q, P, S = frames[1]
        over.upload_async(xc=P, xw=Q)
        r = over.ransac_async("shinji", S, thr3d=0.25, confidence=0.9999, mask=mask_a)
        over.sync()
        votes = over.get_votes(H)
        plain.upload(xc=np.asarray(P), xw=np.asarray(Q))
        w = plain.ransac("shinji", np.asarray(S), thr3d=0.25, confidence=0.9999)
        assert np.array_equal(votes, plain.get_votes(H))
        assert np.array_equal(np.asarray(mask_a), w["mask"])
        if n <= 110000:
            ref = orc.ransac(0, np.asarray(S), thr3d=0.25, confidence=0.9999, full=True, nthreads=os.cpu_count() or 1,
                             xc=np.asarray(P), xw=np.asarray(Q))
            assert np.array_equal(votes, ref["votes"]) and np.array_equal(np.asarray(mask_a), ref["mask"])
        # a second estimator call on the same upload takes the plain path and agrees
        r2 = over.ransac("shinji", np.asarray(S), thr3d=0.25, confidence=0.9999)
        assert (r2["winner"], r2["max_votes"]) == (w["winner"], w["max_votes"]) and np.array_equal(r2["mask"], w["mask"])
        # pageable arrays, other families and small frames are simply copied the plain way
        over.upload(xc=np.asarray(P)[:5000].copy(), xw=np.asarray(Q)[:5000].copy())
        r3 = over.ransac("shinji", rpe.sample_table(1, 5000, 3, 64), thr3d=0.25, confidence=0.99)
        assert r3["winner"] >= 0
