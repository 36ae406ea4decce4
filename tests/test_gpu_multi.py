"""GPU: hypothesis-sharded single frame (config #4) across two contexts — on two GPUs when the box has them, otherwise
both "ranks" on device 0 (same code path: peer blocks, exchange kernel, epochs; only the NVLink hop is missing), so the
exchange is covered on a one-GPU box too. The NCCL variant is exercised by bench.py under torchrun."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu(rpe):
    import ctypes
    n = ctypes.c_int(0)
    rpe.lib.rpe_device_count(ctypes.byref(n))
    return n.value


def _second_device(rpe):
    return 1 if _ngpu(rpe) >= 2 else 0


def test_hypothesis_sharded_frame_two_gpus(rpe, orc):
    from rgbd_pose_estimation_b200 import sharding
    orc.set_math_mode(orc.DET)
    n, H = 20000, 1024
    q, t = rpe.sim_pose(3)
    Q, P, _ = rpe.sim_3d_3d(4, q, t, n)
    S = rpe.sample_table(1, n, 3, H)
    ref = orc.ransac(0, S, thr3d=0.25, confidence=0.9999, full=True, xc=P, xw=Q)
    ctxs = [rpe.Context(0), rpe.Context(_second_device(rpe))]
    votes = np.empty(H, np.int32)
    for r, c in enumerate(ctxs):
        c.upload(xc=P, xw=Q)
        c.generate("shinji", S)
        b, e = sharding.slot_range(r, 2, H)
        c.score("shinji", b, e, thr3d=0.25)
        votes[b:e] = c.get_votes(H)[b:e]
    assert np.array_equal(votes, ref["votes"])
    for c in ctxs:
        c.set_votes(votes)
        out = c.finish("shinji", H, thr3d=0.25, confidence=0.9999)
        assert (out["winner"], out["max_votes"], out["iter_final"]) == (ref["winner"], ref["max_votes"], ref["iter_final"])
        assert np.array_equal(out["mask"], ref["mask"])
        c.close()


def test_partial_range_scoring_single_gpu(rpe, orc, gpu_ctx):
    """rpe_score over two disjoint slot ranges on one device == one full pass (the per-rank work of config #4)."""
    orc.set_math_mode(orc.DET)
    n, H = 5000, 700
    q, t = rpe.sim_pose(5)
    Q, P, _ = rpe.sim_3d_3d(6, q, t, n)
    S = rpe.sample_table(1, n, 3, H)
    ref = orc.ransac(0, S, thr3d=0.25, confidence=0.9999, full=True, xc=P, xw=Q)
    gpu_ctx.upload(xc=P, xw=Q)
    gpu_ctx.generate("shinji", S)
    gpu_ctx.score("shinji", 0, 300, thr3d=0.25)
    gpu_ctx.score("shinji", 300, H, thr3d=0.25)
    assert np.array_equal(gpu_ctx.get_votes(H), ref["votes"])
    out = gpu_ctx.finish("shinji", H, thr3d=0.25, confidence=0.9999)
    assert (out["winner"], out["iter_final"]) == (ref["winner"], ref["iter_final"])
    assert np.array_equal(out["mask"], ref["mask"])


def test_contexts_sharing_the_scorer_lane_do_not_interfere(rpe, orc):
    """Several contexts (streams) of one device run frames asynchronously; their tiled scorers are serialised through the
    per-device scorer lane while everything else overlaps. Every result equals the one a lone blocking context gives."""
    n, H, NCTX, FRAMES = 30000, 512, 4, 5
    frames = []
    for i in range(NCTX * FRAMES):
        q, t = rpe.sim_pose(900 + i)
        Q, P, _ = rpe.sim_3d_3d(1900 + i, q, t, n, noise=0.1, outlier_ratio=0.5)
        frames.append((Q, P, rpe.sample_table(2900 + i, n, 3, H)))
    with rpe.Context(0) as solo:
        want = []
        for Q, P, S in frames:
            solo.upload(xc=P, xw=Q)
            r = solo.ransac("shinji", S, thr3d=0.25, confidence=0.9999, want_mask=False)
            k = solo.refit("kabsch_inliers")
            g = solo.refit("gn", max_iters=3)
            want.append((r, k, g))
    ctxs = [rpe.Context(0) for _ in range(NCTX)]
    try:
        got = []
        for i, (Q, P, S) in enumerate(frames):
            c = ctxs[i % NCTX]
            c.upload(xc=P, xw=Q)  # pageable host arrays: the copies are staged before the call returns
            got.append((c.ransac_async("shinji", S, thr3d=0.25, confidence=0.9999), c.refit_async("kabsch_inliers"),
                        c.refit_async("gn", max_iters=3)))
        for c in ctxs:
            c.sync()
        for (r, k, g), (wr, wk, wg) in zip(got, want):
            assert (r.winner, r.max_votes, r.iter_final) == (wr["winner"], wr["max_votes"], wr["iter_final"])
            assert np.array_equal(np.array(k.q, np.float32).view(np.uint32), wk["q"].view(np.uint32))
            assert np.array_equal(np.array(g.q, np.float32).view(np.uint32), wg["q"].view(np.uint32))
            assert np.array_equal(np.array(g.t, np.float32).view(np.uint32), wg["t"].view(np.uint32))
    finally:
        for c in ctxs:
            c.close()


def test_peer_memory_exchange_one_process_two_gpus(rpe, orc):
    """Two contexts of one process on two GPUs, linked with rpe_peer_import_local; frames enqueued asynchronously on
    both before the host waits (a blocking call on one context would wait for a peer nobody has started)."""
    orc.set_math_mode(orc.DET)
    n, H = 20000, 1024
    q, t = rpe.sim_pose(3)
    Q, P, _ = rpe.sim_3d_3d(4, q, t, n)
    S = rpe.sample_table(1, n, 3, H)
    ref = orc.ransac(0, S, thr3d=0.25, confidence=0.9999, full=True, xc=P, xw=Q)
    ctxs = [rpe.Context(0), rpe.Context(_second_device(rpe))]
    try:
        rpe.Context.peer_link_local(ctxs)
        for c in ctxs:
            c.upload(xc=P, xw=Q)
        for rep in range(3):
            res = [c.ransac_sharded("shinji", S, thr3d=0.25, confidence=0.9999, blocking=False) for c in ctxs]
            for c in ctxs:
                c.sync()
            for c, r in zip(ctxs, res):
                assert (r.winner, r.max_votes, r.iter_final) == (ref["winner"], ref["max_votes"], ref["iter_final"])
                assert np.array_equal(c.get_votes(H), ref["votes"])
        for c in ctxs:
            c.peer_status()
    finally:
        for c in ctxs:
            c.close()


def test_peer_memory_exchange_two_processes(rpe):
    """rpe_ransac_sharded with one process per GPU: CUDA IPC mappings of every rank's exchange block, vote slices written
    straight into the peers' tables over NVLink by the exchange kernel. All ranks get the oracle's result."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(root, "tests", "_peer_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    line = [l for l in p.stdout.splitlines() if l.startswith("PEER_RESULT ")]
    assert p.returncode == 0 and line, p.stdout[-2000:] + p.stderr[-2000:]
    ranks = json.loads(line[0][len("PEER_RESULT "):])
    assert len(ranks) == 2
    for a, b in zip(ranks[0]["results"], ranks[1]["results"]):
        assert a["oracle_ok"]
        for k in ("winner", "max_votes", "iter_final", "votes_crc"):
            assert a[k] == b[k], k


def test_exchange_timeout_is_reported_as_comm_error(rpe):
    """A rank whose peer never shows up: the exchange kernel gives up after the configured time-out, the frame has no
    winner, and the blocking call returns RPE_ERR_COMM (-5); the latch is cleared, so the context keeps working."""
    n, H = 2000, 64
    q, t = rpe.sim_pose(3)
    Q, P, _ = rpe.sim_3d_3d(4, q, t, n)
    S = rpe.sample_table(1, n, 3, H)
    a, b = rpe.Context(0), rpe.Context(_second_device(rpe))
    try:
        rpe.Context.peer_link_local([a, b])
        a.peer_set_timeout_ms(50)
        a.upload(xc=P, xw=Q)
        with pytest.raises(rpe.RpeError) as e:
            a.ransac_sharded("shinji", S, thr3d=0.25, confidence=0.99)  # b never calls
        assert "error -5" in str(e.value)
        a.peer_status()  # latch cleared by the report
        # the pair still works afterwards: re-link (re-arms flags and epochs), both ranks run the frame
        rpe.Context.peer_link_local([a, b])
        b.upload(xc=P, xw=Q)
        ra = a.ransac_sharded("shinji", S, thr3d=0.25, confidence=0.99, blocking=False)
        rb = b.ransac_sharded("shinji", S, thr3d=0.25, confidence=0.99, blocking=False)
        a.sync()
        b.sync()
        assert ra.winner == rb.winner >= 0 and ra.max_votes == rb.max_votes > 0
    finally:
        a.close()
        b.close()
