"""GPU: rpe_seq_* — a batched sequence of frames issued by native threads (BASELINE config #5) gives, frame by frame,
what the CPU path gives for that frame with the sample table rpe_sample_table(seed + frame_index)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

THR, CONF, H = 0.25, 0.999, 128


def _frames(rpe, count, n):
    out = []
    for i in range(count):
        q, t = rpe.sim_pose(100 + i)
        Q, P, _ = rpe.sim_3d_3d(200 + i, q, t, n, noise=0.1, outlier_ratio=0.4 + 0.02 * (i % 5))
        out.append({"xw": Q, "xc": P})
    return out


def _pinned_copy(rpe, a):
    b = rpe.pinned_empty(a.shape, a.dtype)
    b[:] = a
    return b


@pytest.mark.parametrize("contexts,threads", [(1, 1), (3, 2), (5, 3)])
def test_sequence_matches_oracle_per_frame(rpe, orc, contexts, threads):
    orc.set_math_mode(orc.DET)
    n, ring, total, seed = 6000, 7, 23, 77
    frames = _frames(rpe, ring, n)
    host = [{"xw": _pinned_copy(rpe, f["xw"]), "xc": _pinned_copy(rpe, f["xc"]),
             "mask": rpe.pinned_empty((2, n), np.int16)} for f in frames]
    with rpe.Sequence(0, "shinji", H, thr3d=THR, confidence=CONF, refit=("kabsch",), sample_seed=seed,
                      contexts=contexts, threads=threads) as seq:
        seq.set_frames(host)
        first = 5  # a sequence need not start at frame 0
        r0, r1 = seq.run(first, total)
        for i in range(total):
            fi = first + i
            f = frames[fi % ring]
            S = rpe.sample_table(seed + fi, n, 3, H)
            ref = orc.ransac(0, S, thr3d=THR, confidence=CONF, full=True, xc=f["xc"], xw=f["xw"])
            assert (r0[i].winner, r0[i].max_votes, r0[i].iter_final) == (ref["winner"], ref["max_votes"], ref["iter_final"]), i
            ls_q, ls_t, ok = orc.shinji_ls(f["xc"], f["xw"], ref["mask"][1], dt=np.float64)
            assert ok and r1[i].refit_ok == 1
            assert np.abs(np.array(r1[i].t) - ls_t).max() < 1e-4
            assert np.abs(np.abs(np.array(r1[i].q)) - np.abs(ls_q)).max() < 1e-5
        # masks: one pass in which every ring slot (and its mask buffer) is used by exactly one frame — frames that share
        # a buffer may be in flight on different contexts at once, and their copies are not ordered against each other
        first2 = 40
        seq.run(first2, ring)
        for i in range(ring):
            fi = first2 + i
            slot = fi % ring
            S = rpe.sample_table(seed + fi, n, 3, H)
            ref = orc.ransac(0, S, thr3d=THR, confidence=CONF, full=True, xc=frames[slot]["xc"], xw=frames[slot]["xw"])
            assert np.array_equal(host[slot]["mask"], ref["mask"]), slot


def test_sequence_device_frames_explicit_tables_and_gn(rpe, orc):
    """Device-resident frames (rpe_sim_3d_3d_device_to), caller-supplied tables, Kabsch + LM refits; a second run on
    the same sequence object continues to work; results do not depend on threads / contexts."""
    torch = pytest.importorskip("torch")
    orc.set_math_mode(orc.DET)
    n, count = 20000, 6
    dev = torch.device("cuda", 0)
    xw = torch.empty((count, n, 3), dtype=torch.float32, device=dev)
    xc = torch.empty((count, n, 3), dtype=torch.float32, device=dev)
    with rpe.Context(0) as c:
        for i in range(count):
            q, t = rpe.sim_pose(300 + i)
            c.sim_3d_3d_device_to(400 + i, q, t, n, xw[i].data_ptr(), xc[i].data_ptr(), noise=0.05, outlier_ratio=0.3)
        c.sync()
    tables = [_pinned_copy(rpe, rpe.sample_table(900 + i, n, 3, H)) for i in range(count)]
    frames = [{"xw": xw[i].data_ptr(), "xc": xc[i].data_ptr(), "n": n, "samples": tables[i]} for i in range(count)]
    got = []
    for contexts, threads in [(2, 1), (4, 4)]:
        with rpe.Sequence(0, "shinji", H, thr3d=THR, confidence=CONF, refit=("kabsch", "gn"), gn_iters=4,
                          contexts=contexts, threads=threads) as seq:
            seq.set_frames(frames)
            r0, r1 = seq.run(0, count)
            r0b, r1b = seq.run(0, count)
            for i in range(count):
                assert (r0[i].winner, r0[i].max_votes) == (r0b[i].winner, r0b[i].max_votes)
                assert np.array_equal(np.array(r1[i].q), np.array(r1b[i].q))
            got.append([(r0[i].winner, r0[i].max_votes, r0[i].iter_final, tuple(r1[i].q), tuple(r1[i].t)) for i in range(count)])
            for i in range(count):
                P, Q = xc[i].cpu().numpy(), xw[i].cpu().numpy()
                ref = orc.ransac(0, tables[i], thr3d=THR, confidence=CONF, full=True, xc=P, xw=Q)
                assert (r0[i].winner, r0[i].max_votes, r0[i].iter_final) == (ref["winner"], ref["max_votes"], ref["iter_final"])
                assert r1[i].refit_ok == 1 and r1[i].refit_evals >= 1
                # LM over pure 3-D rows converges to the Kabsch closed form (same least-squares objective)
                ls_q, ls_t, _ = orc.shinji_ls(P, Q, ref["mask"][1], dt=np.float64)
                assert np.abs(np.array(r1[i].t) - ls_t).max() < 1e-4
    assert got[0] == got[1]


def test_sequence_errors_are_reported(rpe):
    with pytest.raises(rpe.RpeError):
        rpe.Sequence(0, "shinji", 0)
    with rpe.Sequence(0, "kneip", 64, cos_thr2d=0.999, contexts=1, threads=1, refit=()) as seq:
        n = 500
        q, t = rpe.sim_pose(1)
        Q, P, _ = rpe.sim_3d_3d(2, q, t, n)
        seq.set_frames([{"xw": _pinned_copy(rpe, Q), "xc": _pinned_copy(rpe, P)}])  # kneip needs bearing vectors
        with pytest.raises(rpe.RpeError) as e:
            seq.run(0, 2)
        assert "bearing" in str(e.value)


def test_shared_counter_hands_every_frame_out_once(rpe, orc):
    """rpe_seq_run_shared: two sequences (here on one GPU, two host threads) advance one counter; together they process
    every frame exactly once, and each frame's result is what the oracle gives for that frame's own sample table."""
    import threading
    orc.set_math_mode(orc.DET)
    n, ring, total, seed = 8000, 5, 61, 300
    frames = _frames(rpe, ring, n)
    host = [{"xw": _pinned_copy(rpe, f["xw"]), "xc": _pinned_copy(rpe, f["xc"])} for f in frames]
    counter = np.zeros(1, np.int64)
    out = [None, None]
    seqs = [rpe.Sequence(0, "shinji", H, thr3d=THR, confidence=CONF, refit=("kabsch",), sample_seed=seed, contexts=c, threads=t)
            for c, t in ((3, 2), (2, 1))]
    try:
        for s in seqs:
            s.set_frames(host)

        def work(k):
            out[k] = seqs[k].run_shared(counter, total, total)
        th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        seen = {}
        for k in range(2):
            r0, r1, idx, nd = out[k]
            assert nd >= 1  # both got some work
            for i in range(nd):
                fi = int(idx[i])
                assert fi not in seen
                seen[fi] = (r0[i].winner, r0[i].max_votes, r0[i].iter_final, r1[i].refit_ok)
        assert sorted(seen) == list(range(total)) and int(counter[0]) >= total
        for fi in (0, 7, 33, 60):
            f = frames[fi % ring]
            ref = orc.ransac(0, rpe.sample_table(seed + fi, n, 3, H), thr3d=THR, confidence=CONF, full=True, xc=f["xc"], xw=f["xw"])
            assert seen[fi][:3] == (ref["winner"], ref["max_votes"], ref["iter_final"]) and seen[fi][3] == 1
    finally:
        for s in seqs:
            s.close()
