"""GPU: rpe_set_mask_transfer(1) — asynchronous calls send the inlier matrix as one bit per flag and the collecting
thread (rpe_sync / rpe_poll) expands it in place into the reference's n x cols matrix of 16-bit flags
(setInlier layout, /root/reference/pose/PnPPoseAdapter.hpp:196-237). The caller's buffer must hold exactly what the
plain 16-bit copy delivers, for every column count, for frame sizes that are not multiples of 32, and when no
hypothesis is accepted (all-ones matrix)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pinned(rpe, a):
    b = rpe.pinned_empty(a.shape, a.dtype)
    b[:] = a
    return b


def _frame(rpe, n, seed):
    q, t = rpe.sim_pose(seed)
    d = rpe.sim_2d_3d_nl(seed + 1, q, t, n, n2d=1.0, or2d=0.3, n3d=0.05, or3d=0.3, nnl=float(np.deg2rad(2.0)), ornl=0.3)
    return {k: _pinned(rpe, d[k]) for k in ("bv", "xc", "nc", "xw", "nw")}


TH = dict(thr3d=0.2, cos_thr2d=float(np.cos(np.arctan(np.float32(8.0) / np.float32(585.0)))), cos_thrN=float(np.cos(np.float32(0.1))))


@pytest.mark.parametrize("method,n", [("shinji", 307200), ("shinji", 131077), ("kneip", 140001), ("nl_shinji_kneip", 150031),
                                       ("shinji_kneip", 131072 + 31)])
def test_bit_form_masks_equal_the_plain_copy(rpe, method, n):
    cols = {"shinji": 2, "kneip": 1, "shinji_kneip": 2, "nl_shinji_kneip": 3}[method]
    H = 256
    f = _frame(rpe, n, 900 + n % 97)
    S = _pinned(rpe, rpe.sample_table(3, n, rpe.method_sample_size(rpe.METHODS[method]), H))
    m_plain = rpe.pinned_empty((cols, n), np.int16)
    m_bits = rpe.pinned_empty((cols, n), np.int16)
    with rpe.Context(0) as a, rpe.Context(0) as b:
        b.set_mask_transfer(1)
        for ctx, m in ((a, m_plain), (b, m_bits)):
            m[:] = -7
            ctx.upload_async(**f)
            r = ctx.ransac_async(method, S, confidence=0.99, mask=m, **TH)
            k = ctx.refit_async("gn", max_iters=2)
            ctx.sync()
            assert r.winner >= 0
        assert np.array_equal(np.asarray(m_plain), np.asarray(m_bits))
        assert set(np.unique(np.asarray(m_bits))) <= {0, 1}
        # several frames in flight on one context, collected with poll() while later frames are still running
        masks = [rpe.pinned_empty((cols, n), np.int16) for _ in range(4)]
        res = []
        for m in masks:
            m[:] = -7
            b.upload_async(**f)
            res.append(b.ransac_async(method, S, confidence=0.99, mask=m, **TH))
            b.poll()
        b.sync()
        for m in masks:
            assert np.array_equal(np.asarray(m), np.asarray(m_plain))
        # a blocking call on the same context still copies the matrix itself
        b.upload(**{k: np.asarray(v) for k, v in f.items()})
        w = b.ransac(method, np.asarray(S), confidence=0.99, **TH)
        assert np.array_equal(w["mask"], np.asarray(m_plain))


def test_bit_form_mask_when_nothing_is_accepted(rpe):
    n, H = 140003, 64
    Q = np.full((n, 3), np.nan, np.float32)   # no valid camera point: no hypothesis, the adapters keep setOnes()
    P = np.full((n, 3), np.nan, np.float32)
    S = rpe.sample_table(5, n, 3, H)
    m0 = rpe.pinned_empty((2, n), np.int16)
    m1 = rpe.pinned_empty((2, n), np.int16)
    with rpe.Context(0) as a, rpe.Context(0) as b:
        b.set_mask_transfer(1)
        for ctx, m in ((a, m0), (b, m1)):
            m[:] = -7
            ctx.upload(xc=P, xw=Q)
            r = ctx.ransac_async("shinji", S, thr3d=0.25, confidence=0.99, mask=m)
            ctx.sync()
            assert r.winner < 0
    assert np.array_equal(np.asarray(m0), np.asarray(m1))
    assert np.asarray(m1).min() == 1


@pytest.mark.parametrize("method,n", [("shinji", 307200), ("shinji", 131077), ("nl_shinji", 150031), ("shinji_kneip", 140001)])
def test_constant_column_stays_on_the_device(rpe, method, n):
    """rpe_set_mask_transfer(2): column 0 of a family without the 2-D test is written by the collecting thread, the other
    columns cross the bus; families with the 2-D test are copied whole. Same matrix as the plain copy, blocking and asynchronous."""
    cols = {"shinji": 2, "nl_shinji": 3, "shinji_kneip": 2}[method]
    H = 256
    f = _frame(rpe, n, 700 + n % 89)
    S = _pinned(rpe, rpe.sample_table(3, n, rpe.method_sample_size(rpe.METHODS[method]), H))
    m_plain = rpe.pinned_empty((cols, n), np.int16)
    m_skip = rpe.pinned_empty((cols, n), np.int16)
    with rpe.Context(0) as a, rpe.Context(0) as b:
        b.set_mask_transfer(2)
        for ctx, m in ((a, m_plain), (b, m_skip)):
            m[:] = -7
            ctx.upload_async(**f)
            r = ctx.ransac_async(method, S, confidence=0.99, mask=m, **TH)
            ctx.refit_async("gn", max_iters=2)
            ctx.sync()
            assert r.winner >= 0
        assert np.array_equal(np.asarray(m_plain), np.asarray(m_skip))
        masks = [rpe.pinned_empty((cols, n), np.int16) for _ in range(3)]
        for m in masks:
            m[:] = -7
            b.upload_async(**f)
            b.ransac_async(method, S, confidence=0.99, mask=m, **TH)
            b.poll()
        b.sync()
        for m in masks:
            assert np.array_equal(np.asarray(m), np.asarray(m_plain))
        b.upload(**{k: np.asarray(v) for k, v in f.items()})
        w = b.ransac(method, np.asarray(S), confidence=0.99, **TH)   # blocking call
        assert np.array_equal(w["mask"], np.asarray(m_plain))


def test_constant_column_when_nothing_is_accepted(rpe):
    n, H = 140003, 64
    Q = np.full((n, 3), np.nan, np.float32)
    P = np.full((n, 3), np.nan, np.float32)
    S = rpe.sample_table(5, n, 3, H)
    m = rpe.pinned_empty((2, n), np.int16)
    with rpe.Context(0) as b:
        b.set_mask_transfer(2)
        m[:] = -7
        b.upload(xc=P, xw=Q)
        r = b.ransac_async("shinji", S, thr3d=0.25, confidence=0.99, mask=m)
        b.sync()
        assert r.winner < 0
    assert np.asarray(m).min() == 1 and np.asarray(m).max() == 1
