"""MinimalSolvers.hpp (/root/reference/pose/MinimalSolvers.hpp): ev (:49-83) and ms (:10-46).

CPU part: the host templates of include/rpe/solvers_min.h against numpy.linalg.eigvalsh and against the oracle's
restatement (oracle/solvers.hpp sym3_eigenvalues, nl_2p), bit for bit in DET math mode.
GPU part (-m gpu): the batch kernels (one problem per thread) return the host templates' bits."""
import ctypes as C

import numpy as np
import pytest


def _sym_batch(rng, count, scale=1.0):
    A = rng.normal(size=(count, 3, 3)) * scale
    M = (A + A.transpose(0, 2, 1)) / 2
    M[::7] = np.einsum("ni,ij->nij", rng.normal(size=(len(M[::7]), 3)), np.eye(3))  # exactly diagonal ones (:54 branch)
    M[3::11] = M[3::11] @ M[3::11].transpose(0, 2, 1)                                   # positive semi-definite
    return M


def _orc_ev(orc, M, dt):
    suf = "f" if dt == np.float32 else "d"
    fn = getattr(orc.lib, f"orc_sym3_eigenvalues_{suf}")
    out = np.empty((M.shape[0], 3), dt)
    Mc = np.ascontiguousarray(M, dt).reshape(-1, 9)
    for i in range(M.shape[0]):
        e = np.empty(3, dt)
        fn(Mc[i].ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p))
        out[i] = e
    return out


def _ms_inputs(rpe, rng, count):
    """count two-correspondence problems with a known pose: (A, N_A, B) in both frames, M unused."""
    rows, poses = [], []
    for i in range(count):
        q, t = rpe.sim_pose(1000 + i)
        x, y, z, w = [float(v) for v in q]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        Aw, Bw = rng.normal(size=3) * 2, rng.normal(size=3) * 2
        Nw = rng.normal(size=3)
        Nw /= np.linalg.norm(Nw)
        Mw = rng.normal(size=3)
        Mw /= np.linalg.norm(Mw)
        Ac, Bc = R @ Aw + t, R @ Bw + t
        rows.append(np.concatenate([Aw, Bw, Nw, Mw, Ac, Bc, R @ Nw, R @ Mw]))
        poses.append((R, np.asarray(t, np.float64)))
    return np.asarray(rows, np.float32), poses


def test_ev_host_vs_numpy_and_oracle(rpe, orc):
    rng = np.random.default_rng(3)
    orc.set_math_mode(orc.DET)
    for scale in (1.0, 50.0, 1e-2):
        M = _sym_batch(rng, 400, scale)
        ref = np.sort(np.linalg.eigvalsh(M), axis=1)[:, ::-1]
        for dt, tol in ((np.float64, 1e-9), (np.float32, 2e-4)):
            E = rpe.min_ev_host(M, dtype=dt)
            nondiag = (np.abs(M[:, 0, 1]) ** 2 + np.abs(M[:, 0, 2]) ** 2 + np.abs(M[:, 1, 2]) ** 2) >= 1e-5
            # the trigonometric branch returns them sorted, the reference's diagonal branch returns the diagonal as is
            assert (np.diff(E[nondiag], axis=1) <= 1e-6 * scale).all()
            err = np.abs(np.sort(E, axis=1)[:, ::-1] - ref).max(axis=1) / np.maximum(np.abs(ref).max(axis=1), 1e-30)
            assert err[nondiag].max() < tol, (dt, scale, err.max())
            d = ~nondiag
            assert np.array_equal(E[d], np.stack([M[d, 0, 0], M[d, 1, 1], M[d, 2, 2]], axis=1).astype(dt))
            # the product's template and the oracle's restatement of MinimalSolvers.hpp:49-83: identical bits
            O = _orc_ev(orc, M.astype(dt), dt)
            assert np.array_equal(E.view(np.uint32 if dt == np.float32 else np.uint64),
                                  O.view(np.uint32 if dt == np.float32 else np.uint64))


def test_ms_host_is_nl_2p_and_recovers_the_pose(rpe, orc):
    rng = np.random.default_rng(5)
    orc.set_math_mode(orc.DET)
    rows, poses = _ms_inputs(rpe, rng, 64)
    q, t = rpe.min_ms_host(rows)
    good = 0
    for i, (R, tt) in enumerate(poses):
        r = rows[i]
        oq, ot = orc.nl_2p(r[12:15], r[18:21], r[15:18], r[0:3], r[6:9], r[3:6])
        assert np.array_equal(q[i].view(np.uint32), oq.view(np.uint32)) and np.array_equal(t[i].view(np.uint32), ot.view(np.uint32))
        x, y, z, w = [float(v) for v in q[i]]
        Re = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                       [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                       [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        ang = np.arccos(np.clip((np.trace(Re @ R.T) - 1) / 2, -1, 1))
        if ang < 1e-3 and np.abs(t[i] - tt).max() < 1e-2:
            good += 1
    # the solver's rotation about the normal takes the UNSIGNED angle (AbsoluteOrientationNormal.hpp:119-124):
    # about half of the problems come out with the right handedness, the rest is what RANSAC out-votes
    assert 16 <= good <= 56


@pytest.mark.gpu
def test_device_batches_return_the_host_bits(rpe, gpu_ctx):
    rng = np.random.default_rng(7)
    M = np.concatenate([_sym_batch(rng, 5000, 1.0), _sym_batch(rng, 3000, 30.0)]).astype(np.float32)
    Eh = rpe.min_ev_host(M)
    Ed = gpu_ctx.min_ev(M)
    assert np.array_equal(Eh.view(np.uint32), Ed.view(np.uint32))
    ref = np.sort(np.linalg.eigvalsh(M.astype(np.float64)), axis=1)[:, ::-1]
    err = np.abs(np.sort(Ed, axis=1)[:, ::-1] - ref).max(axis=1) / np.abs(ref).max(axis=1)
    assert err.max() < 2e-4
    rows, _ = _ms_inputs(rpe, rng, 300)
    qh, th = rpe.min_ms_host(rows)
    qd, td = gpu_ctx.min_ms(rows)
    assert np.array_equal(qh.view(np.uint32), qd.view(np.uint32)) and np.array_equal(th.view(np.uint32), td.view(np.uint32))
    launches0 = gpu_ctx.launch_count()
    gpu_ctx.min_ev(M[:10])
    assert gpu_ctx.launch_count() == launches0 + 1
