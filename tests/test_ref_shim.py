"""The oracle restatement against the reference's OWN pose headers (no GPU needed).

oracle/_ref/libref_shim.so is /root/reference/pose/*.hpp — samplers on ::rand(), minimal solvers, the RANSAC / PROSAC
loops, adapters, refits — compiled unmodified from where they lie against the Eigen / Sophus API stand-ins of
oracle/ref_shim/ (Eigen is not installed in this image; the stand-ins carry the arithmetic rules of
oracle/eig_model.hpp and oracle/sophus_model.hpp). Running it next to the oracle on the same inputs and the same
srand() seed pins everything the oracle restates from the reference's sources: the draws, the solver algebra, the
operation order of every inlier test, strict `>` best-keeping, the adaptive `Iter`, mask layout and the refits must
come out bit for bit. What it cannot pin is Eigen's own internals (they are shared by both sides).

Skipped where neither /root/reference nor a prebuilt library exists.
"""
import numpy as np
import pytest

from tests import refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference sources / oracle/_ref/libref_shim.so not present")

F = 585.0
NAMES = {0: "shinji", 1: "kneip", 2: "shinji_kneip", 3: "nl_kneip", 4: "nl_shinji", 5: "nl_shinji_kneip", 6: "kneip_quat"}


def _data(rpe, seed, n, dt, outliers=0.4, nan_every=0):
    q, t = rpe.sim_pose(seed)
    d = rpe.sim_2d_3d_nl(seed + 1, q, t, n, n2d=1.0, or2d=outliers, n3d=0.05, or3d=outliers, nnl=float(np.deg2rad(2.0)),
                         ornl=outliers)
    out = {k: np.ascontiguousarray(d[k]).astype(dt) for k in ("bv", "xc", "nc", "xw", "nw")}
    if dt == np.float64:  # unit vectors to binary64 accuracy: Sophus' ENSURE tolerance is 1e-10 for double
        for k in ("bv", "nc", "nw"):
            out[k] /= np.linalg.norm(out[k], axis=1, keepdims=True)
    if nan_every:  # pixels without depth: the camera point is all-NaN (AOPoseAdapter.hpp:147-152)
        out["xc"][::nan_every] = np.nan
    return out, d["weights"]


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_minimal_solvers_bit_identical(orc, rpe, dt):
    rng = np.random.default_rng(7)
    n_sol = 0
    for it in range(300):
        f5 = rng.normal(size=5) * rng.choice([1.0, 10.0, 0.1])
        assert _same(orc.o4_roots(f5, dt), refshim.o4_roots(f5, dt))  # P3P.hpp:11-60
        d, _ = _data(rpe, 100 + 2 * it, 4, dt, outliers=0.0)
        Xw, bv = d["xw"][:4], d["bv"][:4]
        a, b = orc.kneip_main(Xw[:3], bv[:3], dt), refshim.kneip_main(Xw[:3], bv[:3], dt)  # P3P.hpp:63-232
        assert a[0].shape == b[0].shape and _same(a[0], b[0]) and _same(a[1], b[1])
        n_sol += len(a[0])
        a, b = orc.kneip4(Xw, bv, dt), refshim.kneip4(Xw, bv, dt)  # P3P.hpp:250-294
        assert a[2] == b[2] and (not a[2] or (_same(a[0], b[0]) and _same(a[1], b[1])))
        args = (d["xc"][0], d["nc"][0], d["xc"][1], d["xw"][0], d["nw"][0], d["xw"][1])
        a, b = orc.nl_2p(*args, dt), refshim.nl_2p(*args, dt)  # AbsoluteOrientationNormal.hpp:77-142
        assert _same(a[0], b[0]) and _same(a[1], b[1])
        Xc3 = d["xc"][:3] if it % 3 else rng.normal(size=(3, 3))  # every third: not a rigid motion (SO3 ENSURE path)
        a, b = orc.shinji(d["xw"][:3], Xc3, dt=dt), refshim.shinji(d["xw"][:3], Xc3, dt=dt)  # AbsoluteOrientation.hpp:47-99
        assert _same(a[0], b[0]) and _same(a[1], b[1]) and a[2] == b[2]
        p, ep, K, mi = rng.uniform(0.9, 1.0), rng.uniform(0, 1), int(rng.choice([3, 4])), int(rng.integers(1, 100000))
        assert orc.update_num_iters(p, ep, K, mi, dt) == refshim.update_num_iters(p, ep, K, mi, dt)  # P3P.hpp:296-318
    assert n_sol > 600  # the comparison was not vacuous
    for ep in (0.0, 1.0, 1e-9, 0.5):  # the clamps and the early `return 0`
        for p in (0.0, 1.0, 0.99, 0.9999):
            assert orc.update_num_iters(p, ep, 3, 1000, dt) == refshim.update_num_iters(p, ep, 3, 1000, dt)


def _run_both(orc, rpe, method, seed, dt, n, iters, outliers, nan_every=0, sampler=0, refit=0, use_weights=False,
              thr=(0.2, 8.0, 0.1), conf=0.99):
    d, w3 = _data(rpe, 1000 * method + seed, n, dt, outliers, nan_every)
    thr3d, thr2d, thrN = thr
    ct, cn = refshim.cos_thr(thr2d, F, dt), refshim.cos_nl(thrN, dt)
    m = 3 if method == 0 else 4
    weights = np.ascontiguousarray(w3).astype(dt) if (sampler or use_weights) else None
    if sampler:
        # shinji_prosac sorts the 3-D weights (AOOnlyPoseAdapter.hpp:233-238, column 1); kneip_prosac and
        # shinji_kneip_prosac sort the 2-D weights (PnPPoseAdapter.hpp:239-243, column 0)
        S = orc.prosac_table(seed, n, m, iters, weights[1 if method == 0 else 0])
    else:
        S = orc.sample_table(seed, n, m, iters)
    a = orc.ransac(method, S, thr3d=thr3d, cos_thr=ct, cos_nl=cn, confidence=conf, full=False, dt=dt, **d)
    b = refshim.ransac(method, seed, iters, sampler=sampler, thr3d=thr3d, thr2d=thr2d, focal=F, thrN=thrN, confidence=conf,
                       refit=refit, weights3=weights, dt=dt, **d)
    return d, weights, a, b


def _assert_same_run(a, b, tag):
    cols = a["mask"].shape[0]
    assert b["ensure_failures"] == 0, tag  # a run in which the real Sophus would have aborted proves nothing
    assert a["max_votes"] == b["max_votes"], tag
    assert a["iter_final"] == b["iter_final"], tag
    assert _same(a["q"], b["q"]) and _same(a["t"], b["t"]), tag
    assert np.array_equal(a["mask"], b["mask"][:cols]), tag
    # the `short` index lists of cvtInlier (PnPPoseAdapter.hpp:227-237 ...) have the flag counts as lengths
    flag_rows = {0: [1], 1: [0], 6: [0], 2: [0, 1], 3: [0, 1, 2], 4: [0, 1, 2], 5: [0, 1, 2]}
    for r in flag_rows[tag[1]]:
        if r < cols:
            assert b["n_idx"][r] == int(a["mask"][r].sum()), tag


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("method", range(6))
def test_ransac_loops_bit_identical(orc, rpe, method, dt):
    """shinji_ransac2, kneip_ransac, shinji_kneip_ransac, nl_kneip_ransac, nl_shinji_ransac, nl_shinji_kneip_ransac with
    RandomElements on ::rand(): votes, final Iter, accepted pose and inlier flags."""
    runs = 0
    for seed, n, iters, outliers, nan_every in ((1, 700, 200, 0.4, 0), (2, 900, 300, 0.6, 0), (3, 1100, 400, 0.75, 7),
                                                 (4, 500, 150, 0.2, 3), (5, 1300, 60, 0.5, 0)):
        _, _, a, b = _run_both(orc, rpe, method, seed, dt, n, iters, outliers, nan_every)
        _assert_same_run(a, b, (NAMES[method], method, dt.__name__, seed))
        assert a["max_votes"] > 0
        runs += 1
    assert runs == 5


@pytest.mark.parametrize("method", [0, 2, 6])
def test_prosac_loops_bit_identical(orc, rpe, method):
    """shinji_prosac, shinji_kneip_prosac, kneip_prosac: sortIdx / ProsacSampler / getSortedIdx on ::rand()."""
    for seed, n, iters, outliers in ((1, 700, 200, 0.4), (2, 900, 300, 0.6), (3, 600, 100, 0.3)):
        _, _, a, b = _run_both(orc, rpe, method, seed, np.float32, n, iters, outliers, sampler=1)
        _assert_same_run(a, b, (NAMES[method] + "_prosac", method, "float32", seed))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_refits_bit_identical(orc, rpe, dt):
    """shinji_ls1 / shinji_ls (AbsoluteOrientation.hpp:273-325) and nl_shinji_kneip_ls + find_opt_cc
    (AbsoluteOrientationNormal.hpp:13-46,447-552), the latter with and without setWeights."""
    for seed, n, outliers in ((1, 800, 0.4), (2, 1000, 0.6)):
        d, _, a, b = _run_both(orc, rpe, 0, seed, dt, n, 200, outliers, refit=1)  # shinji_ransac2 + shinji_ls1
        _assert_same_run(a, b, ("shinji+ls1", 0, dt.__name__, seed))
        q, t, ok = orc.shinji_ls(d["xc"], d["xw"], a["mask"][1], dt=dt)
        assert ok and _same(q, b["q_refit"]) and _same(t, b["t_refit"])
        d, _, a, b = _run_both(orc, rpe, 2, seed, dt, n, 200, outliers, refit=1)  # shinji_kneip_ransac + shinji_ls
        _assert_same_run(a, b, ("shinji_kneip+ls", 2, dt.__name__, seed))
        q, t, ok = orc.shinji_ls(d["xc"], d["xw"], a["mask"][1], dt=dt)
        assert ok and _same(q, b["q_refit"]) and _same(t, b["t_refit"])
        for use_w in (False, True):
            d, w, a, b = _run_both(orc, rpe, 5, seed, dt, n, 200, outliers, refit=2, use_weights=use_w)
            _assert_same_run(a, b, ("nl_shinji_kneip+ls", 5, dt.__name__, seed))
            q, t = orc.nl_shinji_kneip_ls(a["q"], a["t"], a["mask"], a["max_votes"], weights3=w, dt=dt, **d)
            assert _same(q, b["q_refit"]) and _same(t, b["t_refit"]), (seed, use_w)
            assert not _same(q, a["q"])  # the refit did move the pose
