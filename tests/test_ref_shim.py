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


def _Rq(q):
    x, y, z, w = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _stats(d, q, t):
    R = _Rq(q)
    y = d["xw"].astype(np.float64) @ R.T + np.asarray(t, np.float64)
    e3 = np.linalg.norm(d["xc"] - y, axis=1)
    yn = y / np.linalg.norm(y, axis=1, keepdims=True)
    c2 = (yn * d["bv"]).sum(1)
    cn = (d["nc"] * (d["nw"].astype(np.float64) @ R.T)).sum(1)
    return {"in3": float((e3 < 0.2).mean()), "med3": float(np.median(e3)), "in2": float((c2 > np.cos(np.arctan(8 / F))).mean()),
            "inn": float((cn > np.cos(0.1)).mean()), "w": d["weights"].mean(axis=1), "zmin": float(d["xc"][:, 2].min()),
            "zmax": float(d["xc"][:, 2].max()), "facing": float((d["nc"][:, 2] < 0).mean())}


def test_reference_simulator_and_ours_draw_from_the_same_distributions(rpe):
    """Simulator.hpp:158-173,175-233,316-367 run as written (Matrix::Random on ::rand(), std::normal_distribution,
    RandomElements for the outlier positions) next to the library's seeded restatement: same inlier fractions per
    modality, noise level, PROSAC weights, depth range and camera-facing normals (the generators differ, so the
    comparison is statistical)."""
    n = 20000
    ref = refshim.sim(2, 11, n, or2d=0.3, or3d=0.4, ornl=0.2)
    q, t = rpe.sim_pose(11)
    ours = rpe.sim_2d_3d_nl(12, q, t, n, or2d=0.3, or3d=0.4, ornl=0.2)
    a, b = _stats(ref, ref["q"], ref["t"]), _stats(ours, q, t)
    for k, tol in (("in3", 0.015), ("in2", 0.015), ("inn", 0.015), ("med3", 0.01), ("zmin", 0.3), ("zmax", 0.3), ("facing", 0.0)):
        assert abs(a[k] - b[k]) <= tol, (k, a[k], b[k])
    assert abs(a["in3"] - 0.6) < 0.02 and abs(a["in2"] - 0.7) < 0.02 and abs(a["inn"] - 0.8) < 0.03
    assert np.allclose(a["w"], b["w"], rtol=0.06), (a["w"], b["w"])
    assert np.abs(np.linalg.norm(ref["bv"], axis=1) - 1).max() < 1e-6


@pytest.mark.parametrize("config", [1, 2, 3])
def test_baseline_configs_on_reference_simulated_inputs(orc, config):
    """BASELINE.json configs #1-#3 end to end the way SimpleMain / TestMain run them: inputs from the reference's own
    Simulator, pose drawn as SimpleMain.cpp:22-23, the reference's estimator and refit on ::rand() — against the oracle
    on the same arrays. (#3 with 30 000 correspondences: the reference's `short` index lists end at 32 767.)"""
    dt = np.float32
    if config == 1:   # SimpleMain.cpp:30-45,76-83: shinji_ransac2 + shinji_ls1, Iter0 = 100 000, confidence 0.9999
        d = refshim.sim(0, 101, 1000, n3d=0.1, or3d=0.5, dt=dt)
        method, n, iters, thr, conf, refit = 0, 1000, 100000, (0.25, 0.0, 0.0), 0.9999, 1
    elif config == 2:  # kneip_ransac on 10 000 2-D / 3-D correspondences with 70 % outliers
        d = refshim.sim(1, 102, 10000, n2d=1.0, or2d=0.7, dt=dt)
        method, n, iters, thr, conf, refit = 1, 10000, 2000, (0.0, 8.0, 0.0), 0.99, 0
    else:              # nl_shinji_kneip_ransac + nl_shinji_kneip_ls, three modalities (Parameters.yml:15-19 thresholds)
        d = refshim.sim(2, 103, 30000, n2d=1.0, or2d=0.3, n3d=0.05, or3d=0.4, ornl=0.2, dt=dt)
        method, n, iters, thr, conf, refit = 5, 30000, 1024, (0.2, 8.0, 0.1), 0.99, 2
    arrs = {k: d[k] for k in ("bv", "xc", "nc", "xw", "nw")}
    seed = 7 + config
    ct, cn = refshim.cos_thr(thr[1], F, dt) if thr[1] else 0.0, refshim.cos_nl(thr[2], dt) if thr[2] else 0.0
    S = orc.sample_table(seed, n, 3 if method == 0 else 4, min(iters, 4096))  # the loop stops long before 4 096 draws
    a = orc.ransac(method, S, thr3d=thr[0], cos_thr=ct, cos_nl=cn, confidence=conf, full=False, dt=dt, **arrs)
    b = refshim.ransac(method, seed, iters, thr3d=thr[0], thr2d=thr[1], focal=F, thrN=thr[2], confidence=conf, refit=refit,
                       dt=dt, **arrs)
    assert a["iters_run"] < S.shape[0]
    # (config #1: the oracle starts from the table length instead of 100 000; the adaptive bound is far below both)
    assert b["iter_final"] < S.shape[0]
    _assert_same_run(a, b, ("config", method, config, seed))
    assert a["max_votes"] == b["max_votes"] and _same(a["q"], b["q"]) and _same(a["t"], b["t"])
    assert np.array_equal(a["mask"], b["mask"][:a["mask"].shape[0]])
    # the accepted pose is the simulated one
    Rg, Re = _Rq(d["q"]), _Rq(b["q_refit"])
    ang = np.arccos(np.clip((np.trace(Rg @ Re.T) - 1) / 2, -1, 1))
    assert ang < 0.02 and np.linalg.norm(b["t_refit"] - d["t"]) < 0.1, (ang, b["t_refit"], d["t"])
    if refit == 1:
        q, t, ok = orc.shinji_ls(arrs["xc"], arrs["xw"], a["mask"][1], dt=dt)
        assert ok and _same(q, b["q_refit"]) and _same(t, b["t_refit"])
    if refit == 2:
        q, t = orc.nl_shinji_kneip_ls(a["q"], a["t"], a["mask"], a["max_votes"], dt=dt, **arrs)
        assert _same(q, b["q_refit"]) and _same(t, b["t_refit"])


def test_reference_kinect_simulator_and_ours_agree_statistically(rpe):
    """simulate_kinect_2d_3d_nl_correspondences (Simulator.hpp:368-436) as written in the reference next to
    rpe_sim_kinect_2d_3d_nl: the axial weights sigma_a(0, min_depth) / sigma_a follow the same law of depth, the axial and
    lateral residuals have the same robust spread per depth band, 3-D outliers come in the same proportion."""
    n = 60000
    ref = refshim.sim(3, 31, n, n2d=1.0, or2d=0.0, or3d=0.2, ornl=0.0)
    q, t = rpe.sim_pose(31)
    ours = rpe.sim_kinect_2d_3d_nl(32, q, t, n, n2d=1.0, or2d=0.0, or3d=0.2, nnl=float(np.deg2rad(2.0)), ornl=0.0)

    def summary(d, q, t):
        R = _Rq(q)
        Pgt = d["xw"].astype(np.float64) @ R.T + np.asarray(t, np.float64)
        e = d["xc"].astype(np.float64) - Pgt
        inl = np.abs(e).max(axis=1) < 0.5
        z, w = Pgt[:, 2], d["weights"][1].astype(np.float64)
        out = {"outliers": float((~inl).mean()), "w_med": float(np.median(w)), "w_max": float(w.max())}
        for lo, hi in ((0.5, 2.0), (3.0, 5.0), (6.0, 8.0)):
            sel = inl & (z > lo) & (z < hi)
            out[f"ax_{lo}"] = 1.4826 * float(np.median(np.abs(e[sel, 2])))
            out[f"lat_{lo}"] = 1.4826 * float(np.median(np.abs(e[sel, 0])))
            out[f"w_{lo}"] = float(np.median(w[sel]))
        return out

    a, b = summary(ref, ref["q"], ref["t"]), summary(ours, q, t)
    assert abs(a["outliers"] - b["outliers"]) < 0.01 and abs(a["outliers"] - 0.2) < 0.01
    assert a["w_max"] <= 1.0 + 1e-6 and b["w_max"] <= 1.0 + 1e-6
    for k in a:
        if k[:3] in ("ax_", "lat", "w_0", "w_3", "w_6", "w_m"):
            assert abs(a[k] / b[k] - 1) < 0.08, (k, a[k], b[k])


def test_reference_ffi_library_ao_and_ao_ransac(orc, rpe):
    """/root/reference/Library.cpp built unmodified (oracle/_ref/libref_library.so): extern "C" ao() = shinji_ls2 over all
    points, ao_ransac() = shinji_ransac2(thr 0.1, 1000 iterations, confidence 0.99999, ::rand()) + shinji_ls1, R_cw written
    row-major. The oracle's flow for them — the one rpe_ao / rpe_ao_ransac are tested against on the GPU
    (tests/test_gpu_cpp_api.py) — gives the same bits."""
    import ctypes
    import os
    so = os.path.join(refshim.ROOT, "oracle", "_ref", "libref_library.so")
    if not os.path.exists(so):
        pytest.skip("libref_library.so not built")
    lib = ctypes.CDLL(so)
    libc = ctypes.CDLL("libc.so.6")
    n = 3000
    q, t = rpe.sim_pose(21)
    Q, P, _ = rpe.sim_3d_3d(22, q, t, n, noise=0.02, outlier_ratio=0.3)
    R, tt = np.empty(9, np.float32), np.empty(3, np.float32)
    devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)  # the reference prints "ao()" etc. on stdout
    try:
        os.dup2(devnull, 1)
        libc.srand(ctypes.c_uint(1))
        lib.ao_ransac(Q.ctypes.data_as(ctypes.c_void_p), P.ctypes.data_as(ctypes.c_void_p), n, R.ctypes.data_as(ctypes.c_void_p),
                      tt.ctypes.data_as(ctypes.c_void_p))
        R1, t1 = R.copy(), tt.copy()
        lib.ao(Q.ctypes.data_as(ctypes.c_void_p), P.ctypes.data_as(ctypes.c_void_p), n, R.ctypes.data_as(ctypes.c_void_p),
               tt.ctypes.data_as(ctypes.c_void_p))
        libc.fflush(None)
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    S = orc.sample_table(1, n, 3, 1000)
    ref = orc.ransac(0, S, thr3d=np.float32(0.1), confidence=np.float32(0.99999), full=False, xc=P, xw=Q, want_arrays=False)
    ls_q, ls_t, ok = orc.shinji_ls(P, Q, ref["mask"][1])
    assert ok and np.array_equal(R1.reshape(3, 3), orc.quat_to_matrix(ls_q)) and np.array_equal(t1, ls_t)
    ls_q, ls_t, ok = orc.shinji_ls(P, Q, None)
    assert ok and np.array_equal(R.reshape(3, 3), orc.quat_to_matrix(ls_q)) and np.array_equal(tt, ls_t)


def test_randomised_sweep_against_the_reference_sources():
    """tools/fuzz_ref_shim.py: random sizes (8 .. 2 500), iteration budgets (1 .. 400), outlier ratios, noise, thresholds,
    confidences, NaN camera points, RANSAC and PROSAC, float and double, with the refits — the oracle against the
    reference's own headers. (7 900 further cases were run when this was written: 0 mismatches.)"""
    import os
    import subprocess
    import sys
    p = subprocess.run([sys.executable, os.path.join(refshim.ROOT, "tools", "fuzz_ref_shim.py"), "300", "2024"], capture_output=True,
                       text=True, timeout=600)
    assert p.returncode == 0 and "300 cases, 0 mismatches" in p.stdout, p.stdout[-3000:] + p.stderr[-2000:]


def test_stale_sample_buffers_are_the_only_deviation_with_invalid_depth(orc, rpe):
    """With all-NaN camera points in the frame the reference's nl_2p (called in every iteration of nl_shinji_ransac /
    nl_shinji_kneip_ransac) reads camera-side sample columns that assign_sample left untouched, i.e. those of an EARLIER
    sample (AbsoluteOrientationNormal.hpp:48-75, :299-315). The product and the default oracle feed it the NaN instead
    (deliberate deviations). The oracle's model of the reference's behaviour (StaleCols) reproduces the
    reference bit for bit, so that IS the whole difference."""
    differs = 0
    for method in (4, 5):
        for seed in (1, 2, 8, 43):  # (with 70 % outliers a stale-column hypothesis out-votes the proper ones in 2 of 118 runs)
            n, iters = 700, 150
            q, t = rpe.sim_pose(6000 + seed)
            g = rpe.sim_2d_3d_nl(6100 + seed, q, t, n, or2d=0.7, or3d=0.7, ornl=0.7)
            d = {k: np.ascontiguousarray(g[k]) for k in ("bv", "xc", "nc", "xw", "nw")}
            d["xc"][::3] = np.nan
            ct, cn = refshim.cos_thr(8.0, F), refshim.cos_nl(0.1)
            S = orc.sample_table(seed, n, 4, iters)
            b = refshim.ransac(method, seed, iters, thr3d=0.2, thr2d=8.0, focal=F, thrN=0.1, confidence=0.99, **d)
            a = orc.ransac(method, S, thr3d=0.2, cos_thr=ct, cos_nl=cn, confidence=0.99, full=False, **d)
            differs += int(not (a["max_votes"] == b["max_votes"] and _same(a["q"], b["q"])))
            orc.set_stale_sample_buffers(True)
            try:
                a = orc.ransac(method, S, thr3d=0.2, cos_thr=ct, cos_nl=cn, confidence=0.99, full=False, **d)
            finally:
                orc.set_stale_sample_buffers(False)
            _assert_same_run(a, b, ("stale", method, seed, 0))
    assert differs >= 1  # the deviation is real on such frames (a third of the pixels without depth)
