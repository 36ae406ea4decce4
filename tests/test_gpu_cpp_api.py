"""GPU: the header-only C++ host API (include/rpe/*.hpp) used exactly like the reference's own drivers
(SimpleMain.cpp, TestMain.cpp), compared with the CPU oracle on the same inputs and the same ::rand() stream."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_libm = ctypes.CDLL("libm.so.6")
_libm.cosf.restype = ctypes.c_float
_libm.cosf.argtypes = [ctypes.c_float]
_libm.atanf.restype = ctypes.c_float
_libm.atanf.argtypes = [ctypes.c_float]


def _cos_thr(px, f=585.0):
    return float(_libm.cosf(_libm.atanf(ctypes.c_float(np.float32(px) / np.float32(f)))))


def _angle(qa, qb):
    a = np.asarray(qa, np.float64) / np.linalg.norm(np.asarray(qa, np.float64))
    b = np.asarray(qb, np.float64) / np.linalg.norm(np.asarray(qb, np.float64))
    av, aw, bv, bw = a[:3], a[3], -b[:3], b[3]
    w = aw * bw - np.dot(av, bv)
    v = aw * bv + bw * av + np.cross(av, bv)
    return 2.0 * np.arctan2(np.linalg.norm(v), abs(w))


@pytest.fixture(scope="module")
def dropin_output(tmp_path_factory, rpe):
    d = tmp_path_factory.mktemp("dropin")
    exe = os.path.join(str(d), "test_dropin")
    libdir = os.path.dirname(rpe.lib_path)
    subprocess.run(["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp"), "-L", libdir, "-lrpe_b200",
                    "-Wl,-rpath," + libdir], check=True)
    f64_file = os.path.join(str(d), "arrays_f64.bin")
    with open(f64_file, "wb") as fh:
        for a in _arrays_f64(rpe, 1000):
            fh.write(np.ascontiguousarray(a, np.float64).tobytes())
    out = subprocess.run([exe, "1000", "100000", f64_file], capture_output=True, text=True, check=True).stdout
    return {j["case"]: j for j in (json.loads(l) for l in out.splitlines() if l.startswith("{"))}


def _arrays_f64(rpe, total):
    """bv, xc, nc, xw, nw in binary64 for the Tp = double section of the drop-in program: the simulator's frame widened,
    perturbed below float resolution and with directions renormalised in binary64."""
    q, t = rpe.sim_pose(11)
    d = rpe.sim_2d_3d_nl(15, q, t, total, n2d=1.0, or2d=0.3, n3d=0.05, or3d=0.3, nnl=float(np.float32(2.0 * np.pi / 180.)),
                         ornl=0.3)
    rng = np.random.default_rng(15)
    out = []
    for k in ("bv", "xc", "nc", "xw", "nw"):
        a = d[k].astype(np.float64) * (1.0 + 1e-9 * rng.standard_normal(d[k].shape))
        if k in ("bv", "nc", "nw"):
            a = a / np.linalg.norm(a, axis=1, keepdims=True)
        out.append(np.ascontiguousarray(a))
    return out


def test_cpp_dropin_matches_oracle(dropin_output, rpe, orc):
    orc.set_math_mode(orc.DET)
    res = dropin_output
    total = 1000
    q, t = rpe.sim_pose(11)
    skip = 0
    # ---- AOOnlyPoseAdapter: shinji_ransac2 (Iter0 = 100 000, conf 0.9999), shinji_ls1, shinji_ls2
    Q, P, W = rpe.sim_3d_3d(12, q, t, total, noise=0.1, outlier_ratio=0.5)
    # like the reference, each estimator leaves ::rand() advanced by the draws of the iterations its early-stopping
    # loop executed (not by the rows the GPU drew ahead): the next estimator's table starts right there
    S = orc.sample_table_skip(1, skip, total, 3, 100000)
    ref = orc.ransac(0, S, thr3d=0.25, confidence=0.9999, full=False, xc=P, xw=Q, want_arrays=False)
    skip += ref["iters_run"] * 3
    got = res["shinji_ransac2"]
    assert (got["max_votes"], got["iter"]) == (ref["max_votes"], ref["iter_final"])
    assert got["n_inliers"] == int(ref["mask"][1].sum())
    assert np.array_equal(np.float32(got["q"]).view(np.uint32), ref["q"].view(np.uint32))
    ls_q, ls_t, _ = orc.shinji_ls(P, Q, ref["mask"][1], dt=np.float64)
    assert _angle(res["shinji_ls1"]["q"], ls_q) < 1e-6 and np.abs(np.array(res["shinji_ls1"]["t"]) - ls_t).max() < 1e-5
    ls_q, ls_t, _ = orc.shinji_ls(P, Q, None, dt=np.float64)
    assert _angle(res["shinji_ls2"]["q"], ls_q) < 1e-6 and np.abs(np.array(res["shinji_ls2"]["t"]) - ls_t).max() < 1e-5
    # ---- PnPPoseAdapter: kneip_ransac, then LM
    Q, U, Pgt, W = rpe.sim_2d_3d(13, q, t, total, noise_px=1.0, outlier_ratio=0.3)
    S = orc.sample_table_skip(1, skip, total, 4, 2000)
    cos_thr = _cos_thr(8.0)
    ref = orc.ransac(1, S, cos_thr=cos_thr, confidence=0.99, full=False, bv=U, xw=Q, want_arrays=False)
    skip += ref["iters_run"] * 4
    got = res["kneip_ransac"]
    assert (got["max_votes"], got["iter"], got["n_inliers"]) == (ref["max_votes"], ref["iter_final"], int(ref["mask"][0].sum()))
    assert np.array_equal(np.float32(got["q"]).view(np.uint32), ref["q"].view(np.uint32))
    tq, tt, _ = orc.refine_gn(ref["q"], ref["t"], ref["mask"], max_iters=8, bv=U, xw=Q)
    assert _angle(res["kneip_ransac+lm"]["q"], tq) < 1e-6
    # ---- NormalAOPoseAdapter: the four multi-modal estimators of TestMain.cpp, then nl_shinji_kneip_ls twice
    d = rpe.sim_2d_3d_nl(14, q, t, total, n2d=1.0, or2d=0.3, n3d=0.05, or3d=0.3, nnl=float(np.float32(2.0 * np.pi / 180.)),
                         ornl=0.3)
    arrs = {k: d[k] for k in ("bv", "xc", "nc", "xw", "nw")}
    cos_nl = float(_libm.cosf(ctypes.c_float(0.1)))
    last = None
    for name, method in [("shinji_kneip_ransac", 2), ("nl_kneip_ransac", 3), ("nl_shinji_ransac", 4), ("nl_shinji_kneip_ransac", 5)]:
        S = orc.sample_table_skip(1, skip, total, 4, 300)
        ref = orc.ransac(method, S, thr3d=0.2, cos_thr=cos_thr, cos_nl=cos_nl, confidence=0.99, full=False, want_arrays=False,
                         **arrs)
        skip += ref["iters_run"] * 4
        got = res[name]
        assert (got["max_votes"], got["iter"]) == (ref["max_votes"], ref["iter_final"]), name
        assert np.array_equal(np.float32(got["q"]).view(np.uint32), ref["q"].view(np.uint32)), name
        last = ref
    q1, t1 = orc.nl_shinji_kneip_ls(last["q"], last["t"], last["mask"], last["max_votes"], dt=np.float64, **arrs)
    got = res["nl_shinji_kneip_ls"]
    assert _angle(got["q"], q1) < 1e-6 and np.abs(np.array(got["t"]) - t1).max() < 1e-5
    # second call starts from the first refit's pose, now with the simulator's dynamic weights (TestMain.cpp:219-221)
    q2, t2 = orc.nl_shinji_kneip_ls(np.float32(got["q"]), np.float32(got["t"]), last["mask"], last["max_votes"],
                                    weights3=d["weights"].astype(np.float64), dt=np.float64, **arrs)
    got2 = res["nl_shinji_kneip_ls_dw"]
    assert _angle(got2["q"], q2) < 1e-6 and np.abs(np.array(got2["t"]) - t2).max() < 1e-5
    assert _angle(got2["q"], q) < 5e-3  # and it is a good pose
    # ---- Tp = double adapters: decided in binary64 on the device, compared with the oracle instantiated for double;
    # the sample stream continues where the float estimators left ::rand()
    bv, xc, nc, xw, nw = _arrays_f64(rpe, total)
    _libm.cos.restype = _libm.atan.restype = ctypes.c_double
    _libm.cos.argtypes = _libm.atan.argtypes = [ctypes.c_double]
    cos_thr64 = _libm.cos(_libm.atan(8.0 / 585.0))
    cos_nl64 = _libm.cos(0.1)
    S = orc.sample_table_skip(1, skip, total, 4, 300)
    ref = orc.ransac(5, S, thr3d=0.2, cos_thr=cos_thr64, cos_nl=cos_nl64, confidence=0.99, full=False, want_arrays=False,
                     dt=np.float64, bv=bv, xc=xc, nc=nc, xw=xw, nw=nw)
    skip += ref["iters_run"] * 4
    got = res["nl_shinji_kneip_ransac_f64"]
    assert (got["max_votes"], got["iter"]) == (ref["max_votes"], ref["iter_final"])
    assert np.array_equal(np.float64(got["q"]).view(np.uint64), ref["q"].view(np.uint64))
    assert np.array_equal(np.float64(got["t"]).view(np.uint64), ref["t"].view(np.uint64))
    rq, rt = orc.nl_shinji_kneip_ls(ref["q"], ref["t"], ref["mask"], ref["max_votes"], dt=np.float64, bv=bv, xc=xc, nc=nc, xw=xw,
                                    nw=nw)
    got = res["nl_shinji_kneip_ls_f64"]
    assert _angle(got["q"], rq) < 1e-6 and np.abs(np.array(got["t"]) - rt).max() < 1e-5
    S = orc.sample_table_skip(1, skip, total, 4, 500)
    ref = orc.ransac(1, S, cos_thr=cos_thr64, confidence=0.99, full=False, want_arrays=False, dt=np.float64, bv=bv, xw=xw)
    got = res["kneip_ransac_f64"]
    assert (got["max_votes"], got["iter"], got["n_inliers"]) == (ref["max_votes"], ref["iter_final"], int(ref["mask"][0].sum()))
    assert np.array_equal(np.float64(got["q"]).view(np.uint64), ref["q"].view(np.uint64))
    orc.set_math_mode(orc.LIBM)


def test_nl_shinji_kneip_ls_through_c_abi(rpe, orc, gpu_ctx):
    orc.set_math_mode(orc.DET)
    n = 20000
    q, t = rpe.sim_pose(81)
    d = rpe.sim_2d_3d_nl(82, q, t, n)
    arrs = {k: d[k] for k in ("bv", "xc", "nc", "xw", "nw")}
    S = rpe.sample_table(1, n, 4, 256)
    cos_thr, cos_nl = _cos_thr(8.0), float(_libm.cosf(ctypes.c_float(0.1)))
    gpu_ctx.upload(**arrs)
    r = gpu_ctx.ransac("nl_shinji_kneip", S, thr3d=0.2, cos_thr2d=cos_thr, cos_thrN=cos_nl, confidence=0.99)
    ref_q, ref_t = orc.nl_shinji_kneip_ls(r["q"], r["t"], r["mask"], r["max_votes"], dt=np.float64, **arrs)
    fit = gpu_ctx.refit("nl_sk_ls")
    assert fit["refit_ok"] == 1
    assert _angle(fit["q"], ref_q) < 1e-6 and np.abs(fit["t"].astype(np.float64) - ref_t).max() < 1e-5
    assert _angle(fit["q"], q) <= _angle(r["q"], q) + 1e-4
    orc.set_math_mode(orc.LIBM)


def test_library_cpp_shim_ao_and_ao_ransac(rpe, orc):
    """rpe_ao / rpe_ao_ransac: same arguments and conventions as the reference's extern "C" ao() / ao_ransac()
    (Library.cpp:17-75): x_w, x_c 3 x n column-major in, R_cw row-major 9 floats + t out; ao = shinji_ls2,
    ao_ransac = shinji_ransac2(thr 0.1, 1000 iterations, confidence 0.99999, unseeded rand()) + shinji_ls1."""
    orc.set_math_mode(orc.DET)
    n = 3000
    q, t = rpe.sim_pose(21)
    Q, P, _ = rpe.sim_3d_3d(22, q, t, n, noise=0.02, outlier_ratio=0.3)
    R = np.empty(9, np.float32)
    tt = np.empty(3, np.float32)
    assert rpe.lib.rpe_ao_ransac(Q.ctypes.data, P.ctypes.data, n, R.ctypes.data, tt.ctypes.data) == 0
    S = orc.sample_table(1, n, 3, 1000)
    ref = orc.ransac(0, S, thr3d=np.float32(0.1), confidence=np.float32(0.99999), full=False, xc=P, xw=Q, want_arrays=False)
    ls_q, ls_t, _ = orc.shinji_ls(P, Q, ref["mask"][1], dt=np.float64)
    Rref = orc.quat_to_matrix(ls_q, np.float64)
    assert np.abs(R.reshape(3, 3) - Rref).max() < 1e-6 and np.abs(tt - ls_t).max() < 1e-5
    # ao(): least squares over all points, so the 30 % outliers bias it — compare with the oracle's shinji_ls2
    assert rpe.lib.rpe_ao(Q.ctypes.data, P.ctypes.data, n, R.ctypes.data, tt.ctypes.data) == 0
    ls_q, ls_t, _ = orc.shinji_ls(P, Q, None, dt=np.float64)
    assert np.abs(R.reshape(3, 3) - orc.quat_to_matrix(ls_q, np.float64)).max() < 1e-6 and np.abs(tt - ls_t).max() < 1e-5
    orc.set_math_mode(orc.LIBM)
