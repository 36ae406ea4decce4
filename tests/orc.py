"""ctypes loader for the CPU oracle (oracle/_build/liboracle.so). TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")


def build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)


if not os.path.exists(SO):
    build()
lib = C.CDLL(SO)

_vp = C.c_void_p
LIBM, DET = 0, 1


class RansacOut(C.Structure):
    _fields_ = [("max_votes", C.c_int), ("iter_final", C.c_int), ("winner", C.c_int), ("iters_run", C.c_int),
                ("evals", C.c_longlong), ("seconds", C.c_double)]


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _arr(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def set_math_mode(m):
    lib.orc_set_math_mode(int(m))


def set_stale_sample_buffers(on):
    """The reference's hoisted sample buffers (oracle/ransac.hpp, StaleCols) — comparison with its own sources only."""
    lib.orc_set_stale_sample_buffers(1 if on else 0)


def rand_seq(seed, n):
    out = np.empty(n, np.int32)
    lib.orc_rand_seq(C.c_uint(seed), n, _p(out))
    return out


def sample_table(seed, n_corr, m, H):
    out = np.empty((H, 4), np.int32)
    lib.orc_sample_table(C.c_uint(seed), n_corr, m, H, _p(out))
    return out


def sample_table_skip(seed, skip, n_corr, m, H):
    out = np.empty((H, 4), np.int32)
    lib.orc_sample_table_skip(C.c_uint(seed), C.c_longlong(skip), n_corr, m, H, _p(out))
    return out


def prosac_table(seed, n_corr, m, H, weights=None):
    out = np.empty((H, 4), np.int32)
    w = _arr(weights, np.float32)
    lib.orc_prosac_table_f(C.c_uint(seed), n_corr, m, H, _p(w), _p(out))
    return out


def _suf(dt):
    return ("f", np.float32, C.c_float) if dt == np.float32 else ("d", np.float64, C.c_double)


def jacobi_svd3(A, dt=np.float32):
    s, npdt, _ = _suf(dt)
    A = _arr(A, npdt)
    U = np.empty((3, 3), npdt)
    V = np.empty((3, 3), npdt)
    S = np.empty(3, npdt)
    getattr(lib, f"orc_jacobi_svd3_{s}")(_p(A), _p(U), _p(S), _p(V))
    return U, S, V


def quat_to_matrix(q, dt=np.float32):
    s, npdt, _ = _suf(dt)
    R = np.empty((3, 3), npdt)
    getattr(lib, f"orc_quat_to_matrix_{s}")(_p(_arr(q, npdt)), _p(R))
    return R


def quat_from_matrix(R, dt=np.float32):
    s, npdt, _ = _suf(dt)
    q = np.empty(4, npdt)
    fn = getattr(lib, f"orc_quat_from_matrix_{s}")
    fn.restype = C.c_int
    ok = fn(_p(_arr(R, npdt)), _p(q))
    return q, bool(ok)


def quat_rotate(q, v, dt=np.float32):
    s, npdt, _ = _suf(dt)
    out = np.empty(3, npdt)
    getattr(lib, f"orc_quat_rotate_{s}")(_p(_arr(q, npdt)), _p(_arr(v, npdt)), _p(out))
    return out


def shinji(Xw, Xc, K=None, cols=None, dt=np.float32):
    """Xw, Xc: (K, 3) arrays (== 3 x K column-major)."""
    s, npdt, _ = _suf(dt)
    Xw = _arr(Xw, npdt)
    Xc = _arr(Xc, npdt)
    K = Xw.shape[0] if K is None else K
    cols = K if cols is None else cols
    q = np.empty(4, npdt)
    t = np.empty(3, npdt)
    fn = getattr(lib, f"orc_shinji_{s}")
    fn.restype = C.c_int
    ok = fn(_p(Xw), _p(Xc), K, cols, _p(q), _p(t))
    return q, t, bool(ok)


def update_num_iters(p, ep, model_points, max_iters, dt=np.float32):
    s, _, ct = _suf(dt)
    fn = getattr(lib, f"orc_update_num_iters_{s}")
    fn.restype = C.c_int
    fn.argtypes = [ct, ct, C.c_int, C.c_int]
    return fn(p, ep, model_points, max_iters)


def o4_roots(f5, dt=np.float32):
    s, npdt, _ = _suf(dt)
    r = np.empty(4, npdt)
    getattr(lib, f"orc_o4_roots_{s}")(_p(_arr(f5, npdt)), _p(r))
    return r


def kneip_main(Xw, bv, dt=np.float32):
    s, npdt, _ = _suf(dt)
    q = np.zeros((4, 4), npdt)
    t = np.zeros((4, 3), npdt)
    fn = getattr(lib, f"orc_kneip_main_{s}")
    fn.restype = C.c_int
    k = fn(_p(_arr(Xw, npdt)), _p(_arr(bv, npdt)), _p(q), _p(t))
    return q[:k], t[:k]


def kneip4(Xw, bv, dt=np.float32):
    s, npdt, _ = _suf(dt)
    q = np.zeros(4, npdt)
    t = np.zeros(3, npdt)
    fn = getattr(lib, f"orc_kneip4_{s}")
    fn.restype = C.c_int
    ok = fn(_p(_arr(Xw, npdt)), _p(_arr(bv, npdt)), _p(q), _p(t))
    return q, t, bool(ok)


def nl_2p(pt1_c, nl1_c, pt2_c, pt1_w, nl1_w, pt2_w, dt=np.float32):
    s, npdt, _ = _suf(dt)
    q = np.zeros(4, npdt)
    t = np.zeros(3, npdt)
    args = [_arr(a, npdt) for a in (pt1_c, nl1_c, pt2_c, pt1_w, nl1_w, pt2_w)]
    getattr(lib, f"orc_nl_2p_{s}")(*[_p(a) for a in args], _p(q), _p(t))
    return q, t


def score(method, q, t, thr3d=0.0, cos_thr=0.0, cos_nl=0.0, bv=None, xc=None, nc=None, xw=None, nw=None,
          dt=np.float32):
    s, npdt, ct = _suf(dt)
    arrs = [_arr(a, npdt) for a in (bv, xc, nc, xw, nw)]
    n = next(a.shape[0] for a in arrs if a is not None)
    cols = 1 if method in (1, 6) else (2 if method in (0, 2) else 3)
    mask = np.zeros((cols, n), np.int16)
    fn = getattr(lib, f"orc_score_{s}")
    fn.restype = C.c_int
    fn.argtypes = [C.c_int] + [_vp] * 5 + [C.c_int, _vp, _vp, ct, ct, ct, _vp]
    votes = fn(method, *[_p(a) for a in arrs], n, _p(_arr(q, npdt)), _p(_arr(t, npdt)), thr3d, cos_thr, cos_nl, _p(mask))
    return votes, mask


def ransac(method, samples, thr3d=0.0, cos_thr=0.0, cos_nl=0.0, confidence=0.99, full=True, nthreads=1, bv=None,
           xc=None, nc=None, xw=None, nw=None, dt=np.float32, want_arrays=True):
    s, npdt, ct = _suf(dt)
    arrs = [_arr(a, npdt) for a in (bv, xc, nc, xw, nw)]
    n = next(a.shape[0] for a in arrs if a is not None)
    samples = np.ascontiguousarray(samples, np.int32)
    H = samples.shape[0]
    S = 2 if method in (2, 4) else (3 if method == 5 else 1)
    cols = 1 if method in (1, 6) else (2 if method in (0, 2) else 3)
    out = RansacOut()
    q = np.zeros(4, npdt)
    t = np.zeros(3, npdt)
    votes = np.zeros(H * S, np.int32) if want_arrays else None
    hyps = np.zeros((H * S, 7), npdt) if want_arrays else None
    mask = np.zeros((cols, n), np.int16)
    fn = getattr(lib, f"orc_ransac_{s}")
    fn.restype = None
    fn.argtypes = [C.c_int] + [_vp] * 5 + [C.c_int, _vp, C.c_int, ct, ct, ct, ct, C.c_int, C.c_int,
                                          C.POINTER(RansacOut), _vp, _vp, _vp, _vp, _vp]
    fn(method, *[_p(a) for a in arrs], n, _p(samples), H, thr3d, cos_thr, cos_nl, confidence, 1 if full else 0,
       nthreads, C.byref(out), _p(q), _p(t), _p(votes), _p(hyps), _p(mask))
    return {"q": q, "t": t, "max_votes": out.max_votes, "iter_final": out.iter_final, "winner": out.winner,
            "iters_run": out.iters_run, "evals": out.evals, "seconds": out.seconds, "votes": votes, "hyps": hyps,
            "mask": mask}


def shinji_ls(xc, xw, flags3d=None, dt=np.float32):
    s, npdt, _ = _suf(dt)
    xc = _arr(xc, npdt)
    xw = _arr(xw, npdt)
    fl = _arr(flags3d, np.int16)
    q = np.zeros(4, npdt)
    t = np.zeros(3, npdt)
    fn = getattr(lib, f"orc_shinji_ls_{s}")
    fn.restype = C.c_int
    ok = fn(_p(xc), _p(xw), xc.shape[0], _p(fl), _p(q), _p(t))
    return q, t, bool(ok)


def refine_gn(q, t, mask, w=(1.0, 1.0, 1.0), max_iters=6, bv=None, xc=None, nc=None, xw=None, nw=None, dt=np.float32):
    s, npdt, ct = _suf(dt)
    arrs = [_arr(a, npdt) for a in (bv, xc, nc, xw, nw)]
    n = next(a.shape[0] for a in arrs if a is not None)
    mask = np.ascontiguousarray(mask, np.int16)
    q = np.array(q, npdt)
    t = np.array(t, npdt)
    info = np.zeros(4, np.float64)
    fn = getattr(lib, f"orc_refine_gn_{s}")
    fn.restype = C.c_int
    fn.argtypes = [_vp] * 5 + [C.c_int, _vp, C.c_int, ct, ct, ct, C.c_int, _vp, _vp, _vp]
    evals = fn(*[_p(a) for a in arrs], n, _p(mask), mask.shape[0], w[0], w[1], w[2], max_iters, _p(q), _p(t), _p(info))
    return q, t, {"cost": info[0], "evals": evals, "accepted": int(info[2]), "mu": info[3]}


def nl_shinji_kneip_ls(q, t, mask3, max_votes, weights3=None, bv=None, xc=None, nc=None, xw=None, nw=None,
                       dt=np.float32):
    s, npdt, _ = _suf(dt)
    arrs = [_arr(a, npdt) for a in (bv, xc, nc, xw, nw)]
    n = arrs[3].shape[0]
    mask3 = np.ascontiguousarray(mask3, np.int16)
    w = _arr(weights3, npdt)
    q = np.array(q, npdt)
    t = np.array(t, npdt)
    getattr(lib, f"orc_nl_shinji_kneip_ls_{s}")(*[_p(a) for a in arrs], n, _p(mask3), _p(w), max_votes, _p(q), _p(t))
    return q, t


for _name in ("orc_det_log", "orc_det_acos", "orc_det_cbrt"):
    getattr(lib, _name).restype = C.c_double
    getattr(lib, _name).argtypes = [C.c_double]
lib.orc_det_atan2.restype = C.c_double
lib.orc_det_atan2.argtypes = [C.c_double, C.c_double]


def det_sincos(a):
    s, c = C.c_double(), C.c_double()
    lib.orc_det_sincos(C.c_double(a), C.byref(s), C.byref(c))
    return s.value, c.value
