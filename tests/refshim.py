"""ctypes loader for oracle/_ref/libref_shim.so: the reference's own pose headers compiled (where /root/reference
exists) against the Eigen / Sophus API stand-ins of oracle/ref_shim/. TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libref_shim.so")
REFERENCE = "/root/reference/pose"


def available():
    """Build on demand where the reference sources are present; a prebuilt library (it travels with gpurun) also counts."""
    if os.path.isdir(REFERENCE):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, stdout=subprocess.DEVNULL)
    return os.path.exists(SO)


_lib = None
_vp = C.c_void_p


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(SO)
    return _lib


class RefOut(C.Structure):
    _fields_ = [("max_votes", C.c_int), ("iter_final", C.c_int), ("n_idx", C.c_int * 3), ("ensure_failures", C.c_longlong)]


def _suf(dt):
    return ("f", np.float32, C.c_float) if dt == np.float32 else ("d", np.float64, C.c_double)


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _arr(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def cos_thr(thr2d, focal, dt=np.float32):
    s, _, ct = _suf(dt)
    fn = getattr(lib(), f"ref_cos_thr_{s}")
    fn.restype = ct
    fn.argtypes = [ct, ct]
    return fn(thr2d, focal)


def cos_nl(thrN, dt=np.float32):
    s, _, ct = _suf(dt)
    fn = getattr(lib(), f"ref_cos_nl_{s}")
    fn.restype = ct
    fn.argtypes = [ct]
    return fn(thrN)


def ransac(method, seed, iter_in, sampler=0, thr3d=0.0, thr2d=0.0, focal=585.0, thrN=0.0, confidence=0.99, refit=0,
           weights3=None, bv=None, xc=None, nc=None, xw=None, nw=None, dt=np.float32):
    s, npdt, ct = _suf(dt)
    arrs = [_arr(a, npdt) for a in (bv, xc, nc, xw, nw)]
    n = next(a.shape[0] for a in arrs if a is not None)
    w = _arr(weights3, npdt)  # (3, n): column-major n x 3
    out = RefOut()
    q, t, qr, tr = np.zeros(4, npdt), np.zeros(3, npdt), np.zeros(4, npdt), np.zeros(3, npdt)
    mask = np.zeros((3, n), np.int16)
    fn = getattr(lib(), f"ref_ransac_{s}")
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_int] + [_vp] * 5 + [C.c_int, _vp, C.c_uint, C.c_int, ct, ct, ct, ct, ct, C.c_int,
                                                    C.POINTER(RefOut), _vp, _vp, _vp, _vp, _vp]
    rc = fn(method, sampler, *[_p(a) for a in arrs], n, _p(w), seed, iter_in, thr3d, thr2d, focal, thrN, confidence, refit,
            C.byref(out), _p(q), _p(t), _p(qr), _p(tr), _p(mask))
    assert rc == 0
    return {"q": q, "t": t, "q_refit": qr, "t_refit": tr, "max_votes": out.max_votes, "iter_final": out.iter_final,
            "n_idx": list(out.n_idx), "ensure_failures": out.ensure_failures, "mask": mask}


def shinji(Xw, Xc, K=None, cols=None, dt=np.float32):
    s, npdt, _ = _suf(dt)
    Xw, Xc = _arr(Xw, npdt), _arr(Xc, npdt)
    K = Xw.shape[0] if K is None else K
    cols = Xw.shape[0] if cols is None else cols
    q, t = np.zeros(4, npdt), np.zeros(3, npdt)
    fn = getattr(lib(), f"ref_shinji_{s}")
    fn.restype = C.c_int
    ok = fn(_p(Xw), _p(Xc), K, cols, _p(q), _p(t))
    return q, t, bool(ok)


def kneip_main(Xw, bv, dt=np.float32):
    s, npdt, _ = _suf(dt)
    q, t = np.zeros((4, 4), npdt), np.zeros((4, 3), npdt)
    fn = getattr(lib(), f"ref_kneip_main_{s}")
    fn.restype = C.c_int
    k = fn(_p(_arr(Xw, npdt)), _p(_arr(bv, npdt)), _p(q), _p(t))
    return q[:k], t[:k]


def kneip4(Xw, bv, dt=np.float32):
    s, npdt, _ = _suf(dt)
    q, t = np.zeros(4, npdt), np.zeros(3, npdt)
    fn = getattr(lib(), f"ref_kneip4_{s}")
    fn.restype = C.c_int
    ok = fn(_p(_arr(Xw, npdt)), _p(_arr(bv, npdt)), _p(q), _p(t))
    return q, t, bool(ok)


def nl_2p(pt1_c, nl1_c, pt2_c, pt1_w, nl1_w, pt2_w, dt=np.float32):
    s, npdt, _ = _suf(dt)
    q, t = np.zeros(4, npdt), np.zeros(3, npdt)
    args = [_arr(a, npdt) for a in (pt1_c, nl1_c, pt2_c, pt1_w, nl1_w, pt2_w)]
    getattr(lib(), f"ref_nl_2p_{s}")(*[_p(a) for a in args], _p(q), _p(t))
    return q, t


def o4_roots(f5, dt=np.float32):
    s, npdt, _ = _suf(dt)
    r = np.zeros(4, npdt)
    getattr(lib(), f"ref_o4_roots_{s}")(_p(_arr(f5, npdt)), _p(r))
    return r


def update_num_iters(p, ep, model_points, max_iters, dt=np.float32):
    s, _, ct = _suf(dt)
    fn = getattr(lib(), f"ref_update_num_iters_{s}")
    fn.restype = C.c_int
    fn.argtypes = [ct, ct, C.c_int, C.c_int]
    return fn(p, ep, model_points, max_iters)


def sim(kind, seed, n, n2d=1.0, or2d=0.3, n3d=0.05, or3d=0.3, nnl=float(np.deg2rad(2.0)), ornl=0.3, min_depth=0.4,
        max_depth=8.0, f=585.0, gaussian=True, dt=np.float32):
    """The reference's own Simulator.hpp (kind 0: simulate_3d_3d, 1: simulate_2d_3d, 2: simulate_2d_3d_nl) with the
    pose drawn as SimpleMain.cpp does, ::rand() and the normal generator seeded with `seed`."""
    s, npdt, ct = _suf(dt)
    q, t = np.zeros(4, npdt), np.zeros(3, npdt)
    out = {k: np.full((n, 3), np.nan, npdt) for k in ("xw", "nw", "xc", "nc", "bv")}
    w = np.zeros((3, n), npdt)
    fn = getattr(lib(), f"ref_sim_{s}")
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_uint, C.c_int] + [ct] * 9 + [C.c_int] + [_vp] * 8
    rc = fn(kind, seed, n, n2d, or2d, n3d, or3d, nnl, ornl, min_depth, max_depth, f, 1 if gaussian else 0, _p(q), _p(t),
            _p(out["xw"]), _p(out["nw"]), _p(out["xc"]), _p(out["nc"]), _p(out["bv"]), _p(w))
    assert rc == 0, rc
    out.update({"q": q, "t": t, "weights": w})
    return out
