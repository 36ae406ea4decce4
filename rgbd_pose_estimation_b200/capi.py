"""ctypes binding of include/rpe_c_api.h.

Everything here is plumbing: argument marshalling into the C-ABI of ``librpe_b200.so``.
No arithmetic of the hot path is done in Python, and nothing under ``oracle/`` is imported.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(_HERE, "librpe_b200.so")

METHODS = {
    "shinji": 0,            # shinji_ransac / shinji_ransac2     AbsoluteOrientation.hpp:101-213
    "kneip": 1,             # kneip_ransac                       P3P.hpp:320-392
    "shinji_kneip": 2,      # shinji_kneip_ransac                AbsoluteOrientation.hpp:367-438
    "nl_kneip": 3,          # nl_kneip_ransac                    AbsoluteOrientationNormal.hpp:215-284
    "nl_shinji": 4,         # nl_shinji_ransac                   AbsoluteOrientationNormal.hpp:286-354
    "nl_shinji_kneip": 5,   # nl_shinji_kneip_ransac             AbsoluteOrientationNormal.hpp:356-445
    "kneip_quat": 6,        # kneip_prosac's scoring form        P3P.hpp:439-453
}
REFITS = {"kabsch_inliers": 0, "kabsch_all": 1, "gn": 2, "nl_sk_ls": 3}


def method_slots(m: int) -> int:
    return 2 if m in (2, 4) else (3 if m == 5 else 1)


def method_mask_cols(m: int) -> int:
    return 1 if m in (1, 6) else (2 if m in (0, 2) else 3)


def method_sample_size(m: int) -> int:
    return 3 if m == 0 else 4


class RpeError(RuntimeError):
    pass


class _Result(C.Structure):
    _fields_ = [
        ("R", C.c_float * 9), ("q", C.c_float * 4), ("t", C.c_float * 3),
        ("max_votes", C.c_int32), ("iter_final", C.c_int32), ("winner", C.c_int32), ("n_slots", C.c_int32),
        ("n_borderline", C.c_int32), ("flags", C.c_int32), ("n_inliers", C.c_int32 * 3), ("refit_ok", C.c_int32),
        ("refit_cost", C.c_double), ("refit_evals", C.c_int32), ("reserved", C.c_int32),
        ("qd", C.c_double * 4), ("td", C.c_double * 3),
    ]

    def to_dict(self):
        return {
            "R": np.array(self.R, dtype=np.float32).reshape(3, 3), "q": np.array(self.q, dtype=np.float32),
            "t": np.array(self.t, dtype=np.float32), "max_votes": int(self.max_votes),
            "iter_final": int(self.iter_final), "winner": int(self.winner), "n_slots": int(self.n_slots),
            "n_borderline": int(self.n_borderline), "flags": int(self.flags),
            "n_inliers": [int(v) for v in self.n_inliers], "refit_ok": int(self.refit_ok),
            "refit_cost": float(self.refit_cost), "refit_evals": int(self.refit_evals),
            "qd": np.array(self.qd, dtype=np.float64), "td": np.array(self.td, dtype=np.float64),
        }


def _load():
    if not os.path.exists(lib_path):
        raise RpeError(
            f"{lib_path} is missing: the CUDA library has not been built. There is no CPU fallback; "
            "run `python -c 'import __graft_entry__ as g; g.build()'` (or `make -C rgbd_pose_estimation_b200/csrc`).")
    return C.CDLL(lib_path)


lib = _load()

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)
_sp = C.POINTER(C.c_int16)
_vp = C.c_void_p


def _sig(name, restype, argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    return fn


_sig("rpe_version", C.c_int, [])
_sig("rpe_status_string", C.c_char_p, [C.c_int])
_sig("rpe_device_count", C.c_int, [C.POINTER(C.c_int)])
_sig("rpe_create", C.c_int, [C.c_int, C.POINTER(_vp)])
_sig("rpe_create_on_stream", C.c_int, [C.c_int, _vp, C.POINTER(_vp)])
_sig("rpe_destroy", C.c_int, [_vp])
_sig("rpe_last_error", C.c_char_p, [_vp])
_sig("rpe_stream", _vp, [_vp])
_sig("rpe_sync", C.c_int, [_vp])
_sig("rpe_launch_count", C.c_longlong, [_vp])
_sig("rpe_host_alloc", C.c_int, [C.c_size_t, C.POINTER(_vp)])
_sig("rpe_host_free", C.c_int, [_vp])
_sig("rpe_upload", C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int])
_sig("rpe_upload_device", C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int])
_sig("rpe_num_correspondences", C.c_int, [_vp])
_sig("rpe_ransac", C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                             C.POINTER(_Result), _vp])
_sig("rpe_ransac_async", C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                   C.POINTER(_Result), _vp])
SAMPLE_FN = C.CFUNCTYPE(C.c_int, _vp, C.c_int, C.c_int, C.POINTER(C.c_int32))
_sig("rpe_ransac_stream", C.c_int, [_vp, C.c_int, SAMPLE_FN, _vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                    C.POINTER(_Result), _vp])
_sig("rpe_upload_f64", C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int])
_sig("rpe_ransac_f64", C.c_int, [_vp, C.c_int, _vp, SAMPLE_FN, _vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                 C.POINTER(_Result), _vp])
_sig("rpe_get_hypotheses_f64", C.c_int, [_vp, C.c_int, _vp, _vp])
_sig("rpe_sim_kinect_2d_3d_nl_device", C.c_int, [_vp, C.c_uint64, _vp, _vp, C.c_int] + [C.c_float] * 8)
_sig("rpe_sim_kinect_2d_3d_nl", C.c_int, [C.c_uint64, _vp, _vp, C.c_int] + [C.c_float] * 8 + [_vp] * 6)
_sig("rpe_set_first_pass_iters", C.c_int, [_vp, C.c_int])
_sig("rpe_set_stale_sample_columns", C.c_int, [_vp, C.c_int])
_sig("rpe_set_upload_overlap", C.c_int, [_vp, C.c_int])
_sig("rpe_peer_export", C.c_int, [_vp, _vp])
_sig("rpe_peer_import", C.c_int, [_vp, C.c_int, C.c_int, _vp])
_sig("rpe_peer_import_local", C.c_int, [_vp, C.c_int, C.c_int, _vp])
_sig("rpe_exchange_votes", C.c_int, [_vp, C.c_int, C.c_int])
_sig("rpe_peer_status", C.c_int, [_vp])
_sig("rpe_peer_set_timeout_ms", C.c_int, [_vp, C.c_int])
_sig("rpe_ransac_sharded", C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                     C.POINTER(_Result), _vp])
_sig("rpe_ransac_sharded_async", C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                           C.POINTER(_Result), _vp])
_sig("rpe_refit", C.c_int, [_vp, C.c_int, _vp, C.c_int, C.POINTER(_Result)])
_sig("rpe_refit_async", C.c_int, [_vp, C.c_int, _vp, C.c_int, C.POINTER(_Result)])
_sig("rpe_set_pose", C.c_int, [_vp, _vp, _vp, C.c_int])
_sig("rpe_set_mask", C.c_int, [_vp, _vp, C.c_int])
_sig("rpe_generate", C.c_int, [_vp, C.c_int, _vp, C.c_int])
_sig("rpe_get_hypotheses", C.c_int, [_vp, _vp, _vp, C.c_int])
_sig("rpe_set_hypotheses", C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int])
_sig("rpe_score", C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float])
_sig("rpe_get_votes", C.c_int, [_vp, _vp, C.c_int])
_sig("rpe_set_votes", C.c_int, [_vp, _vp, C.c_int])
_sig("rpe_votes_device_ptr", _vp, [_vp])
_sig("rpe_finish", C.c_int, [_vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(_Result), _vp])
_sig("rpe_update_num_iters", C.c_int, [C.c_float, C.c_float, C.c_int, C.c_int])
_sig("rpe_sample_table", C.c_int, [C.c_uint32, C.c_int, C.c_int, C.c_int, _vp])
_sig("rpe_prosac_table", C.c_int, [C.c_uint32, C.c_int, C.c_int, C.c_int, _vp, _vp])
_sig("rpe_sampler_create", C.c_int, [C.c_uint32, C.c_int, C.POINTER(_vp)])
_sig("rpe_sampler_rows", C.c_int, [_vp, C.c_int, C.c_int, _vp])
_sig("rpe_sampler_destroy", None, [_vp])
_sig("rpe_sim_pose", C.c_int, [C.c_uint64, C.c_float, C.c_float, _vp, _vp])
_sig("rpe_sim_3d_3d", C.c_int, [C.c_uint64, _vp, _vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                C.c_int, _vp, _vp, _vp])
_sig("rpe_sim_2d_3d", C.c_int, [C.c_uint64, _vp, _vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                C.c_int, _vp, _vp, _vp, _vp])
_sig("rpe_sim_2d_3d_nl", C.c_int, [C.c_uint64, _vp, _vp, C.c_int] + [C.c_float] * 9 + [C.c_int] + [_vp] * 6)
_sig("rpe_sim_3d_3d_device", C.c_int, [_vp, C.c_uint64, _vp, _vp, C.c_int] + [C.c_float] * 5 + [C.c_int])
_sig("rpe_sim_2d_3d_nl_device", C.c_int, [_vp, C.c_uint64, _vp, _vp, C.c_int] + [C.c_float] * 9 + [C.c_int])
_sig("rpe_sim_3d_3d_device_to", C.c_int, [_vp, C.c_uint64, _vp, _vp, C.c_int] + [C.c_float] * 5 + [C.c_int, _vp, _vp])
_sig("rpe_download", C.c_int, [_vp] * 6)
_sig("rpe_sampler_reseed", C.c_int, [_vp, C.c_uint32])


class SeqParams(C.Structure):
    _fields_ = [("device", C.c_int), ("n_contexts", C.c_int), ("n_threads", C.c_int), ("method", C.c_int), ("H", C.c_int),
                ("thr3d", C.c_float), ("cos_thr2d", C.c_float), ("cos_thrN", C.c_float), ("confidence", C.c_float),
                ("refit", C.c_int), ("gn_iters", C.c_int), ("sample_seed", C.c_uint32)]


class SeqFrame(C.Structure):
    _fields_ = [("bv", _vp), ("xc", _vp), ("nc", _vp), ("xw", _vp), ("nw", _vp), ("n", C.c_int), ("on_device", C.c_int),
                ("samples", _vp), ("mask", _vp)]


_sig("rpe_seq_create", C.c_int, [C.POINTER(SeqParams), C.POINTER(_vp)])
_sig("rpe_seq_run", C.c_int, [_vp, C.POINTER(SeqFrame), C.c_int, C.c_longlong, C.c_int, C.POINTER(_Result), C.POINTER(_Result)])
_sig("rpe_seq_run_shared", C.c_int, [_vp, C.POINTER(SeqFrame), C.c_int, _vp, C.c_longlong, C.c_int, C.POINTER(_Result),
                                     C.POINTER(_Result), _vp, C.POINTER(C.c_int)])
_sig("rpe_seq_context", _vp, [_vp, C.c_int])
_sig("rpe_seq_num_contexts", C.c_int, [_vp])
_sig("rpe_seq_last_error", C.c_char_p, [_vp])
_sig("rpe_seq_destroy", C.c_int, [_vp])
_sig("rpe_ao", C.c_int, [_vp, _vp, C.c_int, _vp, _vp])
_sig("rpe_ao_ransac", C.c_int, [_vp, _vp, C.c_int, _vp, _vp])
_sig("rpe_measure_ffma_tflops", C.c_int, [_vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)])
_sig("rpe_last_stage_ms", C.c_int, [_vp, _vp])
_sig("rpe_enable_stage_timing", C.c_int, [_vp, C.c_int])
_sig("rpe_set_mask_transfer", C.c_int, [_vp, C.c_int])
_sig("rpe_poll", C.c_int, [_vp])
_sig("rpe_scorer_time_stats", C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int])
_sig("rpe_scorer_busy_stats", C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int])
_sig("rpe_min_ev", C.c_int, [_vp, _vp, C.c_int, _vp])
_sig("rpe_min_ms", C.c_int, [_vp, _vp, C.c_int, _vp, _vp])
_sig("rpe_min_ev_host", C.c_int, [_vp, C.c_int, _vp])
_sig("rpe_min_ev_host_f64", C.c_int, [_vp, C.c_int, _vp])
_sig("rpe_min_ms_host", C.c_int, [_vp, C.c_int, _vp, _vp])
_sig("rpe_debug_set_packed", C.c_int, [C.c_int])
_sig("rpe_debug_reset", C.c_int, [_vp])
_sig("rpe_debug_set_raw_tiles", C.c_int, [C.c_int])

# every symbol include/rpe_c_api.h declares (tests check the header against this list and the .so)
DECLARED_SYMBOLS = [
    "rpe_version", "rpe_status_string", "rpe_device_count", "rpe_create", "rpe_create_on_stream", "rpe_destroy",
    "rpe_last_error", "rpe_stream", "rpe_sync", "rpe_launch_count", "rpe_host_alloc", "rpe_host_free", "rpe_upload",
    "rpe_upload_device", "rpe_num_correspondences", "rpe_ransac", "rpe_ransac_async", "rpe_ransac_stream", "rpe_set_first_pass_iters", "rpe_set_stale_sample_columns", "rpe_set_upload_overlap", "rpe_upload_f64", "rpe_ransac_f64", "rpe_get_hypotheses_f64", "rpe_refit", "rpe_refit_async",
    "rpe_set_pose", "rpe_set_mask", "rpe_generate", "rpe_get_hypotheses", "rpe_set_hypotheses", "rpe_score",
    "rpe_get_votes", "rpe_set_votes", "rpe_votes_device_ptr", "rpe_peer_export", "rpe_peer_import", "rpe_peer_import_local", "rpe_exchange_votes", "rpe_peer_status", "rpe_peer_set_timeout_ms", "rpe_ransac_sharded", "rpe_ransac_sharded_async", "rpe_finish", "rpe_update_num_iters", "rpe_sample_table",
    "rpe_prosac_table", "rpe_sampler_create", "rpe_sampler_rows", "rpe_sampler_destroy", "rpe_sim_pose", "rpe_sim_3d_3d", "rpe_sim_2d_3d", "rpe_sim_2d_3d_nl", "rpe_sim_kinect_2d_3d_nl", "rpe_sim_kinect_2d_3d_nl_device", "rpe_sim_3d_3d_device",
    "rpe_sim_2d_3d_nl_device", "rpe_min_ev", "rpe_min_ms", "rpe_min_ev_host", "rpe_min_ev_host_f64", "rpe_min_ms_host",
    "rpe_sim_3d_3d_device_to", "rpe_sampler_reseed", "rpe_seq_create", "rpe_seq_run", "rpe_seq_run_shared",
    "rpe_seq_context", "rpe_seq_num_contexts", "rpe_seq_last_error", "rpe_seq_destroy", "rpe_download", "rpe_ao", "rpe_ao_ransac",
    "rpe_measure_ffma_tflops", "rpe_last_stage_ms", "rpe_enable_stage_timing", "rpe_scorer_time_stats", "rpe_scorer_busy_stats", "rpe_set_mask_transfer", "rpe_poll",
]


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    if a is None:
        return None
    a = np.asarray(a)
    if a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
        a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _check(rc, ctx=None):
    if rc != 0:
        msg = lib.rpe_status_string(rc).decode()
        if ctx is not None:
            msg += ": " + lib.rpe_last_error(ctx).decode()
        raise RpeError(f"rpe error {rc}: {msg}")


# ---- host helpers ------------------------------------------------------------------------------
def sample_table(seed: int, n: int, m: int, H: int) -> np.ndarray:
    out = np.empty((H, 4), dtype=np.int32)
    _check(lib.rpe_sample_table(seed, n, m, H, _ptr(out)))
    return out


def prosac_table(seed: int, n: int, m: int, H: int, weights=None) -> np.ndarray:
    out = np.empty((H, 4), dtype=np.int32)
    w = _f32(weights)
    _check(lib.rpe_prosac_table(seed, n, m, H, _ptr(w), _ptr(out)))
    return out


class Sampler:
    """RandomElements<int> that lives across calls (rpe_sampler_*): per-frame sample tables at O(H m) each."""

    def __init__(self, seed: int, n: int):
        self._h = _vp()
        _check(lib.rpe_sampler_create(seed, n, C.byref(self._h)))

    def rows(self, m: int, H: int, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((H, 4), dtype=np.int32)
        _check(lib.rpe_sampler_rows(self._h, m, H, _ptr(out)))
        return out

    def reseed(self, seed: int):
        _check(lib.rpe_sampler_reseed(self._h, seed))

    def close(self):
        if self._h:
            lib.rpe_sampler_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def min_ev_host(M, dtype=np.float32):
    """ev() of MinimalSolvers.hpp on the host: (count, 3, 3) symmetric matrices -> (count, 3) eigenvalues, descending."""
    M = np.ascontiguousarray(M, dtype=dtype).reshape(-1, 9)
    E = np.empty((M.shape[0], 3), dtype)
    fn = lib.rpe_min_ev_host if dtype == np.float32 else lib.rpe_min_ev_host_f64
    _check(fn(_ptr(M), M.shape[0], _ptr(E)))
    return E


def min_ms_host(in24):
    """ms() of MinimalSolvers.hpp on the host: (count, 24) = Aw Bw Nw Mw Ac Bc Nc Mc -> q (count, 4), t (count, 3)."""
    a = np.ascontiguousarray(in24, dtype=np.float32).reshape(-1, 24)
    q = np.empty((a.shape[0], 4), np.float32)
    t = np.empty((a.shape[0], 3), np.float32)
    _check(lib.rpe_min_ms_host(_ptr(a), a.shape[0], _ptr(q), _ptr(t)))
    return q, t


def update_num_iters(p: float, ep: float, model_points: int, max_iters: int) -> int:
    return int(lib.rpe_update_num_iters(p, ep, model_points, max_iters))


def sim_pose(seed: int, max_angle: float = np.pi / 2, t_size: float = 5.0):
    q = np.empty(4, np.float32)
    t = np.empty(3, np.float32)
    _check(lib.rpe_sim_pose(seed, max_angle, t_size, _ptr(q), _ptr(t)))
    return q, t


def sim_3d_3d(seed, q, t, n, noise=0.1, outlier_ratio=0.5, min_depth=0.4, max_depth=8.0, f=585.0, gaussian=True):
    """Points are returned as (n, 3) float32 arrays == column-major 3 x n in memory."""
    Q = np.empty((n, 3), np.float32)
    P = np.empty((n, 3), np.float32)
    W = np.zeros((3, n), np.float32)  # n x 3 column-major
    _check(lib.rpe_sim_3d_3d(seed, _ptr(_f32(q)), _ptr(_f32(t)), n, noise, outlier_ratio, min_depth, max_depth, f,
                             1 if gaussian else 0, _ptr(Q), _ptr(P), _ptr(W)))
    return Q, P, W


def sim_2d_3d(seed, q, t, n, noise_px=1.0, outlier_ratio=0.7, min_depth=0.4, max_depth=8.0, f=585.0, gaussian=True):
    Q = np.empty((n, 3), np.float32)
    U = np.empty((n, 3), np.float32)
    P = np.empty((n, 3), np.float32)
    W = np.zeros((3, n), np.float32)
    _check(lib.rpe_sim_2d_3d(seed, _ptr(_f32(q)), _ptr(_f32(t)), n, noise_px, outlier_ratio, min_depth, max_depth, f,
                             1 if gaussian else 0, _ptr(Q), _ptr(U), _ptr(P), _ptr(W)))
    return Q, U, P, W


def sim_2d_3d_nl(seed, q, t, n, n2d=1.0, or2d=0.3, n3d=0.05, or3d=0.3, nnl=np.deg2rad(2.0), ornl=0.3, min_depth=0.4,
                 max_depth=8.0, f=585.0, gaussian=True):
    Q = np.empty((n, 3), np.float32)
    M = np.empty((n, 3), np.float32)
    P = np.empty((n, 3), np.float32)
    N = np.empty((n, 3), np.float32)
    U = np.empty((n, 3), np.float32)
    W = np.zeros((3, n), np.float32)
    _check(lib.rpe_sim_2d_3d_nl(seed, _ptr(_f32(q)), _ptr(_f32(t)), n, n2d, or2d, n3d, or3d, nnl, ornl, min_depth,
                                max_depth, f, 1 if gaussian else 0, _ptr(Q), _ptr(M), _ptr(P), _ptr(N), _ptr(U),
                                _ptr(W)))
    return {"xw": Q, "nw": M, "xc": P, "nc": N, "bv": U, "weights": W}


def sim_kinect_2d_3d_nl(seed, q, t, n, n2d=1.0, or2d=0.3, or3d=0.3, nnl=np.deg2rad(2.0), ornl=0.3, min_depth=0.4,
                        max_depth=8.0, f=585.0):
    """simulate_kinect_2d_3d_nl_correspondences: camera points with the Kinect lateral / axial noise model."""
    out = {k: np.empty((n, 3), np.float32) for k in ("xw", "nw", "xc", "nc", "bv")}
    w = np.empty((3, n), np.float32)
    _check(lib.rpe_sim_kinect_2d_3d_nl(seed, _ptr(_f32(q)), _ptr(_f32(t)), n, n2d, or2d, or3d, nnl, ornl, min_depth, max_depth,
                                       f, _ptr(out["xw"]), _ptr(out["nw"]), _ptr(out["xc"]), _ptr(out["nc"]), _ptr(out["bv"]),
                                       _ptr(w)))
    out["weights"] = w
    return out


class _PinnedBlock:
    """Owner of one rpe_host_alloc block; arrays made by pinned_empty keep it alive through their `.base` chain and
    the block is returned with rpe_host_free when the last of them is collected."""

    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        _check(lib.rpe_host_alloc(max(nbytes, 1), C.byref(self.ptr)))
        self.buf = (C.c_byte * max(nbytes, 1)).from_address(self.ptr.value)
        weakref.finalize(self, lib.rpe_host_free, C.c_void_p(self.ptr.value))

    @property
    def __array_interface__(self):
        return {"shape": (len(self.buf),), "typestr": "|u1", "data": (self.ptr.value, False), "version": 3}


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """numpy array backed by page-locked host memory from rpe_host_alloc; the memory is freed (rpe_host_free) when
    the array and every view of it are gone. Keep it alive until the asynchronous calls that use it are synchronised."""
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    block = _PinnedBlock(count * dtype.itemsize)
    raw = np.asarray(block)  # base = block
    return raw[:count * dtype.itemsize].view(dtype).reshape(shape)


# ---- batched sequence (rpe_seq_*) ---------------------------------------------------------------
REFIT_BITS = {"kabsch": 1, "gn": 2, "nl_sk_ls": 4}


class Sequence:
    """rpe_seq: frames of a sequence issued by native host threads over several contexts of one GPU (config #5).
    `run(frames, first_frame, n_frames)` blocks until every result is on the host."""

    def __init__(self, device=0, method="shinji", H=1024, thr3d=0.0, cos_thr2d=0.0, cos_thrN=0.0, confidence=0.99,
                 refit=("kabsch", "gn"), gn_iters=3, sample_seed=1, contexts=12, threads=2):
        m = METHODS[method] if isinstance(method, str) else method
        bits = 0
        for r in refit:
            bits |= REFIT_BITS[r]
        self.params = SeqParams(device, contexts, threads, m, H, thr3d, cos_thr2d, cos_thrN, confidence, bits, gn_iters,
                                sample_seed)
        self._h = _vp()
        _check(lib.rpe_seq_create(C.byref(self.params), C.byref(self._h)))
        self.method = m
        self._ring = None
        self._keep = None

    def set_frames(self, frames):
        """frames: list of dicts with keys among bv/xc/nc/xw/nw (numpy (n,3) float32 page-locked arrays, or ints =
        device pointers with 'n' given), optional 'samples' ((H,4) int32 array or device pointer), optional 'mask'."""
        ring = (SeqFrame * len(frames))()
        keep = []
        for i, f in enumerate(frames):
            on_dev = any(isinstance(f.get(k), int) for k in ("bv", "xc", "nc", "xw", "nw"))
            n = f.get("n")
            for k in ("bv", "xc", "nc", "xw", "nw"):
                a = f.get(k)
                if a is None:
                    continue
                if isinstance(a, int):
                    setattr(ring[i], k, a)
                else:
                    if a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
                        raise ValueError("sequence frames must be C-contiguous float32 (n, 3) arrays")
                    n = a.shape[0]
                    keep.append(a)
                    setattr(ring[i], k, a.ctypes.data)
            ring[i].n = int(n)
            ring[i].on_device = 1 if on_dev else 0
            sm = f.get("samples")
            if sm is not None:
                if isinstance(sm, int):
                    ring[i].samples = sm
                else:
                    keep.append(sm)
                    ring[i].samples = sm.ctypes.data
            mk = f.get("mask")
            if mk is not None:
                keep.append(mk)
                ring[i].mask = mk.ctypes.data
        self._ring, self._keep = ring, keep

    def run(self, first_frame, n_frames, want_results=True):
        if self._ring is None:
            raise RpeError("Sequence.set_frames first")
        r0 = (_Result * n_frames)() if want_results else None
        r1 = (_Result * n_frames)() if want_results else None
        rc = lib.rpe_seq_run(self._h, self._ring, len(self._ring), first_frame, n_frames, r0, r1)
        if rc != 0:
            raise RpeError(f"rpe error {rc}: {lib.rpe_status_string(rc).decode()}: {lib.rpe_seq_last_error(self._h).decode()}")
        return r0, r1

    def run_shared(self, counter, total, capacity):
        """Frames handed out by a counter shared with other sequences (rpe_seq_run_shared). `counter`: a numpy int64 array
        of one element (e.g. a view of multiprocessing.shared_memory). Returns (ransac results, final results, frame
        indices, n_done)."""
        if self._ring is None:
            raise RpeError("Sequence.set_frames first")
        r0 = (_Result * capacity)()
        r1 = (_Result * capacity)()
        idx = np.full(capacity, -1, np.int64)
        nd = C.c_int(0)
        rc = lib.rpe_seq_run_shared(self._h, self._ring, len(self._ring), counter.ctypes.data_as(C.c_void_p), total, capacity,
                                    r0, r1, idx.ctypes.data_as(C.c_void_p), C.byref(nd))
        if rc != 0:
            raise RpeError(f"rpe error {rc}: {lib.rpe_status_string(rc).decode()}: {lib.rpe_seq_last_error(self._h).decode()}")
        return r0, r1, idx, nd.value

    def contexts(self):
        return [C.c_void_p(lib.rpe_seq_context(self._h, i)) for i in range(lib.rpe_seq_num_contexts(self._h))]

    def close(self):
        if self._h:
            lib.rpe_seq_destroy(self._h)
            self._h = _vp()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- device context ---------------------------------------------------------------------------
class Context:
    """One rpe_ctx (one GPU, one stream). Arrays are (n, 3) float32 == the reference's 3 x n column-major."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._h = C.c_void_p()
        if stream is None:
            _check(lib.rpe_create(device, C.byref(self._h)))
        else:
            _check(lib.rpe_create_on_stream(device, C.c_void_p(stream), C.byref(self._h)))
        self._keep = []     # host arrays of the last upload (the copy is asynchronous)
        self._pending = []  # result structs / sample tables of enqueued calls: the library writes them at sync time
        self.n = 0

    def close(self):
        if self._h:
            lib.rpe_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def handle(self):
        return self._h

    def upload(self, bv=None, xc=None, nc=None, xw=None, nw=None):
        arrs = [_f32(a) for a in (bv, xc, nc, xw, nw)]
        n = next(a.shape[0] for a in arrs if a is not None)
        for a in arrs:
            if a is not None and a.shape != (n, 3):
                raise ValueError("correspondence arrays must be (n, 3)")
        self._keep = arrs  # host buffers must outlive the asynchronous copy
        self.n = n
        _check(lib.rpe_upload(self._h, *[_ptr(a) for a in arrs], n), self._h)

    def upload_device(self, n, bv=None, xc=None, nc=None, xw=None, nw=None):
        """Device pointers (ints), e.g. torch tensor.data_ptr()."""
        self.n = n
        _check(lib.rpe_upload_device(self._h, *[_ptr(a) for a in (bv, xc, nc, xw, nw)], n), self._h)

    def ransac(self, method, samples, thr3d=0.0, cos_thr2d=0.0, cos_thrN=0.0, confidence=0.99, want_mask=True):
        m = METHODS[method] if isinstance(method, str) else method
        samples = np.ascontiguousarray(samples, dtype=np.int32)
        H = samples.shape[0]
        res = _Result()
        mask = np.empty((method_mask_cols(m), self.n), np.int16) if want_mask else None
        _check(lib.rpe_ransac(self._h, m, _ptr(samples), H, thr3d, cos_thr2d, cos_thrN, confidence, C.byref(res),
                              _ptr(mask)), self._h)
        d = res.to_dict()
        d["mask"] = mask  # (cols, n): row k == column k of the reference's n x cols matrix
        return d

    def sim_3d_3d_device(self, seed, q, t, n, noise=0.1, outlier_ratio=0.5, min_depth=0.4, max_depth=8.0, f=585.0,
                         gaussian=True):
        self.n = n
        _check(lib.rpe_sim_3d_3d_device(self._h, seed, _ptr(_f32(q)), _ptr(_f32(t)), n, noise, outlier_ratio, min_depth,
                                        max_depth, f, 1 if gaussian else 0), self._h)

    def sim_kinect_2d_3d_nl_device(self, seed, q, t, n, n2d=1.0, or2d=0.3, or3d=0.3, nnl=np.deg2rad(2.0), ornl=0.3,
                                   min_depth=0.4, max_depth=8.0, f=585.0):
        self.n = n
        _check(lib.rpe_sim_kinect_2d_3d_nl_device(self._h, seed, _ptr(_f32(q)), _ptr(_f32(t)), n, n2d, or2d, or3d, nnl, ornl,
                                                  min_depth, max_depth, f), self._h)

    def sim_2d_3d_nl_device(self, seed, q, t, n, n2d=1.0, or2d=0.3, n3d=0.05, or3d=0.3, nnl=np.deg2rad(2.0), ornl=0.3,
                            min_depth=0.4, max_depth=8.0, f=585.0, gaussian=True):
        self.n = n
        _check(lib.rpe_sim_2d_3d_nl_device(self._h, seed, _ptr(_f32(q)), _ptr(_f32(t)), n, n2d, or2d, n3d, or3d, nnl, ornl,
                                           min_depth, max_depth, f, 1 if gaussian else 0), self._h)

    def sim_3d_3d_device_to(self, seed, q, t, n, d_xw, d_xc, noise=0.1, outlier_ratio=0.5, min_depth=0.4, max_depth=8.0,
                            f=585.0, gaussian=True):
        """The device-side generator into caller-owned device buffers (ints = device pointers)."""
        _check(lib.rpe_sim_3d_3d_device_to(self._h, seed, _ptr(_f32(q)), _ptr(_f32(t)), n, noise, outlier_ratio, min_depth,
                                           max_depth, f, 1 if gaussian else 0, C.c_void_p(d_xw), C.c_void_p(d_xc)), self._h)

    def min_ev(self, M):
        """ev() batches on the GPU, one matrix per thread (rpe_min_ev)."""
        M = np.ascontiguousarray(M, dtype=np.float32).reshape(-1, 9)
        E = np.empty((M.shape[0], 3), np.float32)
        _check(lib.rpe_min_ev(self._h, _ptr(M), M.shape[0], _ptr(E)), self._h)
        return E

    def min_ms(self, in24):
        """ms() batches on the GPU, one two-correspondence problem per thread (rpe_min_ms)."""
        a = np.ascontiguousarray(in24, dtype=np.float32).reshape(-1, 24)
        q = np.empty((a.shape[0], 4), np.float32)
        t = np.empty((a.shape[0], 3), np.float32)
        _check(lib.rpe_min_ms(self._h, _ptr(a), a.shape[0], _ptr(q), _ptr(t)), self._h)
        return q, t

    def download(self, names=("xc", "xw")):
        order = ("bv", "xc", "nc", "xw", "nw")
        out = {k: (np.empty((self.n, 3), np.float32) if k in names else None) for k in order}
        _check(lib.rpe_download(self._h, *[_ptr(out[k]) for k in order]), self._h)
        return {k: v for k, v in out.items() if v is not None}

    def upload_f64(self, bv=None, xc=None, nc=None, xw=None, nw=None):
        """Binary64 arrays (n, 3): the next ransac_f64 decides everything in binary64 like a Tp = double adapter."""
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (bv, xc, nc, xw, nw)]
        n = next(a.shape[0] for a in arrs if a is not None)
        self.n = n
        self._arrays = arrs
        _check(lib.rpe_upload_f64(self._h, *[_ptr(a) for a in arrs], n), self._h)

    def ransac_f64(self, method, samples, thr3d=0.0, cos_thr2d=0.0, cos_thrN=0.0, confidence=0.99, want_mask=True):
        m = METHODS[method] if isinstance(method, str) else method
        samples = np.ascontiguousarray(samples, dtype=np.int32)
        res = _Result()
        mask = np.empty((method_mask_cols(m), self.n), np.int16) if want_mask else None
        _check(lib.rpe_ransac_f64(self._h, m, _ptr(samples), C.cast(None, SAMPLE_FN), None, samples.shape[0], thr3d, cos_thr2d,
                                  cos_thrN, confidence, C.byref(res), _ptr(mask)), self._h)
        d = res.to_dict()
        d["mask"] = mask
        return d

    def get_hypotheses_f64(self, n_slots):
        hyps = np.empty((n_slots, 7), np.float64)
        valid = np.empty(n_slots, np.int32)
        _check(lib.rpe_get_hypotheses_f64(self._h, n_slots, _ptr(hyps), _ptr(valid)), self._h)
        return hyps, valid

    def ransac_stream(self, method, row_fn, H, thr3d=0.0, cos_thr2d=0.0, cos_thrN=0.0, confidence=0.99, want_mask=True):
        """rpe_ransac_stream: `row_fn(first_iteration, count)` returns the (count, 4) int32 rows of one device pass."""
        m = METHODS[method] if isinstance(method, str) else method
        calls = []

        def thunk(_user, first, count, out):
            rows = np.ascontiguousarray(row_fn(first, count), dtype=np.int32).reshape(count, 4)
            C.memmove(out, rows.ctypes.data, rows.nbytes)
            calls.append((first, count))
            return 0

        cb = SAMPLE_FN(thunk)
        res = _Result()
        mask = np.empty((method_mask_cols(m), self.n), np.int16) if want_mask else None
        _check(lib.rpe_ransac_stream(self._h, m, cb, None, H, thr3d, cos_thr2d, cos_thrN, confidence, C.byref(res),
                                     _ptr(mask)), self._h)
        d = res.to_dict()
        d["mask"] = mask
        d["passes"] = calls
        return d

    def ransac_async(self, method, samples, H=None, thr3d=0.0, cos_thr2d=0.0, cos_thrN=0.0, confidence=0.99, mask=None):
        """Enqueue only. `samples` may be a numpy int32 array (kept alive until sync) or a device pointer (int)
        with H given. Returns the ctypes result struct, valid after :meth:`sync`."""
        m = METHODS[method] if isinstance(method, str) else method
        if isinstance(samples, int):
            sp = C.c_void_p(samples)
        else:
            samples = np.ascontiguousarray(samples, dtype=np.int32)
            H = samples.shape[0]
            self._pending.append(samples)
            sp = _ptr(samples)
        res = _Result()
        self._pending.append(res)  # the library writes into it at sync time
        _check(lib.rpe_ransac_async(self._h, m, sp, H, thr3d, cos_thr2d, cos_thrN, confidence, C.byref(res),
                                    _ptr(mask)), self._h)
        return res

    def refit_async(self, kind, weights=None, max_iters=0):
        k = REFITS[kind] if isinstance(kind, str) else kind
        res = _Result()
        w = _f32(weights)
        if w is not None:
            self._pending.append(w)
        self._pending.append(res)
        _check(lib.rpe_refit_async(self._h, k, _ptr(w), max_iters, C.byref(res)), self._h)
        return res

    def upload_async(self, bv=None, xc=None, nc=None, xw=None, nw=None):
        """Like upload() but does not retain references (caller keeps page-locked arrays alive)."""
        arrs = [bv, xc, nc, xw, nw]
        n = next(a.shape[0] for a in arrs if a is not None)
        self.n = n
        _check(lib.rpe_upload(self._h, *[_ptr(a) for a in arrs], n), self._h)

    def refit(self, kind, weights=None, max_iters=0):
        k = REFITS[kind] if isinstance(kind, str) else kind
        res = _Result()
        w = _f32(weights)
        _check(lib.rpe_refit(self._h, k, _ptr(w), max_iters, C.byref(res)), self._h)
        return res.to_dict()

    def set_pose(self, q, t, max_votes=0):
        _check(lib.rpe_set_pose(self._h, _ptr(_f32(q)), _ptr(_f32(t)), max_votes), self._h)

    def set_mask(self, mask):
        mask = np.ascontiguousarray(mask, dtype=np.int16)
        _check(lib.rpe_set_mask(self._h, _ptr(mask), mask.shape[0]), self._h)

    def generate(self, method, samples, H=None):
        """`samples`: (H, 4) int32 array, or a device pointer (int) with H given."""
        m = METHODS[method] if isinstance(method, str) else method
        if isinstance(samples, int):
            _check(lib.rpe_generate(self._h, m, C.c_void_p(samples), H), self._h)
            return H * method_slots(m)
        samples = np.ascontiguousarray(samples, dtype=np.int32)
        _check(lib.rpe_generate(self._h, m, _ptr(samples), samples.shape[0]), self._h)
        return samples.shape[0] * method_slots(m)

    def get_hypotheses(self, n_slots):
        hyps = np.empty((n_slots, 7), np.float32)
        valid = np.empty(n_slots, np.int32)
        _check(lib.rpe_get_hypotheses(self._h, _ptr(hyps), _ptr(valid), n_slots), self._h)
        return hyps, valid

    def set_hypotheses(self, method, hyps, valid=None):
        m = METHODS[method] if isinstance(method, str) else method
        hyps = _f32(hyps)
        v = None if valid is None else np.ascontiguousarray(valid, dtype=np.int32)
        _check(lib.rpe_set_hypotheses(self._h, m, _ptr(hyps), _ptr(v), hyps.shape[0]), self._h)

    def score(self, method, slot_begin, slot_end, thr3d=0.0, cos_thr2d=0.0, cos_thrN=0.0):
        m = METHODS[method] if isinstance(method, str) else method
        _check(lib.rpe_score(self._h, m, slot_begin, slot_end, thr3d, cos_thr2d, cos_thrN), self._h)

    def set_stale_sample_columns(self, on=True):
        """Opt-in reproduction of the reference's stale sample columns (rpe_set_stale_sample_columns)."""
        _check(lib.rpe_set_stale_sample_columns(self._h, 1 if on else 0), self._h)

    def set_upload_overlap(self, chunks=4):
        """Overlap the upload of page-locked frames with generation + scoring (rpe_set_upload_overlap); 0 = off."""
        _check(lib.rpe_set_upload_overlap(self._h, int(chunks)), self._h)

    def set_first_pass_iters(self, iters):
        _check(lib.rpe_set_first_pass_iters(self._h, iters), self._h)

    def peer_export(self):
        """64-byte CUDA IPC handle of this context's exchange block (hypothesis-sharded single-frame mode)."""
        h = np.zeros(64, np.uint8)
        _check(lib.rpe_peer_export(self._h, _ptr(h)), self._h)
        return h

    def peer_import(self, rank, world, handles):
        handles = np.ascontiguousarray(handles, dtype=np.uint8).reshape(world, 64)
        _check(lib.rpe_peer_import(self._h, rank, world, _ptr(handles)), self._h)

    @staticmethod
    def peer_link_local(ctxs):
        """Contexts of this process, one per GPU: link them for the hypothesis-sharded mode (no IPC)."""
        arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
        for r, c in enumerate(ctxs):
            _check(lib.rpe_peer_import_local(c._h, r, len(ctxs), arr), c._h)

    def exchange_votes(self, slot_begin, slot_end):
        _check(lib.rpe_exchange_votes(self._h, slot_begin, slot_end), self._h)

    def ransac_sharded(self, method, samples, H=None, thr3d=0.0, cos_thr2d=0.0, cos_thrN=0.0, confidence=0.99, blocking=True):
        """One hypothesis-sharded frame (see rpe_ransac_sharded). `samples`: (H, 4) int32 array or a device pointer + H.
        blocking=False returns the ctypes result struct, valid after sync()."""
        m = METHODS[method] if isinstance(method, str) else method
        if isinstance(samples, int):
            sp = C.c_void_p(samples)
        else:
            samples = np.ascontiguousarray(samples, dtype=np.int32)
            H = samples.shape[0]
            self._pending.append(samples)
            sp = _ptr(samples)
        res = _Result()
        if blocking:
            _check(lib.rpe_ransac_sharded(self._h, m, sp, H, thr3d, cos_thr2d, cos_thrN, confidence, C.byref(res), None), self._h)
            return res.to_dict()
        self._pending.append(res)
        _check(lib.rpe_ransac_sharded_async(self._h, m, sp, H, thr3d, cos_thr2d, cos_thrN, confidence, C.byref(res), None),
               self._h)
        return res

    def peer_status(self):
        _check(lib.rpe_peer_status(self._h), self._h)

    def get_votes(self, n_slots):
        v = np.empty(n_slots, np.int32)
        _check(lib.rpe_get_votes(self._h, _ptr(v), n_slots), self._h)
        return v

    def set_votes(self, votes):
        v = np.ascontiguousarray(votes, dtype=np.int32)
        _check(lib.rpe_set_votes(self._h, _ptr(v), v.shape[0]), self._h)

    def votes_device_ptr(self):
        return int(lib.rpe_votes_device_ptr(self._h) or 0)

    def finish(self, method, H, thr3d=0.0, cos_thr2d=0.0, cos_thrN=0.0, confidence=0.99, want_mask=True):
        m = METHODS[method] if isinstance(method, str) else method
        res = _Result()
        mask = np.empty((method_mask_cols(m), self.n), np.int16) if want_mask else None
        _check(lib.rpe_finish(self._h, m, H, thr3d, cos_thr2d, cos_thrN, confidence, C.byref(res), _ptr(mask)), self._h)
        d = res.to_dict()
        d["mask"] = mask
        return d

    def sync(self):
        try:
            _check(lib.rpe_sync(self._h), self._h)
        finally:
            self._keep = []
            self._pending = []

    def debug_reset(self):
        """Every test hook (process-global and of this context) back to the shipped configuration."""
        _check(lib.rpe_debug_reset(self._h), self._h)

    def peer_set_timeout_ms(self, ms):
        _check(lib.rpe_peer_set_timeout_ms(self._h, int(ms)), self._h)

    def launch_count(self):
        return int(lib.rpe_launch_count(self._h))

    def stream(self):
        return int(lib.rpe_stream(self._h) or 0)

    def set_mask_transfer(self, mode):
        """0: asynchronous calls copy the int16 inlier matrix as it is; 1: they send one bit per flag and the collecting
        thread expands it in place (same matrix in the caller's buffer after sync / poll, 1/16 of the D2H bytes); 2: the
        constant 2-D column of a family without the 2-D test stays on the device and the collecting thread writes it."""
        _check(lib.rpe_set_mask_transfer(self._h, int(mode)), self._h)

    def poll(self):
        """Hand over the results of asynchronous calls that have finished, without waiting for the others."""
        _check(lib.rpe_poll(self._h), self._h)

    def enable_stage_timing(self, on=True):
        """False/0: off; True/1: every stage; 2: only the events around the tiled scoring kernel."""
        _check(lib.rpe_enable_stage_timing(self._h, int(on)), self._h)

    def last_stage_ms(self):
        ms = np.zeros(8, np.float32)
        _check(lib.rpe_last_stage_ms(self._h, _ptr(ms)), self._h)
        names = ["upload_pack", "generate", "score", "replay", "mask_refit", "gn", "total", "score_fast"]
        return {k: float(v) for k, v in zip(names, ms)}

    def measure_ffma_tflops(self, ms_target=50):
        a, b = C.c_double(0), C.c_double(0)
        _check(lib.rpe_measure_ffma_tflops(self._h, ms_target, C.byref(a), C.byref(b)), self._h)
        return a.value, b.value
