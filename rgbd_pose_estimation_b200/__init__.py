"""rgbd_pose_estimation_b200 — B200-native robust absolute-pose hot path.

Host side of the C-ABI declared in ``include/rpe_c_api.h`` (ctypes binding; the
C++ header-only mirror of the reference's adapter API lives in ``include/rpe/``).
The compute path is the CUDA library ``librpe_b200.so`` built from ``csrc/``;
there is no CPU fallback and importing :mod:`.capi` fails loudly if the library
has not been built (``python -c "import __graft_entry__ as g; g.build()"``).
"""
from .capi import (  # noqa: F401
    Context,
    Sequence,
    RpeError,
    lib,
    lib_path,
    METHODS,
    REFITS,
    sample_table,
    Sampler,
    prosac_table,
    update_num_iters,
    min_ev_host,
    min_ms_host,
    sim_pose,
    sim_3d_3d,
    sim_2d_3d,
    sim_2d_3d_nl,
    sim_kinect_2d_3d_nl,
    method_slots,
    method_mask_cols,
    method_sample_size,
    pinned_empty,
)

__all__ = [
    "Context", "Sequence", "RpeError", "lib", "lib_path", "METHODS", "REFITS", "sample_table", "prosac_table", "Sampler",
    "update_num_iters", "min_ev_host", "min_ms_host", "sim_pose", "sim_3d_3d", "sim_2d_3d", "sim_2d_3d_nl", "sim_kinect_2d_3d_nl", "method_slots",
    "method_mask_cols", "method_sample_size", "pinned_empty",
]
