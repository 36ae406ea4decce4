"""Host logic of the opt-in uniform-register scorer (score_ur.cu: ur_pick_shape), no GPU needed: the CTA shape the launcher
picks must cover every pair of the frame, keep the warps balanced over the four schedulers for the benchmark frame, stay
inside the shared-memory and packed-counter limits, and config #4 must get the shape the kernel was tuned for."""
import ctypes as C

import numpy as np
import pytest


def _shape(rpe, n, nslots, sms=148):
    out = (C.c_int * 6)()
    ok = rpe.lib.rpe_debug_ur_shape(n, nslots, sms, out)
    return ok, list(out)


def test_config4_shape(rpe):
    ok, (P, T, hs, gx, ppc, tail) = _shape(rpe, 307200, 1024)
    assert ok == 1
    assert (P, T, hs) == (4, 512, 2)       # 16 warps = 4 per scheduler, 2 hypothesis rows
    assert gx == 74 and ppc == 2076 and tail == 28
    assert gx * hs <= 148


@pytest.mark.parametrize("sms", [148, 132, 8])
def test_shapes_cover_the_frame_and_respect_the_limits(rpe, sms):
    rng = np.random.default_rng(3)
    sizes = [1, 2, 3, 37, 1000, 1001, 4097, 65536, 100003, 307200, 640 * 480 * 4] + [int(x) for x in rng.integers(1, 2_000_000, 60)]
    for n in sizes:
        for nslots in (1, 7, 64, 300, 513, 1024):
            ok, (P, T, hs, gx, ppc, tail) = _shape(rpe, n, nslots, sms)
            if not ok:
                continue
            npairs_pad = -(-((n + 1) // 2) // 8) * 8
            assert gx * hs <= max(sms, hs)                        # one wave
            assert gx * ppc >= npairs_pad                        # every pair belongs to a column
            assert ppc % 2 == 0                                  # columns start at a multiple of 4 correspondences
            assert T % 32 == 0 and 64 <= T <= 1024 and P in (2, 3, 4)
            assert T * P + tail >= ppc and 0 <= tail <= 256      # hot loop + tail cover the column
            assert T * P <= 4096                                 # packed 16-bit vote fields cannot overflow
            nh = -(-nslots // hs)
            smem = nh * 4 + 16 + 256 * 48 + max(ppc * 48, (T // 32) * nh * 8)
            assert smem <= 227 * 1024


def test_rejects_more_than_one_hypothesis_column(rpe):
    assert _shape(rpe, 307200, 1025)[0] == 0
    assert _shape(rpe, 0, 10)[0] == 0
