"""world_size-2 (and 3) gloo runs on CPU of the host-side multi-GPU logic: hypothesis-range partition,
ragged all-gather of the vote table, redundant replay — equal to the single-process result. The per-range
votes come from the CPU oracle here (no GPU in this test); on the GPU box the same functions carry the
votes scored by the CUDA kernels (bench.py, tests/test_gpu_multi.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, method, n, H, out_q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rgbd_pose_estimation_b200 as rpe
    from rgbd_pose_estimation_b200 import sharding
    from tests import orc
    orc.set_math_mode(orc.DET)
    q, t = rpe.sim_pose(5)
    d = rpe.sim_2d_3d_nl(6, q, t, n)
    arrs = {k: d[k] for k in ("bv", "xc", "nc", "xw", "nw")}
    S = rpe.sample_table(1, n, 3 if method == 0 else 4, H)
    kw = dict(thr3d=0.2, cos_thr=float(np.cos(np.arctan(8.0 / 585.0))), cos_nl=float(np.cos(0.1)), confidence=0.99)
    full = orc.ransac(method, S, full=True, **kw, **arrs)  # every rank can compute the truth to compare against
    n_slots = full["votes"].shape[0]
    b, e = sharding.slot_range(rank, world, n_slots)
    votes = sharding.gather_votes(dist, full["votes"][b:e], rank, world, n_slots)
    win, best, it = sharding.replay(votes, method, n, 0.99)
    frames = sharding.frame_indices(rank, world, 10)
    out_q.put((rank, bool(np.array_equal(votes, full["votes"])), (win, best, it) == (full["winner"], full["max_votes"], full["iter_final"]), frames))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,method", [(2, 0), (2, 5), (3, 2)])
def test_hypothesis_sharding_gloo(world, method):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, method, 600, 257, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = []
    for rank, votes_ok, replay_ok, frames in res:
        assert votes_ok and replay_ok
        covered += frames
    assert sorted(covered) == list(range(10))


def test_slot_ranges_partition():
    from rgbd_pose_estimation_b200 import sharding
    for n_slots in (1, 7, 1024, 3072, 1000):
        for world in (1, 2, 3, 4, 8):
            r = [sharding.slot_range(k, world, n_slots) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n_slots
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1


def test_host_replay_equals_oracle(orc, rpe):
    from rgbd_pose_estimation_b200 import sharding
    orc.set_math_mode(orc.DET)
    q, t = rpe.sim_pose(9)
    Q, P, _ = rpe.sim_3d_3d(10, q, t, 700)
    S = rpe.sample_table(2, 700, 3, 500)
    full = orc.ransac(0, S, thr3d=0.25, confidence=0.9999, full=True, xc=P, xw=Q)
    assert sharding.replay(full["votes"], 0, 700, 0.9999) == (full["winner"], full["max_votes"], full["iter_final"])
    orc.set_math_mode(orc.LIBM)
