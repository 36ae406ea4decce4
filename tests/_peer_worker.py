"""Worker of tests/test_gpu_multi.py::test_peer_memory_exchange_two_processes (launched with torch.distributed.run,
one process per GPU): a hypothesis-sharded frame through rpe_ransac_sharded, votes exchanged over peer memory."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rgbd_pose_estimation_b200 as rpe  # noqa: E402
from rgbd_pose_estimation_b200 import sharding  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    local = local % max(1, torch.cuda.device_count())  # one-GPU box: both ranks share device 0 (IPC works there too)
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    out = {"rank": rank}
    with rpe.Context(local) as ctx:
        assert sharding.peer_setup(dist, ctx, rank, world)
        ctx.peer_set_timeout_ms(20000)  # two processes time-slicing one GPU exchange slowly; never a false time-out
        results = []
        for trial, (n, H, method, m) in enumerate([(20000, 1024, "shinji", 3), (5000, 300, "shinji", 3), (6000, 256, "nl_shinji_kneip", 4)]):
            q, t = rpe.sim_pose(3 + trial)
            d = rpe.sim_2d_3d_nl(4 + trial, q, t, n)
            arrs = {k: d[k] for k in ("bv", "xc", "nc", "xw", "nw")} if m == 4 else {"xc": d["xc"], "xw": d["xw"]}
            S = rpe.sample_table(1 + trial, n, m, H)
            th = dict(thr3d=0.25, cos_thr2d=float(np.cos(np.arctan(np.float32(8.0) / np.float32(585.0)))),
                      cos_thrN=float(np.cos(np.float32(0.1))))
            ctx.upload(**arrs)
            for rep in range(3):  # several frames back to back: the two alternating tables and the epochs
                r = ctx.ransac_sharded(method, S, confidence=0.99, **th)
            slots = H * rpe.method_slots(rpe.METHODS[method])
            results.append({"winner": r["winner"], "max_votes": r["max_votes"], "iter_final": r["iter_final"],
                            "votes_crc": int(np.bitwise_xor.reduce(ctx.get_votes(slots).astype(np.int64) * (np.arange(slots) + 1)))})
            if rank == 0:
                import orc
                orc.set_math_mode(orc.DET)
                ref = orc.ransac(rpe.METHODS[method], S, thr3d=th["thr3d"], cos_thr=th["cos_thr2d"], cos_nl=th["cos_thrN"],
                                 confidence=0.99, full=True, **arrs)
                results[-1]["oracle_ok"] = bool(np.array_equal(ctx.get_votes(slots), ref["votes"]) and
                                                (r["winner"], r["max_votes"], r["iter_final"]) ==
                                                (ref["winner"], ref["max_votes"], ref["iter_final"]))
        ctx.peer_status()
        out["results"] = results
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print("PEER_RESULT " + json.dumps(gathered))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
