import sys, numpy as np
sys.path.insert(0, '.')
import rgbd_pose_estimation_b200 as rpe
THR, CONF, H = 0.25, 0.999, 128
n, ring, total, seed = 6000, 7, 23, 77
frames = []
for i in range(ring):
    q, t = rpe.sim_pose(100 + i)
    Q, P, _ = rpe.sim_3d_3d(200 + i, q, t, n, noise=0.1, outlier_ratio=0.4 + 0.02 * (i % 5))
    frames.append((Q, P))
def pc(a):
    b = rpe.pinned_empty(a.shape, a.dtype); b[:] = a; return b
host = [{"xw": pc(Q), "xc": pc(P), "mask": rpe.pinned_empty((2, n), np.int16)} for Q, P in frames]
for contexts, threads in [(1,1),(5,1),(2,2),(4,2),(5,3),(5,3),(6,3),(3,3),(8,4)]:
    with rpe.Sequence(0, "shinji", H, thr3d=THR, confidence=CONF, refit=("kabsch",), sample_seed=seed, contexts=contexts, threads=threads) as seq:
        seq.set_frames(host)
        r0, r1 = seq.run(5, total)
        bad = [(i, r1[i].refit_ok, r1[i].winner, r1[i].max_votes, list(r1[i].q), r0[i].winner, r0[i].max_votes, list(r0[i].n_inliers), list(r1[i].n_inliers)) for i in range(total) if r1[i].refit_ok != 1]
        print(contexts, threads, "bad:", bad)
