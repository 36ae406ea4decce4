"""Host-side partitioning for the two multi-GPU modes and the sequential rule on a vote table.

* frames sharded (config #5): rank r takes frames r, r+G, r+2G, ... — no data-path collective;
* hypotheses sharded (config #4): rank r scores the slot range ``slot_range(r, G, n_slots)``; the int32 vote
  table is all-gathered (4 KB at H = 1024) and every rank replays the reference's keep-best / adaptive-stop rule
  (AbsoluteOrientation.hpp:145-151, P3P.hpp:296-318) redundantly, so all ranks agree on the winner.
"""
from __future__ import annotations

import numpy as np

from . import capi

_MODEL_POINTS = {0: 3, 1: 4, 2: 3, 3: 3, 4: 3, 5: 3, 6: 4}
_MODALITIES = {0: 1, 1: 1, 2: 2, 3: 2, 4: 2, 5: 3, 6: 1}


def frame_indices(rank: int, world: int, n_frames: int) -> list[int]:
    return list(range(rank, n_frames, world))


def slot_range(rank: int, world: int, n_slots: int) -> tuple[int, int]:
    """Contiguous, balanced ranges covering [0, n_slots) exactly once."""
    base, rem = divmod(n_slots, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_votes(dist, votes_local: np.ndarray, rank: int, world: int, n_slots: int) -> np.ndarray:
    """All-gather ragged slot ranges with any torch.distributed backend (gloo on CPU, NCCL on GPU tensors)."""
    import torch
    sizes = [slot_range(r, world, n_slots) for r in range(world)]
    width = max(e - b for b, e in sizes)
    mine = torch.full((width,), -1, dtype=torch.int32)
    mine[: votes_local.shape[0]] = torch.from_numpy(np.ascontiguousarray(votes_local, dtype=np.int32))
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = np.empty(n_slots, np.int32)
    for r, (b, e) in enumerate(sizes):
        out[b:e] = parts[r][: e - b].numpy()
    return out


def replay(votes: np.ndarray, method: int, n_corr: int, confidence: float, iter_in: int | None = None):
    """The reference's sequential rule over a complete vote table (slot = iteration*slots + slot, -1 = empty).
    Returns (winner, max_votes, iter_final). Uses the same bit-reproducible rule as the device replay kernel."""
    S = capi.method_slots(method)
    K = _MODEL_POINTS[method]
    m = _MODALITIES[method]
    H = votes.shape[0] // S if iter_in is None else iter_in
    best, win, it = -1, -1, H
    ii = 0
    nf = np.float32(n_corr)
    while ii < it and ii < votes.shape[0] // S:
        for s in range(S):
            v = int(votes[ii * S + s])
            if v < 0:
                continue
            if v > best:
                best, win = v, ii * S + s
                if m == 1:
                    ep = np.float32(n_corr - v) / nf
                else:
                    ep = np.float32(np.float32(n_corr * m - v) / nf) / np.float32(m)
                it = capi.update_num_iters(float(np.float32(confidence)), float(ep), K, it)
        ii += 1
    return win, best, it


def peer_setup(dist, ctx, rank: int, world: int) -> bool:
    """Exchange the CUDA IPC handles of the contexts' vote-exchange blocks over any torch.distributed backend and map
    every peer's block (one process per GPU of a node). Afterwards ``ctx.exchange_votes(b, e)`` /
    ``ctx.ransac_sharded(...)`` move vote slices straight through peer memory.

    Collective-safe: every rank takes part in every collective of this function whatever fails locally, and the return
    value (all ranks succeeded) is the same on all ranks — so a rank that cannot export or map a block never leaves
    the others waiting in a barrier."""
    import torch
    try:
        mine = torch.from_numpy(ctx.peer_export().copy())
        ok = 1
    except Exception:
        mine = torch.zeros(64, dtype=torch.uint8)
        ok = 0
    if dist is None or world == 1:
        if ok:
            ctx.peer_import(0, 1, mine.numpy())
        return bool(ok)
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    parts = [torch.empty(64, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, mine.to(dev))
    flag = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 1:
        try:
            ctx.peer_import(rank, world, torch.stack(parts).cpu().numpy())
        except Exception:
            ok = 0
    else:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return int(flag.item()) == 1
