#!/usr/bin/env python
"""bench.py — headline benchmark of the robust absolute-pose hot path on B200.

Metric (BASELINE.json): hypothesis x correspondence evaluations per second ("hyp-corr evals/s") and
frames/s on the dense-frame workload: one 640x480 RGB-D frame = 307 200 3-D/3-D correspondences scored
against 1 024 RANSAC hypotheses (config #4), whole pipeline per frame:
    sample table -> hypgen_ao -> tiled scorer (+ exact fix-up) -> replay of the adaptive rule -> mask
    -> Kabsch refit (shinji_ls1) -> LM refinement on SE3.
A "step" is `frames_per_step` frames per GPU. At N GPUs frames are sharded (config #5: no data-path
collective, weak scaling); the hypothesis-sharded single-frame mode of config #4 (4 KB NCCL all-gather of
votes) is reported beside it as `single_frame_sharded`.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 5 --warmup 1      # CPU path (oracle port), all host cores

Only the `cpu_baseline` leg and `--impl reference` touch oracle/ (as the thing being timed on the CPU,
never as part of the GPU path).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CORR = 307200
N_HYP = 1024
THR3D = 0.25
CONF = 0.9999
NOISE = 0.1
OUTLIER = 0.5
FLOP_PER_EVAL = 26          # 9 FMA + 3 sub + (1 mul + 2 FMA) of the minimal matrix form
BYTES_PER_CORR = 24         # 2 x (3 x f32) compulsory input per correspondence
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12
METHOD_SHINJI = 0


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def make_frames(rpe, count, n, seed0=1000):
    frames = []
    for i in range(count):
        q, t = rpe.sim_pose(seed0 + 2 * i)
        Q, P, _ = rpe.sim_3d_3d(seed0 + 2 * i + 1, q, t, n, noise=NOISE, outlier_ratio=OUTLIER)
        frames.append({"q": q, "t": t, "xw": Q, "xc": P})
    return frames


class ClockSampler:
    """nvidia-smi clock / throttle sampling DURING the timed region (B200_PROFILING.md recipe). The sampler is
    started early (nvidia-smi needs ~0.5 s to produce its first line); only samples whose timestamp falls inside
    a window [t0, t1] of time.time() are summarised."""

    FIELDS = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass

    def summarise(self, t0, t1):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm, mx, pw = [], [], []
        reasons = set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 10:
                    continue
                try:
                    ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    if ts < t0 - 0.05 or ts > t1 + 0.05:
                        continue
                    sm.append(float(p[2]))
                    mx.append(float(p[3]))
                    pw.append(float(p[4]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[6:10]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        except Exception:
            pass
        if sm:
            out.update({"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "samples": len(sm),
                        "power_w_max": float(max(pw)) if pw else None})
        out["reasons"] = sorted(reasons)
        return out

    def cleanup(self):
        try:
            os.unlink(self.path)
        except Exception:
            pass


def load_measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def load_profile_traffic():
    """dram bytes per launch of the scoring kernel from the committed ncu summary, if present."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "score_kernel_traffic.json")))
        return d.get("dram_bytes_per_launch")
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# CPU path (oracle port of the reference), used by --impl reference and by the cpu_baseline leg
# ------------------------------------------------------------------------------------------------
def cpu_run(frame, samples, nthreads, full=True):
    from tests import orc  # the CPU oracle: the thing being timed here
    orc.set_math_mode(orc.DET)
    return orc.ransac(METHOD_SHINJI, samples, thr3d=THR3D, confidence=CONF, full=full, nthreads=nthreads,
                      xc=frame["xc"], xw=frame["xw"], want_arrays=False)


def time_reference_sources(frame, iters=16):
    """Informational: the reference's OWN shinji_ransac2 loop (oracle/_ref/libref_shim.so: /root/reference/pose/*.hpp
    compiled unmodified against the Eigen / Sophus API stand-in of oracle/ref_shim/), single thread like the
    reference, `iters` iterations with the adaptive stop disabled (confidence 1). Slower than the oracle port, which is
    why the port — multi-threaded — stays the reported baseline. None when the library is not there."""
    try:
        from tests import refshim
        if not os.path.exists(refshim.SO):
            return None
        refshim.ransac(0, 1, 2, thr3d=THR3D, confidence=1.0, xc=frame["xc"], xw=frame["xw"])  # warm-up
        t0 = time.perf_counter()
        r = refshim.ransac(0, 1, iters, thr3d=THR3D, confidence=1.0, xc=frame["xc"], xw=frame["xw"])
        dt = time.perf_counter() - t0
        if r["iter_final"] != iters:
            return None
        return {"value": iters * N_CORR / dt, "unit": "evals/s", "cores": 1,
                "sample": f"{iters} iterations x {N_CORR} correspondences of shinji_ransac2 as written in the reference"}
    except Exception as e:  # never let the informational leg break the line
        return {"unavailable": str(e)[:200]}


def run_reference(args, rank):
    if rank != 0:
        return
    import rgbd_pose_estimation_b200 as rpe
    cores = os.cpu_count() or 1
    h_sample = args.ref_hyp
    frame = make_frames(rpe, 1, N_CORR)[0]
    samples = rpe.sample_table(1, N_CORR, 3, N_HYP)[:h_sample]
    for _ in range(args.warmup):
        cpu_run(frame, samples, cores)
    t0 = time.perf_counter()
    evals = 0
    for _ in range(args.steps):
        r = cpu_run(frame, samples, cores)
        evals += r["evals"]
    dt = time.perf_counter() - t0
    value = evals / dt
    sample_desc = (f"1 frame x {h_sample} of {N_HYP} hypotheses x {N_CORR} correspondences per step, every hypothesis "
                   f"scored (no early stop), hypotheses sharded over {cores} host threads")
    line = {
        "impl": "reference", "metric": "hyp-corr evals/s", "value": value, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "frames_per_s": value / (N_CORR * N_HYP),
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference cannot be compiled here (needs Eigen); this is the oracle port of its CPU path (oracle/README.md)",
    }
    line["cpu_baseline"]["reference_sources_single_thread"] = time_reference_sources(frame)
    print(json.dumps(line))


def workload_config(args, world):
    return {
        "workload": "config4/5: dense 640x480 RGB-D frame, 307200 3-D/3-D correspondences x 1024 hypotheses, "
                    "shinji_ransac2 + shinji_ls1 + LM refinement; frames sharded across GPUs",
        "n_correspondences": N_CORR, "n_hypotheses": N_HYP, "outlier_ratio": OUTLIER, "noise_m": NOISE,
        "thr3d_m": THR3D, "confidence": CONF, "frames_per_step_per_gpu": args.frames_per_step,
        "distinct_frames_per_gpu": args.ring, "contexts_per_gpu": args.contexts, "gn_max_iters": args.gn_iters,
        "l2_policy": f"inputs larger than L2: ring of {args.ring} distinct frames x 7.4 MB per GPU (scored straight from these arrays)",
        "parallelism": f"frames sharded over {world} GPU(s), no data-path collective",
    }


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def bind_near_gpu(torch, local_rank):
    """Pin this rank's host threads (and with them the first-touch placement of its page-locked buffers) to the CPUs
    of the GPU's NUMA node, so that the e2e leg's H2D/D2H copies do not cross the socket interconnect."""
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"pci": bdf, "cpus": len(cpus)}
    except Exception as e:  # no sysfs entry / no permission: leave the affinity alone
        return {"skipped": repr(e)[:80]}
    return {"skipped": "no local cpus"}


def run_gpu(args, rank, world, local_rank):
    import torch
    import rgbd_pose_estimation_b200 as rpe

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    binding = bind_near_gpu(torch, local_rank) if world > 1 else {"skipped": "single GPU"}
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    dev = torch.device("cuda", local_rank)
    # --- data: ring of distinct frames, device-resident for `value`, page-locked host copies for `e2e`
    frames = make_frames(rpe, args.ring, N_CORR, seed0=1000 + 7919 * rank)
    d_xw = [torch.from_numpy(f["xw"]).to(dev) for f in frames]
    d_xc = [torch.from_numpy(f["xc"]).to(dev) for f in frames]
    tables = [rpe.sample_table(1 + i + 100 * rank, N_CORR, 3, N_HYP) for i in range(args.ring)]
    d_tab = [torch.from_numpy(t).to(dev) for t in tables]
    h_xw, h_xc, h_tab = [], [], []
    for f, t in zip(frames, tables):
        a = rpe.pinned_empty((N_CORR, 3), np.float32)
        a[:] = f["xw"]
        b = rpe.pinned_empty((N_CORR, 3), np.float32)
        b[:] = f["xc"]
        c = rpe.pinned_empty((N_HYP, 4), np.int32)
        c[:] = t
        h_xw.append(a)
        h_xc.append(b)
        h_tab.append(c)

    streams = [torch.cuda.Stream(device=dev) for _ in range(args.contexts)]
    ctxs = [rpe.Context(local_rank, stream=s.cuda_stream) for s in streams]
    h_mask = [rpe.pinned_empty((2, N_CORR), np.int16) for _ in ctxs]
    for c in ctxs:
        c.enable_stage_timing(2)  # timed legs: only the two events around the tiled scoring kernel (roofline)

    def frame_device(ci, fi):
        c = ctxs[ci]
        c.upload_device(N_CORR, xc=d_xc[fi].data_ptr(), xw=d_xw[fi].data_ptr())
        r0 = c.ransac_async(METHOD_SHINJI, d_tab[fi].data_ptr(), H=N_HYP, thr3d=THR3D, confidence=CONF)
        r1 = c.refit_async("kabsch_inliers")
        r2 = c.refit_async("gn", max_iters=args.gn_iters)
        return r0, r1, r2

    def frame_host(ci, fi):
        c = ctxs[ci]
        c.upload_async(xc=h_xc[fi], xw=h_xw[fi])
        r0 = c.ransac_async(METHOD_SHINJI, h_tab[fi], thr3d=THR3D, confidence=CONF, mask=h_mask[ci])
        r1 = c.refit_async("kabsch_inliers")
        r2 = c.refit_async("gn", max_iters=args.gn_iters)
        return r0, r1, r2

    def sync_all():
        for c in ctxs:
            c.sync()
            c._keep = []

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def run_steps(step_fn, steps):
        k = 0
        last = None
        for _ in range(steps):
            for _f in range(args.frames_per_step):
                last = step_fn(k % args.contexts, k % args.ring)
                k += 1
        return last

    host_issue_ms = [0.0]  # wall time the host spends enqueueing one frame (no GPU wait unless a queue fills up)

    def timed(step_fn, steps, stage_acc=None):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        k = 0
        last = None
        t_issue = time.perf_counter()
        for _s in range(steps):
            for _f in range(args.frames_per_step):
                last = step_fn(k % args.contexts, k % args.ring)
                k += 1
            if stage_acc is not None and (_s % 4 == 3 or _s == steps - 1):
                sync_all()
                for c in ctxs:
                    stage_acc.append(c.last_stage_ms())
        host_issue_ms[0] = (time.perf_counter() - t_issue) * 1e3 / max(1, steps * args.frames_per_step)
        sync_all()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        barrier()
        return ms, last

    sampler = ClockSampler(local_rank)
    sampler.start()
    # --- FFMA peak of this device, same process/run (roofline denominator)
    ffma_scalar, ffma_packed = ctxs[0].measure_ffma_tflops(100)

    # --- warm-up
    run_steps(frame_device, args.warmup)
    sync_all()
    run_steps(frame_host, max(1, args.warmup // 2))
    sync_all()

    # --- `value`: inputs resident in HBM
    launches0 = sum(c.launch_count() for c in ctxs)
    stage_acc = []
    w0 = time.time()
    ms_dev, last_dev = timed(frame_device, args.steps, stage_acc)
    issue_dev = host_issue_ms[0]
    w1 = time.time()
    launches = sum(c.launch_count() for c in ctxs) - launches0
    # --- `e2e`: host buffers, H2D + D2H inside the timed region
    w2 = time.time()
    ms_e2e, last_e2e = timed(frame_host, args.steps)
    issue_e2e = host_issue_ms[0]
    w3 = time.time()
    sampler.stop()
    clocks = sampler.summarise(w0, w1)
    clocks_e2e = sampler.summarise(w2, w3)
    sampler.cleanup()

    # --- per-stage device times of one frame running alone (all stage events on, one context, inputs in HBM)
    ctxs[0].enable_stage_timing(1)
    stage_alone = []
    for i in range(6):
        frame_device(0, i % args.ring)
        ctxs[0].sync()
        ctxs[0]._keep = []
        if i > 0:
            stage_alone.append(ctxs[0].last_stage_ms())
    ctxs[0].enable_stage_timing(2)

    # --- single blocking frame latency through the C-ABI with host buffers (one context)
    lat = []
    for i in range(5):
        t0 = time.perf_counter()
        c = ctxs[0]
        c.upload_async(xc=h_xc[i % args.ring], xw=h_xw[i % args.ring])
        c.ransac_async(METHOD_SHINJI, h_tab[i % args.ring], thr3d=THR3D, confidence=CONF, mask=h_mask[0])
        c.refit_async("kabsch_inliers")
        c.refit_async("gn", max_iters=args.gn_iters)
        c.sync()
        lat.append((time.perf_counter() - t0) * 1e3)
        c._keep = []

    frames_total = args.steps * args.frames_per_step * world
    evals_total = frames_total * N_CORR * N_HYP
    value = evals_total / (ms_dev * 1e-3)
    e2e_value = evals_total / (ms_e2e * 1e-3)
    if dist is not None:
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())

    # --- roofline of the dominant kernel (tiled scorer), live CUDA-event timing of that kernel alone
    fast_ms = [s["score_fast"] for s in stage_acc if s["score_fast"] > 0]
    stage_mean = {k: float(np.mean([s[k] for s in stage_alone])) for k in stage_alone[0]} if stage_alone else {}
    roofline = None
    if fast_ms:
        k_ms = float(np.mean(fast_ms))
        achieved = FLOP_PER_EVAL * N_CORR * N_HYP / (k_ms * 1e-3) / 1e12
        peak = max(ffma_scalar, ffma_packed)
        peaks = load_measured_peaks()
        hbm_peak = peaks["hbm_gbs"] if peaks else 6650.0
        alg_bytes = BYTES_PER_CORR * N_CORR  # compulsory: every correspondence read once
        roofline = {
            "bound": "fp32", "kernel": "score3d_raw_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak if peak > 0 else None,
            "peak_source": "FFMA/FFMA2 microbenchmark measured in this run (MEASURED_PEAKS.json has no FP32 CUDA-core figure)",
            "frac_of_nominal": achieved / NOMINAL_FP32_TFLOPS, "nominal_peak": NOMINAL_FP32_TFLOPS,
            "ffma_scalar_tflops": ffma_scalar, "ffma2_packed_tflops": ffma_packed,
            "kernel_ms": k_ms, "flop_per_eval": FLOP_PER_EVAL, "evals_per_launch": N_CORR * N_HYP,
            "kernel_ms_note": f"mean over the timed region while {args.contexts} contexts share the GPU (the kernel "
                              "co-runs with other frames' small kernels); kernel_alone_ms is the same kernel with one "
                              "context",
            "kernel_alone_ms": stage_mean.get("score_fast"),
            "frac_alone": (FLOP_PER_EVAL * N_CORR * N_HYP / (stage_mean["score_fast"] * 1e-3) / 1e12 / peak
                           if stage_mean.get("score_fast") and peak > 0 else None),
            "traffic": load_profile_traffic(),
            "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9,
                    "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6.65 TB/s",
                    "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak},
        }

    # --- hypothesis-sharded single frame (config #4) with an NCCL all-gather of the vote table
    sharded = None
    if dist is not None:
        try:
            sharded = run_single_frame_sharded(args, torch, dist, rpe, ctxs[0], streams[0], frames, tables, rank, world, dev)
        except Exception as e:  # never lose the headline line over the auxiliary measurement
            sharded = {"error": repr(e)}

    # --- CPU baseline (rank 0, N=1 only): the oracle port of the reference CPU path on the host cores
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        r_mt = cpu_run(frames[0], tables[0], cores)  # 1 frame, all 1024 hypotheses scored
        h1 = 96
        r_1t = cpu_run(frames[0], tables[0][:h1], 1)
        r_es = cpu_run(frames[0], tables[0], 1, full=False)  # the reference's own early-stopping loop
        gpu_first = None
        c = ctxs[0]
        c.upload_device(N_CORR, xc=d_xc[0].data_ptr(), xw=d_xw[0].data_ptr())
        g = c.ransac(METHOD_SHINJI, tables[0], thr3d=THR3D, confidence=CONF, want_mask=False)
        gpu_first = {"winner": g["winner"], "max_votes": g["max_votes"], "iter_final": g["iter_final"]}
        cpu_baseline = {
            "value": r_mt["evals"] / r_mt["seconds"], "unit": "evals/s", "cores": cores, "kind": "port",
            "sample": f"1 frame x {N_HYP} hypotheses x {N_CORR} correspondences, every hypothesis scored, "
                      f"hypotheses sharded over {cores} host threads ({r_mt['seconds']:.2f} s wall)",
            "frames_per_s": 1.0 / r_mt["seconds"],
            "single_thread": {"value": r_1t["evals"] / r_1t["seconds"], "unit": "evals/s",
                              "sample": f"{h1} hypotheses x {N_CORR} correspondences"},
            "early_stop_loop": {"seconds_per_frame": r_es["seconds"], "iterations_run": r_es["iters_run"],
                                "note": "the reference's literal loop stops at the adaptive Iter; same winner"},
            "cpu_gpu_agree": bool(gpu_first["winner"] == r_mt["winner"] and gpu_first["max_votes"] == r_mt["max_votes"]
                                  and gpu_first["iter_final"] == r_mt["iter_final"]),
        }
        # informational: the reference's own loop (its sources on the Eigen stand-in), single thread; never raises
        cpu_baseline["reference_sources_single_thread"] = time_reference_sources(frames[0])

    if rank == 0:
        h2d = args.frames_per_step * (2 * N_CORR * 12 + N_HYP * 16)
        d2h = args.frames_per_step * (2 * N_CORR * 2 + 3 * 72 + 12)
        line = {
            "metric": "hyp-corr evals/s", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
            "frames_per_s": frames_total / (ms_dev * 1e-3),
            "ms_per_frame_per_gpu": ms_dev / (args.steps * args.frames_per_step),
            "host_issue_ms_per_frame": issue_dev,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "frames_per_s": frames_total / (ms_e2e * 1e-3), "ms_per_step": ms_e2e / args.steps,
                    "single_frame_latency_ms": {"median": float(np.median(lat)), "min": float(min(lat))},
                    "host_issue_ms_per_frame": issue_e2e, "numa_binding_rank0": binding,
                    "clocks": clocks_e2e,
                    "path": "rpe_upload(host pinned) + rpe_ransac_async + rpe_refit_async x2 + mask/pose D2H, "
                            f"{args.contexts} contexts round-robin"},
            "gpu_launches": launches,
            "roofline": roofline,
            "stage_ms_mean": stage_mean,
            "cpu_baseline": cpu_baseline,
            "single_frame_sharded": sharded,
            "last_result": {"max_votes": int(last_dev[0].max_votes), "iter_final": int(last_dev[0].iter_final),
                            "n_borderline": int(last_dev[0].n_borderline), "gn_evals": int(last_dev[2].refit_evals)},
        }
        print(json.dumps(line))
    for c in ctxs:
        c.close()
    if dist is not None:
        dist.destroy_process_group()


class _DevArray:
    """Minimal __cuda_array_interface__ wrapper so torch can alias a device buffer owned by the C library."""

    def __init__(self, ptr, n, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def run_single_frame_sharded(args, torch, dist, rpe, ctx, stream, frames, tables, rank, world, dev, reps=50):
    """Config #4: correspondences replicated, every rank scores H/world hypotheses, one all-gather of the
    int32 vote table (4 KB) over NCCL, then every rank replays the adaptive rule redundantly."""
    # identical frame on every rank: regenerate rank 0's frame 0
    f0 = make_frames(rpe, 1, N_CORR, seed0=1000)[0]
    tab = rpe.sample_table(1, N_CORR, 3, N_HYP)
    from rgbd_pose_estimation_b200 import sharding
    sb, se = sharding.slot_range(rank, world, N_HYP)
    if N_HYP % world:
        return {"skipped": "N_HYP not divisible by world size (in-place all-gather needs equal chunks)"}
    with torch.cuda.stream(stream):
        xw = torch.from_numpy(f0["xw"]).to(dev)
        xc = torch.from_numpy(f0["xc"]).to(dev)
        ctx.upload_device(N_CORR, xc=xc.data_ptr(), xw=xw.data_ptr())
        tab_dev = torch.from_numpy(tab).to(dev)
        ctx.generate(METHOD_SHINJI, tab_dev.data_ptr(), H=N_HYP)
        votes = torch.as_tensor(_DevArray(ctx.votes_device_ptr(), N_HYP), device=dev)
        times = []
        res = None
        for i in range(reps + 5):
            torch.cuda.synchronize()
            dist.barrier()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.generate(METHOD_SHINJI, tab_dev.data_ptr(), H=N_HYP)  # resets the votes; generation is replicated (cheap)
            ctx.score(METHOD_SHINJI, sb, se, thr3d=THR3D)
            dist.all_gather_into_tensor(votes, votes[sb:se].clone())
            res = ctx.finish(METHOD_SHINJI, N_HYP, thr3d=THR3D, confidence=CONF, want_mask=False)
            e1.record(stream)
            torch.cuda.synchronize()
            if i >= 5:
                times.append(e0.elapsed_time(e1))
        tt = torch.tensor([float(np.median(times))], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        nccl = {"ms_per_frame": float(tt.item()), "winner": res["winner"], "max_votes": res["max_votes"],
                "iter_final": res["iter_final"]}
        # --- the same frame with the vote slices exchanged through peer memory (NVLink P2P, CUDA IPC) by our own kernel
        p2p = None
        # peer_setup is collective-safe (same verdict on every rank), so either all ranks run the loop or none does
        if not sharding.peer_setup(dist, ctx, rank, world):
            p2p = {"skipped": "CUDA IPC / peer access unavailable on this box"}
        else:
            times = []
            for i in range(reps + 5):
                torch.cuda.synchronize()
                dist.barrier()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                res2 = ctx.ransac_sharded(METHOD_SHINJI, tab_dev.data_ptr(), H=N_HYP, thr3d=THR3D, confidence=CONF)
                e1.record(stream)
                torch.cuda.synchronize()
                if i >= 5:
                    times.append(e0.elapsed_time(e1))
            # pipelined: frames enqueued back to back, no host wait in between (throughput of the sharded mode)
            torch.cuda.synchronize()
            dist.barrier()
            ea = torch.cuda.Event(enable_timing=True)
            eb = torch.cuda.Event(enable_timing=True)
            ea.record(stream)
            for i in range(reps):
                ctx.ransac_sharded(METHOD_SHINJI, tab_dev.data_ptr(), H=N_HYP, thr3d=THR3D, confidence=CONF, blocking=False)
            ctx.sync()
            ctx._keep = []
            eb.record(stream)
            torch.cuda.synchronize()
            try:
                ctx.peer_status()
                timed_out = 0.0
            except Exception:  # a peer did not publish within the kernel's 2 s bound: report it, on every rank
                timed_out = 1.0
            t2 = torch.tensor([float(np.median(times)), ea.elapsed_time(eb) / reps, timed_out], device=dev, dtype=torch.float64)
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            p2p = {"ms_per_frame": float(t2[0].item()), "ms_per_frame_pipelined": float(t2[1].item()),
                   "exchange_timed_out": bool(t2[2].item() > 0), "winner": res2["winner"], "max_votes": res2["max_votes"],
                   "iter_final": res2["iter_final"],
                   "agrees_with_nccl": bool((res2["winner"], res2["max_votes"], res2["iter_final"]) ==
                                            (res["winner"], res["max_votes"], res["iter_final"]))}
    best = nccl["ms_per_frame"] if not (p2p and "ms_per_frame" in p2p) else min(nccl["ms_per_frame"], p2p["ms_per_frame"])
    return {"ms_per_frame": best, "frames_per_s": 1e3 / best, "evals_per_s": N_CORR * N_HYP / (best * 1e-3),
            "collective": "vote slices (H/G int32 per rank) all-gathered: (a) ncclAllGather, (b) own kernel over peer memory",
            "nccl": nccl, "peer_memory": p2p, "winner": nccl["winner"], "max_votes": nccl["max_votes"],
            "iter_final": nccl["iter_final"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=96)
    ap.add_argument("--ring", type=int, default=24, help="distinct frames resident per GPU (>L2 in total)")
    ap.add_argument("--contexts", type=int, default=12, help="rpe contexts (streams) per GPU, frames round-robin")
    ap.add_argument("--gn-iters", type=int, default=3)
    ap.add_argument("--ref-hyp", type=int, default=128, help="hypotheses per step of the CPU arm (bounded sample)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
