#!/usr/bin/env python3
"""bench.py — RGB-L front-end frames/s on KITTI-sized synthetic frames (BASELINE.json metric).

A "step" = one batch of T consecutive frames of a synthetic RGB-L sequence (1241x376 image + 120k Velodyne
points each, nFeatures=2000, 8 levels: BASELINE.json configs[1]) through the whole hot path: pyramid -> FAST ->
quad-tree -> orientation + rBRIEF -> LiDAR projection -> inverse dilation -> per-keypoint depth (batched), then per
frame SearchByProjection(last frame) -> PoseOptimization (serial in time, on the device).

  value : frames/s with the batch already resident in HBM (rgbl_resident_process), CUDA-event timed
  e2e   : frames/s through the public C ABI with pinned HOST buffers (rgbl_frame_rgbl_batch):
          H2D of images + clouds and D2H of keypoints/descriptors/depths inside the timed region
  --impl reference : the CPU restatement of the reference path (oracle/) on all host cores

Multi-GPU: independent sequences shard across ranks (weak scaling, no data-path collective); one
all_reduce(MAX) of the elapsed time for the throughput report.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from orb_slam3_rgbl_b200 import synthetic as S  # noqa: E402

METRIC = "rgbl_frontend_frames_per_sec_kitti_1241x376"
UNIT = "frames/s"
WORKLOAD = "KITTI-00-like synthetic RGB-L sequence: 1241x376 + 120k Velodyne pts/frame, nFeatures=2000, 8 levels (configs[1]); frame construction + SearchByProjection(last) + PoseOptimization"


def algorithmic_bytes(levels_wh, n_cand, n_kp, n_pts, W, H, n_in):
    """Compulsory bytes per FRAME for each stage (SURVEY.md §8(d) formulas)."""
    Ssum = sum(w * h for w, h in levels_wh)
    S0, S7 = levels_wh[0][0] * levels_wh[0][1], levels_wh[-1][0] * levels_wh[-1][1]
    A = W * H
    return {
        "pyramid": (Ssum - S7) + (Ssum - S0),
        "fast": Ssum + 8 * n_cand,
        "compact": 8 * n_cand,
        "blur": 2 * Ssum,
        "describe": n_kp * (749 + 512 + 32 + 28),
        "depth_project": 16 * n_pts + 4 * n_in,
        "depth_resolve_dilate": 8 * A + 4 * A,   # read index/raw map once, write Processed (+ the zero-fill the reference does)
        "depth_gather": 12 * n_kp,
    }


class ClockSampler:
    """nvidia-smi sampler running during the timed region (recipe in B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_batch_inputs(rank: int, T: int):
    """T consecutive frames of the synthetic sequence `rank` (plane world): images, clouds, P, pose0."""
    seq = S.PlaneSequence(1000 + rank, T + 1)
    return [seq.image(t) for t in range(T)], [seq.cloud(t) for t in range(T)], seq


CAM = (S.KITTI_FX, S.KITTI_FY, S.KITTI_CX, S.KITTI_CY, S.KITTI_BF)
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures (profiles/), else None
TRAFFIC = {"fast": 42294016 + 183552}   # profiles/r01_ncu_full_fast_cells_kernel.txt (32-frame launch)


def cpu_track_step(fv_args_cur, last, pose, sf):
    """oracle SearchByProjection(last frame) + PoseOptimization for one frame (same glue as rgbl_resident_track)."""
    import oracle
    import tracking_data as TD
    xw, ok = TD.chain_unproject(last, pose)
    fv = oracle.FrameView(*fv_args_cur)
    cur_k, cur_ur = fv_args_cur[0], fv_args_cur[1]
    nm, match = oracle.search_by_projection_last(fv, pose, pose, ok.astype(np.uint8), xw, last["d"], last["k"]["octave"], last["k"]["angle"],
                                                 np.ones(len(ok), np.uint8), 15.0)
    m = np.nonzero(match >= 0)[0]
    obs = np.stack([cur_k["x"][m], cur_k["y"][m], cur_ur[m]], 1).astype(np.float32)
    sc = sf[cur_k["octave"][m]]
    inv_s2 = (np.float32(1.0) / (sc * sc).astype(np.float32)).astype(np.float32)
    st = (cur_ur[m] >= 0).astype(np.uint8)
    ni, pose2, _ = oracle.pose_optimize(pose, xw[match[m]], obs, inv_s2, st, *CAM)
    return pose2


def cpu_frame(ex, img, pts, P, mask):
    """Frame construction on the CPU: ORBextractor::operator() + DepthModule::CalculateDepthFromPcd (oracle)."""
    import oracle
    k, d, _ = ex(img)
    dep, ur, _, _ = oracle.depth_from_pcd(pts, P, S.KITTI_W, S.KITTI_H, mask, S.KITTI_BF, k, k)
    return dict(k=k, d=d, depth=dep, ur=ur)


def cpu_baseline_single(imgs, pcs, seq, budget_s=12.0):
    """Oracle (port) on ONE host core over a bounded sample of the same workload (frame construction + tracking)."""
    import oracle
    sys.path.insert(0, str(ROOT / "tests"))
    ex = oracle.Extractor(2000)
    sf = ex.scale_factors.copy()
    mask = S.structuring_element("diamond", 5)
    last = cpu_frame(ex, imgs[0], pcs[0], seq.P, mask)            # warm-up / frame 0
    pose = seq.pose(0)
    n, t0 = 0, time.perf_counter()
    while True:
        t = 1 + n % (len(imgs) - 1)
        if t == 1:
            last = cpu_frame(ex, imgs[0], pcs[0], seq.P, mask); pose = seq.pose(0); n += 1
        cur = cpu_frame(ex, imgs[t], pcs[t], seq.P, mask)
        pose = cpu_track_step((cur["k"], cur["ur"], cur["d"], S.KITTI_W, S.KITTI_H, sf) + CAM, last, pose, sf)
        last = cur
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= 400:
            break
    return n / el, n


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement with every host thread it can use (rank 0 only)."""
    if rank != 0:
        return
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    sys.path.insert(0, str(ROOT / "tests"))
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    frames_per_step = max(8, min(2 * cores, 32))
    imgs, pcs, seq = make_batch_inputs(0, frames_per_step)
    mask = S.structuring_element("diamond", 5)
    exs = [oracle.Extractor(2000) for _ in range(cores)]
    sf = exs[0].scale_factors.copy()

    def work(i):
        return cpu_frame(exs[i % cores], imgs[i % len(imgs)], pcs[i % len(pcs)], seq.P, mask)

    def step(pool):
        # frame construction is independent per frame (all threads); the tracking chain is serial in time
        frs = list(pool.map(work, range(frames_per_step)))
        pose = seq.pose(0)
        for t in range(1, len(frs)):
            pose = cpu_track_step((frs[t]["k"], frs[t]["ur"], frs[t]["d"], S.KITTI_W, S.KITTI_H, sf) + CAM, frs[t - 1], pose, sf)

    with ThreadPoolExecutor(cores) as pool:
        for _ in range(args.warmup):
            step(pool)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(pool)
        el = time.perf_counter() - t0
    fps = frames_per_step * args.steps / el
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": frames_per_step},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{frames_per_step} frames/step x {args.steps} steps, oracle (C++ restatement, scalar): frame construction on {cores} threads, tracking chain serial"},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="frames per step (per GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bow", action="store_true", help="skip the ComputeBoW side measurement")
    ap.add_argument("--multi-sequences", type=int, default=4, help="extra capacity figure: independent sequences tracked concurrently on one GPU (N=1 only; 0/1 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    from orb_slam3_rgbl_b200 import frontend as F

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    from orb_slam3_rgbl_b200.dist import Reporter
    rep = Reporter("nccl", torch.device("cuda", local_rank))      # sequences shard over ranks; NCCL only for the report

    T = args.batch
    imgs, pcs, seq = make_batch_inputs(rank, T)
    P = seq.P
    max_pts = max(p.shape[1] for p in pcs)
    ctx = F.Context(S.KITTI_W, S.KITTI_H, 2000, max_batch=T, max_points=max_pts, device=local_rank)
    prm = F.make_depth_params(bf=S.KITTI_BF)
    batch = F.RgblBatch(ctx, imgs, pcs, P, prm, pinned=True)

    def barrier():
        rep.barrier()
        torch.cuda.synchronize()

    max_over_ranks = rep.max_over_ranks

    # ---- device-resident throughput ("value") ----
    batch.upload()
    pose0 = seq.pose(0)

    def device_steps(k):
        """k steps, software-pipelined the way the reference's tracking thread is: the frame construction of batch i
        (rgbl_resident_process) is issued while the tracking chain of batch i-1 still runs on the tracking stream; every
        batch is fully processed and its poses are read back; the last chain is drained inside the timed region."""
        n = batch.process_resident()
        batch.track_begin(pose0, *CAM, th=15.0)
        for _ in range(k - 1):
            n = batch.process_resident()                 # frame construction of the next batch while the chain runs
            batch.track_begin(pose0, *CAM, th=15.0)      # queued behind the running chain: no host gap between two chains
            batch.track_end()                            # poses of the oldest chain
        poses, nm, ni = batch.track_end()
        return n, poses, nm, ni

    device_steps(args.warmup)
    ctx.profile_enable(True); ctx.profile_reset()
    sampler = ClockSampler(local_rank); sampler.start()
    barrier()
    ctx.timer_mark(0)
    t0 = time.perf_counter()
    n_kp, poses, nm, ni = device_steps(args.steps)
    n_kp = n_kp.copy()
    ctx.timer_mark(1)
    dev_ms = ctx.timer_elapsed_ms()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop()
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    dev_ms = max_over_ranks(dev_ms)
    fps = world * T * args.steps / (dev_ms * 1e-3)

    # ---- end to end through the C ABI with host buffers ("e2e") ----
    def e2e_steps(k):
        batch.run_e2e()                                  # H2D inputs, kernels, D2H keypoints/descriptors/depths
        batch.track_begin(pose0, *CAM, th=15.0)
        for _ in range(k - 1):
            batch.run_e2e()
            batch.track_begin(pose0, *CAM, th=15.0)
            batch.track_end()                            # D2H poses + counts of the oldest chain
        return batch.track_end()

    e2e_steps(2)
    barrier()
    t0 = time.perf_counter()
    e2e_steps(args.steps)
    barrier()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
    e2e_fps = world * T * args.steps / (e2e_ms * 1e-3)
    d2h = batch.d2h_bytes()

    # ---- capacity: several independent sequences on ONE GPU (extra information, not the headline: configs[1] is one
    # sequence per GPU).  A single sequence is bound by the latency of its serial tracking chain (one SM busy); chains of
    # different sequences run side by side on their own streams until the frame construction saturates the device.
    multi = None
    if world == 1 and args.multi_sequences > 1:
        nS = args.multi_sequences
        ctxs, batches, pose0s = [ctx], [batch], [pose0]
        for sidx in range(1, nS):
            im2, pc2, seq2 = make_batch_inputs(100 + sidx, T)
            c2 = F.Context(S.KITTI_W, S.KITTI_H, 2000, max_batch=T, max_points=max(p.shape[1] for p in pc2), device=local_rank)
            b2 = F.RgblBatch(c2, im2, pc2, seq2.P, prm, pinned=False)
            b2.upload()
            ctxs.append(c2); batches.append(b2); pose0s.append(seq2.pose(0))
        batch.upload()

        def multi_rounds(k):
            pend = [False] * nS
            for _ in range(k):
                for i in range(nS):
                    batches[i].process_resident()
                    if pend[i]:
                        batches[i].track_end()
                    batches[i].track_begin(pose0s[i], *CAM, th=15.0); pend[i] = True
            return [batches[i].track_end() for i in range(nS)]

        multi_rounds(2)
        torch.cuda.synchronize()
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record()
        t0 = time.perf_counter()
        res = multi_rounds(args.steps)
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0)
        multi = {"sequences_per_gpu": nS, "value": nS * T * args.steps / (ms * 1e-3), "unit": UNIT, "ms_per_round": ms / args.steps,
                 "timing": "host wall clock around K rounds with device synchronisation on both sides (several library streams)",
                 "inliers_per_frame": [float(np.mean(r[2][1:])) for r in res]}
        for c2 in ctxs[1:]:
            c2.close()

    # ---- Frame::ComputeBoW (keyframes only in the reference, reported on its own, not part of `value`) ----
    bow = None
    if world == 1 and not args.no_bow:
        k_, L_ = 10, 6                                     # the shape of ORBvoc.txt (k = 10, L = 6: 1 111 111 nodes)
        nn = (k_ ** (L_ + 1) - 1) // (k_ - 1)
        rng = np.random.default_rng(7)
        inner = (k_ ** L_ - 1) // (k_ - 1)
        cb = np.minimum(np.arange(nn + 1, dtype=np.int64) * k_, inner * k_).astype(np.int32)
        ci = np.arange(1, nn, dtype=np.int32)              # BFS numbering: children of i are k i + 1 .. k i + k
        nd = rng.integers(0, 256, (nn, 32), dtype=np.uint8)
        wid = np.full(nn, -1, np.int32); wid[inner:] = np.arange(nn - inner, dtype=np.int32)
        nw = np.zeros(nn, np.float64); nw[inner:] = rng.uniform(0.5, 12.0, nn - inner)
        voc = F.ORBVocabulary(ctx, cb, ci, nd, nw, wid, L_)
        batch.upload(); batch.process_resident()
        for f in range(3):
            voc.transform_resident(f)
        t0 = time.perf_counter()
        res = [voc.transform_resident(f) for f in range(T)]
        gpu_ms = 1e3 * (time.perf_counter() - t0) / T
        voc.close()
        bow = {"gpu_ms_per_frame": gpu_ms, "words_per_frame": float(np.mean([len(r[0][0]) for r in res])),
               "vocabulary": f"synthetic k={k_} L={L_} ({nn} nodes, {nn * 32 / 1e6:.0f} MB of node descriptors), TF-IDF / L1, levelsup 4",
               "timing": "host wall clock per rgbl_resident_compute_bow call (descriptors already in HBM; includes the D2H of both maps)"}
        if not args.no_cpu_baseline:
            from oracle import compute_bow as cpu_bow
            vd = dict(child_begin=cb, child_index=ci, node_desc=nd, node_weight=nw, word_id=wid, levels=L_)
            descs = [o[1] for o in batch.download()[:4]]
            t0 = time.perf_counter()
            for dsc in descs:
                cpu_bow(vd, dsc)
            bow["cpu_ms_per_frame"] = 1e3 * (time.perf_counter() - t0) / len(descs)
            bow["cpu"] = "oracle (std::map restatement of DBoW2 transform), one core"

    # ---- Optimizer::LocalBundleAdjustment (mapping thread, reported on its own) ----
    lba = None
    if world == 1 and not args.no_bow:
        sys.path.insert(0, str(ROOT / "tests"))
        import ba_data
        prob = ba_data.make_problem(42, n_kf=20, n_fixed=4, n_points=2500, outlier_frac=0.02)
        F.local_bundle_adjustment(ctx, *ba_data.args(prob))
        t0 = time.perf_counter()
        for _ in range(3):
            gpo, gpt, ger, git = F.local_bundle_adjustment(ctx, *ba_data.args(prob))
        gpu_ms = 1e3 * (time.perf_counter() - t0) / 3
        lba = {"gpu_ms": gpu_ms, "key_frames": 20, "fixed": 4, "points": 2500, "edges": int(len(prob["e_point"])), "lm_iterations": int(git),
               "timing": "host wall clock per rgbl_local_bundle_adjustment call with host arrays (uploads, ~12 launches + one scalar read-back per LM trial, downloads)"}
        if not args.no_cpu_baseline:
            from oracle import local_bundle_adjustment as cpu_lba
            t0 = time.perf_counter()
            rpo = cpu_lba(*ba_data.args(prob))[0]
            lba["cpu_ms"] = 1e3 * (time.perf_counter() - t0)
            lba["cpu"] = "oracle (dense Schur restatement of the g2o problem), one core"
            lba["max_pose_diff_vs_cpu"] = float(np.abs(gpo - rpo).max())

    # ---- online use: ONE new frame at a time through the individual C-ABI calls a tracking thread makes (host buffers in and
    # out of every call): Frame construction -> SearchByProjection(last frame) -> PoseOptimization.  Latency, not throughput.
    online = None
    if world == 1 and not args.no_bow:
        def unproject_identity_rotation(fr, pose_):        # Frame::UnprojectStereo (src/Frame.cc:1137-1150); the synthetic camera does not rotate
            f32 = np.float32
            z = fr["dep"]; zz = np.where(z > 0, z, f32(1)).astype(f32)
            x = ((fr["k"]["x"] - f32(S.KITTI_CX)) * zz * f32(1.0 / np.float32(S.KITTI_FX))).astype(f32)
            y = ((fr["k"]["y"] - f32(S.KITTI_CY)) * zz * f32(1.0 / np.float32(S.KITTI_FY))).astype(f32)
            return (np.stack([x, y, zz], 1) - np.asarray(pose_[4:7], f32)).astype(f32), (z > 0)

        c1 = F.Context(S.KITTI_W, S.KITTI_H, 2000, max_batch=1, max_points=max_pts, device=local_rank)
        matcher = F.ORBmatcher(c1, 0.9, True)
        sfac = F.orb_tables(2000)["scale"][:8].astype(np.float32)
        t_stage = np.zeros(3); n_on = 0; pose = seq.pose(0); last = None
        for t in range(min(T, 14)):
            t0 = time.perf_counter()
            (k, d, dep, ur), = F.frame_rgbl_batch(c1, [imgs[t]], [pcs[t]], P, prm)
            t1 = time.perf_counter()
            cur = dict(k=k, d=d, dep=dep, ur=ur)
            if last is not None:
                xw, ok = unproject_identity_rotation(last, pose)
                gfv = F.FrameView(k, ur, d, S.KITTI_W, S.KITTI_H, sfac, *CAM)
                _, match = matcher.SearchByProjectionLastFrame(gfv, pose, pose, ok.astype(np.uint8), xw, last["d"], last["k"]["octave"],
                                                               last["k"]["angle"], np.ones(len(ok), np.uint8), 15.0)
                t2 = time.perf_counter()
                m = np.nonzero(match >= 0)[0]
                obs = np.stack([k["x"][m], k["y"][m], ur[m]], 1)
                inv_s2 = (1.0 / sfac[k["octave"][m]] ** 2).astype(np.float32)
                _, pose, _ = F.Optimizer.PoseOptimization(c1, pose, xw[match[m]], obs, inv_s2, (ur[m] >= 0).astype(np.uint8), *CAM)
                t3 = time.perf_counter()
                if t >= 3:
                    t_stage += [t1 - t0, t2 - t1, t3 - t2]; n_on += 1
            last = cur
        c1.close()
        if n_on:
            ms = 1e3 * t_stage / n_on
            online = {"per_frame_ms": float(ms.sum()), "frame_construction_ms": float(ms[0]), "search_by_projection_ms": float(ms[1]),
                      "pose_optimization_ms": float(ms[2]), "frames": n_on,
                      "note": "one frame per call, host arrays in and out of each C-ABI call, Python glue (unprojection, edge assembly in numpy) included"}

    if rank == 0:
        # roofline of the dominant kernel (per-stage CUDA-event time / launches, measured above)
        levels = []
        from orb_slam3_rgbl_b200 import _lib
        t = F.orb_tables(2000)
        for l in range(8):
            levels.append((int(np.rint(np.float32(S.KITTI_W) * t["inv_scale"][l])), int(np.rint(np.float32(S.KITTI_H) * t["inv_scale"][l]))))
        # candidates: measure from the last processed batch
        n_cand = 0
        try:
            ex = F.ORBextractor(2000, 1.2, 8, 12, 7, S.KITTI_W, S.KITTI_H, ctx=ctx)
            n_cand = int(np.mean([sum(len(ex.level_candidates(l, f)) for l in range(8)) for f in range(min(T, 4))]))
        except Exception:
            pass
        n_in = int(0.13 * max_pts)
        per_frame = algorithmic_bytes(levels, n_cand, float(np.mean(n_kp)), max_pts, S.KITTI_W, S.KITTI_H, n_in)
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            peak = float(json.loads(peaks_path.read_text())["hbm_gbs"]); peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)"
        else:
            peak = 6650.0; peak_src = "fallback 6.65 TB/s (B200_PROFILING.md)"
        kernels = {}
        other = {}
        for name, st in prof.items():
            if name.startswith("_") or st["calls"] == 0:
                continue
            if name not in per_frame:
                other[name] = {"ms_per_step": st["ms"] / args.steps, "launches_per_step": st["launches"] / args.steps}
                continue
            ms_per_call = st["ms"] / st["calls"]
            gbs = per_frame[name] * T / (ms_per_call * 1e-3) / 1e9
            kernels[name] = {"ms_per_step": ms_per_call, "launches_per_step": st["launches"] / st["calls"],
                             "algorithmic_MB_per_step": per_frame[name] * T / 1e6, "achieved_GBs": gbs, "frac": gbs / peak}
        def roof(name, kd, note=None):
            r = {"bound": "hbm", "kernel": name, "achieved": kd["achieved_GBs"], "peak": peak, "unit": "GB/s",
                 "frac": kd["frac"], "traffic": TRAFFIC.get(name), "peak_source": peak_src,
                 "avg_launch_ms": kd["ms_per_step"] / max(kd["launches_per_step"], 1),
                 "algorithmic_bytes_per_launch": kd["algorithmic_MB_per_step"] * 1e6 / max(kd["launches_per_step"], 1)}
            if note:
                r["note"] = note
            return r

        # tracking chain (grid + collect + resolve + edges + PoseOptimization per frame): 64 B per evaluated descriptor pair
        # and 64 B per edge per LM evaluation; latency-bound by construction (serial in time, one CTA)
        if "match" in other:
            pairs = 30000.0; evals = 60.0
            bytes_step = (T - 1) * (64.0 * pairs + 64.0 * float(np.mean(nm[1:])) * evals)
            ms = other["match"]["ms_per_step"]
            kernels["tracking_chain"] = {"ms_per_step": ms, "launches_per_step": other["match"]["launches_per_step"],
                                         "algorithmic_MB_per_step": bytes_step / 1e6, "achieved_GBs": bytes_step / (ms * 1e-3) / 1e9,
                                         "frac": bytes_step / (ms * 1e-3) / 1e9 / peak}
        dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"]) if kernels else None
        roofline = None
        if dom:
            roofline = roof(dom, kernels[dom], "latency-bound serial chain (pose_optimize_kernel ~80 % of it): one persistent CTA per frame, FP64 LM"
                            if dom == "tracking_chain" else None)
        stream = {k: v for k, v in kernels.items() if k != "tracking_chain"}
        dom_s = max(stream, key=lambda k: stream[k]["ms_per_step"]) if stream else None
        roofline_streaming = roof(dom_s, stream[dom_s], "dominant frame-construction kernel; FAST is ALU/issue-bound (see DESIGN.md 4)") if dom_s else None
        working_set_mb = (2 * 1.74 + 4 * 4 * max_pts / 1e6 + 2 * 4 * S.KITTI_W * S.KITTI_H / 1e6) * T
        line = {"metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic",
                "config": {"workload": WORKLOAD,
                           "frames_per_step_per_gpu": T, "parallelism": f"sequences sharded x{world}",
                           "l2": f"inputs larger than L2: ~{working_set_mb:.0f} MB touched per step vs 126 MB L2",
                           "timing": "CUDA events on the library stream around K steps, max over ranks",
                           "pipeline": "frame construction of batch i+1 overlaps the tracking chain of batch i (two streams); the last chain drains inside the timed region; per-stage times below are measured under that overlap"},
                "e2e": {"value": e2e_fps, "unit": UNIT, "h2d_bytes_per_step": batch.h2d_bytes, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": int(prof["_total_launches"]),
                "clocks": clocks, "roofline": roofline, "roofline_frame_construction": roofline_streaming, "kernels": kernels, "latency_bound_stages": other,
                "tracking": {"matches_per_frame": float(np.mean(nm[1:])), "inliers_per_frame": float(np.mean(ni[1:])),
                             "pose_x_error_m_last_frame": float(abs(poses[-1, 4] - seq.pose(T - 1)[4]))},
                "multi_sequence_capacity": multi, "compute_bow": bow, "local_bundle_adjustment": lba, "online_single_frame": online,
                "host_quadtree_ms_per_step": prof["_host_quadtree_ms"] / args.steps,
                "wall_ms_per_step": wall_ms / args.steps}
        if world == 1 and not args.no_cpu_baseline:
            v, n = cpu_baseline_single(imgs, pcs, seq)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{n} frames of the same workload, oracle (C++ restatement of ORBextractor + DepthModule + SearchByProjection + PoseOptimization, scalar, -O3) on one core"}
        print(json.dumps(line))
    ctx.close()
    rep.close()


if __name__ == "__main__":
    main()
