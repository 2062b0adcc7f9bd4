#!/usr/bin/env python
"""Benchmark of the ICP scan-matching hot path (BASELINE.json metric: ICP pairs/sec, 64-beam point-to-plane).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

N = 1  BASELINE configs[1]: one batch of `--pairs` consecutive keyframe pairs (`--pairs` + 1 synthetic OS1-64 scans) per
       step: preprocessing of every scan (radius/height filter, Morton sort, hash grid, k-NN normals) + point-to-plane
       ICP of every pair to Open3D's default convergence criteria.  Extra lines in the same JSON object: configs[2]
       (128-beam point-to-point), voxel 0.2, the reference's own one-pair-per-call pattern through the drop-in
       KeyFrameManager, small loop-closure batches, configs[3] on one GPU (the strong-scaling reference), map building.
N > 1  BASELINE configs[3]: ONE seeded global list of `--lc-pairs` loop-closure pairs (64-beam, initial guesses perturbed
       by N(0, 0.2 m) / N(0, 2 deg)) sorted for scan-cache reuse and sharded over the ranks in contiguous blocks; every
       rank uploads / preprocesses only the scans its pairs touch; the 160-byte result records are all-gathered from
       device memory over NCCL and delivered to rank 0's host.  Strong scaling: the list does not grow with N.
  value : pairs/s with the raw scans already resident in HBM when the timed region starts (device time, CUDA events on
          the engine's stream, max over ranks)
  e2e   : pairs/s through the C-ABI from pinned HOST buffers: H2D of every scan + preprocessing + ICP + D2H of the
          result records, every step (host wall clock between synchronisations, max over ranks)
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "icp_pairs_per_sec"
UNIT = "pairs/s"
TOL_T, TOL_REL = 1e-4, 1e-5          # north_star: transforms within 1e-4 m / rad, fitness / rmse within 1e-5 relative


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=99, help="N=1: keyframe pairs per step (BASELINE.md config 2: the 99 pairs of a 100-scan sequence)")
    ap.add_argument("--ref-pairs", type=int, default=6, help="pairs per step of the CPU reference arm (bounded sample)")
    ap.add_argument("--cpu-pairs", type=int, default=6, help="pairs of the cpu_baseline sample (N=1 only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="N=1: headline only")
    ap.add_argument("--lc-pairs", type=int, default=10000, help="N>1: loop-closure pairs of the global list (BASELINE config 4)")
    ap.add_argument("--lc-scans", type=int, default=464, help="keyframes of the loop-closure trajectory (two laps of the synthetic corridor)")
    ap.add_argument("--lc-batch", type=int, default=2048, help="N>1: pairs per device batch (the gather of batch k overlaps batch k+1)")
    ap.add_argument("--lc-pairs-n1", type=int, default=0, help="N=1 extra: leading pairs of the sorted global list timed on one GPU (0 = the whole list)")
    ap.add_argument("--c5-scans", type=int, default=5000, help="N=1 extra: keyframes of the configs[4] pipeline run (0 = skip)")
    ap.add_argument("--lc-balance", default="history", choices=["history", "count"],
                    help="N>1: contiguous shard bounds by the previous batch's per-pair passes (equal work) or by pair count")
    ap.add_argument("--parity-pairs", type=int, default=4, help="N>1: pairs per rank checked against the oracle after the timed regions")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        if os.environ.get("ARVC_BENCH_NO_SMI"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("ARVC_BENCH_SMI_MS", "250")], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        """Summary of the samples that arrived inside [t0, t1] (falls back to the nearest ones for a short window)."""
        inside = [s for (t, s) in self.samples if t0 <= t <= t1]
        if not inside and self.samples:
            mid = 0.5 * (t0 + t1)
            inside = [min(self.samples, key=lambda ts: abs(ts[0] - mid))[1]]
        return self._summarise(inside)

    def stop(self):
        if not self.proc:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

    def _summarise(self, lines):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in lines:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peak():
    peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError, TypeError):
        pass
    return peak, src


def load_traffic():
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


# ---------------------------------------------------------------------------------------------- workloads
def config2(pairs):
    return {"workload": "configs[1]: batched consecutive-keyframe point-to-plane ICP, 64-beam (OS1-64-like 64x1024) synthetic scans",
            "pairs_per_step": pairs, "scans_per_step": pairs + 1, "method": "icppointplane", "voxel_size": None,
            "max_corr_dist": 10.0, "criteria": "rel_fitness=1e-6 rel_rmse=1e-6 max_iter=30", "normals": "radius=0.3 max_nn=300",
            "l2": "inputs larger than L2 (every step re-streams %d MB of scans, grids and normals)" % (12 * (pairs + 1)),
            "parallelism": "1 GPU"}


def config4(n_pairs, n_scans, world, batch):
    return {"workload": "configs[3]: loop-closing candidate ICP batch, %d pairs of 64-beam keyframes <= 5 m apart (seed 777, initial guess = "
                        "ground truth perturbed by N(0, 0.2 m) / N(0, 2 deg)), point-to-plane" % n_pairs,
            "pairs_per_step": n_pairs, "trajectory_keyframes": n_scans, "method": "icppointplane", "voxel_size": None,
            "max_corr_dist": 10.0, "criteria": "rel_fitness=1e-6 rel_rmse=1e-6 max_iter=30", "normals": "radius=0.3 max_nn=300",
            "l2": "inputs larger than L2 (every rank re-streams the scans, grids and normals of its shard every step)",
            "parallelism": "one global pair list sorted by (target, source), contiguous shards x%d, device batches of <= %d pairs, "
                           "NCCL all-gather of 160 B records from device memory" % (world, batch)}


class LoopClosureWorkload:
    """BASELINE config 4: global list of loop-closure pairs over a two-lap trajectory, sorted for scan-cache reuse.
    A rank materialises (ray-casts) only the scans of its own shard; scan k always has seed 10000 + k."""

    def __init__(self, n_scans, n_pairs, workers):
        from lidar_slam_arvc_b200 import sharding, synth
        self.synth, self.workers = synth, workers
        self.world_model = synth.World(1234)
        self.poses = synth.loop_trajectory(self.world_model, n_scans, step=0.5, start=0.0)
        pairs = synth.loop_closure_pairs(self.poses, n_pairs, radius=5.0, min_gap=20, seed=777, sigma_t=0.2, sigma_rot_deg=2.0)
        tg = np.array([p[0] for p in pairs], dtype=np.int64)
        sr = np.array([p[1] for p in pairs], dtype=np.int64)
        order = sharding.sort_pairs_for_cache(tg, sr)
        self.tg, self.sr = tg[order], sr[order]
        self.init = np.array([pairs[k][2] for k in order])
        self.scans = {}

    def materialise(self, scan_ids):
        todo = [int(k) for k in scan_ids if int(k) not in self.scans]
        jobs = [(self.world_model, self.synth.OS1_64, self.poses[k], 10000 + k) for k in todo]
        if self.workers > 1 and len(jobs) > 2:
            import multiprocessing as mp
            with mp.get_context("fork").Pool(min(self.workers, len(jobs)), initializer=self.synth._pool_init) as pool:
                out = pool.map(self.synth._scan_job, jobs, chunksize=max(1, len(jobs) // (4 * self.workers)))
        else:
            out = [self.synth._scan_job(j) for j in jobs]
        for k, s in zip(todo, out):
            self.scans[k] = s


# ---------------------------------------------------------------------------------------------- CPU arms
def oracle_consecutive(seq, n_pairs):
    """The oracle (C++/OpenMP float64 restatement of the reference's Open3D CPU path) on all host threads:
    preprocessing of every scan once + ICP of consecutive pairs, like run_scanmatcher.py:191-213.  Returns the rate and
    the results (kept: the parity gate of the same run compares every one of them with the GPU's)."""
    from oracle import oracle as orc
    orc.set_num_threads(os.cpu_count() or 1)    # torchrun exports OMP_NUM_THREADS=1: use every host thread anyway
    pre = [orc.preprocess(seq.scans[0])]        # steady state of consecutive matching: one new scan per pair
    results = []
    t0 = time.perf_counter()
    for k in range(n_pairs):
        pre.append(orc.preprocess(seq.scans[k + 1]))
        tgt, ntgt = pre[k]
        src, _ = pre[k + 1]
        results.append(orc.icp(src, tgt, ntgt, seq.relative_odo(k, k + 1), orc.P2PLANE))
    dt = time.perf_counter() - t0
    return n_pairs / dt, dt, results


def oracle_loop_closure(wl, idx, threads):
    """The oracle on loop-closure pairs `idx` of the sorted global list: preprocessing of every scan the sample touches
    (once) + ICP of every pair, like loopclosing.py:154-184 with a keyframe cache."""
    from oracle import oracle as orc
    orc.set_num_threads(threads)
    pre, results = {}, []
    t0 = time.perf_counter()
    for k in idx:
        for sid in (int(wl.tg[k]), int(wl.sr[k])):
            if sid not in pre:
                pre[sid] = orc.preprocess(wl.scans[sid])
        tgt, ntgt = pre[int(wl.tg[k])]
        src, _ = pre[int(wl.sr[k])]
        results.append(orc.icp(src, tgt, ntgt, wl.init[k], orc.P2PLANE))
    dt = time.perf_counter() - t0
    return len(idx) / dt, dt, results, len(pre)


def parity_report(records, refs, label):
    """GPU records against oracle results of the same pairs: the north_star tolerances, asserted (the run fails)."""
    rep = {"pairs": len(refs), "max_abs_dT": 0.0, "max_rel_fitness": 0.0, "max_rel_rmse": 0.0, "updates_equal": True, "what": label}
    for rec, ref in zip(records, refs):
        rep["max_abs_dT"] = max(rep["max_abs_dT"], float(np.abs(np.asarray(rec["T"]) - ref.transformation).max()))
        rep["max_rel_fitness"] = max(rep["max_rel_fitness"], float(abs(rec["fitness"] - ref.fitness) / max(abs(ref.fitness), 1e-300)))
        rep["max_rel_rmse"] = max(rep["max_rel_rmse"], float(abs(rec["rmse"] - ref.inlier_rmse) / max(abs(ref.inlier_rmse), 1e-300)))
        rep["updates_equal"] = rep["updates_equal"] and bool(rec["updates"] == ref.updates)
    rep["ok"] = bool(rep["max_abs_dT"] < TOL_T and rep["max_rel_fitness"] <= TOL_REL and rep["max_rel_rmse"] <= TOL_REL and rep["updates_equal"])
    return rep


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port: Open3D is not installable here),
    all host threads, a bounded sample of the SAME workload per step (rate = steady state, like our arm).  Rank 0 only."""
    if rank != 0:
        return
    from lidar_slam_arvc_b200 import synth
    from oracle import oracle as orc
    sp = max(1, args.ref_pairs)
    if args.gpus <= 1:
        seq = synth.Sequence(sp + 1, synth.OS1_64, start=30.0)
        step = lambda: oracle_consecutive(seq, sp)                                           # noqa: E731
        cfg = config2(args.pairs)
        sample = "%d consecutive 64-beam pairs per step (%d new scans preprocessed + %d ICPs), oracle C++/OpenMP" % (sp, sp, sp)
        scaling = "weak"
    else:
        wl = LoopClosureWorkload(args.lc_scans, args.lc_pairs, max(1, (os.cpu_count() or 1)))
        idx = list(range(sp))
        wl.materialise(np.concatenate([wl.tg[idx], wl.sr[idx]]))
        step = lambda: oracle_loop_closure(wl, idx, os.cpu_count() or 1)                      # noqa: E731
        cfg = config4(args.lc_pairs, args.lc_scans, args.gpus, args.lc_batch)
        sample = "first %d pairs of the sorted global list per step (their scans preprocessed once + %d ICPs), oracle C++/OpenMP" % (sp, sp)
        scaling = "strong"
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.steps * sp / dt
    cores = orc.num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample + " on %d threads" % cores},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "note": "a step of this arm is a bounded sample (%d pairs) of the workload our arm runs in full per step; both values are "
                    "steady-state pairs/s of the same per-pair work" % sp}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- helpers of our arm
def split_profile(prof_raw):
    """{name: (launches, ms)} with the per-pass search kernels merged into 'icp_pass' (+ their per-pass averages)."""
    prof, passes = {}, {}
    for k, v in prof_raw.items():
        if k.startswith("icp_pass_"):
            passes[k[-2:]] = round(v[1] / v[0], 4)
            c, t = prof.get("icp_pass", (0, 0.0))
            prof["icp_pass"] = (c + v[0], t + v[1])
        elif k.startswith("icp_far_"):
            c, t = prof.get("icp_far", (0, 0.0))
            prof["icp_far"] = (c + v[0], t + v[1])
        else:
            prof[k] = v
    return prof, passes


def roofline_dict(prof, passes, n_pts_scans, pass_points, steps, region_ms, method_p2plane=True):
    """SURVEY.md §8(d) byte model for the dominant kernel of a profiled region (per-launch CUDA events on the engine's
    stream).  n_pts_scans: points of every preprocessed scan; pass_points: sum over pairs of executed passes x source points."""
    peak, peak_src = load_peak()
    if not prof:
        return None
    name, (n_launch, tot_ms) = max(prof.items(), key=lambda kv: kv[1][1])
    if name == "icp_pass":        # search kernel: the source point (16 B) + the matched target record (16 B) per point and pass
        alg = 32.0 * pass_points * steps
        model = "32 B x source points x executed passes (source record + matched target record)"
    elif name.startswith("normals"):
        alg = 32.0 * float(np.sum(n_pts_scans)) * steps
        model = "32 B x points (16 B record read + 16 B normal written, SURVEY 8d)"
    elif name == "icp_accum":
        alg = (48.0 if method_p2plane else 32.0) * pass_points * steps
        model = "48 B x source points x executed passes (source + target record + target normal)"
    else:
        alg = 36.0 * float(np.sum(n_pts_scans)) * steps
        model = "36 B x points (grid build, SURVEY 8d)"
    achieved = alg / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0
    tj = load_traffic()
    key = "normals" if name.startswith("normals") else name
    traffic, util = None, None
    if key in tj and n_launch:
        util = {k: v for k, v in tj[key].items() if k.endswith("_pct")}
        if key == "normals" and "dram_bytes_per_scan" in tj[key]:
            traffic = tj[key]["dram_bytes_per_scan"] * len(n_pts_scans)
        elif key == "icp_pass" and "dram_bytes_per_pair_pass" in tj[key]:
            # captured at ~63k points per scan: scale by the points this launch searched
            traffic = tj[key]["dram_bytes_per_pair_pass"] * (pass_points / 63000.0) * steps / n_launch
    return {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": peak_src, "byte_model": model, "launches": n_launch, "kernel_ms_total": tot_ms,
            "kernel_share_of_region": tot_ms / region_ms if region_ms > 0 else None,
            "algorithmic_bytes_per_launch": alg / max(n_launch, 1), "ncu_utilisation_pct": util,
            "all_kernels_ms": {k: round(v[1], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
            "icp_pass_avg_ms_by_pass": passes,
            "note": "the per-pair working set is L2-resident and the kernels are issue / latency bound: the HBM fraction is low by construction (DESIGN.md)"}


def pin_scans(torch, scans):
    out = []
    for s in scans:
        t = torch.empty((len(s), 3), dtype=torch.float32).pin_memory()
        t.copy_(torch.from_numpy(np.ascontiguousarray(s, dtype=np.float32)))
        out.append(t)
    return out


def growing_chunks(n, head=6):
    """Chunk bounds 6, 12, 24, ... : the upload of chunk c + 1 (copy stream) overlaps the preprocessing of chunk c."""
    bounds = [0]
    while bounds[-1] < n:
        bounds.append(min(n, bounds[-1] + head * 2 ** (len(bounds) - 1)))
    if len(bounds) > 2 and bounds[-1] - bounds[-2] < head:
        bounds.pop(-2)
    return bounds


# ---------------------------------------------------------------------------------------------- N = 1
def run_single(args, torch, engine, synth, local_rank):
    dev = torch.device("cuda", local_rank)
    P = args.pairs
    seq = synth.Sequence(P + 1, synth.OS1_64, start=30.0, workers=max(1, os.cpu_count() or 1))
    ids = np.arange(P + 1, dtype=np.int64)
    tg, sr = ids[:-1], ids[1:]
    init = np.array([seq.relative_odo(int(a), int(b)) for a, b in zip(tg, sr)])
    pinned = pin_scans(torch, seq.scans)
    h2d_bytes = int(sum(t.numel() * 4 for t in pinned) + init.nbytes + tg.nbytes + sr.nbytes)

    eng = engine.Engine(local_rank)
    pp = eng.make_preprocess_params()
    ip = eng.make_icp_params(engine.P2PLANE)
    stream = torch.cuda.ExternalStream(eng.stream_handle(), device=dev)

    def upload_all():
        for k, t in enumerate(pinned):
            eng.upload_ptr(k, t.data_ptr(), t.shape[0])

    def hot_path():
        eng.invalidate(ids)
        eng.preprocess(ids, pp)
        return eng.icp_batch(tg, sr, init, ip)

    bounds = growing_chunks(len(ids))

    def e2e_path():
        """Host scans -> records: every chunk is uploaded (copy stream) and then preprocessed (compute stream, waits for
        its own uploads only); the ICP batch is the same single call."""
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            for k in range(lo, hi):
                eng.upload_ptr(k, pinned[k].data_ptr(), pinned[k].shape[0])
            eng.preprocess(ids[lo:hi], pp)
        return eng.icp_batch(tg, sr, init, ip)

    def barrier():
        eng.sync()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)      # started well before the timed regions: nvidia-smi start-up stalls the driver
    clocks.start()
    for _ in range(max(args.warmup, 3)):
        upload_all()
        hot_path()
        e2e_path()
    eng.sync()
    n_pts = np.array([eng.info(int(k))["n_points"] for k in ids])

    # ---- value: scans resident in HBM, device time; the product path (device-terminated ICP loop), no per-kernel events
    upload_all()
    barrier()
    l0 = eng.kernel_launches()
    tw0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    step_ms = []
    for _ in range(args.steps):
        ts0 = time.perf_counter()
        own = hot_path()
        step_ms.append(round((time.perf_counter() - ts0) * 1e3, 2))
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = eng.kernel_launches() - l0
    value = P * args.steps / (dev_ms * 1e-3)

    # ---- the same steps again with one CUDA-event pair around every kernel launch (the per-kernel times of `roofline`);
    # the ICP loop is enqueued pass by pass here, because the kernels inside a graph cannot be bracketed by events
    eng.set_option("icp_loop_graph", 0)
    hot_path()
    barrier()
    eng.profile_enable(True)
    pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pv0.record(stream)
    for _ in range(args.steps):
        hot_path()
    pv1.record(stream)
    barrier()
    prof_ms = pv0.elapsed_time(pv1)
    prof, icp_passes = split_profile(eng.profile_report())
    eng.profile_enable(False)
    eng.set_option("icp_loop_graph", 1)
    pass_points = float(sum(int(own["passes"][k]) * int(n_pts[k + 1]) for k in range(P)))
    roofline = roofline_dict(prof, icp_passes, n_pts, pass_points, args.steps, prof_ms)
    roofline["profiled_region"] = {"ms_per_step": prof_ms / args.steps, "note": "same K steps, per-launch CUDA events, ICP passes enqueued one by one; "
                                   "the `value` region above runs without them (one graph launch per batch)"}

    # ---- e2e: host buffers -> result records on the host, every step
    e2e_path()
    barrier()
    t0 = time.perf_counter()
    e2e_step_ms = []
    for _ in range(args.steps):
        ts0 = time.perf_counter()
        e2e_path()
        e2e_step_ms.append(round((time.perf_counter() - ts0) * 1e3, 2))
    barrier()
    e2e_s = time.perf_counter() - t0
    time.sleep(0.25)
    clk = clocks.window(tw0, t0 + e2e_s)
    e2e_value = P * args.steps / e2e_s
    d2h_bytes = int(164 * P)      # 160-byte record per pair + status word

    counters = [eng.get_counters(int(k)) for k in ids[:8]]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config2(P), "ms_per_pair": dev_ms / (args.steps * P),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": int(launches), "clocks": clk, "step_ms": step_ms, "e2e_step_ms": e2e_step_ms, "roofline": roofline,
            "mean_icp_updates": float(np.mean(own["updates"])), "icp_updates": [int(u) for u in own["updates"]], "points_per_scan": int(n_pts.mean()),
            "normals_paths": {"per_point_kernel_share": float(np.mean([c["normals_per_point"] / max(c["n_points"], 1) for c in counters])),
                              "canonical_resummation_share": float(np.mean([c["normals_redone"] / max(c["n_points"], 1) for c in counters]))},
            "notes": {"value_region": "one CUDA graph per batch: 4 unrolled passes + a WHILE node that ends the iteration on the device with the "
                                      "last convergence; gpu_launches counts the kernels that actually ran",
                      "e2e_region": "uploads on the engine's copy stream in chunks of %s scans: chunk c is preprocessed while chunk c+1 is in flight" % [b - a for a, b in zip(bounds[:-1], bounds[1:])],
                      "reference_arm": "bench.py --impl reference times --ref-pairs pairs per step (a bounded sample); both arms report steady-state pairs/s"}}

    if not args.no_extras:
        eng.invalidate(ids)
        extras_single(args, torch, engine, synth, eng, seq, ids, tg, sr, init, pp, ip, stream, line, dev)
        clocks.stop()
    else:
        clocks.stop()

    if not args.no_cpu_baseline:
        from oracle import oracle as orc
        cp = max(1, min(args.cpu_pairs, P))
        v, dt, refs = oracle_consecutive(seq, cp)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                                "sample": "first %d pairs of the same batch (%d new scans preprocessed + %d ICPs) in %.1f s, oracle C++/OpenMP"
                                          % (cp, cp, cp, dt)}
        line["parity_check"] = parity_report([own[k] for k in range(cp)], refs, "every pair of the cpu_baseline sample against the GPU result of the timed region")
    print(json.dumps(line), flush=True)
    eng.close()
    if "parity_check" in line and not line["parity_check"]["ok"]:
        raise SystemExit("parity check failed: %s" % line["parity_check"])


def timed_steps(torch, stream, eng, fn, steps, warm=1):
    for _ in range(warm):
        out = fn()
    eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        out = fn()
    e1.record(stream)
    eng.sync()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def extras_single(args, torch, engine, synth, eng, seq, ids, tg, sr, init, pp, ip, stream, line, dev):
    P = len(tg)
    nv = max(2, args.steps // 3)
    # ---- BASELINE.md config 2 is reported for voxel_size None and 0.2: same batch with voxel down-sampling on
    ppv = eng.make_preprocess_params(voxel_size=0.2)

    def voxel_step():
        eng.invalidate(ids)
        eng.preprocess(ids, ppv)
        return eng.icp_batch(tg, sr, init, ip)
    ms, rv = timed_steps(torch, stream, eng, voxel_step, nv, warm=2)
    line["voxel_0p2"] = {"value": P / (ms * 1e-3), "unit": UNIT, "steps": nv, "ms_per_step": ms,
                         "points_per_scan": int(np.mean([eng.info(int(k))["n_points"] for k in ids[:8]])),
                         "mean_icp_updates": float(np.mean(rv["updates"])), "note": "same pairs, voxel_size 0.2 (float64 records path), device time"}
    eng.invalidate(ids)

    # ---- the whole sequence as one map, voxel_size 0.2, ground-truth poses, host array out
    ppm = eng.make_preprocess_params(0.5, 35.0, -120.0, 120.0, voxel_size=0.2, want_normals=False)
    Tm = np.stack([seq.poses[int(k)] for k in ids])
    eng.map_build(ids, Tm, ppm)
    tm0 = time.perf_counter()
    for _ in range(nv):
        eng.invalidate(ids)
        mxyz, moff = eng.map_build(ids, Tm, ppm)
    tm = (time.perf_counter() - tm0) / nv
    line["map_build"] = {"keyframes": len(ids), "raw_points_per_s": float(sum(len(seq.scans[int(k)]) for k in ids)) / tm, "map_points": int(moff[-1]),
                         "ms": tm * 1e3, "d2h_bytes": int(moff[-1]) * 24, "note": "filter + voxel 0.2 + transform + concatenation of the batch, "
                         "map delivered to a pageable host array, wall clock"}
    eng.invalidate(ids)

    # ---- the reference's own call pattern (run_scanmatcher.py:196-213): one pair per call through the drop-in
    # KeyFrameManager - add_keyframe, load_pointcloud (PCD file), pre_process, compute_transformation, unload_pointcloud
    line["sequential_dropin"] = sequential_dropin(seq, engine, dev.index, n_scans=min(len(seq.scans), 40))

    # ---- small batches, as loop closing issues them (loopclosing.py:80-99: <= 2 x number_of_triplets pairs per call)
    eng.preprocess(ids, pp)
    small = {}
    for nb in (1, 2, 8, 40):
        nb = min(nb, P)
        ts = []
        for rep in range(6):
            t0 = time.perf_counter()
            r = eng.icp_batch(tg[:nb], sr[:nb], init[:nb], ip)
            ts.append(time.perf_counter() - t0)
        small[str(nb)] = {"ms_per_call": float(np.median(ts[1:]) * 1e3), "ms_per_pair": float(np.median(ts[1:]) * 1e3 / nb),
                          "passes_max": int(r["passes"].max())}
    line["small_batches"] = {"by_pairs_per_call": small, "note": "registration only (scans preprocessed), host wall clock per call incl. result delivery; "
                             "the loop ends on the device with the last convergence, so a call costs its own passes only"}
    for k in ids:
        eng.free(int(k))

    # ---- BASELINE configs[2]: 128-beam (~260k points) point-to-point
    line["config3_128beam_p2p"] = config3_line(args, torch, engine, synth, eng, stream)

    # ---- BASELINE configs[3] on ONE GPU: the leading pairs of the same sorted global list the N > 1 runs shard
    line["config4_single_gpu"] = config4_single(args, torch, engine, eng, stream)

    # ---- BASELINE configs[4]: scan-matcher + loop closing on the GPU feeding a host pose graph (gtsam stand-in, timing only)
    if args.c5_scans > 2:
        line["config5_pipeline"] = config5_line(args, torch, engine, synth, eng)


def sequential_dropin(seq, engine, device, n_scans):
    from lidar_slam_arvc_b200 import euroc_synth, runtime
    from lidar_slam_arvc_b200.homogeneousmatrix import HomogeneousMatrix
    dropin = os.path.join(ROOT, "lidar_slam_arvc_b200", "dropin")
    sys.path.insert(0, dropin)
    eng = engine.Engine(device)          # the drop-in numbers its scans itself: keep them apart from the bench's own
    runtime.set_engine(eng)
    try:
        import keyframemanager.keyframemanager as kfm
        with tempfile.TemporaryDirectory() as d:
            sub = type("S", (), {})()
            sub.scans, sub.odometry = seq.scans[:n_scans], seq.odometry[:n_scans]
            scan_times = euroc_synth.write_euroc_tree(d, sub)
            odo = [HomogeneousMatrix(seq.relative_odo(i, i + 1)) for i in range(n_scans - 1)]
            ts = []
            with contextlib.redirect_stdout(io.StringIO()):
                km = kfm.KeyFrameManager(directory=d, scan_times=scan_times, voxel_size=None, method="icppointplane")
                km.add_keyframe(0)
                km.load_pointcloud(0)
                km.pre_process(0)
                for i in range(n_scans - 1):
                    t0 = time.perf_counter()
                    km.add_keyframe(i + 1)
                    km.load_pointcloud(i + 1)
                    km.pre_process(i + 1)
                    km.compute_transformation(i, i + 1, Tij=odo[i])
                    km.unload_pointcloud(i)
                    ts.append(time.perf_counter() - t0)
                km.unload_pointcloud(n_scans - 1)
        ts = np.array(ts[2:]) * 1e3
        return {"ms_per_pair_median": float(np.median(ts)), "ms_per_pair_min": float(ts.min()), "pairs_per_s": float(1e3 / np.median(ts)), "pairs": len(ts),
                "note": "unchanged call sequence of run_scanmatcher.py:196-213 on the drop-in KeyFrameManager: PCD read from disk (page cache), upload, "
                        "preprocess of ONE scan, ONE pair per call, unload; host wall clock per loop iteration"}
    finally:
        runtime.set_engine(None)
        eng.close()
        sys.path.remove(dropin)


def config3_line(args, torch, engine, synth, eng, stream):
    n3 = 13
    seq3 = synth.Sequence(n3, synth.OS_128, start=30.0, workers=max(1, os.cpu_count() or 1))
    ids3 = np.arange(1000, 1000 + n3, dtype=np.int64)
    tg3, sr3 = ids3[:-1], ids3[1:]
    init3 = np.array([seq3.relative_odo(k, k + 1) for k in range(n3 - 1)])
    pin3 = pin_scans(torch, seq3.scans)
    pp3 = eng.make_preprocess_params(want_normals=False)
    ip3 = eng.make_icp_params(engine.P2P)
    for k, t in zip(ids3, pin3):
        eng.upload_ptr(int(k), t.data_ptr(), t.shape[0])

    def step3():
        eng.invalidate(ids3)
        eng.preprocess(ids3, pp3)
        return eng.icp_batch(tg3, sr3, init3, ip3)
    steps3 = max(3, args.steps // 2)
    ms, own3 = timed_steps(torch, stream, eng, step3, steps3, warm=2)
    npts3 = np.array([eng.info(int(k))["n_points"] for k in ids3])
    eng.set_option("icp_loop_graph", 0)
    step3()
    eng.sync()
    eng.profile_enable(True)
    pms, _ = timed_steps(torch, stream, eng, step3, steps3, warm=0)
    prof3, passes3 = split_profile(eng.profile_report())
    eng.profile_enable(False)
    eng.set_option("icp_loop_graph", 1)
    pass_points = float(sum(int(own3["passes"][k]) * int(npts3[k + 1]) for k in range(n3 - 1)))
    rl = roofline_dict(prof3, passes3, npts3, pass_points, steps3, pms * steps3, method_p2plane=False)
    # e2e: host scans -> records, every step
    def e2e3():
        for k, t in zip(ids3, pin3):
            eng.upload_ptr(int(k), t.data_ptr(), t.shape[0])
        eng.preprocess(ids3, pp3)
        return eng.icp_batch(tg3, sr3, init3, ip3)
    e2e3()
    eng.sync()
    t0 = time.perf_counter()
    for _ in range(steps3):
        e2e3()
    eng.sync()
    e2e_ms = (time.perf_counter() - t0) / steps3 * 1e3
    out = {"workload": "configs[2]: point-to-point ICP, 128-beam (128x2048) synthetic scans, %d consecutive pairs per step" % (n3 - 1),
           "value": (n3 - 1) / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "ms_per_pair": ms / (n3 - 1), "steps": steps3,
           "e2e": {"value": (n3 - 1) / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(sum(t.numel() * 4 for t in pin3))},
           "points_per_scan": int(npts3.mean()), "mean_icp_updates": float(np.mean(own3["updates"])), "roofline": rl}
    for k in ids3:
        eng.free(int(k))
    return out


def config5_line(args, torch, engine, synth, eng):
    from lidar_slam_arvc_b200 import pipeline
    n5 = args.c5_scans
    t0 = time.perf_counter()
    seq5 = synth.Sequence(n5, synth.OS1_64, start=0.0, workers=max(1, os.cpu_count() or 1))
    t_gen = time.perf_counter() - t0
    odo = [seq5.relative_odo(k, k + 1) for k in range(n5 - 1)]
    pin5 = pin_scans(torch, seq5.scans)
    # (no arvc_ctx_reserve here: this context's pool already holds the memory the earlier lines used; reserving on top of a
    # populated pool was measured to slow the first batches down - it is for a fresh context)
    eng.sync()
    t0 = time.perf_counter()
    l0 = eng.kernel_launches()
    rel, recs = pipeline.scan_matcher(eng, seq5.scans, odo, batch=100, pinned=pin5)
    eng.sync()
    t_front = time.perf_counter() - t0
    front_launches = eng.kernel_launches() - l0
    rep = pipeline.run_backend(eng, seq5.scans, rel, odo, skip_loop_closing=50, skip_optimization=50, number_of_triplets_loop_closing=20,
                               distance_backwards=7.0, radius_threshold=5.0, seed=0)
    # how well the registrations did (ground truth is known for synthetic data)
    gt = np.array([seq5.relative_gt(k, k + 1) for k in range(n5 - 1)])
    err_t = np.linalg.norm(rel[:, :3, 3] - gt[:, :3, 3], axis=1)
    total = t_front + rep["total_s"]
    return {"workload": "configs[4]: %d synthetic 64-beam keyframes (%.1f laps of the loop corridor): GPU scan-matcher (device batches of 100 keyframes from "
                        "pinned host scans) + drop-in LoopClosing.loop_closing_triangle every 50 steps (20 triplets, one device batch per invocation, "
                        "keyframes cached on the device) + host pose graph" % (n5, n5 / 232.0),
            "scans": n5, "total_s": total, "scans_per_s": n5 / total,
            "front_end": {"s": t_front, "pairs": n5 - 1, "pairs_per_s": (n5 - 1) / t_front, "gpu_launches": int(front_launches),
                          "median_translation_error_m": float(np.median(err_t)), "mean_icp_updates": float(np.mean(recs["updates"]))},
            "back_end": rep,
            "shares": {"front_end_gpu": t_front / total, "loop_closing_engine": rep["loop_closing_engine_s"] / total,
                       "loop_closing_candidate_search_host": rep["loop_closing_candidate_search_s"] / total,
                       "pose_graph_host": rep["optimize_s"] / total, "other_host": rep["other_host_s"] / total},
            "scan_generation_s": t_gen,
            "note": "gtsam is not installed in this image: the host back end is lidar_slam_arvc_b200.pipeline.PoseGraphStandIn (one sparse linear solve for "
                    "the positions per optimize(), rotations fixed) - a timing stand-in with the same duck-typed surface, no accuracy claim; the reference's "
                    "unchanged host loop (run_graphSLAM.py:229-267) is what the back-end share measures"}


def config4_single(args, torch, engine, eng, stream):
    """The WHOLE global list of the N > 1 runs on one GPU, same device batches: the strong-scaling reference."""
    wl = LoopClosureWorkload(args.lc_scans, args.lc_pairs, max(1, os.cpu_count() or 1))
    n4 = len(wl.tg) if args.lc_pairs_n1 <= 0 else min(args.lc_pairs_n1, len(wl.tg))
    off = 1 << 20                                     # scan ids apart from everything else in this context
    tg4, sr4, init4 = wl.tg[:n4] + off, wl.sr[:n4] + off, wl.init[:n4]
    scans4 = np.unique(np.concatenate([wl.tg[:n4], wl.sr[:n4]]))
    wl.materialise(scans4)
    pin4 = pin_scans(torch, [wl.scans[int(k)] for k in scans4])
    ids4 = scans4 + off
    pp = eng.make_preprocess_params()
    ip = eng.make_icp_params(engine.P2PLANE)
    B = max(1, args.lc_batch)
    for k, t in zip(ids4, pin4):
        eng.upload_ptr(int(k), t.data_ptr(), t.shape[0])

    def step4():
        eng.invalidate(ids4)
        eng.preprocess(ids4, pp)
        tickets = [eng.icp_batch_async(tg4[s0:s0 + B], sr4[s0:s0 + B], init4[s0:s0 + B], ip) for s0 in range(0, n4, B)]
        return np.concatenate([eng.icp_batch_finish(t) for t in tickets])
    steps4 = 2
    ms, own4 = timed_steps(torch, stream, eng, step4, steps4, warm=1)
    out = {"workload": "configs[3] on one GPU: %d of the %d loop-closure pairs of the sorted global list (%d scans touched), device batches of <= %d pairs"
                       % (n4, len(wl.tg), len(ids4), B),
           "value": n4 / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "ms_per_pair": ms / n4, "steps": steps4,
           "mean_icp_updates": float(np.mean(own4["updates"])), "max_icp_updates": int(own4["updates"].max()),
           "note": "the strong-scaling reference for the N > 1 lines (same list, same code path, scans resident, device time)"}
    for k in ids4:
        eng.free(int(k))
    return out


# ---------------------------------------------------------------------------------------------- N > 1
def run_sharded(args, torch, dist, engine, sharding, rank, world, local_rank):
    dev = torch.device("cuda", local_rank)
    wl = LoopClosureWorkload(args.lc_scans, args.lc_pairs, max(1, (os.cpu_count() or 1) // world))
    G = len(wl.tg)
    if G < world:
        raise SystemExit("no loop-closure candidates: --lc-scans %d must cover more than one lap (232 keyframes) of the synthetic corridor" % args.lc_scans)
    # Shard bounds can move between steps (see `rebalance`), so a rank keeps host copies of every scan the list touches;
    # it UPLOADS and preprocesses only the scans of its current shard.
    all_scans = sharding.scans_of_pairs(wl.tg, wl.sr)
    wl.materialise(all_scans)
    pinned = dict(zip([int(k) for k in all_scans], pin_scans(torch, [wl.scans[int(k)] for k in all_scans])))

    eng = engine.Engine(local_rank)
    pp = eng.make_preprocess_params()
    ip = eng.make_icp_params(engine.P2PLANE)
    stream = torch.cuda.ExternalStream(eng.stream_handle(), device=dev)
    B = max(1, args.lc_batch)
    gather = sharding.DeviceGather(eng, dev, B)
    state = {"bounds": [sharding.shard_bounds(G, world, r)[0] for r in range(world)] + [G], "resident": set()}

    busy_buf = torch.zeros(world, dtype=torch.float64, device=dev)

    def rebalance(records, bounds, busy_ms):
        """Run-time load balancer.  Cost history of the batch that has just been gathered: every rank holds all records
        (passes per pair) and learns every rank's measured busy time (one 8-byte all-gather); a pair's cost estimate is
        (3 + passes) x the time its rank needed per such unit, and the contiguous shard bounds of the NEXT batch equalise
        the summed estimates.  The first batch is split by pair count.  A loop-closing back end sees the same places again
        and again (run_graphSLAM.py:259-263), which is what makes a cost history meaningful; here the list repeats exactly."""
        if args.lc_balance != "history":
            return
        mine = torch.tensor([busy_ms], dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(busy_buf, mine)
        state["bounds"] = sharding.rebalanced_bounds(records["passes"], bounds, busy_buf.cpu().numpy())

    def step(upload, ev=None):
        """One pass over this rank's shard of the global list.  Batch b + 1 is enqueued before batch b is collected, so the
        all-gather and the device -> host copy of batch b overlap the kernels of batch b + 1."""
        bounds = state["bounds"]
        lo, hi = bounds[rank], bounds[rank + 1]
        sizes = [bounds[r + 1] - bounds[r] for r in range(world)]
        tg, sr, init = wl.tg[lo:hi], wl.sr[lo:hi], wl.init[lo:hi]
        my_scans = sharding.scans_of_pairs(tg, sr)
        ev = ev or (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record(stream)
        for k in state["resident"] - set(int(x) for x in my_scans):      # scans that left the shard with the last rebalancing
            eng.free(k)
        if upload:
            ch = growing_chunks(len(my_scans))
            for a, b in zip(ch[:-1], ch[1:]):
                for k in my_scans[a:b]:
                    eng.upload_ptr(int(k), pinned[int(k)].data_ptr(), pinned[int(k)].shape[0])
                eng.preprocess(my_scans[a:b], pp)
        else:
            for k in my_scans:
                if int(k) not in state["resident"]:
                    eng.upload_ptr(int(k), pinned[int(k)].data_ptr(), pinned[int(k)].shape[0])
            eng.invalidate(my_scans)
            eng.preprocess(my_scans, pp)
        state["resident"] = set(int(x) for x in my_scans)
        n_batches = max((n + B - 1) // B for n in sizes)
        pending, parts = [], []
        for b in range(n_batches):
            s0, s1 = min(b * B, len(tg)), min((b + 1) * B, len(tg))
            ticket = eng.icp_batch_async(tg[s0:s1], sr[s0:s1], init[s0:s1].reshape(-1, 4, 4), ip)
            if b == n_batches - 1:
                ev[1].record(stream)                      # end of this rank's own work of the step (before its last gather)
            pending.append((ticket, gather.start(ticket, sharding.batch_counts(bounds, b, B))))
            if len(pending) > 1:
                t, slot = pending.pop(0)
                parts.append(gather.finish(slot))
                eng.icp_batch_finish(t)
        while pending:
            t, slot = pending.pop(0)
            parts.append(gather.finish(slot))
            eng.icp_batch_finish(t)
        # the gathered batches, back in the order of the global list
        out = sharding.assemble_global(parts, bounds, B)
        info = {"pairs": hi - lo, "scans": len(my_scans), "bounds": list(bounds), "h2d": int(sum(pinned[int(k)].numel() * 4 for k in my_scans))}
        rebalance(out, bounds, ev[0].elapsed_time(ev[1]))
        return out, info

    def barrier():
        eng.sync()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(max(args.warmup, 3)):
        step(False)
        step(True)
    eng.sync()

    # ---- value: scans resident in HBM
    step(False)
    barrier()
    l0 = eng.kernel_launches()
    tw0 = time.perf_counter()
    busy_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    step_ms = []
    for s in range(args.steps):
        ts0 = time.perf_counter()
        records, info = step(False, busy_ev[s])
        step_ms.append(round((time.perf_counter() - ts0) * 1e3, 1))
    ev1.record(stream)
    barrier()
    dev_ms_own = ev0.elapsed_time(ev1)
    busy_ms = float(np.mean([a.elapsed_time(b) for a, b in busy_ev]))
    launches = eng.kernel_launches() - l0
    t_ms = torch.tensor([dev_ms_own], dtype=torch.float64, device=dev)
    dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    dev_ms = float(t_ms.item())
    value = G * args.steps / (dev_ms * 1e-3)

    # ---- e2e: host scans -> all records on every rank's host, every step
    step(True)
    barrier()
    t0 = time.perf_counter()
    e2e_step_ms = []
    for _ in range(args.steps):
        ts0 = time.perf_counter()
        records, info = step(True)
        e2e_step_ms.append(round((time.perf_counter() - ts0) * 1e3, 1))
    barrier()
    e2e_s = time.perf_counter() - t0
    time.sleep(0.25)
    clocks.stop()
    clk = clocks.window(tw0, t0 + e2e_s)
    t_s = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    dist.all_reduce(t_s, op=dist.ReduceOp.MAX)
    e2e_value = G * args.steps / float(t_s.item())

    # ---- roofline of the dominant kernel: rank 0 runs the first device batch of its shard once more with per-launch CUDA
    # events (passes enqueued one by one: kernels inside a graph cannot be bracketed), after the timed regions
    rl = None
    if rank == 0:
        lo0, hi0 = info["bounds"][0], info["bounds"][1]
        nb = min(B, hi0 - lo0)
        if nb > 0:
            tgp, srp, initp = wl.tg[lo0:lo0 + nb], wl.sr[lo0:lo0 + nb], wl.init[lo0:lo0 + nb]
            eng.set_option("icp_loop_graph", 0)
            eng.icp_batch(tgp, srp, initp, ip)
            eng.sync()
            eng.profile_enable(True)
            t0p = time.perf_counter()
            recp = eng.icp_batch(tgp, srp, initp, ip)
            eng.sync()
            region_ms = (time.perf_counter() - t0p) * 1e3
            profp, passesp = split_profile(eng.profile_report())
            eng.profile_enable(False)
            eng.set_option("icp_loop_graph", 1)
            npts = {int(k): eng.info(int(k))["n_points"] for k in np.unique(srp)}
            pass_points = float(sum(int(recp["passes"][k]) * npts[int(srp[k])] for k in range(nb)))
            rl = roofline_dict(profp, passesp, np.array(list(npts.values())), pass_points, 1, region_ms)
            if rl is not None:
                rl["sample"] = "rank 0, the first %d pairs of its shard, one registration batch (scans resident, no preprocessing in it)" % nb

    # ---- parity: a sample of THIS rank's loop-closure pairs against the oracle (same run, after the timed regions)
    lo, hi = info["bounds"][rank], info["bounds"][rank + 1]
    npar = min(args.parity_pairs, hi - lo)
    sample = [lo + int(v) for v in np.linspace(0, hi - lo - 1, npar).round()] if npar else []
    _, cpu_dt, refs, _ = oracle_loop_closure(wl, sample, max(1, (os.cpu_count() or 1) // world))
    rep = parity_report([records[k] for k in sample], refs, "rank %d: pairs %s of the global list (its shard)" % (rank, sample))
    mine = {"rank": rank, "pairs": int(info["pairs"]), "scans": int(info["scans"]), "busy_ms_per_step": busy_ms, "device_ms_per_step": dev_ms_own / args.steps,
            "mean_icp_updates": float(np.mean(records["updates"][lo:hi])) if hi > lo else 0.0, "launches": int(launches), "h2d_bytes_per_step": info["h2d"],
            "parity": rep}
    infos = [None] * world
    dist.all_gather_object(infos, mine)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config4(G, args.lc_scans, world, B), "ms_per_pair": dev_ms / (args.steps * G),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(sum(i["h2d_bytes_per_step"] for i in infos) + wl.init.nbytes + 16 * G),
                        "d2h_bytes_per_step": int(160 * G * world + 164 * G), "ms_per_step": float(t_s.item()) / args.steps * 1e3},
                "gpu_launches": int(sum(i["launches"] for i in infos)), "clocks": clk, "step_ms": step_ms, "e2e_step_ms": e2e_step_ms,
                "per_rank": [{k: v for k, v in i.items() if k != "parity"} for i in infos],
                "per_gpu_rate_pairs_per_s": float(np.mean([i["pairs"] / (i["busy_ms_per_step"] * 1e-3) for i in infos if i["pairs"]])),
                "sharding": {"balance": args.lc_balance, "bounds": info["bounds"],
                             "note": "contiguous shards of the (target, source)-sorted list; 'history': bounds re-placed after every batch from the passes "
                                     "the gathered records report (equal summed cost), 'count': equal pair counts"},
                "gather": {"records_bytes_per_rank_per_batch": int(160 * min(B, max(i["pairs"] for i in infos))), "batches_per_step": int(max((i["pairs"] + B - 1) // B for i in infos)),
                           "exposed_ms_per_step": float(max(0.0, dev_ms_own / args.steps - busy_ms)),
                           "note": "all_gather_into_tensor straight from the engine's device records on the engine's stream + one D2H on every rank; "
                                   "exposed = rank 0's step time minus its own kernels' time (gather of the last batch + waiting for the slowest rank)"},
                "mean_icp_updates": float(np.mean(records["updates"])),
                "parity_check": {"ok": all(i["parity"]["ok"] for i in infos), "per_rank": [i["parity"] for i in infos]},
                "roofline": rl,
                "notes": {"strong_scaling_reference": "the N=1 line's `config4_single_gpu` runs the leading pairs of the same list on one GPU; "
                                                      "per_gpu_rate_pairs_per_s is the same quantity measured inside this run (pairs / own busy time, mean over ranks)",
                          "roofline": "the dominant kernel of a registration-only batch (the search), timed on rank 0 after the timed regions; the N=1 "
                                      "line carries the roofline of the headline workload (normals dominate there)",
                          "reference_arm": "bench.py --impl reference --gpus N times --ref-pairs loop-closure pairs of the same list per step"}}
        print(json.dumps(line), flush=True)
    gather.close()
    del gather
    eng.close()
    ok = all(i["parity"]["ok"] for i in infos)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit("parity check failed: %s" % [i["parity"] for i in infos])


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import __graft_entry__ as entry
    entry.build()
    from lidar_slam_arvc_b200 import engine, sharding, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ICP engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1 or os.environ.get("ARVC_BENCH_FORCE_SHARDED"):      # (the variable: one-rank run of the sharded code path, for A/B)
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        os.environ.setdefault("RANK", "0")
        os.environ.setdefault("WORLD_SIZE", "1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        run_sharded(args, torch, dist, engine, sharding, rank, world, local_rank)
    else:
        run_single(args, torch, engine, synth, local_rank)


if __name__ == "__main__":
    main()
