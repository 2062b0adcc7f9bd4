#!/usr/bin/env python
"""Benchmark of the ICP scan-matching hot path (BASELINE.json metric: ICP pairs/sec, 64-beam point-to-plane).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path over one batch of `--pairs` consecutive keyframe pairs per GPU
(`--pairs`+1 synthetic OS1-64 scans): preprocessing of every scan (radius/height filter, Morton sort, hash grid,
k-NN normals) + point-to-plane ICP of every pair to Open3D's default convergence criteria.
  value : pairs/s, raw scans already resident in HBM when the timed region starts (device time, CUDA events)
  e2e   : pairs/s through the C-ABI from pinned HOST buffers: H2D of every scan + preprocessing + ICP +
          D2H of the result records, every step (host wall clock between synchronisations, max over ranks)
Weak scaling: every rank owns its own batch; the only collective is the all-gather of 160-byte result records.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "icp_pairs_per_sec"
UNIT = "pairs/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=99, help="keyframe pairs per GPU per step (BASELINE.md config 2: the 99 pairs of a 100-scan sequence)")
    ap.add_argument("--ref-pairs", type=int, default=6, help="pairs per step of the CPU reference arm (bounded sample)")
    ap.add_argument("--cpu-pairs", type=int, default=6, help="pairs of the cpu_baseline sample (N=1 only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rank-stride", type=float, default=0.0,
                    help="metres between the trajectory starts of consecutive ranks; 0 = the same synthetic sequence on every "
                         "rank, i.e. exactly equal per-GPU work (weak scaling); > 0 gives every rank its own scans")
    ap.add_argument("--no-voxel", action="store_true", help="skip the extra voxel_size 0.2 measurement")
    return ap.parse_args()


def workload_config(pairs, n_gpus):
    return {"workload": "configs[1]: batched consecutive-keyframe point-to-plane ICP, 64-beam (OS1-64-like 64x1024) synthetic scans",
            "pairs_per_gpu_per_step": pairs, "scans_per_gpu_per_step": pairs + 1, "method": "icppointplane", "voxel_size": None,
            "max_corr_dist": 10.0, "criteria": "rel_fitness=1e-6 rel_rmse=1e-6 max_iter=30", "normals": "radius=0.3 max_nn=300",
            "l2": "inputs larger than L2 (every step re-streams %d MB of scans, grids and normals per GPU)" % (12 * (pairs + 1)),
            "parallelism": "pairs sharded x%d, all-gather of 160 B records" % n_gpus,
            "per_rank_data": "every rank processes its own copy of the same synthetic sequence (equal per-GPU work); "
                             "--rank-stride > 0 gives every rank different scans"}


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


# ---------------------------------------------------------------------------------------------- CPU arms
def cpu_pairs_per_sec(seq, n_pairs):
    """The oracle (C++/OpenMP float64 restatement of the reference's Open3D CPU path) on all host threads:
    preprocessing of every scan once + ICP of consecutive pairs, like run_scanmatcher.py:191-213."""
    from oracle import oracle as orc
    orc.set_num_threads(os.cpu_count() or 1)    # torchrun exports OMP_NUM_THREADS=1: use every host thread anyway
    pre = [orc.preprocess(seq.scans[0])]        # steady state of consecutive matching: one new scan per pair
    t0 = time.perf_counter()
    for k in range(n_pairs):
        pre.append(orc.preprocess(seq.scans[k + 1]))
        tgt, ntgt = pre[k]
        src, _ = pre[k + 1]
        orc.icp(src, tgt, ntgt, seq.relative_odo(k, k + 1), orc.P2PLANE)
    dt = time.perf_counter() - t0
    return n_pairs / dt, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port: Open3D is not installable here),
    all host threads, bounded sample per step.  Rank 0 only."""
    if rank != 0:
        return
    from lidar_slam_arvc_b200 import synth
    from oracle import oracle as orc
    sp = max(1, args.ref_pairs)
    seq = synth.Sequence(sp + 1, synth.OS1_64, start=30.0)
    for _ in range(args.warmup):
        cpu_pairs_per_sec(seq, sp)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pairs_per_sec(seq, sp)
    dt = time.perf_counter() - t0
    value = args.steps * sp / dt
    cores = orc.num_threads()
    cfg = workload_config(args.pairs, args.gpus)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d consecutive 64-beam pairs per step (%d new scans preprocessed + %d ICPs), oracle C++/OpenMP on %d threads"
                                       % (sp, sp, sp, cores)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- ours
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    entry.build()
    from lidar_slam_arvc_b200 import engine, sharding, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ICP engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    P = args.pairs
    seq = synth.Sequence(P + 1, synth.OS1_64, start=30.0 + args.rank_stride * rank, workers=max(1, (os.cpu_count() or 1) // max(world, 1)))
    ids = np.arange(P + 1, dtype=np.int64)
    tg, sr = ids[:-1], ids[1:]
    init = np.array([seq.relative_odo(int(a), int(b)) for a, b in zip(tg, sr)])
    pinned = []
    for s in seq.scans:
        t = torch.empty((len(s), 3), dtype=torch.float32).pin_memory()
        t.copy_(torch.from_numpy(s))
        pinned.append(t)
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
        rec = eng.icp_batch(tg, sr, init, ip)
        return sharding.gather_records(rec, device=dev, counts=[P] * world) if world > 1 else rec

    # e2e: geometrically growing chunks (6, 12, 24, ... scans) - the upload of chunk c+1 runs on the engine's copy stream
    # while chunk c is preprocessed, also when several ranks share the host's PCIe / memory bandwidth
    bounds, head = [0], 6
    while bounds[-1] < len(ids):
        bounds.append(min(len(ids), bounds[-1] + head * 2 ** (len(bounds) - 1)))
    if len(bounds) > 2 and bounds[-1] - bounds[-2] < head:
        bounds.pop(-2)

    def e2e_path():
        """Host scans -> records: every chunk is uploaded (copy stream) and then preprocessed (compute stream, waits for
        its own uploads only); the ICP batch is the same single call."""
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            for k in range(lo, hi):
                eng.upload_ptr(k, pinned[k].data_ptr(), pinned[k].shape[0])
            eng.preprocess(ids[lo:hi], pp)
        rec = eng.icp_batch(tg, sr, init, ip)
        return sharding.gather_records(rec, device=dev, counts=[P] * world) if world > 1 else rec

    def barrier():
        eng.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)      # started well before the timed regions: nvidia-smi start-up stalls the driver
    clocks.start()
    # ---- warm-up (also fills the stream-ordered memory pool)
    for _ in range(max(args.warmup, 1)):
        upload_all()
        rec = hot_path()
        rec = e2e_path()
    eng.sync()
    n_pts = np.array([eng.info(int(k))["n_points"] for k in ids])

    # ---- value: scans resident in HBM, device time
    upload_all()
    barrier()
    eng.profile_enable(not os.environ.get("ARVC_BENCH_NO_PROFILE"))
    l0 = eng.kernel_launches()
    tw0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    step_ms = []
    for _ in range(args.steps):
        ts0 = time.perf_counter()
        rec_local = hot_path()
        step_ms.append(round((time.perf_counter() - ts0) * 1e3, 1))
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    if os.environ.get("ARVC_BENCH_RANK_LOG"):
        print("[rank %d] device %.2f ms for %d steps, host step ms %s" % (rank, dev_ms, args.steps, step_ms), file=sys.stderr, flush=True)
    launches = eng.kernel_launches() - l0
    prof_raw = eng.profile_report()
    prof, icp_passes = {}, {}
    for k, v in prof_raw.items():                      # the engine reports every ICP pass separately
        if k.startswith("icp_pass_"):
            icp_passes[k[-2:]] = round(v[1] / v[0], 4)
            c, t = prof.get("icp_pass", (0, 0.0))
            prof["icp_pass"] = (c + v[0], t + v[1])
        else:
            prof[k] = v
    eng.profile_enable(False)
    tw1 = time.perf_counter()
    own = eng.icp_batch(tg, sr, init, ip)          # this rank's own records (cached preprocessing), for the byte model
    t_ms = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    dev_ms = float(t_ms.item())
    value = world * P * args.steps / (dev_ms * 1e-3)

    # ---- e2e: host buffers -> result records on the host, every step
    rec_all = e2e_path()      # untimed: back from the resident-scan path to the upload path (re-sizes the engine's scratch blocks)
    barrier()
    t0 = time.perf_counter()
    e2e_step_ms = []
    for _ in range(args.steps):
        ts0 = time.perf_counter()
        rec_all = e2e_path()
        e2e_step_ms.append(round((time.perf_counter() - ts0) * 1e3, 1))
    barrier()
    e2e_s = time.perf_counter() - t0
    time.sleep(0.25)
    clocks.stop()
    clk = clocks.window(tw0, t0 + e2e_s)
    t_s = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_s, op=dist.ReduceOp.MAX)
    e2e_value = world * P * args.steps / float(t_s.item())
    d2h_bytes = int(464 * P + (160 * P * world if world > 1 else 0))   # per-pair state read-back (+ gathered records)

    # ---- roofline of the dominant kernel (device events recorded around every launch of the timed region)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (KeyError, ValueError, TypeError):
        pass
    total_kernel_ms = sum(v[1] for v in prof.values())
    dom = max(prof.items(), key=lambda kv: kv[1][1]) if prof else ("none", (0, 0.0))
    name, (n_launch, tot_ms) = dom
    # SURVEY.md §8(d) byte model: normals 32*M per scan; ICP 48*N_s per pair and executed pass (point-to-plane)
    if name == "icp_pass":
        alg_bytes = float(sum(int(own["passes"][k]) * 48 * int(n_pts[k + 1]) for k in range(P))) * args.steps
    elif name == "normals":
        alg_bytes = float(32 * n_pts.sum()) * args.steps
    else:
        alg_bytes = float(36 * n_pts.sum()) * args.steps
    achieved = alg_bytes / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and n_launch:      # DRAM bytes per unit from the committed ncu --set full capture, scaled to this launch
        tj = json.load(open(tpath))
        if name == "normals" and "normals" in tj:
            traffic = tj["normals"]["dram_bytes_per_scan"] * len(ids)
        elif name == "icp_pass" and "icp_pass" in tj:
            traffic = tj["icp_pass"]["dram_bytes_per_pair_pass"] * float(sum(int(x) for x in own["passes"])) * args.steps / n_launch
    ncu_util = None
    if os.path.exists(tpath):
        ncu_util = {k: v for k, v in json.load(open(tpath)).get(name, {}).items() if k.endswith("_pct")}
    roofline = {"bound": "hbm", "kernel": name, "ncu_utilisation_pct": ncu_util, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "launches": n_launch, "kernel_ms_total": tot_ms,
                "kernel_share_of_device_time": tot_ms / dev_ms if dev_ms > 0 else None,
                "algorithmic_bytes_per_launch": alg_bytes / max(n_launch, 1),
                "all_kernels_ms": {k: round(v[1], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
                "icp_pass_avg_ms_by_pass": icp_passes,
                "note": "working set per pair is L2-resident and the search is FP64/LSU bound; see DESIGN.md"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(P, world), "ms_per_pair": dev_ms / (args.steps * P),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": float(t_s.item()) / args.steps * 1e3},
            "gpu_launches": int(launches), "clocks": clk, "step_ms": step_ms, "e2e_step_ms": e2e_step_ms, "roofline": roofline,
            "mean_icp_updates": float(np.mean(own["updates"])), "icp_updates": [int(u) for u in own["updates"]], "points_per_scan": int(n_pts.mean()),
            "notes": {"value_region": "carries one CUDA-event pair per kernel launch (the per-kernel times of `roofline`), ~1 % overhead",
                      "e2e_region": "uploads on the engine's copy stream in chunks of %s scans: chunk c is preprocessed while chunk c+1 is in flight" % [b - a for a, b in zip(bounds[:-1], bounds[1:])]}}

    # ---- extra (BASELINE.md config 2 is reported for voxel_size None and 0.2): same batch with voxel down-sampling on
    if world == 1 and not args.no_voxel:
        ppv = eng.make_preprocess_params(voxel_size=0.2)
        for _ in range(2):
            eng.invalidate(ids); eng.preprocess(ids, ppv); rv = eng.icp_batch(tg, sr, init, ip)
        eng.sync()
        tv0 = time.perf_counter()
        nv = max(2, args.steps // 3)
        for _ in range(nv):
            eng.invalidate(ids); eng.preprocess(ids, ppv); rv = eng.icp_batch(tg, sr, init, ip)
        tv = time.perf_counter() - tv0
        line["voxel_0p2"] = {"value": P * nv / tv, "unit": UNIT, "steps": nv, "points_per_scan": int(np.mean([eng.info(int(k))["n_points"] for k in ids[:8]])),
                             "mean_icp_updates": float(np.mean(rv["updates"])), "note": "same pairs, voxel_size 0.2 (float64 records path), wall clock"}
        eng.invalidate(ids)
        # ---- extra: the whole sequence as one map, voxel_size 0.2, ground-truth poses, host array out
        ppm = eng.make_preprocess_params(0.5, 35.0, -120.0, 120.0, voxel_size=0.2, want_normals=False)
        Tm = np.stack([seq.poses[int(k)] for k in ids])
        eng.map_build(ids, Tm, ppm)
        tm0 = time.perf_counter()
        for _ in range(nv):
            eng.invalidate(ids); mxyz, moff = eng.map_build(ids, Tm, ppm)
        tm = (time.perf_counter() - tm0) / nv
        raw_pts = float(sum(len(seq.scans[int(k)]) for k in ids))
        line["map_build"] = {"keyframes": len(ids), "raw_points_per_s": raw_pts / tm, "map_points": int(moff[-1]), "ms": tm * 1e3,
                             "d2h_bytes": int(moff[-1]) * 24, "note": "filter + voxel 0.2 + transform + concatenation of the batch, "
                             "map delivered to a pageable host array, wall clock"}
        eng.invalidate(ids)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        cp = max(1, min(args.cpu_pairs, P))
        v, dt = cpu_pairs_per_sec(seq, cp)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                                "sample": "first %d pairs of the same batch (%d new scans preprocessed + %d ICPs) in %.1f s, oracle C++/OpenMP"
                                          % (cp, cp, cp, dt)}
        # parity asserted in the same run on the sampled pairs
        tgt, ntgt = orc.preprocess(seq.scans[0])
        src, _ = orc.preprocess(seq.scans[1])
        ref = orc.icp(src, tgt, ntgt, init[0], orc.P2PLANE)
        line["parity_check"] = {"pair": 0, "max_abs_dT": float(np.abs(own["T"][0] - ref.transformation).max()),
                                "rmse_rel": float(abs(own["rmse"][0] - ref.inlier_rmse) / ref.inlier_rmse),
                                "updates_equal": bool(own["updates"][0] == ref.updates)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
