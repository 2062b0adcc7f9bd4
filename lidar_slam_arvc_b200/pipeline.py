"""BASELINE configs[4]: the full pipeline around the hot path - scan-matcher front end + loop closing on the GPU, feeding
a host pose graph - for timing.

Reference call stacks:
  run_scanmatcher.py:188-213  consecutive keyframes registered one pair at a time          -> `scan_matcher` (device batches)
  run_graphSLAM.py:229-267    per step: initial estimate + SM / ODO edges; every `skip_optimization` steps optimize();
                              every `skip_loop_closing` steps LoopClosing.loop_closing_triangle -> `run_backend`
The loop-closing object is the drop-in `graphslam.loopclosing.LoopClosing` (same gates, same random draws), driving the
drop-in `KeyFrameManager`; its registrations are ONE device batch per invocation and its keyframes stay preprocessed
on the device between invocations.

gtsam is not installed in this image, and the north_star leaves "the gtSAM pose-graph optimisation on the
host unchanged": `PoseGraphStandIn` only stands in for it so that the host side of the loop has a realistic shape and
cost - same duck-typed surface LoopClosing uses (`current_estimate.atPose3(i).matrix()`, `.exists(i)`, `T0_gps`,
`add_edge`), `optimize()` = one sparse linear solve for the positions with the rotations of the chained estimate held
fixed.  It is NOT a replacement for gtsam's nonlinear optimisation and no accuracy claim is attached to it.
"""
import contextlib
import io
import os
import sys
import time

import numpy as np

from . import runtime
from .engine import P2PLANE
from .homogeneousmatrix import HomogeneousMatrix

_DROPIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")


@contextlib.contextmanager
def dropin_modules(engine):
    """Import the drop-in `keyframemanager` / `graphslam` packages (the reference's import paths) bound to `engine`."""
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("config", "keyframemanager", "graphslam")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, _DROPIN)
    runtime.set_engine(engine)
    try:
        import graphslam.loopclosing as lc
        import keyframemanager.keyframemanager as kfm
        yield kfm, lc
    finally:
        runtime.set_engine(None)
        sys.path.remove(_DROPIN)
        for k in [k for k in sys.modules if k.split(".")[0] in ("config", "keyframemanager", "graphslam")]:
            del sys.modules[k]
        sys.modules.update(saved)


# ---------------------------------------------------------------------------------------------- host back-end stand-in
class _Pose3:
    def __init__(self, M):
        self._M = M

    def matrix(self):
        return self._M


class _Values:
    """Poses in one preallocated array [capacity, 4, 4] (a run adds thousands, optimize() rewrites all positions at once)."""

    def __init__(self):
        self.buf = np.zeros((1024, 4, 4))
        self.n = 0

    def append(self, M):
        if self.n == len(self.buf):
            self.buf = np.concatenate([self.buf, np.zeros_like(self.buf)])
        self.buf[self.n] = M
        self.n += 1

    @property
    def poses(self):
        return self.buf[:self.n]

    def exists(self, i):
        return 0 <= i < self.n

    def atPose3(self, i):
        return _Pose3(self.buf[i])


class PoseGraphStandIn:
    """See the module docstring: a timing stand-in for graphslam.graphSLAM.GraphSLAM (gtsam), not a substitute."""

    WEIGHTS = {"SM": 1.0, "ODO": 0.05}

    def __init__(self, T0=None, T0_gps=None):
        self.T0 = np.eye(4) if T0 is None else np.asarray(getattr(T0, "array", T0), dtype=np.float64)
        self.T0_gps = T0_gps if T0_gps is not None else HomogeneousMatrix(np.eye(4))
        self.current_estimate = _Values()
        self.edges = []                                   # (i, j, 4x4, kind) - what the tests and callers read
        self._ei, self._ej, self._et, self._ew = [], [], [], []
        self.n_optimizations = 0

    def init_graph(self):
        self.current_estimate = _Values()
        self.current_estimate.append(self.T0)

    def add_initial_estimate(self, atb, k):
        est = self.current_estimate
        assert k == est.n
        est.append(est.buf[k - 1] @ np.asarray(atb.array))

    def add_edge(self, atb, i, j, kind):
        A = np.array(atb.array)
        self.edges.append((int(i), int(j), A, kind))
        self._ei.append(int(i)); self._ej.append(int(j)); self._et.append(A[:3, 3]); self._ew.append(np.sqrt(self.WEIGHTS.get(kind, 1.0)))

    def optimize(self):
        """Positions from one sparse least-squares solve (rotations fixed): sum_e w_e |p_j - p_i - R_i t_ij|^2 + prior on p_0."""
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        P = self.current_estimate.poses
        n = len(P)
        if n < 2 or not self.edges:
            return
        ei, ej = np.array(self._ei), np.array(self._ej)
        ok = (ei < n) & (ej < n)
        ei, ej = ei[ok], ej[ok]
        t = np.array(self._et)[ok]
        w = np.array(self._ew)[ok]
        m = len(ei)
        rows = np.concatenate([np.arange(m), np.arange(m), [m]])
        cols = np.concatenate([ej, ei, [0]])
        vals = np.concatenate([w, -w, [1e3]])
        A = sp.csr_matrix((vals, (rows, cols)), shape=(m + 1, n))
        rhs = np.vstack([w[:, None] * np.einsum("kab,kb->ka", P[ei, :3, :3], t), 1e3 * P[0, :3, 3][None]])
        lu = spla.splu((A.T @ A).tocsc())
        P[:, :3, 3] = lu.solve(A.T @ rhs)
        self.n_optimizations += 1

    def get_solution_transforms_lidar(self):
        Tinv = np.linalg.inv(np.asarray(self.T0_gps.array))
        return [HomogeneousMatrix(P @ Tinv) for P in self.current_estimate.poses]


# ---------------------------------------------------------------------------------------------- front end
def scan_matcher(engine, scans, odometry_rel, batch=100, method=P2PLANE, pinned=None):
    """run_scanmatcher.py:188-213 for a whole sequence, `batch` keyframes per device batch (one-scan halo between batches):
    upload -> preprocess -> registration of the consecutive pairs -> free.  Returns (relative transforms [n-1,4,4], records)."""
    n = len(scans)
    pp = engine.make_preprocess_params(want_normals=method == P2PLANE)
    ip = engine.make_icp_params(method)
    base = 1 << 40                                   # scan ids of the front end, apart from the drop-in's own numbering
    rel, recs = [], []
    resident = set()

    def up(k):
        if k in resident:
            return
        if pinned is not None:
            engine.upload_ptr(base + k, pinned[k].data_ptr(), pinned[k].shape[0])
        else:
            engine.upload(base + k, scans[k])
        resident.add(k)

    lo = 0
    pending = None
    while lo < n - 1:
        hi = min(n, lo + batch)
        for k in range(lo, hi):
            up(k)
        ids = np.arange(lo, hi, dtype=np.int64) + base
        engine.preprocess(ids, pp)
        init = np.asarray(odometry_rel[lo:hi - 1], dtype=np.float64)
        ticket = engine.icp_batch_async(ids[:-1], ids[1:], init, ip)
        if pending is not None:                      # collect batch b - 1 while batch b runs
            r = engine.icp_batch_finish(pending[0])
            recs.append(r)
            for k in pending[1]:
                engine.free(base + k)
                resident.discard(k)
        pending = (ticket, list(range(lo, hi - 1)))
        lo = hi - 1
    if pending is not None:
        recs.append(engine.icp_batch_finish(pending[0]))
        for k in list(resident):
            engine.free(base + k)
    recs = np.concatenate(recs)
    return np.array(recs["T"]), recs


# ---------------------------------------------------------------------------------------------- back end loop
def run_backend(engine, scans, sm_rel, odo_rel, skip_loop_closing=50, skip_optimization=50, number_of_triplets_loop_closing=20,
                distance_backwards=7.0, radius_threshold=5.0, seed=0, quiet=True):
    """run_graphSLAM.py:229-267 with the drop-in LoopClosing / KeyFrameManager on `engine` and the stand-in pose graph.
    Returns the timing shares and counts."""
    t_lc = t_opt = t_host = 0.0
    n_lc_calls = n_lc_edges = 0
    out = io.StringIO()
    with dropin_modules(engine) as (kfm, lc):
        class MemoryKeyFrameManager(kfm.KeyFrameManager):
            """Keyframes come from host arrays instead of PCD files (5k files are not written for a timing run)."""

            gpu_s = 0.0

            def pre_process_many(self, indices):
                ta = time.perf_counter()
                super().pre_process_many(indices)
                MemoryKeyFrameManager.gpu_s += time.perf_counter() - ta

            def compute_transformations(self, pairs, Tijs):
                ta = time.perf_counter()
                out = super().compute_transformations(pairs, Tijs)
                MemoryKeyFrameManager.gpu_s += time.perf_counter() - ta
                MemoryKeyFrameManager.n_pairs = getattr(MemoryKeyFrameManager, "n_pairs", 0) + len(pairs)
                return out

            def load_pointcloud(self, i):
                kf = self.keyframes[i]
                if not kf._on_device:
                    kf.set_points(scans[i])
                self._resident.pop(i, None)
                self._resident[i] = True
                while len(self._resident) > self.max_resident_keyframes:
                    old = next(iter(self._resident))
                    self._resident.pop(old)
                    self.keyframes[old].unload_pointcloud()

        with (contextlib.redirect_stdout(out) if quiet else contextlib.nullcontext()):
            t0 = time.perf_counter()
            graph = PoseGraphStandIn()
            graph.init_graph()
            dassoc = lc.LoopClosing(graph, distance_backwards=distance_backwards, radius_threshold=radius_threshold)
            km = MemoryKeyFrameManager(directory="<memory>", scan_times=list(range(len(scans))), voxel_size=None, method="icppointplane")
            km.add_keyframes(keyframe_sampling=1)
            np.random.seed(seed)
            n = len(sm_rel)
            launches0 = engine.kernel_launches()
            for i in range(n):
                atb_sm, atb_odo = HomogeneousMatrix(sm_rel[i]), HomogeneousMatrix(odo_rel[i])
                graph.add_initial_estimate(atb_sm, i + 1)
                graph.add_edge(atb_sm, i, i + 1, 'SM')
                graph.add_edge(atb_odo, i, i + 1, 'ODO')
                if i % skip_optimization == 0:
                    ta = time.perf_counter()
                    graph.optimize()
                    t_opt += time.perf_counter() - ta
                if (i % skip_loop_closing) == 0 or (n - i) < 2:
                    ta = time.perf_counter()
                    added = dassoc.loop_closing_triangle(current_index=i, number_of_triplets_loop_closing=number_of_triplets_loop_closing,
                                                         keyframe_manager=km)
                    t_lc += time.perf_counter() - ta
                    n_lc_calls += 1
                    n_lc_edges += len(added) if added else 0
            ta = time.perf_counter()
            graph.optimize()
            t_opt += time.perf_counter() - ta
            total = time.perf_counter() - t0
            resident = len(km._resident)
            launches = engine.kernel_launches() - launches0
            for i in list(km._resident):
                km.unload_pointcloud(i)
    t_host = total - t_lc - t_opt
    gpu_s = MemoryKeyFrameManager.gpu_s
    return {"total_s": total, "loop_closing_s": t_lc, "loop_closing_engine_s": gpu_s, "loop_closing_candidate_search_s": t_lc - gpu_s,
            "loop_closing_pairs": getattr(MemoryKeyFrameManager, "n_pairs", 0),
            "optimize_s": t_opt, "other_host_s": t_host, "loop_closing_calls": n_lc_calls,
            "loop_closure_edges": n_lc_edges, "edges": len(graph.edges), "optimizations": graph.n_optimizations,
            "keyframes_resident_at_end": resident, "gpu_launches": int(launches)}
