"""ctypes binding of libarvc_icp.so (C-ABI in include/arvc_icp.h).

This is the only module that talks to the native library.  There is no CPU fallback: if the library is
missing or no CUDA device is present, construction raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ARVC_LIB_VARIANT") or os.path.join(_HERE, "libarvc_icp.so")   # the variable is for developer A/B builds only

P2P, P2PLANE = 0, 1

# every symbol include/arvc_icp.h declares (checked by tests/test_abi.py without a GPU)
EXPORTED_SYMBOLS = [
    "arvc_ctx_create", "arvc_ctx_destroy", "arvc_last_error", "arvc_sync", "arvc_stream", "arvc_version",
    "arvc_kernel_launches", "arvc_scan_upload_f32", "arvc_scan_upload_f64", "arvc_scan_free", "arvc_scan_preprocess", "arvc_scan_preprocess_ahead",
    "arvc_scan_info", "arvc_scan_get_points", "arvc_scan_get_filter_indices", "arvc_scan_get_voxels",
    "arvc_scan_get_nn_counts", "arvc_icp_batch", "arvc_icp_batch_async", "arvc_icp_batch_finish", "arvc_icp_trace",
    "arvc_host_alloc", "arvc_host_free", "arvc_profile_enable", "arvc_profile_report", "arvc_scan_invalidate", "arvc_lzf_decompress",
    "arvc_map_build", "arvc_scan_fit_plane", "arvc_scan_split_plane", "arvc_ctx_set_option", "arvc_scan_get_counters",
    "arvc_icp_batch_device_records", "arvc_scan_wait_upload", "arvc_ctx_host_alloc", "arvc_ctx_reserve", "arvc_scan_get_neighbors",
]


class PreprocessParams(ctypes.Structure):
    _fields_ = [("min_radius2", ctypes.c_double), ("max_radius2", ctypes.c_double), ("min_height", ctypes.c_double),
                ("max_height", ctypes.c_double), ("voxel_size", ctypes.c_double), ("normal_radius", ctypes.c_double),
                ("max_nn", ctypes.c_int32), ("want_normals", ctypes.c_int32), ("grid_cell", ctypes.c_double),
                ("grid_max_dist", ctypes.c_double)]


class IcpParams(ctypes.Structure):
    _fields_ = [("max_corr_dist", ctypes.c_double), ("rel_fitness", ctypes.c_double), ("rel_rmse", ctypes.c_double),
                ("max_iter", ctypes.c_int32), ("method", ctypes.c_int32)]


class ResultRecord(ctypes.Structure):
    _fields_ = [("pair", ctypes.c_int32), ("updates", ctypes.c_int32), ("n_corr", ctypes.c_int32), ("passes", ctypes.c_int32),
                ("T", ctypes.c_double * 16), ("fitness", ctypes.c_double), ("rmse", ctypes.c_double)]


RESULT_DTYPE = np.dtype([("pair", "<i4"), ("updates", "<i4"), ("n_corr", "<i4"), ("passes", "<i4"), ("T", "<f8", (4, 4)),
                         ("fitness", "<f8"), ("rmse", "<f8")])
assert RESULT_DTYPE.itemsize == ctypes.sizeof(ResultRecord) == 160

_lib = None


def load_library():
    """dlopen the engine.  Raises (loudly) when it has not been built: `python -c 'import __graft_entry__ as g; g.build()'`."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("native CUDA library %s is missing; build it with __graft_entry__.build() "
                           "(make -C lidar_slam_arvc_b200/csrc). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, i64p, dp, ip, fp = c.c_void_p, c.POINTER(c.c_int64), c.POINTER(c.c_double), c.POINTER(c.c_int32), c.POINTER(c.c_float)
    lib.arvc_ctx_create.argtypes = [c.c_int, c.POINTER(vp)]
    lib.arvc_ctx_destroy.argtypes = [vp]
    lib.arvc_ctx_destroy.restype = None
    lib.arvc_last_error.argtypes = [vp]
    lib.arvc_last_error.restype = c.c_char_p
    lib.arvc_sync.argtypes = [vp]
    lib.arvc_stream.argtypes = [vp]
    lib.arvc_stream.restype = vp
    lib.arvc_kernel_launches.argtypes = [vp]
    lib.arvc_kernel_launches.restype = c.c_int64
    lib.arvc_scan_upload_f32.argtypes = [vp, c.c_int64, vp, c.c_int]
    lib.arvc_scan_upload_f64.argtypes = [vp, c.c_int64, vp, c.c_int]
    lib.arvc_scan_free.argtypes = [vp, c.c_int64]
    lib.arvc_scan_wait_upload.argtypes = [vp, c.c_int64]
    lib.arvc_scan_preprocess.argtypes = [vp, c.c_int, i64p, c.POINTER(PreprocessParams)]
    lib.arvc_scan_preprocess_ahead.argtypes = [vp, c.c_int, i64p, c.POINTER(PreprocessParams)]
    lib.arvc_scan_info.argtypes = [vp, c.c_int64, ip, ip, ip, ip]
    lib.arvc_scan_get_points.argtypes = [vp, c.c_int64, dp, dp]
    lib.arvc_scan_get_filter_indices.argtypes = [vp, c.c_int64, ip]
    lib.arvc_scan_get_voxels.argtypes = [vp, c.c_int64, ip, ip]
    lib.arvc_scan_get_nn_counts.argtypes = [vp, c.c_int64, ip]
    lib.arvc_scan_get_counters.argtypes = [vp, c.c_int64, ip]
    lib.arvc_scan_get_neighbors.argtypes = [vp, c.c_int64, c.c_int, ip, ip, ip]
    lib.arvc_ctx_set_option.argtypes = [vp, c.c_char_p, c.c_int]
    lib.arvc_icp_batch.argtypes = [vp, c.c_int, i64p, i64p, dp, c.POINTER(IcpParams), dp, dp, dp, ip, ip]
    lib.arvc_icp_batch_async.argtypes = [vp, c.c_int, i64p, i64p, dp, c.POINTER(IcpParams), c.POINTER(c.c_uint64)]
    lib.arvc_icp_batch_finish.argtypes = [vp, c.c_uint64, vp]
    lib.arvc_icp_batch_device_records.argtypes = [vp, c.c_uint64, c.POINTER(vp), ip]
    lib.arvc_icp_trace.argtypes = [vp, c.c_int64, c.c_int64, dp, c.POINTER(IcpParams), ip, dp, dp, dp, ip, c.POINTER(ResultRecord)]
    lib.arvc_map_build.argtypes = [vp, c.c_int, i64p, dp, c.POINTER(PreprocessParams), dp, c.c_int64, i64p]
    lib.arvc_scan_fit_plane.argtypes = [vp, c.c_int64, c.c_double, c.c_double, c.c_int, c.c_uint64, dp, ip]
    lib.arvc_scan_split_plane.argtypes = [vp, c.c_int64, dp, c.c_double, c.c_int64, c.c_int64, ip, ip]
    lib.arvc_profile_enable.argtypes = [vp, c.c_int]
    lib.arvc_profile_report.argtypes = [vp, c.c_char_p, c.c_size_t]
    lib.arvc_scan_invalidate.argtypes = [vp, c.c_int64]
    lib.arvc_lzf_decompress.argtypes = [c.c_char_p, c.c_size_t, vp, c.c_size_t]
    lib.arvc_lzf_decompress.restype = c.c_longlong
    lib.arvc_host_alloc.argtypes = [c.c_size_t]
    lib.arvc_host_alloc.restype = vp
    lib.arvc_ctx_host_alloc.argtypes = [vp, c.c_size_t]
    lib.arvc_ctx_host_alloc.restype = vp
    lib.arvc_ctx_reserve.argtypes = [vp, c.c_size_t]
    lib.arvc_host_free.argtypes = [vp]
    lib.arvc_host_free.restype = None
    _lib = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def _i64p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


class EngineError(RuntimeError):
    pass


def lzf_decompress(data, out_size):
    """LZF-decompress `data` (bytes) into a new bytes object of exactly `out_size` bytes (host helper, no GPU needed)."""
    lib = load_library()
    out = ctypes.create_string_buffer(max(int(out_size), 1))
    n = lib.arvc_lzf_decompress(bytes(data), len(data), ctypes.cast(out, ctypes.c_void_p), int(out_size))
    if n != out_size:
        raise ValueError("malformed LZF stream (got %d of %d bytes)" % (n, out_size))
    return out.raw[:out_size]


class PinnedPool:
    """Pooled page-locked host buffers (arvc_host_alloc) handed out as numpy arrays: the load path parses a PCD file
    straight into one, so that the upload is a true asynchronous copy (keyframe.py:41-45 replacement).
    Thread-safe: the read-ahead thread of the drop-in allocates, the main thread releases."""

    def __init__(self, engine):
        import threading
        self.lib = load_library()
        self.ctx = engine.h
        self.free = []                 # (capacity, pointer)
        self.live = {}                 # pointer -> capacity
        self.lock = threading.Lock()

    def empty(self, n_rows, dtype=np.float32):
        """An uninitialised [n_rows, 3] array in pinned memory."""
        dtype = np.dtype(dtype)
        nbytes = max(int(n_rows) * 3 * dtype.itemsize, 1)
        with self.lock:
            best = None
            for k, (cap, ptr) in enumerate(self.free):
                if cap >= nbytes and (best is None or cap < self.free[best][0]):
                    best = k
            if best is not None:
                cap, ptr = self.free.pop(best)
            else:
                cap = 1 << 16
                while cap < nbytes:
                    cap <<= 1
                ptr = self.lib.arvc_ctx_host_alloc(self.ctx, cap)
                if not ptr:
                    raise EngineError("arvc_host_alloc(%d) failed" % cap)
            self.live[ptr] = cap
        buf = (ctypes.c_char * cap).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=int(n_rows) * 3).reshape(int(n_rows), 3)
        return arr, ptr

    def release(self, ptr):
        with self.lock:
            cap = self.live.pop(ptr, None)
            if cap is not None:
                self.free.append((cap, ptr))

    def close(self):
        with self.lock:
            for cap, ptr in self.free:
                self.lib.arvc_host_free(ptr)
            self.free = []


class Engine:
    """One context = one GPU + one stream.  Not thread-safe (calls are serialised by the caller, like the reference)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = ctypes.c_void_p()
        rc = self.lib.arvc_ctx_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise EngineError("arvc_ctx_create failed (%d): %s" % (rc, self.lib.arvc_last_error(None).decode()))
        self.h = h
        self.device = int(device)
        self._n_raw = {}                  # raw size of every uploaded scan (output capacity of map_build)
        self._pinned = None

    @property
    def pinned(self):
        """The context's pool of page-locked staging buffers (created on first use)."""
        if self._pinned is None:
            self._pinned = PinnedPool(self)
        return self._pinned

    def close(self):
        if getattr(self, "h", None):
            self.lib.arvc_ctx_destroy(self.h)      # synchronises: no copy from a staging buffer is in flight afterwards
            if self._pinned is not None:
                self._pinned.close()
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise EngineError("engine error %d: %s" % (rc, self.lib.arvc_last_error(self.h).decode()))

    # ---- scans
    def upload(self, scan_id, xyz):
        """xyz: [n,3] float32 (PCD payload) or float64.  A pinned buffer must stay alive until sync()."""
        a = np.asarray(xyz)
        if a.dtype == np.float64:
            a = np.ascontiguousarray(a)
            self._ck(self.lib.arvc_scan_upload_f64(self.h, int(scan_id), a.ctypes.data, len(a)))
            self._n_raw[int(scan_id)] = len(a)
        else:
            a = np.ascontiguousarray(a, dtype=np.float32)
            self._ck(self.lib.arvc_scan_upload_f32(self.h, int(scan_id), a.ctypes.data, len(a)))
            self._n_raw[int(scan_id)] = len(a)
        return a

    def upload_ptr(self, scan_id, ptr, n):
        self._ck(self.lib.arvc_scan_upload_f32(self.h, int(scan_id), ctypes.c_void_p(ptr), int(n)))
        self._n_raw[int(scan_id)] = int(n)

    def wait_upload(self, scan_id):
        """Host wait for the copy of one scan (its pinned staging buffer may be reused afterwards)."""
        self._ck(self.lib.arvc_scan_wait_upload(self.h, int(scan_id)))

    def free(self, scan_id):
        self._ck(self.lib.arvc_scan_free(self.h, int(scan_id)))
        self._n_raw.pop(int(scan_id), None)

    @staticmethod
    def make_preprocess_params(min_radius=0.5, max_radius=35, min_height=-1.0, max_height=50.0, voxel_size=None,
                               normal_radius=0.3, max_nn=300, want_normals=True, grid_cell=0.0, grid_max_dist=10.0):
        # `radius ** 2` evaluated by Python exactly as keyframe.py:92 does
        return PreprocessParams(float(min_radius ** 2), float(max_radius ** 2), float(min_height), float(max_height),
                                float("nan") if voxel_size is None else float(voxel_size), float(normal_radius), int(max_nn),
                                1 if want_normals else 0, float(grid_cell), float(grid_max_dist))

    def preprocess(self, scan_ids, params):
        ids = np.ascontiguousarray(scan_ids, dtype=np.int64).reshape(-1)
        self._ck(self.lib.arvc_scan_preprocess(self.h, len(ids), _i64p(ids), ctypes.byref(params)))

    def reserve(self, n_bytes):
        """Grow the device memory pool now (arvc_ctx_reserve) instead of piecemeal while kernels run."""
        self._ck(self.lib.arvc_ctx_reserve(self.h, ctypes.c_size_t(int(n_bytes))))

    def preprocess_ahead(self, scan_ids, params):
        """preprocess() of scans needed next, on the engine's look-ahead stream (overlaps a registration in flight)."""
        ids = np.ascontiguousarray(scan_ids, dtype=np.int64).reshape(-1)
        self._ck(self.lib.arvc_scan_preprocess_ahead(self.h, len(ids), _i64p(ids), ctypes.byref(params)))

    def info(self, scan_id):
        v = [ctypes.c_int32() for _ in range(4)]
        self._ck(self.lib.arvc_scan_info(self.h, int(scan_id), *[ctypes.byref(x) for x in v]))
        return {"n_raw": v[0].value, "n_filtered": v[1].value, "n_points": v[2].value, "has_normals": bool(v[3].value)}

    def get_points(self, scan_id, normals=False):
        n = self.info(scan_id)["n_points"]
        xyz = np.empty((n, 3))
        nrm = np.empty((n, 3)) if normals else None
        self._ck(self.lib.arvc_scan_get_points(self.h, int(scan_id), _dp(xyz), _dp(nrm) if normals else None))
        return (xyz, nrm) if normals else xyz

    def map_build(self, scan_ids, transforms, params):
        """KeyFrameManager.build_map (keyframemanager.py:154-184) for a batch: every scan filtered / down-sampled with
        `params`, moved by its 4x4 and concatenated in order.  Returns (xyz [total,3] float64, offsets [n+1] int64)."""
        ids = np.ascontiguousarray(scan_ids, dtype=np.int64).reshape(-1)
        T = np.ascontiguousarray(transforms, dtype=np.float64).reshape(-1, 4, 4)
        if len(T) != len(ids):
            raise ValueError("map_build: one 4x4 transform per scan")
        offsets = np.zeros(len(ids) + 1, dtype=np.int64)
        cap = int(sum(self._n_raw.get(int(k), 0) for k in ids))         # the filter only removes points
        xyz = np.empty((max(cap, 1), 3))
        self._ck(self.lib.arvc_map_build(self.h, len(ids), _i64p(ids), _dp(T), ctypes.byref(params), _dp(xyz), cap, _i64p(offsets)))
        return xyz[:int(offsets[-1])], offsets

    def fit_plane(self, scan_id, max_z=-0.5, dist_threshold=0.01, iterations=1000, seed=0):
        """KeyFrame.calculate_plane (keyframe.py:417-436) on the preprocessed cloud: ([a, b, c, d], inliers)."""
        pl = np.zeros(4)
        n_in = ctypes.c_int32()
        self._ck(self.lib.arvc_scan_fit_plane(self.h, int(scan_id), float(max_z), float(dist_threshold), int(iterations), int(seed),
                                              _dp(pl), ctypes.byref(n_in)))
        return pl, n_in.value

    def split_plane(self, src_id, plane_model, threshold, near_id, far_id):
        """KeyFrame.segment_plane (keyframe.py:438-461): two new scans (near the plane / the rest), returns their sizes."""
        pl = np.ascontiguousarray(plane_model, dtype=np.float64).reshape(4)
        a, b = ctypes.c_int32(), ctypes.c_int32()
        self._ck(self.lib.arvc_scan_split_plane(self.h, int(src_id), _dp(pl), float(threshold), int(near_id), int(far_id),
                                                ctypes.byref(a), ctypes.byref(b)))
        self._n_raw[int(near_id)], self._n_raw[int(far_id)] = a.value, b.value
        return a.value, b.value

    def get_filter_indices(self, scan_id):
        n = self.info(scan_id)["n_filtered"]
        idx = np.empty(max(n, 1), dtype=np.int32)
        self._ck(self.lib.arvc_scan_get_filter_indices(self.h, int(scan_id), _ip(idx)))
        return idx[:n]

    def get_voxels(self, scan_id):
        n = self.info(scan_id)["n_points"]
        keys = np.empty((max(n, 1), 3), dtype=np.int32)
        cnt = np.empty(max(n, 1), dtype=np.int32)
        self._ck(self.lib.arvc_scan_get_voxels(self.h, int(scan_id), _ip(keys), _ip(cnt)))
        return keys[:n], cnt[:n]

    def get_nn_counts(self, scan_id):
        n = self.info(scan_id)["n_points"]
        cnt = np.empty(max(n, 1), dtype=np.int32)
        self._ck(self.lib.arvc_scan_get_nn_counts(self.h, int(scan_id), _ip(cnt)))
        return cnt[:n]

    def get_neighbors(self, scan_id, point_ids, max_nn):
        """Neighbour index sets of the normals of `point_ids` (cloud order): list of sorted int arrays.  Needs
        set_option("normals_tap", 1) before the scan is preprocessed."""
        ids = np.ascontiguousarray(point_ids, dtype=np.int32).reshape(-1)
        out = np.full((max(len(ids), 1), int(max_nn)), -1, dtype=np.int32)
        cnt = np.zeros(max(len(ids), 1), dtype=np.int32)
        self._ck(self.lib.arvc_scan_get_neighbors(self.h, int(scan_id), len(ids), _ip(ids), _ip(out), _ip(cnt)))
        return [np.sort(out[k, :cnt[k]]) for k in range(len(ids))]

    COUNTER_NAMES = ("n_filtered", "n_points", "error_flags", "grid_cells", "normals_redone", "normals_per_point",
                     "normals_blocks_handed_back", "normals_points_handed_back", "normals_trial_blocks", "dbg_tile_records",
                     "dbg_neighbours", "dbg_points_one_sweep", "dbg_points_sweep_then_select", "dbg_points_trial", "dbg_records_streamed", "dbg_blocks_too_wide")

    def get_counters(self, scan_id):
        """Device counters of a preprocessed scan (see arvc_scan_get_counters) as a dict."""
        c = np.zeros(16, dtype=np.int32)
        self._ck(self.lib.arvc_scan_get_counters(self.h, int(scan_id), _ip(c)))
        return {k: int(v) for k, v in zip(self.COUNTER_NAMES, c)}

    def set_option(self, name, value):
        """Engine switch (arvc_ctx_set_option), e.g. ("icp_loop_graph", 0) for the unconditional max_iter + 1 enqueue."""
        self._ck(self.lib.arvc_ctx_set_option(self.h, name.encode(), int(value)))

    # ---- registration
    @staticmethod
    def make_icp_params(method=P2PLANE, max_corr_dist=10.0, rel_fitness=1e-6, rel_rmse=1e-6, max_iter=30):
        return IcpParams(float(max_corr_dist), float(rel_fitness), float(rel_rmse), int(max_iter), int(method))

    def icp_batch_async(self, tgt_ids, src_ids, init_T, params):
        t = np.ascontiguousarray(tgt_ids, dtype=np.int64).reshape(-1)
        s = np.ascontiguousarray(src_ids, dtype=np.int64).reshape(-1)
        T = np.ascontiguousarray(init_T, dtype=np.float64).reshape(-1, 16)
        assert len(t) == len(s) == len(T)
        ticket = ctypes.c_uint64()
        self._ck(self.lib.arvc_icp_batch_async(self.h, len(t), _i64p(t), _i64p(s), _dp(T), ctypes.byref(params), ctypes.byref(ticket)))
        return ticket.value, len(t)

    def icp_batch_device_records(self, ticket):
        """(device address, n_pairs) of a pending batch's 160-byte records; valid until icp_batch_finish(ticket)."""
        tk, n = ticket
        ptr, cnt = ctypes.c_void_p(), ctypes.c_int32()
        self._ck(self.lib.arvc_icp_batch_device_records(self.h, ctypes.c_uint64(tk), ctypes.byref(ptr), ctypes.byref(cnt)))
        return ptr.value or 0, cnt.value

    def icp_batch_finish(self, ticket):
        tk, n = ticket
        rec = np.zeros(max(n, 1), dtype=RESULT_DTYPE)
        self._ck(self.lib.arvc_icp_batch_finish(self.h, ctypes.c_uint64(tk), rec.ctypes.data))
        return rec[:n]

    def icp_batch(self, tgt_ids, src_ids, init_T, params):
        """Returns a structured array (RESULT_DTYPE): T[4,4], fitness, rmse, updates, n_corr, passes per pair."""
        return self.icp_batch_finish(self.icp_batch_async(tgt_ids, src_ids, init_T, params))

    def icp_trace(self, tgt_id, src_id, init_T, params):
        ns = self.info(src_id)["n_points"]
        passes = params.max_iter + 1
        corr = np.full((passes, max(ns, 1)), -2, dtype=np.int32)
        trT = np.zeros((passes, 4, 4))
        trf = np.zeros(passes)
        trr = np.zeros(passes)
        npass = ctypes.c_int32()
        rec = ResultRecord()
        T = np.ascontiguousarray(init_T, dtype=np.float64).reshape(16)
        # the C side packs rows with stride n_src
        flat = np.full(passes * max(ns, 1), -2, dtype=np.int32)
        self._ck(self.lib.arvc_icp_trace(self.h, int(tgt_id), int(src_id), _dp(T), ctypes.byref(params), _ip(flat), _dp(trT),
                                         _dp(trf), _dp(trr), ctypes.byref(npass), ctypes.byref(rec)))
        k = npass.value
        corr = flat[:k * ns].reshape(k, ns) if ns > 0 else np.zeros((k, 0), dtype=np.int32)
        out = np.zeros(1, dtype=RESULT_DTYPE)
        ctypes.memmove(out.ctypes.data, ctypes.byref(rec), 160)
        return {"corr": corr, "T": trT[:k], "fitness": trf[:k], "rmse": trr[:k], "passes": k, "result": out[0]}

    def invalidate(self, scan_ids):
        for k in np.atleast_1d(scan_ids):
            self._ck(self.lib.arvc_scan_invalidate(self.h, int(k)))

    def profile_enable(self, on=True):
        self._ck(self.lib.arvc_profile_enable(self.h, 1 if on else 0))

    def profile_report(self):
        """{kernel name: (launches, total_ms)} since profile_enable(); synchronises."""
        buf = ctypes.create_string_buffer(1 << 16)
        self._ck(self.lib.arvc_profile_report(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.split(",")
            out[name] = (int(cnt), float(ms))
        return out

    def sync(self):
        self._ck(self.lib.arvc_sync(self.h))

    def kernel_launches(self):
        return int(self.lib.arvc_kernel_launches(self.h))

    def stream_handle(self):
        return self.lib.arvc_stream(self.h)
