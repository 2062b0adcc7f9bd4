"""Load path of the drop-in KeyFrame (reference: keyframemanager/keyframe.py:41-45, where
`o3d.io.read_point_cloud` reads one PCD file synchronously per `load_pointcloud()` call).

`ScanLoader` reads PCD files into page-locked buffers of the engine's pool, so that the host -> device copy that
follows is a true asynchronous copy on the engine's copy stream, and it reads AHEAD: while the caller registers the
pair (i, i + 1) - a synchronous call - one helper thread already reads and parses the file of scan i + 2.  The
reference's loop (run_scanmatcher.py:196-213) calls load_pointcloud(i + 2) next, which then only enqueues the copy.
With an engine that has no pinned pool (the CPU test double) the loader degrades to plain numpy arrays.

The read-ahead goes one step further when the caller registers one pair per call: `stage_ahead()` - called by the
registration between enqueueing its batch and waiting for it - uploads the scan that was read ahead under a fresh scan
id and preprocesses it on the engine's look-ahead stream (`arvc_scan_preprocess_ahead`), with the parameters of the
caller's last pre_process().  The keyframe that loads that file next adopts the staged scan: its own pre_process() then
finds the work done, and the GPU has spent the registration's latency-bound passes on something useful.
"""
import os
from concurrent.futures import ThreadPoolExecutor

from .pcd import read_pcd_xyz


class ScanLoader:
    def __init__(self, engine):
        self.engine = engine
        self.pool = getattr(engine, "pinned", None)
        self.executor = ThreadPoolExecutor(max_workers=1, thread_name_prefix="arvc-readahead")
        self.pending = {}                # filename -> Future of (array, pinned handle)
        self.staged = {}                 # filename -> (scan id, array, pinned handle): uploaded + preprocessed ahead
        self.ahead_params = None         # parameters of the caller's last plain pre_process(); None = do not stage
        self.stats = {"read_ahead_hits": 0, "reads": 0, "staged": 0, "staged_hits": 0}
        self.prewarmed = False
        self.prewarm_buffers = 8

    def _read(self, filename):
        handle = [None]

        def alloc(n_rows, dtype):
            arr, handle[0] = self.pool.empty(n_rows, dtype)
            return arr

        xyz = read_pcd_xyz(filename, alloc=alloc if self.pool is not None else None)
        if self.pool is not None and not self.prewarmed and handle[0] is not None:
            # cudaMallocHost waits for the GPU to go idle: allocating a staging buffer while a registration batch runs would
            # stall the very load it is meant to overlap.  The scans of a sequence have one size, so a handful of buffers of
            # that size are allocated now, with the first scan, and recycled from then on.
            self.prewarmed = True
            spare = [self.pool.empty(len(xyz) + len(xyz) // 8, xyz.dtype)[1] for _ in range(self.prewarm_buffers)]
            for h in spare:
                self.pool.release(h)
        return xyz, handle[0]

    def prefetch(self, filename):
        """Start reading `filename` in the background (no-op when it is already pending or does not exist)."""
        if filename in self.pending or not os.path.exists(filename):
            return
        if len(self.pending) >= 4:       # a caller that never comes back for its read-ahead must not pile up buffers
            old = next(iter(self.pending))
            self.release(self.pending.pop(old).result()[1])
        self.pending[filename] = self.executor.submit(self._read, filename)

    def stage_ahead(self, new_scan_id, wait_s=0.0005):
        """Upload + preprocess (look-ahead stream) what has been read ahead.  `new_scan_id`: allocator of engine scan ids."""
        if self.ahead_params is None or not self.pending or not hasattr(self.engine, "preprocess_ahead"):
            return
        for filename in list(self.pending):
            fut = self.pending[filename]
            try:
                xyz, handle = fut.result(timeout=wait_s)      # normally long done: it was started before the pre_process
            except Exception:
                continue                                      # still reading (or failed): the ordinary path will handle it
            del self.pending[filename]
            while len(self.staged) >= 2:                      # a caller that changed its mind must not pile up scans
                self._drop_staged(next(iter(self.staged)))
            sid = new_scan_id()
            self.engine.upload(sid, xyz)
            self.engine.preprocess_ahead([sid], self.ahead_params)
            self.staged[filename] = (sid, xyz, handle)
            self.stats["staged"] += 1

    def take_staged(self, filename):
        """(scan id, xyz, pinned handle) of a scan staged by stage_ahead(), or None; the caller owns all three afterwards."""
        st = self.staged.pop(filename, None)
        if st is not None:
            self.stats["reads"] += 1
            self.stats["read_ahead_hits"] += 1
            self.stats["staged_hits"] += 1
        return st

    def _drop_staged(self, filename):
        sid, _, handle = self.staged.pop(filename)
        if handle is not None:
            self.engine.wait_upload(sid)
        self.engine.free(sid)
        self.release(handle)

    def fetch(self, filename):
        """(xyz, pinned handle or None): the read-ahead result when there is one, else a synchronous read."""
        self.stats["reads"] += 1
        fut = self.pending.pop(filename, None)
        if fut is not None:
            self.stats["read_ahead_hits"] += 1
            return fut.result()
        return self._read(filename)

    def release(self, handle):
        if handle is not None and self.pool is not None:
            self.pool.release(handle)

    def close(self):
        for fut in self.pending.values():
            self.release(fut.result()[1])
        self.pending = {}
        for filename in list(self.staged):
            try:
                self._drop_staged(filename)
            except Exception:
                self.staged.pop(filename, None)
        self.executor.shutdown(wait=True)
