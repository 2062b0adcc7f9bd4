"""TEST DOUBLE (tests/ only): an object with the Engine interface whose arithmetic is the CPU oracle.

It exists so that the host-side logic of the drop-in KeyFrame / KeyFrameManager classes — and the reference's
UNMODIFIED drivers on top of them — can be exercised in the CPU-only container.  It is never importable from the
product package and is injected explicitly with lidar_slam_arvc_b200.runtime.set_engine().
"""
import numpy as np

from lidar_slam_arvc_b200.engine import P2PLANE, RESULT_DTYPE, Engine
from oracle import oracle as orc


class OracleEngine:
    make_preprocess_params = staticmethod(Engine.make_preprocess_params)
    make_icp_params = staticmethod(Engine.make_icp_params)

    def __init__(self):
        self.raw = {}
        self.pre = {}
        self.pre_params = {}
        self.calls = []

    def upload(self, scan_id, xyz):
        self.raw[int(scan_id)] = np.asarray(xyz)
        self.pre.pop(int(scan_id), None)
        self.pre_params.pop(int(scan_id), None)

    def free(self, scan_id):
        self.raw.pop(int(scan_id), None)
        self.pre.pop(int(scan_id), None)
        self.pre_params.pop(int(scan_id), None)

    def preprocess_ahead(self, scan_ids, p):
        """The look-ahead variant: same results; the double only records that it was used."""
        self.calls.append(("preprocess_ahead", len(np.atleast_1d(scan_ids))))
        self.preprocess(scan_ids, p, _count=False)

    def preprocess(self, scan_ids, p, _count=True):
        if _count:
            self.calls.append(("preprocess", len(np.atleast_1d(scan_ids))))
        for k in np.atleast_1d(scan_ids):
            if int(k) in self.pre and self.pre_params.get(int(k)) == bytes(p):
                continue                                  # like the engine: same scan, same parameters -> nothing to do
            self.pre_params[int(k)] = bytes(p)
            pts = self.raw[int(k)].astype(np.float64)
            d = pts[:, 0] ** 2 + pts[:, 1] ** 2
            with np.errstate(invalid="ignore"):
                keep = (d < p.max_radius2) & (d > p.min_radius2) & (pts[:, 2] > p.min_height) & (pts[:, 2] < p.max_height)
            pts = pts[keep]
            if p.voxel_size > 0:
                pts, _, _ = orc.voxel_down_sample(pts, p.voxel_size)
            nrm = orc.estimate_normals(pts, p.normal_radius, p.max_nn) if p.want_normals else None
            self.pre[int(k)] = (pts, nrm, int(keep.sum()))

    def info(self, scan_id):
        pts, nrm, nf = self.pre[int(scan_id)]
        return {"n_raw": len(self.raw[int(scan_id)]), "n_filtered": nf, "n_points": len(pts), "has_normals": nrm is not None}

    def get_points(self, scan_id, normals=False):
        pts, nrm, _ = self.pre[int(scan_id)]
        return (pts, nrm) if normals else pts

    def fit_plane(self, scan_id, max_z=-0.5, dist_threshold=0.01, iterations=1000, seed=0):
        self.calls.append(("fit_plane", int(scan_id)))
        return orc.fit_plane(self.pre[int(scan_id)][0], max_z, dist_threshold, iterations, seed)

    def split_plane(self, src_id, plane_model, threshold, near_id, far_id):
        self.calls.append(("split_plane", int(src_id)))
        pts = self.pre[int(src_id)][0]
        near, far = orc.segment_plane(pts, plane_model, threshold)
        self.upload(near_id, pts[near])
        self.upload(far_id, pts[far])
        return len(near), len(far)

    def map_build(self, scan_ids, transforms, p):
        self.calls.append(("map_build", len(scan_ids)))
        self.preprocess(scan_ids, p)
        parts = [orc.transform_points(self.pre[int(k)][0], T) for k, T in zip(scan_ids, np.asarray(transforms).reshape(-1, 4, 4))]
        offsets = np.concatenate([[0], np.cumsum([len(x) for x in parts])]).astype(np.int64)
        return (np.concatenate(parts) if parts else np.zeros((0, 3))), offsets

    def upload_ptr(self, scan_id, ptr, n):
        raise NotImplementedError("the test double takes arrays")

    def kernel_launches(self):
        return 0

    def icp_batch_async(self, tgt_ids, src_ids, init_T, p):
        return self.icp_batch(tgt_ids, src_ids, init_T, p)

    def icp_batch_finish(self, ticket):
        return ticket

    def icp_batch(self, tgt_ids, src_ids, init_T, p):
        self.calls.append(("icp_batch", len(tgt_ids)))
        init_T = np.asarray(init_T, dtype=np.float64).reshape(-1, 4, 4)
        out = np.zeros(len(tgt_ids), dtype=RESULT_DTYPE)
        for k, (t, s) in enumerate(zip(tgt_ids, src_ids)):
            tp, tn, _ = self.pre[int(t)]
            sp, _, _ = self.pre[int(s)]
            r = orc.icp(sp, tp, tn if p.method == P2PLANE else None, init_T[k], p.method, p.max_corr_dist, p.rel_fitness,
                        p.rel_rmse, p.max_iter)
            out[k] = (k, r.updates, r.n_corr, r.passes, r.transformation, r.fitness, r.inlier_rmse)
        return out
