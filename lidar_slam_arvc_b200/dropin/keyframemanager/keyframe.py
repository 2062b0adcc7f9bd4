"""KeyFrame: one LiDAR scan, its preprocessing and pairwise registration, backed by the CUDA engine.

Mirror of the reference's keyframemanager/keyframe.py public surface for the methods the scan-matching hot
path uses (`icppointpoint`, `icppointplane`); the Open3D calls are replaced by calls through the C-ABI:
    load_pointcloud            keyframe.py:41-45     PCD payload -> device (arvc_scan_upload_*)
    filter_radius_height       keyframe.py:74-94     )
    pre_process / preprocess_* keyframe.py:113-162   ) arvc_scan_preprocess
    local_registration_simple  keyframe.py:231-260   arvc_icp_batch (one pair)
    unload_pointcloud          keyframe.py:61-72     arvc_scan_free
    preprocess_icp2planes / calculate_plane / segment_plane / local_registration_two_planes
                               keyframe.py:164-189, 417-461, 262-295   arvc_scan_fit_plane / _split_plane + one batch of two pairs
The reference's `fpfh` method (Open3D feature matching + RANSAC global registration) is out of scope:
it raises NotImplementedError instead of silently doing something else.
"""
import numpy as np

from config import ICP_PARAMETERS
from lidar_slam_arvc_b200 import runtime
from lidar_slam_arvc_b200.engine import P2P, P2PLANE, Engine
from lidar_slam_arvc_b200.homogeneousmatrix import result_type


class PointCloud:
    """Host view of a cloud (`np.asarray(pcd.points)`, `.normals`), standing in for o3d.geometry.PointCloud."""

    def __init__(self, points=None, normals=None):
        self.points = np.zeros((0, 3)) if points is None else points
        self.normals = normals

    def has_normals(self):
        return self.normals is not None

    def transform(self, T):
        T = np.asarray(T, dtype=np.float64)
        self.points = np.asarray(self.points, dtype=np.float64) @ T[:3, :3].T + T[:3, 3]
        if self.normals is not None:
            self.normals = self.normals @ T[:3, :3].T
        return self

    def __add__(self, other):
        return PointCloud(np.vstack([np.asarray(self.points, dtype=np.float64), np.asarray(other.points, dtype=np.float64)]))

    def __len__(self):
        return len(self.points)


def merge_two_planes(T_ground, T_non_ground):
    """keyframe.py:282-292: t2v(n=3) of both results; tx, ty, gamma from the non-ground one, tz, alpha, beta from the ground."""
    H = result_type()
    t1 = H(np.asarray(T_ground)).t2v(n=3)
    t2 = H(np.asarray(T_non_ground)).t2v(n=3)
    return H(np.array([t2[0], t2[1], t1[2]]), [float(t1[3]), float(t1[4]), float(t2[5])])


class KeyFrame():
    def __init__(self, directory, scan_time, voxel_size):
        self.directory = directory
        self.scan_time = scan_time
        self.voxel_size = voxel_size
        self.fpfh_threshold = 5
        self.pointcloud = None
        self.pointcloud_fpfh = None
        self.voxel_size_normals_ground_plane = 0.5
        self.voxel_size_normals = 0.3            # the radius actually used for normals (keyframe.py:33)
        self.max_radius = ICP_PARAMETERS.max_radius
        self.min_radius = ICP_PARAMETERS.min_radius
        self.max_height = ICP_PARAMETERS.max_height
        self.min_height = ICP_PARAMETERS.min_height
        self.plane_model = None                  # the ground plane of the last icp2planes preprocessing (keyframe.py:173)
        self.fixed_plane_model = None            # extension: a caller-supplied model (cf. the fixed [0, 0, 1, 0.69] of keyframe.py:436)
        self.pre_processed = False               # never set by the reference either (keyframe.py:39,114)
        self.last_result = None                  # extension: fitness / inlier_rmse / iterations of the last registration
        self._scan_id = runtime.new_scan_id()
        self._pinned_handle = None
        self._loaded_from = None
        self._filtered_cache = None
        self._on_device = False
        self._preprocessed_on_device = False
        # 'icp2planes': the ground / non-ground parts live on the device as two more scans
        self._scan_id_ground = runtime.new_scan_id()
        self._scan_id_non_ground = runtime.new_scan_id()
        self._planes_on_device = False
        self._plane_clouds = [None, None]
        self.plane_seed = 0                      # extension: seed of the reproducible RANSAC (Open3D's is unseeded)

    # ------------------------------------------------------------------ load / unload
    def pcd_filename(self):
        return self.directory + '/robot0/lidar/data/' + str(self.scan_time) + '.pcd'

    def load_pointcloud(self):
        """keyframe.py:41-45.  The file is parsed into a page-locked buffer (by the read-ahead thread when the previous
        load announced it) and the upload is only enqueued: the call returns while the copy runs on the copy stream.
        A keyframe that is still resident is not read again - the reference re-reads the same file on every call
        (loopclosing.py:163,177), which yields the same points - so loop closing finds its keyframes preprocessed."""
        filename = self.pcd_filename()
        print('Reading pointcloud: ', filename)
        if self._on_device and self._loaded_from == filename:
            return
        staged = runtime.get_loader().take_staged(filename)
        if staged is not None:
            # uploaded and preprocessed ahead, while the previous pair was being registered: adopt that device scan
            self._release_staging()
            if self._on_device:
                runtime.get_engine().free(self._scan_id)
            self._scan_id, xyz, self._pinned_handle = staged
            self.pointcloud = PointCloud(xyz)
            self._on_device = True
            self._preprocessed_on_device = False
            self._filtered_cache = None
        else:
            xyz, handle = runtime.get_loader().fetch(filename)
            self.set_points(xyz, _pinned_handle=handle)
        self._loaded_from = filename

    def set_points(self, xyz, _pinned_handle=None):
        """Extension: hand the PCD payload over directly ([n,3] float32 or float64)."""
        self._release_staging()
        self.pointcloud = PointCloud(xyz)
        runtime.get_engine().upload(self._scan_id, xyz)
        self._pinned_handle = _pinned_handle
        self._loaded_from = None
        self._on_device = True
        self._preprocessed_on_device = False
        self._filtered_cache = None

    def _release_staging(self):
        """Give the page-locked staging buffer back to the pool - after its copy has finished, and only then."""
        if getattr(self, "_pinned_handle", None) is not None:
            if self._on_device:
                runtime.get_engine().wait_upload(self._scan_id)
            runtime.get_loader().release(self._pinned_handle)
        self._pinned_handle = None

    def unload_pointcloud(self):
        print('Removing pointclouds from memory (filtered, planes, fpfh): ')
        self._release_staging()
        self._loaded_from = None
        if self._on_device:
            runtime.get_engine().free(self._scan_id)
        self._on_device = False
        self._preprocessed_on_device = False
        self._filtered_cache = None
        self.pointcloud = None
        self._drop_planes()
        self.pointcloud_fpfh = None

    def _drop_planes(self):
        if self._planes_on_device:
            runtime.get_engine().free(self._scan_id_ground)
            runtime.get_engine().free(self._scan_id_non_ground)
        self._planes_on_device = False
        self._plane_clouds = [None, None]

    def _plane_cloud(self, which):
        if not self._planes_on_device:
            return None
        if self._plane_clouds[which] is None:
            pts, nrm = runtime.get_engine().get_points((self._scan_id_ground, self._scan_id_non_ground)[which], normals=True)
            self._plane_clouds[which] = PointCloud(pts, nrm)
        return self._plane_clouds[which]

    @property
    def pointcloud_ground_plane(self):
        return self._plane_cloud(0)

    @property
    def pointcloud_non_ground_plane(self):
        return self._plane_cloud(1)

    # ------------------------------------------------------------------ preprocessing
    def _params(self, want_normals, radii=None, heights=None, voxel=True):
        min_radius, max_radius = (self.min_radius, self.max_radius) if radii is None else (radii[0], radii[1])
        min_height, max_height = (self.min_height, self.max_height) if heights is None else (heights[0], heights[1])
        return Engine.make_preprocess_params(min_radius, max_radius, min_height, max_height,
                                             self.voxel_size if voxel else None, self.voxel_size_normals,
                                             ICP_PARAMETERS.max_nn, want_normals,
                                             grid_max_dist=ICP_PARAMETERS.distance_threshold)

    def _require_loaded(self):
        if not self._on_device:
            raise RuntimeError("KeyFrame %s: load_pointcloud() first" % str(self.scan_time))

    def _preprocess(self, params):
        self._require_loaded()
        runtime.get_engine().preprocess([self._scan_id], params)
        self._last_params = params
        self._preprocessed_on_device = True
        self._filtered_cache = None

    @property
    def pointcloud_filtered(self):
        """np.asarray(kf.pointcloud_filtered.points) / .normals, downloaded on demand in the reference's point order."""
        if not self._preprocessed_on_device:
            return None
        if self._filtered_cache is None:
            eng = runtime.get_engine()
            if eng.info(self._scan_id)["has_normals"]:
                pts, nrm = eng.get_points(self._scan_id, normals=True)
            else:
                pts, nrm = eng.get_points(self._scan_id), None
            self._filtered_cache = PointCloud(pts, nrm)
        return self._filtered_cache

    @pointcloud_filtered.setter
    def pointcloud_filtered(self, value):
        self._filtered_cache = value

    def filter_radius_height(self, radii=None, heights=None):
        self._filter_bounds = (radii, heights)
        self._preprocess(self._params(False, radii, heights, voxel=False))
        return self.pointcloud_filtered

    def down_sample(self):
        """Voxel down-sampling of the cloud the last filter_radius_height() produced (keyframe.py:108-111)."""
        if self.voxel_size is None:
            return
        radii, heights = getattr(self, "_filter_bounds", (None, None))
        self._preprocess(self._params(False, radii, heights))

    def pre_process(self, method=False):
        if self.pre_processed:
            print('Already preprocessed, exiting')
            return
        if method == 'icppointpoint':
            self.preprocess_icp_point_point()
        elif method == 'icppointplane':
            self.preprocess_icp_point_plane()
        elif method == 'icp2planes':
            self.preprocess_icp2planes()
        elif method == 'fpfh':
            raise NotImplementedError("method 'fpfh' is outside the B200 hot path (Open3D global registration)")

    def preprocess_icp_point_point(self):
        self._preprocess(self._params(False))
        runtime.get_loader().ahead_params = self._last_params      # what a scan read ahead will be preprocessed with

    def preprocess_icp_point_plane(self):
        self._preprocess(self._params(True))
        runtime.get_loader().ahead_params = self._last_params

    def preprocess_icp2planes(self):
        """keyframe.py:164-189: filter -> [voxel] -> normals, ground-plane model, split into the points within 0.4 m of
        the plane and the rest, normals of both parts (radius 0.5 / max_nn_gd on the ground, 0.3 / max_nn elsewhere).
        Like the reference, the plane is recomputed on every call (keyframe.py:173) - unless the caller pinned one in
        `fixed_plane_model` (the reference hints at a fixed model, keyframe.py:436)."""
        self._preprocess(self._params(True))
        runtime.get_loader().ahead_params = self._last_params      # the first step of the next keyframe can go ahead too
        self.plane_model = np.asarray(self.fixed_plane_model, dtype=np.float64) if self.fixed_plane_model is not None else self.calculate_plane()
        self.segment_plane(self.plane_model, _download=False)
        eng = runtime.get_engine()
        for sid, radius, max_nn in ((self._scan_id_ground, self.voxel_size_normals_ground_plane, ICP_PARAMETERS.max_nn_gd),
                                    (self._scan_id_non_ground, self.voxel_size_normals, ICP_PARAMETERS.max_nn)):
            p = Engine.make_preprocess_params(0.0, self.max_radius, self.min_height, self.max_height, None, radius, max_nn, True,
                                              grid_max_dist=ICP_PARAMETERS.distance_threshold)
            p.min_radius2 = -1.0          # the parts already passed the filter: nothing may be dropped a second time
            eng.preprocess([sid], p)

    def calculate_plane(self, pcd=None, height=-0.5, thresholdA=0.01):
        """keyframe.py:417-436: ground plane [a, b, c, d] from the filtered points below `height` (RANSAC, 1000 iterations,
        inlier distance thresholdA) - reproducible for a given `plane_seed`."""
        if pcd is not None:
            raise NotImplementedError("calculate_plane works on this keyframe's filtered cloud")
        if not self._preprocessed_on_device:
            raise RuntimeError("pre_process / filter_radius_height first")
        plane_model, _ = runtime.get_engine().fit_plane(self._scan_id, height, thresholdA, 1000, self.plane_seed)
        a, b, c, d = plane_model
        print(f"Plane model calculated: {a:.2f}x + {b:.2f}y + {c:.2f}z + {d:.2f} = 0")
        return plane_model

    def segment_plane(self, plane_model, pcd=None, thresholdB=0.4, _download=True):
        """keyframe.py:438-461: (points within thresholdB of the plane, the others), order preserved."""
        if pcd is not None:
            raise NotImplementedError("segment_plane works on this keyframe's filtered cloud")
        if not self._preprocessed_on_device:
            raise RuntimeError("pre_process / filter_radius_height first")
        self._drop_planes()
        eng = runtime.get_engine()
        eng.split_plane(self._scan_id, plane_model, thresholdB, self._scan_id_ground, self._scan_id_non_ground)
        self._planes_on_device = True
        if not _download:
            return None
        p = Engine.make_preprocess_params(0.0, self.max_radius, self.min_height, self.max_height, None, want_normals=False)
        p.min_radius2 = -1.0
        eng.preprocess([self._scan_id_ground, self._scan_id_non_ground], p)
        out = (PointCloud(eng.get_points(self._scan_id_ground)), PointCloud(eng.get_points(self._scan_id_non_ground)))
        return out

    # ------------------------------------------------------------------ registration
    def local_registration_simple(self, other, initial_transform, option='pointpoint'):
        """ICP with target = self, source = other, init = initial_transform (4x4 ndarray or None)."""
        if initial_transform is None:
            initial_transform = np.eye(4)
        print("Apply point-to-plane ICP. Local registration")
        if option == 'pointpoint':
            method = P2P
        elif option == 'pointplane':
            method = P2PLANE
        else:
            print('UNKNOWN OPTION. Should be pointpoint or pointplane')
            raise UnboundLocalError("local variable 'reg_p2p' referenced before assignment")   # what the reference does (keyframe.py:253-255)
        if not (self._preprocessed_on_device and other._preprocessed_on_device):
            raise RuntimeError("pre_process() both keyframes before registering them")
        eng = runtime.get_engine()
        ip = eng.make_icp_params(method, ICP_PARAMETERS.distance_threshold, ICP_PARAMETERS.relative_fitness,
                                 ICP_PARAMETERS.relative_rmse, ICP_PARAMETERS.max_iteration)
        # enqueue, then - while the device iterates - stage the scan the caller will ask for next, then wait
        ticket = eng.icp_batch_async([self._scan_id], [other._scan_id], np.asarray(initial_transform, dtype=np.float64)[None], ip)
        try:
            runtime.get_loader().stage_ahead(runtime.new_scan_id)
        finally:
            rec = eng.icp_batch_finish(ticket)[0]
        self.last_result = rec
        print('Registration result: fitness=%.6e, inlier_rmse=%.6e, correspondence_set size=%d, iterations=%d'
              % (rec["fitness"], rec["rmse"], rec["n_corr"], rec["updates"]))
        return result_type()(np.array(rec["T"]))

    def local_registration_two_planes(self, other, initial_transform):
        """keyframe.py:262-295: point-to-plane ICP of the ground parts and of the non-ground parts (one device batch of two
        pairs); x, y, gamma come from the non-ground solution, z, alpha, beta from the ground solution."""
        print("Apply point-to-plane ICP. Local registration in two phases")
        if initial_transform is None:
            initial_transform = np.eye(4)
        if not (self._planes_on_device and other._planes_on_device):
            raise RuntimeError("pre_process('icp2planes') both keyframes before registering them")
        eng = runtime.get_engine()
        ip = eng.make_icp_params(P2PLANE, ICP_PARAMETERS.distance_threshold, ICP_PARAMETERS.relative_fitness,
                                 ICP_PARAMETERS.relative_rmse, ICP_PARAMETERS.max_iteration)
        init = np.asarray(initial_transform, dtype=np.float64)
        ticket = eng.icp_batch_async([self._scan_id_ground, self._scan_id_non_ground],
                                     [other._scan_id_ground, other._scan_id_non_ground], np.stack([init, init]), ip)
        try:
            runtime.get_loader().stage_ahead(runtime.new_scan_id)
        finally:
            rec = eng.icp_batch_finish(ticket)
        self.last_result = rec[1]
        return merge_two_planes(np.array(rec[0]["T"]), np.array(rec[1]["T"]))

    def global_registration(self, other):
        raise NotImplementedError("fpfh global registration is outside the B200 hot path")

    # ------------------------------------------------------------------ used by the map-building / viewer callers
    def transform(self, T):
        """keyframe.py:399-400: the filtered cloud moved by T (Open3D PointCloud::Transform), computed on the device by the
        map-building kernel; the keyframe's own cloud stays in the sensor frame (Open3D transforms in place)."""
        params = getattr(self, "_last_params", None)
        if not self._preprocessed_on_device or params is None:
            raise RuntimeError("filter_radius_height() / pre_process() first")
        xyz, _ = runtime.get_engine().map_build([self._scan_id], np.asarray(T, dtype=np.float64)[None], params)
        return PointCloud(xyz)
