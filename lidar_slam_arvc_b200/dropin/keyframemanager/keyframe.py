"""KeyFrame: one LiDAR scan, its preprocessing and pairwise registration, backed by the CUDA engine.

Mirror of the reference's keyframemanager/keyframe.py public surface for the methods the scan-matching hot
path uses (`icppointpoint`, `icppointplane`); the Open3D calls are replaced by calls through the C-ABI:
    load_pointcloud            keyframe.py:41-45     PCD payload -> device (arvc_scan_upload_*)
    filter_radius_height       keyframe.py:74-94     )
    pre_process / preprocess_* keyframe.py:113-162   ) arvc_scan_preprocess
    local_registration_simple  keyframe.py:231-260   arvc_icp_batch (one pair)
    unload_pointcloud          keyframe.py:61-72     arvc_scan_free
The reference's `icp2planes` and `fpfh` methods depend on Open3D's unseeded RANSAC and are out of scope
(SURVEY.md §2 #17): they raise NotImplementedError instead of silently doing something else.
"""
import numpy as np

from config import ICP_PARAMETERS
from lidar_slam_arvc_b200 import runtime
from lidar_slam_arvc_b200.engine import P2P, P2PLANE, Engine
from lidar_slam_arvc_b200.homogeneousmatrix import result_type
from lidar_slam_arvc_b200.pcd import read_pcd_xyz


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


class KeyFrame():
    def __init__(self, directory, scan_time, voxel_size):
        self.directory = directory
        self.scan_time = scan_time
        self.voxel_size = voxel_size
        self.fpfh_threshold = 5
        self.pointcloud = None
        self.pointcloud_ground_plane = None
        self.pointcloud_non_ground_plane = None
        self.pointcloud_fpfh = None
        self.voxel_size_normals_ground_plane = 0.5
        self.voxel_size_normals = 0.3            # the radius actually used for normals (keyframe.py:33)
        self.max_radius = ICP_PARAMETERS.max_radius
        self.min_radius = ICP_PARAMETERS.min_radius
        self.max_height = ICP_PARAMETERS.max_height
        self.min_height = ICP_PARAMETERS.min_height
        self.plane_model = None
        self.pre_processed = False               # never set by the reference either (keyframe.py:39,114)
        self.last_result = None                  # extension: fitness / inlier_rmse / iterations of the last registration
        self._scan_id = runtime.new_scan_id()
        self._filtered_cache = None
        self._on_device = False
        self._preprocessed_on_device = False

    # ------------------------------------------------------------------ load / unload
    def load_pointcloud(self):
        filename = self.directory + '/robot0/lidar/data/' + str(self.scan_time) + '.pcd'
        print('Reading pointcloud: ', filename)
        self.set_points(read_pcd_xyz(filename))

    def set_points(self, xyz):
        """Extension: hand the PCD payload over directly ([n,3] float32 or float64)."""
        self.pointcloud = PointCloud(xyz)
        runtime.get_engine().upload(self._scan_id, xyz)
        self._on_device = True
        self._preprocessed_on_device = False
        self._filtered_cache = None

    def unload_pointcloud(self):
        print('Removing pointclouds from memory (filtered, planes, fpfh): ')
        if self._on_device:
            runtime.get_engine().free(self._scan_id)
        self._on_device = False
        self._preprocessed_on_device = False
        self._filtered_cache = None
        self.pointcloud = None
        self.pointcloud_ground_plane = None
        self.pointcloud_non_ground_plane = None
        self.pointcloud_fpfh = None

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
        elif method in ('icp2planes', 'fpfh'):
            raise NotImplementedError("method '%s' is outside the B200 hot path (Open3D RANSAC based)" % method)

    def preprocess_icp_point_point(self):
        self._preprocess(self._params(False))

    def preprocess_icp_point_plane(self):
        self._preprocess(self._params(True))

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
        rec = eng.icp_batch([self._scan_id], [other._scan_id], np.asarray(initial_transform, dtype=np.float64)[None], ip)[0]
        self.last_result = rec
        print('Registration result: fitness=%.6e, inlier_rmse=%.6e, correspondence_set size=%d, iterations=%d'
              % (rec["fitness"], rec["rmse"], rec["n_corr"], rec["updates"]))
        return result_type()(np.array(rec["T"]))

    def local_registration_two_planes(self, other, initial_transform):
        raise NotImplementedError("icp2planes is outside the B200 hot path (SURVEY.md §8 f-2)")

    def global_registration(self, other):
        raise NotImplementedError("fpfh global registration is outside the B200 hot path")

    # ------------------------------------------------------------------ used by the map-building / viewer callers
    def transform(self, T):
        pc = self.pointcloud_filtered
        return PointCloud(np.array(pc.points), None if pc.normals is None else np.array(pc.normals)).transform(T)
