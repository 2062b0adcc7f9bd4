"""KeyFrameManager: indexable container of KeyFrames that dispatches `method` to the registration routine.

Same public surface as the reference's keyframemanager/keyframemanager.py:9-75 (constructor, add_keyframe(s),
load/unload_pointcloud, pre_process, compute_transformation); the GUI methods (draw_*, visualize_*) are not part
of the hot path.  Extension for batched callers (loop closing): `pre_process_many` and
`compute_transformations`, which issue ONE device batch for many independent pairs.
"""
import collections
import os

import numpy as np

from config import ICP_PARAMETERS
from keyframemanager.keyframe import KeyFrame, PointCloud, merge_two_planes
from lidar_slam_arvc_b200 import runtime
from lidar_slam_arvc_b200.engine import P2P, P2PLANE
from lidar_slam_arvc_b200.homogeneousmatrix import result_type


class KeyFrameManager():
    def __init__(self, directory, scan_times, voxel_size, method='icppointplane'):
        """given a list of scan times (ROS times), each pcd is read on demand"""
        self.directory = directory
        self.scan_times = scan_times
        self.keyframes = []
        self.voxel_size = voxel_size
        self.method = method
        self.show_registration_result = False
        self._resident = collections.OrderedDict()      # keyframe positions with a device copy, least recently loaded first
        self.max_resident_keyframes = int(os.environ.get("ARVC_MAX_RESIDENT_KEYFRAMES", "4096"))   # ~12 MB of HBM each at 64 beams

    def add_keyframes(self, keyframe_sampling):
        for i in range(0, len(self.scan_times), keyframe_sampling):
            print("Keyframemanager: Adding Keyframe: ", i, "out of: ", len(self.scan_times), end='\r')
            self.add_keyframe(i)

    def add_keyframe(self, index):
        print('Adding keyframe with scan_time: ', self.scan_times[index])
        kf = KeyFrame(directory=self.directory, scan_time=self.scan_times[index], voxel_size=self.voxel_size)
        kf._scan_index = index
        self.keyframes.append(kf)

    def load_pointclouds(self):
        for i in range(0, len(self.keyframes)):
            print("Keyframemanager: Loading Pointcloud: ", i, "out of: ", len(self.keyframes), end='\r')
            self.keyframes[i].load_pointcloud()

    def load_pointcloud(self, i):
        self.keyframes[i].load_pointcloud()
        self._read_ahead(i)
        # Loop closing never unloads (loopclosing.py:163-178), and resident keyframes are what makes its second visit of a
        # keyframe free - but device memory is finite: beyond `max_resident_keyframes` the least recently loaded ones are
        # dropped (they come back, bit-identical, with their next load_pointcloud).
        self._resident.pop(i, None)
        self._resident[i] = True
        while len(self._resident) > self.max_resident_keyframes:
            old = next(iter(self._resident))
            self._resident.pop(old)
            self.keyframes[old].unload_pointcloud()

    def _read_ahead(self, i):
        """Announce the file a sequential caller will ask for next (run_scanmatcher.py:196-213 adds and loads keyframe
        i + 1 right after registering i): the loader's helper thread reads it while the GPU registers the current pair."""
        if i != len(self.keyframes) - 1:
            return                                   # not the newest keyframe: random access (loop closing), no guess
        idx = getattr(self.keyframes[i], "_scan_index", None)
        if idx is None:
            return
        prev = getattr(self.keyframes[i - 1], "_scan_index", None) if i > 0 else None
        step = idx - prev if prev is not None and idx > prev else 1
        if idx + step < len(self.scan_times):
            runtime.get_loader().prefetch(self.directory + '/robot0/lidar/data/' + str(self.scan_times[idx + step]) + '.pcd')

    def unload_pointcloud(self, i):
        self.keyframes[i].unload_pointcloud()
        self._resident.pop(i, None)

    def pre_process(self, index):
        self.keyframes[index].pre_process(method=self.method)

    def compute_transformation(self, i, j, Tij):
        """ICP with target = keyframes[i], source = keyframes[j], initial guess Tij (HomogeneousMatrix); returns iTj."""
        if self.method == 'icppointpoint':
            transform = self.keyframes[i].local_registration_simple(self.keyframes[j], initial_transform=Tij.array,
                                                                    option='pointpoint')
        elif self.method == 'icppointplane':
            transform = self.keyframes[i].local_registration_simple(self.keyframes[j], initial_transform=Tij.array,
                                                                    option='pointplane')
        elif self.method == 'icp2planes':
            transform = self.keyframes[i].local_registration_two_planes(self.keyframes[j], initial_transform=Tij.array)
        elif self.method == 'fpfh':
            transform = self.keyframes[i].global_registration(self.keyframes[j])
        else:
            print('Unknown registration method')
            transform = None
        return transform

    # ------------------------------------------------------------------ batched extensions
    def pre_process_many(self, indices):
        """pre_process() of many keyframes in one device batch (same results as calling pre_process one by one)."""
        kfs = [self.keyframes[i] for i in indices]
        if not kfs:
            return
        if self.method == 'icp2planes':                 # plane model + split are per keyframe
            for kf in kfs:
                kf.pre_process(method=self.method)
            return
        if self.method not in ('icppointpoint', 'icppointplane'):
            raise NotImplementedError("batched preprocessing supports icppointpoint / icppointplane / icp2planes")
        for kf in kfs:
            kf._require_loaded()
        runtime.get_engine().preprocess([kf._scan_id for kf in kfs], kfs[0]._params(self.method == 'icppointplane'))
        for kf in kfs:
            kf._preprocessed_on_device = True
            kf._filtered_cache = None
            kf._last_params = kfs[0]._params(self.method == 'icppointplane')

    def compute_transformations(self, pairs, Tijs):
        """[(i, j), ...] and initial guesses -> list of iTj, one device batch.  Each keyframe's `last_result`-style record
        is returned alongside: (transforms, records)."""
        eng = runtime.get_engine()
        if self.method == 'icp2planes':                 # two point-to-plane problems per pair, still one batch
            ip = eng.make_icp_params(P2PLANE, ICP_PARAMETERS.distance_threshold, ICP_PARAMETERS.relative_fitness,
                                     ICP_PARAMETERS.relative_rmse, ICP_PARAMETERS.max_iteration)
            tg, sr, init = [], [], []
            for (i, j), T in zip(pairs, Tijs):
                a = getattr(T, "array", T)
                a = np.eye(4) if a is None else np.asarray(a, dtype=np.float64)
                tg += [self.keyframes[i]._scan_id_ground, self.keyframes[i]._scan_id_non_ground]
                sr += [self.keyframes[j]._scan_id_ground, self.keyframes[j]._scan_id_non_ground]
                init += [a, a]
            rec = eng.icp_batch(tg, sr, np.array(init).reshape(-1, 4, 4), ip)
            return [merge_two_planes(np.array(rec[2 * k]["T"]), np.array(rec[2 * k + 1]["T"])) for k in range(len(pairs))], rec
        if self.method not in ('icppointpoint', 'icppointplane'):
            raise NotImplementedError("batched registration supports icppointpoint / icppointplane / icp2planes")
        method = P2P if self.method == 'icppointpoint' else P2PLANE
        ip = eng.make_icp_params(method, ICP_PARAMETERS.distance_threshold, ICP_PARAMETERS.relative_fitness,
                                 ICP_PARAMETERS.relative_rmse, ICP_PARAMETERS.max_iteration)
        tg = [self.keyframes[i]._scan_id for i, _ in pairs]
        sr = [self.keyframes[j]._scan_id for _, j in pairs]
        init = np.array([np.eye(4) if (T is None or getattr(T, "array", T) is None) else np.asarray(getattr(T, "array", T), dtype=np.float64)
                         for T in Tijs]).reshape(-1, 4, 4)
        rec = eng.icp_batch(tg, sr, init, ip)
        H = result_type()
        return [H(np.array(r["T"])) for r in rec], rec

    # ------------------------------------------------------------------ map building without the GUI
    def build_map(self, global_transforms, keyframe_sampling=10, radii=None, heights=None):
        """keyframemanager.py:154-184 as ONE device batch: every keyframe is filtered with
        radii / heights, down-sampled when voxel_size is set, moved by its sampled global transform and written at its
        offset of a single map array.  Returns the map as a PointCloud (the reference also opens a viewer).  Unlike
        Open3D's in-place transform, the keyframes' own `pointcloud_filtered` stay in the sensor frame."""
        if radii is None:
            radii = [0.5, 35.0]
        if heights is None:
            heights = [-120.0, 120.0]
        sampled = [global_transforms[i] for i in range(0, len(global_transforms), keyframe_sampling)]
        kfs = list(self.keyframes)
        if not kfs:
            return PointCloud()
        for kf in kfs:
            kf._require_loaded()
        T = np.array([np.asarray(sampled[i].array, dtype=np.float64) for i in range(len(kfs))])
        params = kfs[0]._params(False, radii, heights)
        xyz, offsets = runtime.get_engine().map_build([kf._scan_id for kf in kfs], T, params)
        for kf in kfs:
            kf._filter_bounds = (radii, heights)
            kf._last_params = params
            kf._preprocessed_on_device = True
            kf._filtered_cache = None
        self.map_offsets = offsets
        return PointCloud(xyz)
