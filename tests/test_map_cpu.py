"""SURVEY.md §8 f-4 on CPU: the oracle's build_map restatement (keyframemanager.py:154-184 minus the viewer) against
plain numpy, and the drop-in KeyFrameManager.build_map host logic on the oracle test double."""
import os
import sys

import numpy as np

from lidar_slam_arvc_b200 import euroc_synth, runtime, synth
from lidar_slam_arvc_b200.homogeneousmatrix import HomogeneousMatrix
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "lidar_slam_arvc_b200", "dropin")


def numpy_map(scans, Ts, radii, heights):
    parts = []
    for s, T in zip(scans, Ts):
        p = s.astype(np.float64)
        r2 = p[:, 0] ** 2 + p[:, 1] ** 2
        p = p[(r2 < radii[1] ** 2) & (r2 > radii[0] ** 2) & (p[:, 2] > heights[0]) & (p[:, 2] < heights[1])]
        parts.append(p @ T[:3, :3].T + T[:3, 3])
    return np.vstack(parts)


def test_oracle_build_map_matches_numpy():
    seq = synth.Sequence(5, synth.TINY_16, start=12.0)
    xyz, off = orc.build_map(seq.scans, seq.poses, radii=(0.5, 35.0), heights=(-120.0, 120.0))
    ref = numpy_map(seq.scans, seq.poses, (0.5, 35.0), (-120.0, 120.0))
    assert off[0] == 0 and off[-1] == len(xyz) == len(ref) and np.all(np.diff(off) > 0)
    np.testing.assert_allclose(xyz, ref, rtol=0, atol=1e-12)
    # keyframe_sampling picks every k-th transform for consecutive keyframes (keyframemanager.py:165-167)
    xyz2, _ = orc.build_map(seq.scans[:3], seq.poses, keyframe_sampling=2)
    ref2 = numpy_map(seq.scans[:3], [seq.poses[0], seq.poses[2], seq.poses[4]], (0.5, 35.0), (-120.0, 120.0))
    np.testing.assert_allclose(xyz2, ref2, rtol=0, atol=1e-12)
    # voxel size: every keyframe is down-sampled on its own before it is moved
    xyz3, off3 = orc.build_map(seq.scans, seq.poses, voxel_size=0.5)
    assert off3[-1] < off[-1]
    p0 = seq.scans[0].astype(np.float64)
    p0 = p0[orc.filter_radius_height(p0, 0.5, 35.0, -120.0, 120.0)]
    v0, _, _ = orc.voxel_down_sample(p0, 0.5)
    np.testing.assert_allclose(xyz3[:off3[1]], v0 @ seq.poses[0][:3, :3].T + seq.poses[0][:3, 3], rtol=0, atol=1e-12)
    # a projective last row is honoured (Open3D divides by w)
    T = np.eye(4); T[3, 3] = 2.0
    np.testing.assert_allclose(orc.transform_points(np.array([[2.0, 4.0, 6.0]]), T), [[1.0, 2.0, 3.0]])


def test_dropin_build_map_is_one_batch(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_engine import OracleEngine
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    fake = OracleEngine()
    runtime.set_engine(fake)
    try:
        for m in [k for k in sys.modules if k.split(".")[0] in ("config", "keyframemanager")]:
            del sys.modules[m]
        sys.path.insert(0, DROPIN)
        import keyframemanager.keyframemanager as kfm
        seq = synth.Sequence(6, synth.TINY_16, start=40.0)
        d = str(tmp_path / "euroc")
        times = euroc_synth.write_euroc_tree(d, seq)
        for voxel in (None, 0.4):
            km = kfm.KeyFrameManager(directory=d, scan_times=times, voxel_size=voxel)
            km.add_keyframes(keyframe_sampling=2)                     # keyframes 0, 2, 4
            km.load_pointclouds()
            fake.calls.clear()
            gt = [HomogeneousMatrix(T) for T in seq.poses]
            cloud = km.build_map(gt, keyframe_sampling=2)
            assert [c for c in fake.calls if c[0] == "map_build"] == [("map_build", 3)]
            want, off = orc.build_map([seq.scans[0], seq.scans[2], seq.scans[4]], seq.poses, voxel_size=voxel, keyframe_sampling=2)
            np.testing.assert_array_equal(cloud.points, want)
            np.testing.assert_array_equal(km.map_offsets, off)
            # the keyframes keep the map's filter bounds: down_sample() afterwards works on that cloud (keyframe.py:108-111)
            kf = km.keyframes[1]
            kf.filter_radius_height(radii=[1.0, 20.0], heights=[-0.5, 3.0])
            kf.down_sample()
            p = seq.scans[2].astype(np.float64)
            p = p[orc.filter_radius_height(p, 1.0, 20.0, -0.5, 3.0)]
            if voxel is not None:
                p, _, _ = orc.voxel_down_sample(p, voxel)
            np.testing.assert_array_equal(kf.pointcloud_filtered.points, p)
            # KeyFrame.transform (keyframe.py:399-400): the filtered cloud in the global frame, keyframe itself untouched
            moved = kf.transform(seq.poses[2])
            np.testing.assert_array_equal(moved.points, orc.transform_points(p, seq.poses[2]))
            np.testing.assert_array_equal(kf.pointcloud_filtered.points, p)
    finally:
        runtime.set_engine(None)
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k not in saved_mods:
                del sys.modules[k]
