"""Generate golden vectors from the REFERENCE'S OWN numpy code, run in the authoring container.

Run:  python tests/golden/make_golden.py      (needs /root/reference; the GPU box never runs this)

What the reference itself pins on the hot path (everything else is Open3D, absent — parity unpinned):
  * KeyFrame.filter_radius_height          keyframemanager/keyframe.py:74-94   (own numpy)
  * artelib SE(3) helpers used on the return value of compute_transformation
    (HomogeneousMatrix.inv/__mul__/pos/Q, tools.rot2quaternion/rot2euler/euler2rot)
      artelib/homogeneousmatrix.py:16-107, artelib/tools.py:110-275
open3d and matplotlib are absent here, so they are stubbed in sys.modules: the stub PointCloud only
carries `.points`, which is all keyframe.py:88-93 touches.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def install_stubs():
    o3d = types.ModuleType("open3d")
    o3d.geometry = types.SimpleNamespace()
    o3d.utility = types.SimpleNamespace()

    class PointCloud:
        def __init__(self, points=None):
            self.points = points

        def select_by_index(self, index, invert=False):          # order-preserving, like Open3D's
            pts = np.asarray(self.points)
            if invert:
                mask = np.ones(len(pts), dtype=bool)
                mask[np.asarray(index, dtype=np.int64)] = False
                return PointCloud(pts[mask])
            return PointCloud(pts[np.asarray(index, dtype=np.int64)])

    o3d.geometry.PointCloud = PointCloud
    o3d.utility.Vector3dVector = lambda a: np.array(a, dtype=np.float64)
    sys.modules["open3d"] = o3d
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    return PointCloud


def main():
    PointCloud = install_stubs()
    sys.path.insert(0, REF)
    from keyframemanager.keyframe import KeyFrame          # the reference's own class
    from artelib.homogeneousmatrix import HomogeneousMatrix
    from artelib import tools

    rng = np.random.default_rng(20261018)
    # points around every threshold of the predicate, float32 payload widened to float64 like a PCD load
    n = 4000
    pts = rng.uniform(-40, 40, size=(n, 3)).astype(np.float32)
    pts[:, 2] = rng.uniform(-3, 60, size=n).astype(np.float32)
    edge = []
    for r in (0.5, 35.0):
        for a in np.linspace(0, 2 * np.pi, 25):
            for eps in (-1e-6, 0.0, 1e-6):
                edge.append([(r + eps) * np.cos(a), (r + eps) * np.sin(a), 0.3])
    for z in (-1.0, 50.0):
        for eps in (-1e-6, 0.0, 1e-6):
            edge.append([3.0, 4.0, z + eps])
    edge += [[35.0, 0.0, 0.0], [0.0, 35.0, 0.0], [0.5, 0.0, 0.0], [0.3, 0.4, 0.0], [21.0, 28.0, 1.0],
             [np.nan, 1.0, 1.0], [1.0, np.inf, 1.0], [2.0, 2.0, np.nan]]
    pts = np.vstack([pts, np.array(edge, dtype=np.float32)])
    kf = KeyFrame(directory="", scan_time=0, voxel_size=None)
    kf.pointcloud = PointCloud(pts.astype(np.float64))
    with np.errstate(invalid="ignore"):
        out = kf.filter_radius_height()
    kept_default = np.asarray(out.points)
    with np.errstate(invalid="ignore"):
        out2 = kf.filter_radius_height(radii=[1.25, 20.5], heights=[-0.5, 2.75])
    kept_custom = np.asarray(out2.points)
    np.savez_compressed(os.path.join(OUT, "filter_radius_height.npz"), points_f32=pts, kept_default=kept_default,
                        kept_custom=kept_custom, custom_radii=np.array([1.25, 20.5]), custom_heights=np.array([-0.5, 2.75]),
                        defaults=np.array([kf.min_radius, kf.max_radius, kf.min_height, kf.max_height], dtype=np.float64))

    # SE(3) helper goldens
    mats, invs, prods, quats, eulers, eulers2 = [], [], [], [], [], []
    for _ in range(32):
        e = rng.uniform(-np.pi, np.pi, 3)
        e[1] = rng.uniform(-1.4, 1.4)
        R = tools.euler2rot(e)
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = rng.uniform(-10, 10, 3)
        H = HomogeneousMatrix(T)
        mats.append(T)
        invs.append(H.inv().array)
        prods.append((H * H.inv() * H).array)
        quats.append(np.array(tools.rot2quaternion(T)))
        eulers.append(np.array(tools.rot2euler(T)[0]))
        eulers2.append(np.array(tools.rot2euler(T.copy())[1]))
    gimbal_mats, gimbal_e1, gimbal_e2 = [], [], []
    for k in range(8):                                   # beta = +-pi/2: the degenerate branch of rot2euler
        e = rng.uniform(-np.pi, np.pi, 3)
        e[1] = np.pi / 2 * (1 if k % 2 == 0 else -1)
        R = np.array(tools.euler2rot(e))[:3, :3]
        e1, e2 = tools.rot2euler(R.copy())
        gimbal_mats.append(R); gimbal_e1.append(np.array(e1)); gimbal_e2.append(np.array(e2))
    np.savez_compressed(os.path.join(OUT, "se3_helpers.npz"), mats=np.array(mats), invs=np.array(invs),
                        prods=np.array(prods), quats=np.array(quats), eulers=np.array(eulers), eulers2=np.array(eulers2),
                        gimbal_mats=np.array(gimbal_mats), gimbal_e1=np.array(gimbal_e1), gimbal_e2=np.array(gimbal_e2))
    # 'icp2planes' pieces the reference implements itself: segment_plane (keyframe.py:438-461, own numpy) and the
    # component merge of local_registration_two_planes (keyframe.py:282-292: t2v(n=3) of both results, tx ty gamma from
    # the non-ground solution, tz alpha beta from the ground solution, HomogeneousMatrix(position, Euler))
    from artelib.euler import Euler
    cloud = rng.uniform(-30, 30, size=(3000, 3))
    cloud[:, 2] = rng.uniform(-1.2, 3.0, size=3000)
    plane = np.array([0.01, -0.02, 0.999, 0.69])
    cloud[:40, 2] = (-(plane[0] * cloud[:40, 0] + plane[1] * cloud[:40, 1] + plane[3]) / plane[2]) + np.linspace(-0.41, 0.41, 40)
    kf2 = KeyFrame(directory="", scan_time=0, voxel_size=None)
    near_pc, far_pc = kf2.segment_plane(plane, pcd=PointCloud(cloud))
    t2v3, merged = [], []
    for k in range(32):
        t2v3.append(HomogeneousMatrix(np.array(mats[k])).t2v(n=3))
    for k in range(16):
        t1, t2 = t2v3[k], t2v3[k + 16]
        merged.append(HomogeneousMatrix(np.array([t2[0], t2[1], t1[2]]), Euler([t1[3], t1[4], t2[5]])).array)
    np.savez_compressed(os.path.join(OUT, "two_planes.npz"), cloud=cloud, plane=plane, near=np.asarray(near_pc.points),
                        far=np.asarray(far_pc.points), mats=np.array(mats), t2v3=np.array(t2v3), merged=np.array(merged))
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
