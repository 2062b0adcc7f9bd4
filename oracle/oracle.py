"""ctypes front-end of the C++ oracle (``icp_oracle.cpp``).  TEST INFRASTRUCTURE ONLY (see package docstring).

Every function cites the reference call site it stands in for (paths relative to /root/reference).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_icp.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    """Compile the oracle with the committed Makefile (g++ only, no reference sources involved)."""
    src = os.path.join(_HERE, "icp_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_num_threads.restype = ctypes.c_int
        _lib.orc_icp.restype = ctypes.c_int
        _lib.orc_correspondences.restype = ctypes.c_int
        _lib.orc_filter_radius_height.restype = ctypes.c_int
        _lib.orc_voxel_down_sample.restype = ctypes.c_int
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _pts(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert a.ndim == 2 and a.shape[1] == 3
    return a


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(ctypes.c_int(int(n)))


def filter_radius_height(points, min_radius=0.5, max_radius=35, min_height=-1.0, max_height=50.0):
    """keyframe.py:74-94.  Returns the kept indices (ascending).  Radii are squared the Python way
    (``radius ** 2``) exactly like keyframe.py:92."""
    p = _pts(points)
    keep = np.empty(len(p), dtype=np.int32)
    m = lib().orc_filter_radius_height(_d(p), len(p), ctypes.c_double(float(min_radius ** 2)),
                                       ctypes.c_double(float(max_radius ** 2)), ctypes.c_double(float(min_height)),
                                       ctypes.c_double(float(max_height)), _i(keep))
    return keep[:m].copy()


def voxel_down_sample(points, voxel_size):
    """keyframe.py:111/151/159 → Open3D PointCloud::VoxelDownSample.  Returns (points[M,3], keys[M,3] int32,
    counts[M]) ordered by key (ix, iy, iz)."""
    p = _pts(points)
    n = len(p)
    out = np.empty((max(n, 1), 3))
    keys = np.empty((max(n, 1), 3), dtype=np.int32)
    cnt = np.empty(max(n, 1), dtype=np.int32)
    m = lib().orc_voxel_down_sample(_d(p), n, ctypes.c_double(float(voxel_size)), _d(out), _i(keys), _i(cnt))
    if m < 0:
        raise ValueError("voxel_size <= 0")
    return out[:m].copy(), keys[:m].copy(), cnt[:m].copy()


def knn_hybrid(points, queries, radius, max_nn):
    """Open3D KDTreeFlann::SearchHybrid for every query.  Returns (idx[nq,max_nn] (-1 pad), d2, count)."""
    p = _pts(points)
    q = _pts(queries)
    idx = np.empty((len(q), max_nn), dtype=np.int32)
    d2 = np.empty((len(q), max_nn))
    cnt = np.empty(len(q), dtype=np.int32)
    lib().orc_knn_hybrid(_d(p), len(p), _d(q), len(q), ctypes.c_double(radius), int(max_nn), _i(idx), _d(d2), _i(cnt))
    return idx, d2, cnt


def estimate_normals(points, radius=0.3, max_nn=300, return_cov=False):
    """keyframe.py:160-162 → Open3D EstimateNormals(KDTreeSearchParamHybrid(radius, max_nn)), fast eigen."""
    p = _pts(points)
    n = len(p)
    nrm = np.empty((n, 3))
    cov = np.empty((n, 9))
    cnt = np.empty(n, dtype=np.int32)
    lib().orc_estimate_normals(_d(p), n, ctypes.c_double(radius), int(max_nn), _d(nrm), _d(cov), _i(cnt))
    if return_cov:
        return nrm, cov.reshape(n, 3, 3), cnt
    return nrm


def normal_from_covariance(cov):
    c = np.ascontiguousarray(cov, dtype=np.float64).reshape(9)
    out = np.empty(3)
    lib().orc_normal_from_covariance(_d(c), _d(out))
    return out


def correspondences(source, target, T=None, max_dist=10.0):
    """One Open3D GetRegistrationResultAndCorrespondences pass.  Returns (corr[ns] int32 (-1 = none), d2, fitness, rmse)."""
    s = _pts(source)
    t = _pts(target)
    T = np.ascontiguousarray(np.eye(4) if T is None else T, dtype=np.float64)
    corr = np.empty(max(len(s), 1), dtype=np.int32)
    d2 = np.empty(max(len(s), 1))
    fit = ctypes.c_double()
    rmse = ctypes.c_double()
    lib().orc_correspondences(_d(s), len(s), _d(t), len(t), _d(T), ctypes.c_double(max_dist), _i(corr), _d(d2),
                              ctypes.byref(fit), ctypes.byref(rmse))
    return corr[:len(s)], d2[:len(s)], fit.value, rmse.value


class IcpResult:
    __slots__ = ("transformation", "fitness", "inlier_rmse", "updates", "passes", "n_corr", "correspondences",
                 "trace_T", "trace_fitness", "trace_rmse")


P2P, P2PLANE = 0, 1


def icp(source, target, target_normals=None, init=None, method=P2PLANE, max_corr_dist=10.0, rel_fitness=1e-6,
        rel_rmse=1e-6, max_iter=30):
    """keyframe.py:246-252 → Open3D registration_icp with default ICPConvergenceCriteria(1e-6, 1e-6, 30)."""
    s = _pts(source)
    t = _pts(target)
    nrm = None
    if method == P2PLANE:
        nrm = _pts(target_normals)
        assert len(nrm) == len(t)
    T0 = np.ascontiguousarray(np.eye(4) if init is None else init, dtype=np.float64)
    outT = np.empty((4, 4))
    fit = ctypes.c_double()
    rmse = ctypes.c_double()
    upd = ctypes.c_int()
    nc = ctypes.c_int()
    corr = np.empty(max(len(s), 1), dtype=np.int32)
    trT = np.zeros((max_iter + 1, 4, 4))
    trf = np.zeros(max_iter + 1)
    trr = np.zeros(max_iter + 1)
    passes = lib().orc_icp(_d(s), len(s), _d(t), len(t), _d(nrm) if nrm is not None else None, _d(T0), int(method),
                           ctypes.c_double(max_corr_dist), ctypes.c_double(rel_fitness), ctypes.c_double(rel_rmse),
                           int(max_iter), _d(outT), ctypes.byref(fit), ctypes.byref(rmse), ctypes.byref(upd),
                           ctypes.byref(nc), _i(corr), _d(trT), _d(trf), _d(trr))
    if passes < 0:
        raise RuntimeError("oracle icp failed: %d" % passes)
    r = IcpResult()
    r.transformation = outT
    r.fitness = fit.value
    r.inlier_rmse = rmse.value
    r.updates = upd.value
    r.passes = passes
    r.n_corr = nc.value
    r.correspondences = corr[:len(s)]
    r.trace_T = trT[:passes]
    r.trace_fitness = trf[:passes]
    r.trace_rmse = trr[:passes]
    return r


def p2plane_system(src_transformed, target, normals, corr):
    s = _pts(src_transformed)
    t = _pts(target)
    n = _pts(normals)
    c = np.ascontiguousarray(corr, dtype=np.int32)
    JTJ = np.empty((6, 6))
    JTr = np.empty(6)
    upd = np.empty((4, 4))
    lib().orc_p2plane_system(_d(s), _d(t), _d(n), _i(c), len(s), _d(JTJ), _d(JTr), _d(upd))
    return JTJ, JTr, upd


def p2p_update(src_transformed, target, corr):
    s = _pts(src_transformed)
    t = _pts(target)
    c = np.ascontiguousarray(corr, dtype=np.int32)
    upd = np.empty((4, 4))
    lib().orc_p2p_update(_d(s), _d(t), _i(c), len(s), _d(upd))
    return upd


def ldlt_solve6(A, b):
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.empty(6)
    lib().orc_ldlt_solve6(_d(A), _d(b), _d(x))
    return x


def vec6_to_mat4(v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    T = np.empty((4, 4))
    lib().orc_vec6_to_mat4(_d(v), _d(T))
    return T


def svd3(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    U = np.empty((3, 3))
    s = np.empty(3)
    V = np.empty((3, 3))
    lib().orc_svd3(_d(A), _d(U), _d(s), _d(V))
    return U, s, V


def transform_points(points, T):
    """keyframe.py:399-400 (`pointcloud_filtered.transform(T)`): homogeneous product, then division by w."""
    p = _pts(points)
    out = np.empty_like(p)
    T = np.ascontiguousarray(T, dtype=np.float64).reshape(4, 4)
    lib().orc_transform_points(_d(p), len(p), _d(T), _d(out))
    return out


def build_map(scans_f32, transforms, voxel_size=None, radii=(0.5, 35.0), heights=(-120.0, 120.0), keyframe_sampling=1):
    """keyframemanager.py:154-184 build_map without the GUI: per keyframe filter_radius_height(radii, heights) ->
    down_sample -> transform(sampled_transforms[i]) -> concatenate.  Returns (points [sum,3], offsets [n+1])."""
    sampled = [transforms[i] for i in range(0, len(transforms), keyframe_sampling)]
    parts, offsets = [], [0]
    for s, T in zip(scans_f32, sampled):
        p = np.asarray(s, dtype=np.float32).astype(np.float64) if np.asarray(s).dtype == np.float32 else np.asarray(s, dtype=np.float64)
        p = p[filter_radius_height(p, radii[0], radii[1], heights[0], heights[1])]
        if voxel_size is not None:
            p, _, _ = voxel_down_sample(p, voxel_size)
        parts.append(transform_points(p, T))
        offsets.append(offsets[-1] + len(p))
    return (np.concatenate(parts) if parts else np.zeros((0, 3))), np.array(offsets, dtype=np.int64)


def fit_plane(points, max_z=-0.5, dist_threshold=0.01, iterations=1000, seed=0):
    """keyframe.py:417-436 calculate_plane (deterministic RANSAC convention, see icp_oracle.cpp).  Returns ([a,b,c,d], inliers)."""
    p = _pts(points)
    pl = np.zeros(4)
    lib().orc_fit_plane.restype = ctypes.c_int
    n_in = lib().orc_fit_plane(_d(p), len(p), ctypes.c_double(max_z), ctypes.c_double(dist_threshold), int(iterations),
                               ctypes.c_ulonglong(int(seed)), _d(pl))
    return pl, int(n_in)


def segment_plane(points, plane_model, threshold=0.4):
    """keyframe.py:438-461 (the reference's own numpy): indices of the points near the plane and of the others."""
    points = np.asarray(points, dtype=np.float64)
    a, b, c, d = [float(v) for v in plane_model]
    dist = np.abs(a * points[:, 0] + b * points[:, 1] + c * points[:, 2] + d) / np.sqrt(a * a + b * b + c * c)
    near = dist < threshold
    return np.where(near)[0], np.where(~near)[0]


def preprocess_two_planes(points_f32, plane_model=None, voxel_size=None, normal_radius=0.3, normal_radius_ground=0.5, max_nn=300, max_nn_gd=300,
                          seed=0):
    """keyframe.py:164-189 preprocess_icp2planes: filter -> [voxel] -> plane model -> split -> normals of both parts (radius
    0.5 / max_nn_gd on the ground, 0.3 / max_nn elsewhere).  Returns (plane, (ground pts, normals), (other pts, normals))."""
    p, _ = preprocess(points_f32, voxel_size=voxel_size, method="icppointpoint")
    if plane_model is None:
        plane_model, _ = fit_plane(p, seed=seed)
    near, far = segment_plane(p, plane_model)
    g, o = p[near], p[far]
    return plane_model, (g, estimate_normals(g, normal_radius_ground, max_nn_gd)), (o, estimate_normals(o, normal_radius, max_nn))


def preprocess(points_f32, voxel_size=None, method="icppointplane", min_radius=0.5, max_radius=35, min_height=-1.0,
               max_height=50.0, normal_radius=0.3, max_nn=300):
    """keyframe.py:148-162: filter → optional voxel → normals (point-plane only).  Input is the float32 PCD
    payload, widened exactly to float64 as Open3D does.  Returns (points[M,3] f64, normals or None)."""
    p = np.asarray(points_f32, dtype=np.float32).astype(np.float64)
    keep = filter_radius_height(p, min_radius, max_radius, min_height, max_height)
    p = p[keep]
    if voxel_size is not None:
        p, _, _ = voxel_down_sample(p, voxel_size)
    nrm = None
    if method == "icppointplane":
        nrm = estimate_normals(p, normal_radius, max_nn)
    return p, nrm
