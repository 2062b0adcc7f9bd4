"""CPU tests that pin the oracle (oracle/icp_oracle.cpp) — against the reference's own numpy where the
reference has any (golden vectors), against an independent numpy/scipy implementation, and against
known answers.  SURVEY.md §4 / §8(c)."""
import os

import numpy as np
import pytest

from lidar_slam_arvc_b200 import synth
from oracle import numpy_ref as ref
from oracle import oracle as orc


# ------------------------------------------------------------------ golden: the reference's own code
def test_filter_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "filter_radius_height.npz"))
    pts = g["points_f32"].astype(np.float64)
    keep = orc.filter_radius_height(pts)
    np.testing.assert_array_equal(pts[keep], g["kept_default"])
    r, h = g["custom_radii"], g["custom_heights"]
    keep2 = orc.filter_radius_height(pts, r[0], r[1], h[0], h[1])
    np.testing.assert_array_equal(pts[keep2], g["kept_custom"])
    # and the numpy twin
    np.testing.assert_array_equal(ref.filter_radius_height(pts), keep)


def test_filter_strict_inequalities():
    pts = np.array([[35.0, 0, 0], [0.5, 0, 0], [3, 4, -1.0], [3, 4, 50.0], [3, 4, 0.0], [np.nan, 0, 0]], dtype=np.float64)
    with np.errstate(invalid="ignore"):
        assert list(orc.filter_radius_height(pts)) == [4]
    assert len(orc.filter_radius_height(np.zeros((0, 3)))) == 0


# ------------------------------------------------------------------ voxel down-sample
def test_voxel_matches_numpy_dict():
    rng = np.random.default_rng(0)
    pts = rng.uniform(-5, 5, size=(3000, 3)).astype(np.float32).astype(np.float64)
    out, keys, cnt = orc.voxel_down_sample(pts, 0.5)
    d = ref.voxel_down_sample(pts, 0.5)
    assert len(out) == len(d)
    assert (np.diff(keys.astype(np.int64) @ np.array([1 << 40, 1 << 20, 1]), axis=0) > 0).all()   # sorted by key
    for p, k, c in zip(out, keys, cnt):
        m, cc = d[tuple(int(v) for v in k)]
        assert c == cc
        np.testing.assert_allclose(p, m, rtol=0, atol=1e-12)
    assert cnt.sum() == len(pts)


def test_voxel_face_aligned_keys():
    # points exactly on cell faces: origin = min - v/2, key = floor((p - origin)/v)
    v = 0.25
    pts = np.array([[0, 0, 0], [0.125, 0, 0], [0.25, 0, 0], [0.375, 0, 0], [0.124999, 0, 0]], dtype=np.float64)
    out, keys, cnt = orc.voxel_down_sample(pts, v)
    exp = np.floor((pts - (pts.min(0) - v / 2)) / v).astype(int)
    got = {tuple(k): c for k, c in zip(keys, cnt)}
    uk, uc = np.unique(exp, axis=0, return_counts=True)
    assert got == {tuple(k): c for k, c in zip(uk, uc)}
    with pytest.raises(ValueError):
        orc.voxel_down_sample(pts, 0.0)


# ------------------------------------------------------------------ hybrid k-NN
def test_knn_hybrid_vs_bruteforce_and_scipy():
    rng = np.random.default_rng(1)
    pts = rng.normal(size=(1500, 3)).astype(np.float32).astype(np.float64)
    q = pts[:200]
    idx, d2, cnt = orc.knn_hybrid(pts, q, 0.4, 20)
    sets = ref.knn_hybrid_sets(pts, q, 0.4, 20)
    for i in range(len(q)):
        diff = pts - q[i]
        bd2 = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]
        order = np.lexsort((np.arange(len(pts)), bd2))
        order = order[bd2[order] < 0.4 * 0.4][:20]
        assert cnt[i] == len(order)
        np.testing.assert_array_equal(idx[i, :cnt[i]], order)
        np.testing.assert_array_equal(d2[i, :cnt[i]], bd2[order])
        np.testing.assert_array_equal(np.sort(sets[i]), np.sort(order))
        assert idx[i, 0] == i and d2[i, 0] == 0.0          # the query point itself is included


def test_knn_radius_cut_is_strict_and_ties_lowest_index():
    pts = np.array([[0, 0, 0], [0.3, 0, 0], [0, 0.3, 0], [0.1, 0, 0], [-0.1, 0, 0], [0, 0.1, 0]], dtype=np.float64)
    idx, d2, cnt = orc.knn_hybrid(pts, pts[:1], 0.3, 10)
    # d = 0.3 exactly is excluded when (0.3*0.3 computed) is not < r*r
    assert set(idx[0, :cnt[0]]) == {0, 3, 4, 5}
    # ties at distance 0.1: ascending index
    assert list(idx[0, :cnt[0]]) == [0, 3, 4, 5]
    idx, d2, cnt = orc.knn_hybrid(pts, pts[:1], 0.3, 2)
    assert list(idx[0, :cnt[0]]) == [0, 3]                   # k-th place tie -> lowest index


# ------------------------------------------------------------------ normals
def test_eigen_solver_edge_cases():
    np.testing.assert_array_equal(orc.normal_from_covariance(np.eye(3)), [0, 0, 1])          # C = I  -> (0,0,1)
    np.testing.assert_array_equal(orc.normal_from_covariance(np.zeros((3, 3))), [0, 0, 1])   # zero -> fallback
    np.testing.assert_array_equal(orc.normal_from_covariance(np.diag([0.1, 2.0, 3.0])), [1, 0, 0])
    np.testing.assert_array_equal(orc.normal_from_covariance(np.diag([2.0, 0.1, 3.0])), [0, 1, 0])
    np.testing.assert_array_equal(orc.normal_from_covariance(np.diag([2.0, 3.0, 0.1])), [0, 0, 1])
    np.testing.assert_array_equal(orc.normal_from_covariance(np.diag([1.0, 1.0, 2.0])), [0, 0, 1])   # tie -> z


def test_eigen_solver_vs_lapack():
    rng = np.random.default_rng(2)
    for _ in range(300):
        A = rng.normal(size=(3, 3))
        w = np.sort(rng.uniform(0.01, 1.0, 3)) * np.array([1.0, 2.0, 4.0])
        Q, _ = np.linalg.qr(A)
        C = Q @ np.diag(w) @ Q.T
        C = (C + C.T) / 2
        n = orc.normal_from_covariance(C)
        assert abs(np.linalg.norm(n) - 1) < 1e-12
        assert abs(abs(n @ Q[:, 0]) - 1) < 1e-9


def test_normals_on_plane_and_isolated_points():
    rng = np.random.default_rng(3)
    # tilted plane, noise free: normal must be the plane normal (up to sign)
    nrm = np.array([0.3, -0.2, 0.933])
    nrm /= np.linalg.norm(nrm)
    u = np.cross(nrm, [1, 0, 0]); u /= np.linalg.norm(u)
    v = np.cross(nrm, u)
    ab = rng.uniform(-1, 1, size=(2000, 2))
    pts = (ab[:, :1] * u + ab[:, 1:] * v + np.array([5.0, 2.0, 1.0])).astype(np.float32).astype(np.float64)
    iso = np.array([[50.0, 50, 50], [60.0, 60, 60]])
    allp = np.vstack([pts, iso])
    n, cov, cnt = orc.estimate_normals(allp, 0.3, 300, return_cov=True)
    ok = cnt[:2000] >= 3
    assert ok.mean() > 0.99
    assert (np.abs(np.abs(n[:2000][ok] @ nrm) - 1) < 1e-5).all()
    np.testing.assert_array_equal(n[2000:], [[0, 0, 1], [0, 0, 1]])     # < 3 neighbours -> C = I -> (0,0,1)
    np.testing.assert_array_equal(cnt[2000:], [1, 1])
    np.testing.assert_array_equal(cov[2000], np.eye(3))


def test_normals_vs_numpy_scipy_on_lidar_scan():
    seq = synth.Sequence(1, synth.TINY_16)
    p, _ = orc.preprocess(seq.scans[0], method="icppointpoint")
    n, cov, cnt = orc.estimate_normals(p, 0.3, 30, return_cov=True)       # small k exercises the k-cap
    n2, gaps, cnt2 = ref.estimate_normals(p, 0.3, 30)
    np.testing.assert_array_equal(cnt, cnt2)
    assert (cnt == 30).any() and (cnt < 30).any()
    good = (gaps > 1e-3) & (cnt >= 3)
    assert good.sum() > 100
    assert (np.abs(np.abs((n[good] * n2[good]).sum(1)) - 1) < 1e-7).all()
    for i in np.where(cnt >= 3)[0][:50]:
        idx, _, c = orc.knn_hybrid(p, p[i:i + 1], 0.3, 30)
        np.testing.assert_allclose(cov[i], ref.covariance(p, idx[0, :c[0]]), rtol=0, atol=1e-10)


# ------------------------------------------------------------------ small dense algebra
def test_ldlt_and_rotation_update():
    rng = np.random.default_rng(4)
    for _ in range(50):
        J = rng.normal(size=(40, 6))
        A, b = J.T @ J, rng.normal(size=6)
        np.testing.assert_allclose(orc.ldlt_solve6(A, b), np.linalg.solve(A, b), rtol=1e-9, atol=1e-12)
    x = np.array([0.1, -0.2, 0.3, 1, 2, 3.0])
    T = orc.vec6_to_mat4(x)
    np.testing.assert_allclose(T[:3, :3], ref.rot_zyx(0.1, -0.2, 0.3), atol=1e-15)
    np.testing.assert_array_equal(T[:3, 3], [1, 2, 3])
    np.testing.assert_array_equal(T[3], [0, 0, 0, 1])


def test_svd3_vs_lapack():
    rng = np.random.default_rng(5)
    mats = [rng.normal(size=(3, 3)) for _ in range(100)]
    mats += [np.outer(rng.normal(size=3), rng.normal(size=3)), np.diag([3.0, 2.0, 0.0]), np.zeros((3, 3))]
    for A in mats:
        U, s, V = orc.svd3(A)
        np.testing.assert_allclose(U @ np.diag(s) @ V.T, A, atol=1e-12)
        np.testing.assert_allclose(U.T @ U, np.eye(3), atol=1e-12)
        np.testing.assert_allclose(V.T @ V, np.eye(3), atol=1e-12)
        np.testing.assert_allclose(s, np.linalg.svd(A, compute_uv=False), atol=1e-12)


# ------------------------------------------------------------------ correspondences and ICP
def _pair(sensor=synth.TINY_16, start=30.0):
    seq = synth.Sequence(2, sensor, start=start)
    tgt, ntgt = orc.preprocess(seq.scans[0])
    src, _ = orc.preprocess(seq.scans[1])
    return seq, src, tgt, ntgt


def test_correspondences_vs_bruteforce():
    seq, src, tgt, _ = _pair()
    T = seq.relative_odo(0, 1)
    corr, d2, fit, rmse = orc.correspondences(src, tgt, T, 1.0)
    s = src @ T[:3, :3].T + T[:3, 3]
    for i in range(0, len(s), 7):
        diff = s[i] - tgt
        b = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]
        j = int(np.argmin(b))                       # argmin returns the lowest index among ties
        if b[j] < 1.0:
            assert corr[i] == j and abs(d2[i] - b[j]) <= 1e-12 * max(1.0, b[j])
        else:
            assert corr[i] == -1
    K = (corr >= 0).sum()
    assert 0 < K < len(src)                           # the 1 m cut-off bites on this pair
    assert fit == K / len(src)
    assert abs(rmse - np.sqrt(d2[corr >= 0].sum() / K)) < 1e-14
    # empty inputs
    c0, _, f0, r0 = orc.correspondences(np.zeros((0, 3)), tgt, T, 1.0)
    assert len(c0) == 0 and f0 == 0 and r0 == 0


def test_icp_exact_recovery_noise_free():
    rng = np.random.default_rng(6)
    tgt = rng.uniform(-3, 3, size=(3000, 3)).astype(np.float32).astype(np.float64)
    tgt[:1000, 2] = 0            # some structure: a plane plus a volume cloud
    Tgt = synth.pose_matrix(0.05, -0.03, 0.02, 0.02, -0.01, 0.015)
    src = (tgt - Tgt[:3, 3]) @ Tgt[:3, :3]          # src = Tgt^-1 * tgt  ->  ICP must return Tgt
    n = orc.estimate_normals(tgt, 0.8, 30)
    for method, nn in ((orc.P2P, None), (orc.P2PLANE, n)):
        r = orc.icp(src, tgt, nn, np.eye(4), method, max_corr_dist=10.0)
        np.testing.assert_allclose(r.transformation, Tgt, atol=1e-9)
        assert r.fitness == 1.0 and r.inlier_rmse < 1e-9


@pytest.mark.parametrize("method", ["pointplane", "pointpoint"])
def test_icp_vs_numpy_scipy(method):
    seq, src, tgt, ntgt = _pair()
    init = seq.relative_odo(0, 1)
    m = orc.P2PLANE if method == "pointplane" else orc.P2P
    r = orc.icp(src, tgt, ntgt if m == orc.P2PLANE else None, init, m)
    T, fit, rmse, passes, corr = ref.icp(src, tgt, ntgt, init, method)
    assert passes == r.passes
    np.testing.assert_allclose(r.transformation, T, atol=1e-9)
    assert abs(fit - r.fitness) < 1e-12 and abs(rmse - r.inlier_rmse) < 1e-10
    assert (corr == r.correspondences).mean() > 0.9999
    # the criteria are Open3D's defaults: stop as soon as both deltas are < 1e-6, at most 30 updates
    assert r.updates <= 30 and r.passes == r.updates + 1


def test_icp_p2plane_invariant_to_normal_sign():
    seq, src, tgt, ntgt = _pair()
    init = seq.relative_odo(0, 1)
    flip = np.where(np.arange(len(tgt)) % 2 == 0, -1.0, 1.0)[:, None]
    r1 = orc.icp(src, tgt, ntgt, init, orc.P2PLANE)
    r2 = orc.icp(src, tgt, ntgt * flip, init, orc.P2PLANE)
    np.testing.assert_allclose(r1.transformation, r2.transformation, atol=1e-10)
    assert r1.passes == r2.passes


def test_icp_edge_cases():
    seq, src, tgt, ntgt = _pair()
    # nothing within the cut-off: fitness = rmse = 0, transformation = init after 1 update pass (both deltas 0 -> stop)
    far = np.eye(4)
    far[:3, 3] = (500, 0, 0)
    r = orc.icp(src, tgt, ntgt, far, orc.P2PLANE, max_corr_dist=1.0)
    assert r.fitness == 0 and r.inlier_rmse == 0 and r.n_corr == 0
    np.testing.assert_array_equal(r.transformation, far)
    assert r.passes == 2
    # max_iter = 0: one evaluation pass only
    r = orc.icp(src, tgt, ntgt, seq.relative_odo(0, 1), orc.P2PLANE, max_iter=0)
    assert r.passes == 1 and r.updates == 0
    # empty source
    r = orc.icp(np.zeros((0, 3)), tgt, ntgt, np.eye(4), orc.P2PLANE)
    assert r.fitness == 0 and r.n_corr == 0
