"""Independent numpy + scipy restatement of the hot path — TEST INFRASTRUCTURE ONLY.

Second implementation used to cross-check ``icp_oracle.cpp`` so that the oracle is not self-certified
(SURVEY.md §8c).  It deliberately uses different building blocks: ``scipy.spatial.cKDTree`` for neighbour
search, ``numpy.linalg.eigh`` / ``svd`` / ``solve`` for the dense algebra.  Pure-numpy, small inputs only.
"""
import numpy as np
from scipy.spatial import cKDTree


def filter_radius_height(points, min_radius=0.5, max_radius=35, min_height=-1.0, max_height=50.0):
    """1:1 with the reference's own numpy (keyframemanager/keyframe.py:88-92)."""
    points = np.asarray(points, dtype=np.float64)
    [x, y, z] = points[:, 0], points[:, 1], points[:, 2]
    r2 = x ** 2 + y ** 2
    idx2 = np.where((r2 < max_radius ** 2) & (r2 > min_radius ** 2) & (z > min_height) & (z < max_height))
    return idx2[0]


def voxel_down_sample(points, voxel_size):
    """Open3D VoxelDownSample as a dict {key: (mean, count)} (order-free comparison)."""
    points = np.asarray(points, dtype=np.float64)
    origin = points.min(axis=0) - voxel_size * 0.5
    keys = np.floor((points - origin) / voxel_size).astype(np.int64)
    out = {}
    for k, p in zip(map(tuple, keys), points):
        if k in out:
            out[k][0] += p
            out[k][1] += 1
        else:
            out[k] = [p.copy(), 1]
    return {k: (s / c, c) for k, (s, c) in out.items()}


def knn_hybrid_sets(points, queries, radius, max_nn):
    """Neighbour index sets of SearchHybrid via cKDTree (k-NN then strict radius cut).  Ties at the k-th
    distance are resolved towards the lowest index, like the oracle."""
    points = np.asarray(points, dtype=np.float64)
    tree = cKDTree(points, leafsize=15)
    res = []
    r2 = radius * radius
    for q in np.asarray(queries, dtype=np.float64):
        cand = np.array(tree.query_ball_point(q, radius * (1 + 1e-9) + 1e-12), dtype=np.int64)
        if len(cand) == 0:
            res.append(np.empty(0, dtype=np.int64))
            continue
        d = points[cand] - q
        dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        keep = d2 < r2
        cand, d2 = cand[keep], d2[keep]
        order = np.lexsort((cand, d2))[:max_nn]
        res.append(cand[order])
    return res


def covariance(points, idx):
    p = points[idx]
    mu = p.mean(axis=0)
    return (p[:, :, None] * p[:, None, :]).mean(axis=0) - np.outer(mu, mu)


def normal_eigh(cov):
    """Smallest-eigenvalue eigenvector by LAPACK (sign-free reference for well-separated eigenvalues)."""
    w, v = np.linalg.eigh(cov)
    return v[:, 0], w


def estimate_normals(points, radius=0.3, max_nn=300):
    points = np.asarray(points, dtype=np.float64)
    sets = knn_hybrid_sets(points, points, radius, max_nn)
    nrm = np.zeros_like(points)
    gaps = np.zeros(len(points))
    for i, s in enumerate(sets):
        if len(s) >= 3:
            n, w = normal_eigh(covariance(points, s))
            nrm[i] = n
            gaps[i] = (w[1] - w[0]) / max(w[2], 1e-300)
        else:
            nrm[i] = (0, 0, 1)
            gaps[i] = 0.0
    return nrm, gaps, np.array([len(s) for s in sets])


def rot_zyx(a, b, g):
    ca, sa, cb, sb, cg, sg = np.cos(a), np.sin(a), np.cos(b), np.sin(b), np.cos(g), np.sin(g)
    Rx = np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]])
    Ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])
    Rz = np.array([[cg, -sg, 0], [sg, cg, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def correspondences(src_t, tree, tgt, max_dist):
    d, j = tree.query(src_t, k=1)
    # recompute d2 the oracle's way and apply the strict cut
    diff = src_t - tgt[np.minimum(j, len(tgt) - 1)]
    d2 = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]
    ok = np.isfinite(d) & (d2 < max_dist * max_dist)
    return np.where(ok, j, -1), np.where(ok, d2, 0.0)


def icp(source, target, target_normals=None, init=None, method="pointplane", max_corr_dist=10.0, rel_fitness=1e-6,
        rel_rmse=1e-6, max_iter=30):
    """Open3D RegistrationICP restated with numpy/scipy blocks."""
    src = np.asarray(source, dtype=np.float64)
    tgt = np.asarray(target, dtype=np.float64)
    T = np.eye(4) if init is None else np.array(init, dtype=np.float64)
    tree = cKDTree(tgt, leafsize=15)
    pcd = src @ T[:3, :3].T + T[:3, 3]

    def evaluate(p):
        corr, d2 = correspondences(p, tree, tgt, max_corr_dist)
        K = int((corr >= 0).sum())
        if K == 0:
            return corr, 0.0, 0.0
        return corr, K / len(p), float(np.sqrt(d2.sum() / K))

    corr, fit, rmse = evaluate(pcd)
    passes = 1
    for _ in range(max_iter):
        m = corr >= 0
        upd = np.eye(4)
        if m.any():
            s = pcd[m]
            t = tgt[corr[m]]
            if method == "pointplane":
                n = target_normals[corr[m]]
                r = ((s - t) * n).sum(axis=1)
                J = np.hstack([np.cross(s, n), n])
                x = np.linalg.solve(J.T @ J, -(J.T @ r))
                upd[:3, :3] = rot_zyx(x[0], x[1], x[2])
                upd[:3, 3] = x[3:]
            else:
                ms, mt = s.mean(axis=0), t.mean(axis=0)
                sigma = (t - mt).T @ (s - ms) / len(s)
                U, _, Vt = np.linalg.svd(sigma)
                S = np.eye(3)
                if np.linalg.det(U) * np.linalg.det(Vt) < 0:
                    S[2, 2] = -1
                R = U @ S @ Vt
                upd[:3, :3] = R
                upd[:3, 3] = mt - R @ ms
        T = upd @ T
        pcd = pcd @ upd[:3, :3].T + upd[:3, 3]
        bfit, brmse = fit, rmse
        corr, fit, rmse = evaluate(pcd)
        passes += 1
        if abs(bfit - fit) < rel_fitness and abs(brmse - rmse) < rel_rmse:
            break
    return T, fit, rmse, passes, corr
