"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle on the same seeded
inputs.  Bar (BASELINE.json north_star): filter / voxel keys / correspondence indices bit-exact; final
transforms within 1e-4 rad and 1e-4 m; fitness and RMSE within 1e-5 relative."""
import os

import numpy as np
import pytest

from lidar_slam_arvc_b200 import engine, synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TOL_T = 1e-4          # metres / radians (north_star)
TOL_REL = 1e-5        # fitness, rmse relative (north_star)


@pytest.fixture(scope="module")
def eng():
    e = engine.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def seq16():
    return synth.Sequence(4, synth.TINY_16, start=30.0)


@pytest.fixture(scope="module")
def seq32():
    return synth.Sequence(3, synth.SMALL_32, start=12.0)


def rot_angle(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1) / 2
    return float(np.arccos(np.clip(c, -1, 1)))


def assert_transform_close(Tg, To):
    assert np.linalg.norm(Tg[:3, 3] - To[:3, 3]) < TOL_T
    assert rot_angle(Tg[:3, :3], To[:3, :3]) < TOL_T
    np.testing.assert_array_equal(Tg[3], [0, 0, 0, 1])


def assert_rel(a, b):
    assert abs(a - b) <= TOL_REL * max(abs(b), 1e-300), (a, b)


# ------------------------------------------------------------------------------------------ filter
def test_filter_bit_exact_golden(eng, golden_dir):
    g = np.load(os.path.join(golden_dir, "filter_radius_height.npz"))
    pts = g["points_f32"]
    eng.upload(900, pts)
    eng.preprocess([900], eng.make_preprocess_params(want_normals=False))
    idx = eng.get_filter_indices(900)
    np.testing.assert_array_equal(pts[idx].astype(np.float64), g["kept_default"])
    np.testing.assert_array_equal(eng.get_points(900), g["kept_default"])       # cloud order = filter order
    r, h = g["custom_radii"], g["custom_heights"]
    eng.preprocess([900], eng.make_preprocess_params(r[0], r[1], h[0], h[1], want_normals=False))
    np.testing.assert_array_equal(eng.get_points(900), g["kept_custom"])
    eng.free(900)


def test_filter_and_info_on_scan(eng, seq32):
    s = seq32.scans[0]
    eng.upload(1, s)
    eng.preprocess([1], eng.make_preprocess_params(want_normals=False))
    keep = orc.filter_radius_height(s.astype(np.float64))
    np.testing.assert_array_equal(eng.get_filter_indices(1), keep)
    info = eng.info(1)
    assert info == {"n_raw": len(s), "n_filtered": len(keep), "n_points": len(keep), "has_normals": False}
    # float64 upload path gives the same cloud
    eng.upload(2, s.astype(np.float64))
    eng.preprocess([2], eng.make_preprocess_params(want_normals=False))
    np.testing.assert_array_equal(eng.get_points(2), eng.get_points(1))
    eng.free(1)
    eng.free(2)


# ------------------------------------------------------------------------------------------ voxel
@pytest.mark.parametrize("voxel", [0.2, 0.5, 0.05])
def test_voxel_keys_and_means(eng, seq32, voxel):
    s = seq32.scans[1]
    eng.upload(3, s)
    eng.preprocess([3], eng.make_preprocess_params(voxel_size=voxel, want_normals=False))
    p = s.astype(np.float64)
    p = p[orc.filter_radius_height(p)]
    out, keys, cnt = orc.voxel_down_sample(p, voxel)
    gk, gc = eng.get_voxels(3)
    np.testing.assert_array_equal(gk, keys)            # integer voxel keys bit-exact, same (key-sorted) order
    np.testing.assert_array_equal(gc, cnt)
    np.testing.assert_array_equal(eng.get_points(3), out)   # means: same summation order -> bit-exact
    eng.free(3)


# ------------------------------------------------------------------------------------------ normals
@pytest.mark.parametrize("max_nn,radius", [(300, 0.3), (20, 0.3), (5, 1.0)])
def test_normals_vs_oracle(eng, seq32, max_nn, radius):
    s = seq32.scans[0]
    eng.upload(4, s)
    eng.preprocess([4], eng.make_preprocess_params(normal_radius=radius, max_nn=max_nn))
    pts, nrm = eng.get_points(4, normals=True)
    on, cov, cnt = orc.estimate_normals(pts, radius, max_nn, return_cov=True)
    np.testing.assert_array_equal(eng.get_nn_counts(4), cnt)          # same neighbour count after k / radius cut
    if max_nn <= 20:
        assert (cnt == max_nn).any()                                   # the k-cap bites
    # same formulas in float64 on both sides; ill-conditioned neighbourhoods (two smallest eigenvalues nearly equal,
    # e.g. collinear ring points) are re-summed in the oracle's order on the device, so they agree as well
    w = np.linalg.eigvalsh(cov)
    gap = (w[:, 1] - w[:, 0]) / np.maximum(w[:, 2], 1e-300)
    err = np.minimum(np.linalg.norm(nrm - on, axis=1), np.linalg.norm(nrm + on, axis=1))
    assert (np.abs(np.linalg.norm(nrm, axis=1) - 1) < 1e-9).all()
    assert (err[gap > 1e-9] < 1e-6).all()
    assert (err < 1e-6).mean() > 0.999
    few = cnt < 3
    np.testing.assert_array_equal(nrm[few], np.tile([0.0, 0.0, 1.0], (few.sum(), 1)))
    eng.free(4)


def test_normals_duplicates_and_tiny_clouds(eng):
    # many exact duplicates stress the k-th-distance tie path (lowest index wins) and the pathological bucket path
    rng = np.random.default_rng(0)
    base = rng.uniform(1.0, 1.6, size=(40, 3)).astype(np.float32)
    pts = np.vstack([base] * 12 + [np.array([[5.0, 5.0, 1.0], [5.0, 5.1, 1.0]], dtype=np.float32)])
    eng.upload(5, pts)
    eng.preprocess([5], eng.make_preprocess_params(normal_radius=0.3, max_nn=25))
    gp, gn = eng.get_points(5, normals=True)
    on, cov, cnt = orc.estimate_normals(gp, 0.3, 25, return_cov=True)
    np.testing.assert_array_equal(eng.get_nn_counts(5), cnt)
    w = np.linalg.eigvalsh(cov)
    ok = (w[:, 1] - w[:, 0]) / np.maximum(w[:, 2], 1e-300) > 1e-9
    assert (np.minimum(np.linalg.norm(gn - on, axis=1), np.linalg.norm(gn + on, axis=1))[ok] < 1e-6).all()
    np.testing.assert_array_equal(gn[-2:], [[0, 0, 1], [0, 0, 1]])
    eng.free(5)


def _assert_neighbor_sets(e, scan_id, pts, radius, max_nn, sample):
    """Neighbour index SETS of the normals, straight from the kernels' tap, against Open3D's hybrid search (the oracle)."""
    got = e.get_neighbors(scan_id, sample, max_nn)
    idx, _, cnt = orc.knn_hybrid(pts, pts[sample], radius, max_nn)
    for k in range(len(sample)):
        np.testing.assert_array_equal(got[k], np.sort(idx[k, :cnt[k]]), err_msg="point %d" % sample[k])


def test_normals_neighbor_sets_exact():
    """SURVEY.md §8(b) parity tap for row a4: the hybrid k-NN index sets are the oracle's, for the block kernel, the
    per-point kernel it hands points back to, the trial-radius path (k bites) and duplicate-heavy clouds (ties by index)."""
    e = engine.Engine(0)
    e.set_option("normals_tap", 1)
    try:
        seq = synth.Sequence(1, synth.SMALL_32, start=12.0)
        rng = np.random.default_rng(1)
        for max_nn, radius in [(300, 0.3), (20, 0.3), (5, 1.0)]:
            e.upload(1, seq.scans[0])
            e.preprocess([1], e.make_preprocess_params(normal_radius=radius, max_nn=max_nn))
            pts = e.get_points(1)
            sample = rng.choice(len(pts), size=400, replace=False).astype(np.int32)
            _assert_neighbor_sets(e, 1, pts, radius, max_nn, sample)
            c = e.get_counters(1)
            assert c["normals_per_point"] < c["n_points"]                       # the block kernel served (most of) them
        # float64 records (voxel means): the per-point kernel
        e.upload(2, seq.scans[0])
        e.preprocess([2], e.make_preprocess_params(voxel_size=0.1, normal_radius=0.3, max_nn=30))
        pts = e.get_points(2)
        _assert_neighbor_sets(e, 2, pts, 0.3, 30, rng.choice(len(pts), size=300, replace=False).astype(np.int32))
        # duplicate-heavy cloud: many exact distance ties at the k-th neighbour, broken by the lowest index
        base = rng.uniform(1.0, 1.6, size=(40, 3)).astype(np.float32)
        dup = np.vstack([base] * 12 + [np.array([[5.0, 5.0, 1.0], [5.0, 5.1, 1.0]], dtype=np.float32)])
        e.upload(3, dup)
        e.preprocess([3], e.make_preprocess_params(normal_radius=0.3, max_nn=25))
        pts = e.get_points(3)
        _assert_neighbor_sets(e, 3, pts, 0.3, 25, np.arange(len(pts), dtype=np.int32))
        # a full-size 64-beam scan at the reference's parameters (sampled)
        big = synth.Sequence(1, synth.OS1_64, start=30.0)
        e.upload(4, big.scans[0])
        e.preprocess([4], e.make_preprocess_params())
        pts = e.get_points(4)
        _assert_neighbor_sets(e, 4, pts, 0.3, 300, rng.choice(len(pts), size=300, replace=False).astype(np.int32))
        # the tap does not change the result: same normals as a context without it
        _, n_tap = e.get_points(4, normals=True)
        e2 = engine.Engine(0)
        e2.upload(4, big.scans[0])
        e2.preprocess([4], e2.make_preprocess_params())
        _, n_plain = e2.get_points(4, normals=True)
        e2.close()
        assert np.abs(n_tap - n_plain).max() < 1e-9
    finally:
        e.close()


# ------------------------------------------------------------------------------------------ ICP
def _run_pair(eng, seq, i, j, method, voxel=None, f64=False, init=None, **icp_kw):
    want_n = method == engine.P2PLANE
    for k in (i, j):
        eng.upload(k, seq.scans[k].astype(np.float64) if f64 else seq.scans[k])
    pp = eng.make_preprocess_params(voxel_size=voxel, want_normals=want_n)
    eng.preprocess([i, j], pp)
    init = seq.relative_odo(i, j) if init is None else init
    ip = eng.make_icp_params(method, **icp_kw)
    tr = eng.icp_trace(i, j, init, ip)
    tgt, tn = (eng.get_points(i, normals=True) if want_n else (eng.get_points(i), None))
    src = eng.get_points(j)
    # the device clouds ARE the oracle's clouds
    otgt, _ = orc.preprocess(seq.scans[i], voxel_size=voxel, method="icppointpoint")
    np.testing.assert_array_equal(tgt, otgt)
    on = orc.estimate_normals(otgt) if want_n else None
    osrc, _ = orc.preprocess(seq.scans[j], voxel_size=voxel, method="icppointpoint")
    ref = orc.icp(osrc, otgt, on, init, orc.P2PLANE if want_n else orc.P2P, **icp_kw)
    return tr, ref, src, tgt, tn


@pytest.mark.parametrize("method", [engine.P2PLANE, engine.P2P])
def test_icp_correspondences_exact_and_result(eng, seq16, method):
    tr, ref, src, tgt, _ = _run_pair(eng, seq16, 0, 1, method)
    # every pass: correspondence indices bit-exact against the oracle evaluated at the same transformation
    for k in range(tr["passes"]):
        corr, d2, fit, rmse = orc.correspondences(src, tgt, tr["T"][k], 10.0)
        np.testing.assert_array_equal(tr["corr"][k], corr)
        assert tr["fitness"][k] == fit
        assert_rel(tr["rmse"][k], rmse)
    res = tr["result"]
    assert res["passes"] == ref.passes and res["updates"] == ref.updates
    assert_transform_close(res["T"], ref.transformation)
    assert_rel(res["fitness"], ref.fitness)
    assert_rel(res["rmse"], ref.inlier_rmse)
    assert res["n_corr"] == ref.n_corr
    # and the whole trajectory of the iteration matches the oracle's
    np.testing.assert_allclose(tr["T"], ref.trace_T, atol=1e-9)
    np.testing.assert_allclose(tr["rmse"], ref.trace_rmse, rtol=1e-9)


@pytest.mark.parametrize("method", [engine.P2PLANE, engine.P2P])
def test_icp_small_cutoff_and_bad_init(eng, seq16, method):
    # a 0.5 m cut-off leaves many source points without correspondence; the init is 0.4 m / 3 deg off
    init = seq16.relative_gt(1, 2) @ synth.pose_matrix(0.3, -0.25, 0.05, np.deg2rad(3.0), 0.01, -0.01)
    tr, ref, src, tgt, _ = _run_pair(eng, seq16, 1, 2, method, init=init, max_corr_dist=0.5)
    assert 0 < tr["result"]["n_corr"] < len(src)
    for k in range(tr["passes"]):
        corr, _, fit, _ = orc.correspondences(src, tgt, tr["T"][k], 0.5)
        np.testing.assert_array_equal(tr["corr"][k], corr)
        assert tr["fitness"][k] == fit
    assert tr["result"]["passes"] == ref.passes
    assert_transform_close(tr["result"]["T"], ref.transformation)
    assert_rel(tr["result"]["rmse"], ref.inlier_rmse)


@pytest.mark.parametrize("method,voxel,f64", [(engine.P2PLANE, 0.2, False), (engine.P2P, 0.3, False), (engine.P2PLANE, None, True)])
def test_icp_wide_records(eng, seq32, method, voxel, f64):
    tr, ref, src, tgt, _ = _run_pair(eng, seq32, 0, 1, method, voxel=voxel, f64=f64)
    for k in (0, tr["passes"] - 1):
        corr, _, fit, _ = orc.correspondences(src, tgt, tr["T"][k], 10.0)
        np.testing.assert_array_equal(tr["corr"][k], corr)
    assert tr["result"]["passes"] == ref.passes
    assert_transform_close(tr["result"]["T"], ref.transformation)
    assert_rel(tr["result"]["fitness"], ref.fitness)
    assert_rel(tr["result"]["rmse"], ref.inlier_rmse)


def test_icp_batch_matches_single_and_is_deterministic(eng, seq16):
    for k in range(4):
        eng.upload(k, seq16.scans[k])
    eng.preprocess([0, 1, 2, 3], eng.make_preprocess_params())
    ip = eng.make_icp_params(engine.P2PLANE)
    tg, sr = [0, 1, 2, 0, 3], [1, 2, 3, 2, 0]
    init = np.array([seq16.relative_odo(a, b) for a, b in zip(tg, sr)])
    r1 = eng.icp_batch(tg, sr, init, ip)
    r2 = eng.icp_batch(tg, sr, init, ip)
    np.testing.assert_array_equal(r1["T"], r2["T"])                     # bit-reproducible run to run
    for k, (a, b) in enumerate(zip(tg, sr)):
        single = eng.icp_batch([a], [b], init[k:k + 1], ip)[0]
        np.testing.assert_array_equal(single["T"], r1["T"][k])
        tgt, tn = eng.get_points(a, normals=True)
        ref = orc.icp(eng.get_points(b), tgt, orc.estimate_normals(tgt), init[k], orc.P2PLANE)
        assert r1["updates"][k] == ref.updates
        assert_transform_close(r1["T"][k], ref.transformation)
        assert_rel(r1["rmse"][k], ref.inlier_rmse)


def test_icp_edge_cases(eng, seq16):
    for k in range(2):
        eng.upload(k, seq16.scans[k])
    eng.upload(7, np.zeros((0, 3), dtype=np.float32))                            # empty scan
    eng.upload(8, np.full((50, 3), 1000.0, dtype=np.float32))                    # everything filtered out
    eng.preprocess([0, 1, 7, 8], eng.make_preprocess_params())
    assert eng.info(7)["n_points"] == 0 and eng.info(8)["n_points"] == 0
    ip = eng.make_icp_params(engine.P2PLANE, max_corr_dist=1.0)
    far = np.eye(4)
    far[:3, 3] = (500.0, 0, 0)
    r = eng.icp_batch([0, 0, 0], [1, 7, 8], np.array([far, np.eye(4), np.eye(4)]), ip)
    # nothing within the cut-off: fitness = rmse = 0 and the transformation is the init (Open3D returns identity updates)
    assert r["fitness"][0] == 0 and r["rmse"][0] == 0 and r["n_corr"][0] == 0 and r["passes"][0] == 2
    np.testing.assert_array_equal(r["T"][0], far)
    assert (r["fitness"][1:] == 0).all() and (r["n_corr"][1:] == 0).all()
    # max_iter = 0: evaluation only
    r0 = eng.icp_batch([0], [1], seq16.relative_odo(0, 1)[None], eng.make_icp_params(engine.P2PLANE, max_iter=0))[0]
    assert r0["passes"] == 1 and r0["updates"] == 0
    np.testing.assert_array_equal(r0["T"], seq16.relative_odo(0, 1))
    # empty target
    r1 = eng.icp_batch([7], [1], np.eye(4)[None], ip)[0]
    assert r1["fitness"] == 0 and r1["n_corr"] == 0


def test_error_behaviour(eng, seq16):
    eng.upload(0, seq16.scans[0])
    eng.upload(1, seq16.scans[1])
    eng.preprocess([0, 1], eng.make_preprocess_params(want_normals=False))
    with pytest.raises(engine.EngineError, match="normals"):
        eng.icp_batch([0], [1], np.eye(4)[None], eng.make_icp_params(engine.P2PLANE))
    with pytest.raises(engine.EngineError, match="not uploaded"):
        eng.icp_batch([0], [12345], np.eye(4)[None], eng.make_icp_params(engine.P2P))
    with pytest.raises(engine.EngineError, match="unknown method"):
        eng.icp_batch([0], [1], np.eye(4)[None], eng.make_icp_params(7))
    with pytest.raises(engine.EngineError, match="unknown scan"):
        eng.preprocess([777], eng.make_preprocess_params())
    eng.free(0)
    with pytest.raises(engine.EngineError):
        eng.icp_batch([0], [1], np.eye(4)[None], eng.make_icp_params(engine.P2P))


def test_full_size_64_beam_pair(eng):
    """BASELINE config 2 shape: one OS1-64 pair, point-to-plane, against the oracle."""
    seq = synth.Sequence(2, synth.OS1_64, start=30.0)
    tr, ref, src, tgt, tn = _run_pair(eng, seq, 0, 1, engine.P2PLANE)
    for k in range(tr["passes"]):                                       # EVERY pass: correspondence indices bit-exact
        corr, _, fit, rmse = orc.correspondences(src, tgt, tr["T"][k], 10.0)
        np.testing.assert_array_equal(tr["corr"][k], corr)
        assert tr["fitness"][k] == fit
    on, cov, cnt = orc.estimate_normals(tgt, return_cov=True)
    np.testing.assert_array_equal(eng.get_nn_counts(0), cnt)
    # normals: every point with a resolvable smallest eigen-direction agrees to 1e-6 (gap-conditioned, like the small test)
    w = np.linalg.eigvalsh(cov)
    gap = (w[:, 1] - w[:, 0]) / np.maximum(w[:, 2], 1e-300)
    err = np.minimum(np.linalg.norm(tn - on, axis=1), np.linalg.norm(tn + on, axis=1))
    assert (err[gap > 1e-9] < 1e-6).all(), (err[gap > 1e-9].max(), int((err[gap > 1e-9] >= 1e-6).sum()))
    assert (err < 1e-6).mean() > 0.999
    assert tr["result"]["passes"] == ref.passes
    assert_transform_close(tr["result"]["T"], ref.transformation)
    assert_rel(tr["result"]["fitness"], ref.fitness)
    assert_rel(tr["result"]["rmse"], ref.inlier_rmse)


# ------------------------------------------------------------------------------------------ BASELINE configs 2-4 shapes
def test_128_beam_point_to_point_pair(eng):
    """BASELINE config 3 shape: 128-beam (~260k points) point-to-point pair against the oracle."""
    seq = synth.Sequence(2, synth.OS_128, start=30.0)
    tr, ref, src, tgt, _ = _run_pair(eng, seq, 0, 1, engine.P2P)
    assert len(src) > 200000
    for k in sorted({0, 1, 2, tr["passes"] // 2, tr["passes"] - 2, tr["passes"] - 1}):      # cold, warm, middle and last passes
        corr, _, fit, _ = orc.correspondences(src, tgt, tr["T"][k], 10.0)
        np.testing.assert_array_equal(tr["corr"][k], corr)
        assert tr["fitness"][k] == fit
    assert tr["result"]["passes"] == ref.passes
    assert_transform_close(tr["result"]["T"], ref.transformation)
    assert_rel(tr["result"]["fitness"], ref.fitness)
    assert_rel(tr["result"]["rmse"], ref.inlier_rmse)
    eng.free(0)
    eng.free(1)


def test_loop_closure_batch_properties(eng):
    """BASELINE config 4 shape (scaled down): loop-closure pairs with perturbed initial guesses, one batch.
    Size-independent properties: a sample agrees with the oracle; a pair's result is independent of the rest of the
    batch and of its position in it; sharding the batch does not change any result."""
    from lidar_slam_arvc_b200 import sharding
    seq = synth.Sequence(24, synth.SMALL_32, start=0.0, step=1.0)
    pairs = synth.loop_closure_pairs(seq.poses, 40, radius=5.0, min_gap=3, seed=777, sigma_t=0.1, sigma_rot_deg=1.0)
    assert len(pairs) == 40
    for k in range(len(seq.scans)):
        eng.upload(100 + k, seq.scans[k])
    eng.preprocess([100 + k for k in range(len(seq.scans))], eng.make_preprocess_params())
    order = sharding.sort_pairs_for_cache([p[0] for p in pairs], [p[1] for p in pairs])
    tg = np.array([100 + pairs[k][0] for k in order])
    sr = np.array([100 + pairs[k][1] for k in order])
    init = np.array([pairs[k][2] for k in order])
    ip = eng.make_icp_params(engine.P2PLANE)
    res = eng.icp_batch(tg, sr, init, ip)
    assert (res["fitness"] > 0.9).all() and (res["updates"] <= 30).all()
    # oracle on a sample
    for k in (0, 7, 23, 39):
        tgt, tn = eng.get_points(int(tg[k]), normals=True)
        ref = orc.icp(eng.get_points(int(sr[k])), tgt, orc.estimate_normals(tgt), init[k], orc.P2PLANE)
        assert res["updates"][k] == ref.updates
        assert_transform_close(res["T"][k], ref.transformation)
        assert_rel(res["rmse"][k], ref.inlier_rmse)
    # a pair's result does not depend on what else is in the batch or on its position in it
    perm = np.random.default_rng(5).permutation(len(tg))
    shuffled = eng.icp_batch(tg[perm], sr[perm], init[perm], ip)
    np.testing.assert_array_equal(shuffled["T"], res["T"][perm])
    np.testing.assert_array_equal(shuffled["updates"], res["updates"][perm])
    # sharding: two "ranks" process contiguous halves; concatenation equals the single batch bit for bit
    parts = []
    for r in range(2):
        lo, hi = sharding.shard_bounds(len(tg), 2, r)
        parts.append(eng.icp_batch(tg[lo:hi], sr[lo:hi], init[lo:hi], ip))
    merged = np.concatenate(parts)
    np.testing.assert_array_equal(merged["T"], res["T"])
    np.testing.assert_array_equal(merged["rmse"], res["rmse"])
    # batches of 24 pairs and more run the search / far-query kernels compiled for more resident blocks (other register
    # budgets, same source): the list twice over is such a batch, and every result is the same bit for bit
    big = eng.icp_batch(np.tile(tg, 2), np.tile(sr, 2), np.tile(init, (2, 1, 1)), ip)
    for half in (big[:len(tg)], big[len(tg):]):
        np.testing.assert_array_equal(half["T"], res["T"])
        np.testing.assert_array_equal(half["rmse"], res["rmse"])
        np.testing.assert_array_equal(half["n_corr"], res["n_corr"])
        np.testing.assert_array_equal(half["updates"], res["updates"])
    for k in range(len(seq.scans)):
        eng.free(100 + k)


@pytest.mark.parametrize("src_f64,tgt_f64", [(True, False), (False, True)])
def test_icp_mixed_record_types(eng, seq16, src_f64, tgt_f64):
    """Source and target stored with different record widths (float64 upload vs float32 PCD payload)."""
    eng.upload(40, seq16.scans[0].astype(np.float64) if tgt_f64 else seq16.scans[0])
    eng.upload(41, seq16.scans[1].astype(np.float64) if src_f64 else seq16.scans[1])
    eng.preprocess([40, 41], eng.make_preprocess_params())
    init = seq16.relative_odo(0, 1)
    for method in (engine.P2PLANE, engine.P2P):
        tr = eng.icp_trace(40, 41, init, eng.make_icp_params(method))
        tgt, tn = eng.get_points(40, normals=True)
        src = eng.get_points(41)
        for k in (0, tr["passes"] - 1):
            corr, _, fit, _ = orc.correspondences(src, tgt, tr["T"][k], 10.0)
            np.testing.assert_array_equal(tr["corr"][k], corr)
        ref = orc.icp(src, tgt, orc.estimate_normals(tgt) if method == engine.P2PLANE else None, init,
                      orc.P2PLANE if method == engine.P2PLANE else orc.P2P)
        assert tr["result"]["passes"] == ref.passes
        assert_transform_close(tr["result"]["T"], ref.transformation)
        assert_rel(tr["result"]["rmse"], ref.inlier_rmse)
    eng.free(40)
    eng.free(41)


@pytest.mark.parametrize("voxel", [None, 0.3])
def test_map_build_bit_exact(eng, seq32, voxel):
    """filter -> [voxel] -> rigid transform -> concatenation of a batch of keyframes, against the
    oracle's restatement of keyframemanager.py:154-184 (bit-exact: same operation order, no FMA)."""
    ids = [7000 + k for k in range(len(seq32.scans))]
    for k, s in zip(ids, seq32.scans):
        eng.upload(k, s)
    eng.upload(7100, seq32.scans[1].astype(np.float64) * 1.0000001)          # float64 records
    eng.upload(7101, np.zeros((0, 3), dtype=np.float32))                      # empty keyframe
    order = ids + [7100, 7101, ids[0]]                                        # a keyframe may appear twice
    scans = list(seq32.scans) + [seq32.scans[1].astype(np.float64) * 1.0000001, np.zeros((0, 3), dtype=np.float32), seq32.scans[0]]
    rng = np.random.default_rng(5)
    Ts = [synth.pose_matrix(*rng.uniform(-20, 20, 3), *rng.uniform(-3, 3, 3)) for _ in order]
    p = eng.make_preprocess_params(0.5, 35.0, -120.0, 120.0, voxel_size=voxel, want_normals=False)
    xyz, off = eng.map_build(order, Ts, p)
    want, woff = orc.build_map(scans, Ts, voxel_size=voxel)
    np.testing.assert_array_equal(off, woff)
    np.testing.assert_array_equal(xyz, want)
    assert off[-2] == off[-3]                                                 # the empty keyframe contributes nothing
    # too small an output: error + valid offsets, nothing written past the capacity
    import ctypes
    buf = np.full((10, 3), -7.0)
    o2 = np.zeros(len(order) + 1, dtype=np.int64)
    ids_a = np.array(order, dtype=np.int64)
    T_a = np.ascontiguousarray(Ts)
    rc = eng.lib.arvc_map_build(eng.h, len(order), ids_a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                T_a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.byref(p),
                                buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 10, o2.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    assert rc != 0 and b"capacity" in eng.lib.arvc_last_error(eng.h)
    np.testing.assert_array_equal(o2, woff)
    assert (buf == -7.0).all()
    # the scans stay usable for registration afterwards (re-preprocessed with the ICP parameters)
    eng.preprocess(ids[:2], eng.make_preprocess_params())
    r = eng.icp_batch([ids[0]], [ids[1]], seq32.relative_odo(0, 1)[None], eng.make_icp_params())
    assert r["fitness"][0] > 0.9
    for k in ids + [7100, 7101]:
        eng.free(k)


@pytest.mark.parametrize("voxel", [None, 0.25])
def test_plane_fit_and_split_bit_exact(eng, seq32, voxel):
    """SURVEY.md §8 f-2 pieces: reproducible RANSAC plane (keyframe.py:417-436) and the split by plane distance
    (keyframe.py:438-461), against the oracle - same hash samples, same operation order, so bit-exact."""
    eng.upload(7300, seq32.scans[0])
    eng.preprocess([7300], eng.make_preprocess_params(voxel_size=voxel, want_normals=False))
    pts = eng.get_points(7300)
    for seed in (0, 5):
        pl, n_in = eng.fit_plane(7300, -0.5, 0.01, 1000, seed)
        wpl, wn = orc.fit_plane(pts, -0.5, 0.01, 1000, seed)
        np.testing.assert_array_equal(pl, wpl)
        assert n_in == wn > 100
    n_near, n_far = eng.split_plane(7300, pl, 0.4, 7301, 7302)
    near, far = orc.segment_plane(pts, pl, 0.4)
    assert (n_near, n_far) == (len(near), len(far)) and n_near > 0 and n_far > 0
    p = eng.make_preprocess_params(0.0, 35.0, -1.0, 50.0, want_normals=True, normal_radius=0.5)
    p.min_radius2 = -1.0
    eng.preprocess([7301, 7302], p)
    g, gn = eng.get_points(7301, normals=True)
    r = eng.get_points(7302)
    np.testing.assert_array_equal(g, pts[near])
    np.testing.assert_array_equal(r, pts[far])
    on = orc.estimate_normals(pts[near], 0.5, 300)
    err = np.minimum(np.linalg.norm(gn - on, axis=1), np.linalg.norm(gn + on, axis=1))
    assert err.max() < 1e-6
    # a plane far away from every point: everything lands in the second scan; no usable points -> error
    assert eng.split_plane(7300, [0.0, 0.0, 1.0, -100.0], 0.4, 7301, 7302) == (0, len(pts))
    assert eng.info(7302)["n_raw"] == len(pts) and eng.info(7301)["n_raw"] == 0
    with pytest.raises(engine.EngineError):
        eng.fit_plane(7300, -50.0, 0.01, 100, 0)
    for k in (7300, 7301, 7302):
        eng.free(k)
