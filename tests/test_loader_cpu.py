"""Load path of the drop-in on CPU (no pinned pool without a GPU): read-ahead bookkeeping of ScanLoader and the keyframe
cache of KeyFrame.load_pointcloud, with the oracle test double as engine."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
DROPIN = os.path.join(ROOT, "lidar_slam_arvc_b200", "dropin")

from lidar_slam_arvc_b200 import euroc_synth, loader, pcd, runtime, synth  # noqa: E402


def test_scan_loader_read_ahead_and_release(tmp_path):
    from fake_engine import OracleEngine
    files = []
    rng = np.random.default_rng(0)
    clouds = [rng.normal(size=(50 + k, 3)).astype(np.float32) for k in range(4)]
    for k, c in enumerate(clouds):
        files.append(str(tmp_path / ("s%d.pcd" % k)))
        pcd.write_pcd_xyz(files[-1], c)
    ld = loader.ScanLoader(OracleEngine())
    assert ld.pool is None                                   # the test double has no pinned pool: plain numpy arrays
    a, h = ld.fetch(files[0])
    np.testing.assert_array_equal(a, clouds[0])
    assert h is None and (ld.stats["read_ahead_hits"], ld.stats["reads"]) == (0, 1)
    ld.prefetch(files[1])
    ld.prefetch(files[1])                                    # announced twice: one read
    ld.prefetch(str(tmp_path / "missing.pcd"))               # a file that does not exist is not an error here
    assert list(ld.pending) == [files[1]]
    b, _ = ld.fetch(files[1])
    np.testing.assert_array_equal(b, clouds[1])
    assert (ld.stats["read_ahead_hits"], ld.stats["reads"]) == (1, 2) and not ld.pending
    for f in files:                                          # a caller that never comes back: bounded backlog
        ld.prefetch(f)
    ld.prefetch(files[0])
    assert len(ld.pending) <= 4
    ld.close()
    assert not ld.pending


def test_keyframe_cache_and_read_ahead_through_the_dropin(tmp_path):
    from fake_engine import OracleEngine
    saved_path, saved_mods = list(sys.path), set(sys.modules)
    sys.path.insert(0, DROPIN)
    eng = OracleEngine()
    runtime.set_engine(eng)
    try:
        import keyframemanager.keyframemanager as kfm
        seq = synth.Sequence(4, synth.TINY_16, start=30.0)
        d = str(tmp_path / "euroc")
        times = euroc_synth.write_euroc_tree(d, seq)
        km = kfm.KeyFrameManager(directory=d, scan_times=times, voxel_size=None, method="icppointplane")
        ld = runtime.get_loader()
        km.add_keyframe(0)
        km.load_pointcloud(0)                                # sequential caller: scan 1 is read ahead
        assert ld.stats["reads"] == 1 and len(ld.pending) == 1
        km.add_keyframe(1)
        km.load_pointcloud(1)
        assert (ld.stats["read_ahead_hits"], ld.stats["reads"]) == (1, 2)
        np.testing.assert_array_equal(km.keyframes[1].pointcloud.points, seq.scans[1])
        km.pre_process(0)
        n_pre = len([c for c in eng.calls if c[0] == "preprocess"])
        km.load_pointcloud(0)                                # still resident: neither read nor uploaded again
        assert ld.stats["reads"] == 2 and km.keyframes[0]._preprocessed_on_device
        km.pre_process(0)
        assert len([c for c in eng.calls if c[0] == "preprocess"]) == n_pre + 1      # (the engine's own cache decides; the double counts calls)
        km.unload_pointcloud(0)
        km.load_pointcloud(0)                                # unloaded: read again
        assert ld.stats["reads"] == 3 and not km.keyframes[0]._preprocessed_on_device
        # LRU cap on resident keyframes
        km.max_resident_keyframes = 2
        km.add_keyframe(2)
        km.load_pointcloud(2)
        assert list(km._resident) == [0, 2] and km.keyframes[1].pointcloud is None
    finally:
        runtime.set_engine(None)
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k not in saved_mods:
                del sys.modules[k]


def test_sequential_caller_gets_the_next_scan_preprocessed_ahead(tmp_path):
    """run_scanmatcher.py:196-213 on the drop-in: while a pair is being registered the loader uploads and preprocesses the
    scan read ahead (engine.preprocess_ahead); the keyframe that loads it next adopts that device scan, and its own
    pre_process() finds the work done.  Transforms equal those of a run without the look-ahead."""
    from fake_engine import OracleEngine
    from lidar_slam_arvc_b200.homogeneousmatrix import HomogeneousMatrix
    saved_path, saved_mods = list(sys.path), set(sys.modules)
    sys.path.insert(0, DROPIN)
    seq = synth.Sequence(5, synth.TINY_16, start=30.0)
    d = str(tmp_path / "euroc")
    times = euroc_synth.write_euroc_tree(d, seq)
    odo = [HomogeneousMatrix(seq.relative_odo(i, i + 1)) for i in range(4)]
    results = {}
    try:
        for ahead in (True, False):
            eng = OracleEngine()
            if not ahead:
                eng.preprocess_ahead = None              # hasattr() stays true, so switch the loader off instead
            runtime.set_engine(eng)
            import keyframemanager.keyframemanager as kfm
            ld = runtime.get_loader()
            if not ahead:
                ld.stage_ahead = lambda *a, **k: None
            else:
                patient = ld.stage_ahead                 # the product waits 0.5 ms for the read-ahead thread and otherwise skips
                ld.stage_ahead = lambda alloc: patient(alloc, wait_s=10.0)      # the look-ahead; a test must not depend on that race
            km = kfm.KeyFrameManager(directory=d, scan_times=times, voxel_size=None, method="icppointplane")
            km.add_keyframe(0)
            km.load_pointcloud(0)
            km.pre_process(0)
            Ts = []
            for i in range(4):
                km.add_keyframe(i + 1)
                km.load_pointcloud(i + 1)
                km.pre_process(i + 1)
                Ts.append(km.compute_transformation(i, i + 1, Tij=odo[i]).array.copy())
                km.unload_pointcloud(i)
            km.unload_pointcloud(4)
            results[ahead] = np.array(Ts)
            if ahead:
                # scans 2, 3, 4 were staged during the registrations of (0,1), (1,2), (2,3) and adopted afterwards
                assert ld.stats["staged"] == 3 and ld.stats["staged_hits"] == 3 and not ld.staged
                assert [c for c in eng.calls if c[0] == "preprocess_ahead"] == [("preprocess_ahead", 1)] * 3
                assert not eng.raw and not eng.pre       # everything unloaded again: no staged scan left behind
            else:
                assert ld.stats["staged"] == 0
            runtime.set_engine(None)
        np.testing.assert_array_equal(results[True], results[False])
    finally:
        runtime.set_engine(None)
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k not in saved_mods:
                del sys.modules[k]
