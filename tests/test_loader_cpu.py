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
    assert h is None and ld.stats == {"read_ahead_hits": 0, "reads": 1}
    ld.prefetch(files[1])
    ld.prefetch(files[1])                                    # announced twice: one read
    ld.prefetch(str(tmp_path / "missing.pcd"))               # a file that does not exist is not an error here
    assert list(ld.pending) == [files[1]]
    b, _ = ld.fetch(files[1])
    np.testing.assert_array_equal(b, clouds[1])
    assert ld.stats == {"read_ahead_hits": 1, "reads": 2} and not ld.pending
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
        assert ld.stats == {"read_ahead_hits": 1, "reads": 2}
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
