"""The C ray caster of the synthetic world (csrc/synth_cast.c) against the numpy implementation it accelerates: the scans
the benchmark and the parity tests run on must not depend on which one produced them."""
import numpy as np
import pytest

from lidar_slam_arvc_b200 import synth


@pytest.mark.skipif(synth._cast_lib() is None, reason="libarvc_synth.so not built (python -c 'import __graft_entry__ as g; g.build()')")
def test_c_caster_is_bit_identical_to_numpy():
    w = synth.World(1234)
    poses = synth.loop_trajectory(w, 40, start=30.0, step=3.0)
    for sensor in (synth.TINY_16, synth.SMALL_32):
        d_s = synth.sensor_directions(sensor)
        for T in poses[::7]:
            d_w = d_s @ T[:3, :3].T
            a, b = w.cast(T[:3, 3], d_w), w.cast_numpy(T[:3, 3], d_w)
            np.testing.assert_array_equal(a, b)
            assert np.isfinite(a).mean() > 0.9
    # axis-aligned rays (zero direction components: 1 / 0 = inf, 0 * inf = NaN inside the slab test)
    d = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, -1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.0, 1.0]])
    o = np.array([19.0, 0.3, 0.69])
    with np.errstate(all="ignore"):
        np.testing.assert_array_equal(w.cast(o, d), w.cast_numpy(o, d))


def test_sequence_is_reproducible_with_and_without_workers():
    a = synth.Sequence(4, synth.TINY_16, start=30.0)
    b = synth.Sequence(4, synth.TINY_16, start=30.0, workers=2)
    for x, y in zip(a.scans, b.scans):
        np.testing.assert_array_equal(x, y)
