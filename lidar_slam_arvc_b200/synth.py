"""Seeded synthetic LiDAR scans of the shapes BASELINE.json names.

The reference ships no data (its datasets live on the authors' disk, run_scanmatcher.py:146), so the
benchmark and the parity tests ray-cast a fixed world: ground plane (z = -0.69 m in the sensor frame, cf.
the reference's own ground model keyframe.py:436), a walled loop corridor with 4 m walls, random boxes
and vertical cylinders.  Pure numpy; runs on the host only.
"""
import ctypes
import os
from dataclasses import dataclass

import numpy as np

SENSOR_HEIGHT = 0.69

_CAST_LIB = [False]


def _cast_lib():
    """libarvc_synth.so (built next to libarvc_icp.so by __graft_entry__.build()), or None: then numpy does the casting."""
    if _CAST_LIB[0] is False:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libarvc_synth.so")
        lib = None
        if os.path.exists(path) and not os.environ.get("ARVC_SYNTH_NUMPY"):
            try:
                lib = ctypes.CDLL(path)
                dp = ctypes.POINTER(ctypes.c_double)
                lib.arvc_synth_cast.argtypes = [dp, dp, ctypes.c_long, dp, ctypes.c_int, ctypes.c_double, dp, ctypes.c_int, dp, ctypes.c_int, dp]
                lib.arvc_synth_cast.restype = None
            except OSError:
                lib = None
        _CAST_LIB[0] = lib
    return _CAST_LIB[0]


@dataclass(frozen=True)
class Sensor:
    beams: int
    azimuth_steps: int
    elev_min_deg: float
    elev_max_deg: float
    min_range: float = 0.3
    max_range: float = 100.0
    sigma_range: float = 0.01
    dropout: float = 0.02

    @property
    def n_rays(self):
        return self.beams * self.azimuth_steps


OS1_64 = Sensor(64, 1024, -16.6, 16.6)           # 65 536 rays  (config 1/2/4)
OS_128 = Sensor(128, 2048, -22.5, 22.5)          # 262 144 rays (config 3)
TINY_16 = Sensor(16, 256, -16.6, 16.6)           # 4 096 rays   (unit tests)
SMALL_32 = Sensor(32, 512, -16.6, 16.6)          # 16 384 rays  (unit tests)


class World:
    """Static scene: ground z=0, outer room + inner island (loop corridor), boxes, cylinders."""

    def __init__(self, seed=1234, n_boxes=40, n_cylinders=20, outer=(44.0, 24.0), inner=(34.0, 14.0), wall_h=4.0):
        rng = np.random.default_rng(seed)
        self.wall_h = wall_h
        ox, oy = outer[0] / 2, outer[1] / 2
        ix, iy = inner[0] / 2, inner[1] / 2
        self.outer, self.inner = (ox, oy), (ix, iy)
        segs = []
        for (hx, hy) in ((ox, oy), (ix, iy)):
            c = [(-hx, -hy), (hx, -hy), (hx, hy), (-hx, hy)]
            for k in range(4):
                segs.append((*c[k], *c[(k + 1) % 4]))
        self.segments = np.array(segs, dtype=np.float64)           # [8,4] x0 y0 x1 y1
        # obstacles live in the corridor, hugging the walls so that the centre line stays free
        boxes, cyls = [], []

        def corridor_point(margin):
            while True:
                x = rng.uniform(-ox + margin, ox - margin)
                y = rng.uniform(-oy + margin, oy - margin)
                inside_inner = abs(x) < ix + margin and abs(y) < iy + margin
                # keep a 2 m wide free lane around the centre-line rectangle of the 5 m corridor
                cx, cy = (ox + ix) / 2, (oy + iy) / 2
                on_lane = (abs(abs(x) - cx) < 1.0 and abs(y) < cy + 1.0) or (abs(abs(y) - cy) < 1.0 and abs(x) < cx + 1.0)
                if not inside_inner and not on_lane:
                    return x, y

        for _ in range(n_boxes):
            x, y = corridor_point(0.4)
            sx, sy, sz = rng.uniform(0.3, 1.0), rng.uniform(0.3, 1.0), rng.uniform(0.3, 2.5)
            boxes.append((x - sx / 2, y - sy / 2, 0.0, x + sx / 2, y + sy / 2, sz))
        for _ in range(n_cylinders):
            x, y = corridor_point(0.4)
            cyls.append((x, y, rng.uniform(0.1, 0.4), rng.uniform(1.5, 3.5)))
        self.boxes = np.array(boxes, dtype=np.float64).reshape(-1, 6)
        self.cylinders = np.array(cyls, dtype=np.float64).reshape(-1, 4)

    # ---- ray casting: origins o [3], directions d [N,3] (world frame, unit) -> range t [N] (inf = no hit)
    def cast(self, o, d):
        """First-hit range per ray.  Uses the C caster (csrc/synth_cast.c -> libarvc_synth.so, bit-identical, ~50x faster)
        when it has been built, else the numpy implementation below."""
        lib = _cast_lib()
        if lib is not None:
            o = np.ascontiguousarray(o, dtype=np.float64)
            d = np.ascontiguousarray(d, dtype=np.float64)
            t = np.empty(len(d))
            dp = ctypes.POINTER(ctypes.c_double)
            lib.arvc_synth_cast(o.ctypes.data_as(dp), d.ctypes.data_as(dp), len(d), self.segments.ctypes.data_as(dp), len(self.segments),
                                ctypes.c_double(self.wall_h), self.boxes.ctypes.data_as(dp), len(self.boxes),
                                self.cylinders.ctypes.data_as(dp), len(self.cylinders), t.ctypes.data_as(dp))
            return t
        return self.cast_numpy(o, d)

    def cast_numpy(self, o, d):
        n = len(d)
        t = np.full(n, np.inf)
        with np.errstate(divide="ignore", invalid="ignore"):
            # ground
            tg = -o[2] / d[:, 2]
            tg = np.where((d[:, 2] < 0) & (tg > 0), tg, np.inf)
            t = np.minimum(t, tg)
            # walls: 2-D ray/segment intersection, then height check
            for (x0, y0, x1, y1) in self.segments:
                ex, ey = x1 - x0, y1 - y0
                den = d[:, 0] * ey - d[:, 1] * ex
                wx, wy = x0 - o[0], y0 - o[1]
                tt = (wx * ey - wy * ex) / den
                u = (wx * d[:, 1] - wy * d[:, 0]) / den
                z = o[2] + tt * d[:, 2]
                ok = (np.abs(den) > 1e-12) & (tt > 0) & (u >= 0) & (u <= 1) & (z >= 0) & (z <= self.wall_h)
                t = np.minimum(t, np.where(ok, tt, np.inf))
            # boxes: slab method, vectorised over rays x boxes
            if len(self.boxes):
                inv = 1.0 / d                                                    # [N,3]
                lo = (self.boxes[None, :, 0:3] - o[None, None, :]) * inv[:, None, :]
                hi = (self.boxes[None, :, 3:6] - o[None, None, :]) * inv[:, None, :]
                tmin = np.minimum(lo, hi).max(axis=2)
                tmax = np.maximum(lo, hi).min(axis=2)
                ok = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 0)
                t = np.minimum(t, np.where(ok, tmin, np.inf).min(axis=1))
            # vertical cylinders (side surface only; tops are above the sensor)
            if len(self.cylinders):
                cx = self.cylinders[None, :, 0] - o[0]
                cy = self.cylinders[None, :, 1] - o[1]
                r = self.cylinders[None, :, 2]
                h = self.cylinders[None, :, 3]
                a = (d[:, 0] ** 2 + d[:, 1] ** 2)[:, None]
                b = d[:, 0:1] * cx + d[:, 1:2] * cy
                c = cx * cx + cy * cy - r * r
                disc = b * b - a * c
                tt = (b - np.sqrt(np.where(disc > 0, disc, np.nan))) / a
                z = o[2] + tt * d[:, 2:3]
                ok = (disc > 0) & (tt > 0) & (z >= 0) & (z <= h)
                t = np.minimum(t, np.where(ok, tt, np.inf).min(axis=1))
        return t


def sensor_directions(sensor):
    el = np.deg2rad(np.linspace(sensor.elev_min_deg, sensor.elev_max_deg, sensor.beams))
    az = np.linspace(0.0, 2 * np.pi, sensor.azimuth_steps, endpoint=False)
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    d = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :], np.broadcast_to(se, (sensor.beams, sensor.azimuth_steps))],
                 axis=-1)
    return d.reshape(-1, 3)


def pose_matrix(x, y, z, yaw, pitch=0.0, roll=0.0):
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    R = np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                  [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                  [-sp, cp * sr, cp * cr]])
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = (x, y, z)
    return T


def scan_from_pose(world, sensor, T_world_sensor, seed):
    """One scan in the SENSOR frame, float32 [n,3]; drop-outs and out-of-range rays are simply absent."""
    rng = np.random.default_rng(seed)
    d_s = sensor_directions(sensor)
    d_w = d_s @ T_world_sensor[:3, :3].T
    t = world.cast(T_world_sensor[:3, 3], d_w)
    t = t + rng.normal(0.0, sensor.sigma_range, size=t.shape)
    keep = np.isfinite(t) & (t > sensor.min_range) & (t < sensor.max_range) & (rng.random(t.shape) >= sensor.dropout)
    return (d_s[keep] * t[keep, None]).astype(np.float32)


def loop_trajectory(world, n_poses, step=0.5, start=0.0):
    """Poses on the rounded-rectangle centre line of the loop corridor, `step` metres apart, traversed
    repeatedly (so that loop-closure candidates exist).  Returns list of 4x4 (world <- sensor)."""
    cx = (world.outer[0] + world.inner[0]) / 2
    cy = (world.outer[1] + world.inner[1]) / 2
    rad = 1.5
    sx, sy = cx - rad, cy - rad
    # perimeter pieces: straight (2*sx), arc, straight (2*sy), arc, ...
    pieces = [("L", (-sx, -cy), 0.0, 2 * sx), ("A", (sx, -sy), -np.pi / 2, rad), ("L", (cx, -sy), np.pi / 2, 2 * sy),
              ("A", (sx, sy), 0.0, rad), ("L", (sx, cy), np.pi, 2 * sx), ("A", (-sx, sy), np.pi / 2, rad),
              ("L", (-cx, sy), -np.pi / 2, 2 * sy), ("A", (-sx, -sy), np.pi, rad)]
    lengths = [p[3] if p[0] == "L" else p[3] * np.pi / 2 for p in pieces]
    total = sum(lengths)
    poses = []
    for k in range(n_poses):
        s = (start + k * step) % total
        for p, L in zip(pieces, lengths):
            if s <= L:
                break
            s -= L
        if p[0] == "L":
            yaw = p[2]
            x = p[1][0] + s * np.cos(yaw)
            y = p[1][1] + s * np.sin(yaw)
        else:
            a = p[2] + s / p[3]
            x = p[1][0] + p[3] * np.cos(a)
            y = p[1][1] + p[3] * np.sin(a)
            yaw = a + np.pi / 2
        # gentle deterministic roll/pitch wobble so that all 6 DoF are exercised
        poses.append(pose_matrix(x, y, SENSOR_HEIGHT, yaw, 0.01 * np.sin(0.37 * k), 0.01 * np.cos(0.23 * k)))
    return poses


def noisy_odometry(poses, seed=4321, sigma_xy=0.02, sigma_yaw_deg=0.5):
    """Odometry prior = ground truth + per-step noise (integrated).  Returns list of 4x4 global poses."""
    rng = np.random.default_rng(seed)
    out = [poses[0].copy()]
    for a, b in zip(poses[:-1], poses[1:]):
        rel = np.linalg.inv(a) @ b
        n = pose_matrix(rng.normal(0, sigma_xy), rng.normal(0, sigma_xy), 0.0, np.deg2rad(rng.normal(0, sigma_yaw_deg)))
        out.append(out[-1] @ rel @ n)
    return out


def _pool_init():
    lib = _cast_lib()
    if lib is not None:
        lib.arvc_synth_set_threads(1)          # the pool is the parallelism: one casting thread per worker process


def _scan_job(args):
    world, sensor, T, seed = args
    return scan_from_pose(world, sensor, T, seed)


class Sequence:
    """n_scans consecutive keyframes: scans (sensor frame, float32), GT poses, odometry poses.
    `workers` > 1 ray-casts the scans in a process pool (same results, the per-scan seed fixes the noise)."""

    def __init__(self, n_scans, sensor=OS1_64, world=None, seed_world=1234, seed_odo=4321, step=0.5, start=0.0, workers=1):
        self.world = world or World(seed_world)
        self.sensor = sensor
        self.poses = loop_trajectory(self.world, n_scans, step=step, start=start)
        self.odometry = noisy_odometry(self.poses, seed_odo)
        jobs = [(self.world, sensor, T, 10000 + k) for k, T in enumerate(self.poses)]
        if workers > 1 and n_scans > 2:
            import multiprocessing as mp
            with mp.get_context("fork").Pool(min(workers, n_scans), initializer=_pool_init) as pool:
                self.scans = pool.map(_scan_job, jobs, chunksize=max(1, n_scans // (4 * workers)))
        else:
            self.scans = [_scan_job(j) for j in jobs]

    def relative_gt(self, i, j):
        return np.linalg.inv(self.poses[i]) @ self.poses[j]

    def relative_odo(self, i, j):
        return np.linalg.inv(self.odometry[i]) @ self.odometry[j]


def loop_closure_pairs(poses, n_pairs, radius=5.0, min_gap=20, seed=777, sigma_t=0.2, sigma_rot_deg=2.0):
    """Config 4: (i, j, init) with ||p_i - p_j|| < radius (loopclosing.py:10 radius_threshold) and |i-j| >= min_gap;
    init = ground truth perturbed by N(0, sigma_t) / N(0, sigma_rot)."""
    rng = np.random.default_rng(seed)
    P = np.array([T[:3, 3] for T in poses])
    cand = []
    for i in range(len(poses)):
        d = np.linalg.norm(P - P[i], axis=1)
        js = np.where((d < radius) & (np.arange(len(poses)) - i >= min_gap))[0]
        cand.extend((i, int(j)) for j in js)
    if not cand:
        return []
    pick = rng.choice(len(cand), size=n_pairs, replace=len(cand) < n_pairs)
    out = []
    for k in pick:
        i, j = cand[k]
        gt = np.linalg.inv(poses[i]) @ poses[j]
        r = np.deg2rad(rng.normal(0, sigma_rot_deg, 3))
        n = pose_matrix(*rng.normal(0, sigma_t, 3), r[2], r[1], r[0])
        out.append((i, j, gt @ n))
    return out
