"""Writer of a synthetic EuRoC/ASL directory in the layout the reference's drivers read (eurocreader/eurocreader.py,
run_scanmatcher.py:101-125,159-166):
    <dir>/robot0/lidar/data.csv                 '#timestamp [ns]'
    <dir>/robot0/lidar/data/<timestamp>.pcd     binary PCD, FIELDS x y z (float32)
    <dir>/robot0/odom/data.csv                  '#timestamp [ns],x,y,z,qx,qy,qz,qw'
    <dir>/robot0/scanmatcher_parameters.yaml    start_index, delta_time, voxel_size, method
"""
import os

import numpy as np
import yaml

from .homogeneousmatrix import rot2quaternion
from .pcd import write_pcd_xyz


def write_euroc_tree(directory, seq, delta_time=0.5, voxel_size=None, method="icppointplane", t0_ns=1_700_000_000_000_000_000):
    lidar_dir = os.path.join(directory, "robot0", "lidar", "data")
    os.makedirs(lidar_dir, exist_ok=True)
    os.makedirs(os.path.join(directory, "robot0", "odom"), exist_ok=True)
    times = [int(t0_ns + round(k * delta_time * 1e9)) for k in range(len(seq.scans))]
    with open(os.path.join(directory, "robot0", "lidar", "data.csv"), "w") as f:
        f.write("#timestamp [ns]\n")
        for t in times:
            f.write("%d\n" % t)
    for t, s in zip(times, seq.scans):
        write_pcd_xyz(os.path.join(lidar_dir, "%d.pcd" % t), s)
    with open(os.path.join(directory, "robot0", "odom", "data.csv"), "w") as f:
        f.write("#timestamp [ns],x,y,z,qx,qy,qz,qw\n")
        for t, T in zip(times, seq.odometry):
            q = rot2quaternion(T)          # [qw, qx, qy, qz]
            f.write("%d,%.17g,%.17g,%.17g,%.17g,%.17g,%.17g,%.17g\n" % (t, T[0, 3], T[1, 3], T[2, 3], q[1], q[2], q[3], q[0]))
    with open(os.path.join(directory, "robot0", "scanmatcher_parameters.yaml"), "w") as f:
        yaml.safe_dump({"start_index": 0, "delta_time": float(delta_time), "voxel_size": voxel_size, "method": method}, f)
    return np.array(times, dtype=np.int64)
