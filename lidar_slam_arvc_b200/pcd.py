"""Minimal PCD (Point Cloud Data) reader / writer for the load path.

Stands in for `o3d.io.read_point_cloud(filename)` at keyframemanager/keyframe.py:41-45 of the reference: the
scan-matcher only needs the x/y/z fields.  Supports DATA ascii and DATA binary with float32 or float64
coordinates; `binary_compressed` (LZF) is not implemented yet and raises.
Like Open3D's reader with default arguments, NaN/inf points are kept (the radius/height filter drops them).
"""
import numpy as np

_NP = {("F", 4): "<f4", ("F", 8): "<f8", ("I", 1): "<i1", ("I", 2): "<i2", ("I", 4): "<i4", ("I", 8): "<i8",
       ("U", 1): "<u1", ("U", 2): "<u2", ("U", 4): "<u4", ("U", 8): "<u8"}


def read_pcd_xyz(filename):
    """Returns an [n,3] array: float32 when the file stores float32 coordinates (the usual case), else float64."""
    with open(filename, "rb") as f:
        header = {}
        while True:
            line = f.readline()
            if not line:
                raise ValueError("%s: truncated PCD header" % filename)
            line = line.decode("ascii", errors="replace").strip()
            if not line or line.startswith("#"):
                continue
            key, _, val = line.partition(" ")
            header[key.upper()] = val.split()
            if key.upper() == "DATA":
                break
        fields = header.get("FIELDS")
        if fields is None or not all(a in fields for a in ("x", "y", "z")):
            raise ValueError("%s: PCD without x y z fields" % filename)
        sizes = [int(s) for s in header["SIZE"]]
        types = header["TYPE"]
        counts = [int(c) for c in header.get("COUNT", ["1"] * len(fields))]
        n = int(header["POINTS"][0]) if "POINTS" in header else int(header["WIDTH"][0]) * int(header["HEIGHT"][0])
        data = header["DATA"][0].lower()
        names, formats = [], []
        for fld, sz, tp, cnt in zip(fields, sizes, types, counts):
            dt = _NP[(tp.upper(), sz)]
            if cnt == 1:
                names.append(fld)
                formats.append(dt)
            else:
                for c in range(cnt):
                    names.append("%s_%d" % (fld, c))
                    formats.append(dt)
        if data == "binary":
            rec = np.dtype({"names": names, "formats": formats})
            arr = np.frombuffer(f.read(rec.itemsize * n), dtype=rec, count=n)
            cols = [arr["x"], arr["y"], arr["z"]]
        elif data == "ascii":
            raw = np.loadtxt(f, dtype=np.float64, ndmin=2) if n > 0 else np.zeros((0, len(names)))
            cols = [raw[:n, names.index(a)] for a in ("x", "y", "z")]
            ftype = formats[names.index("x")]
            cols = [c.astype(ftype) for c in cols]
        elif data == "binary_compressed":
            raise NotImplementedError("%s: DATA binary_compressed (LZF) is not supported yet" % filename)
        else:
            raise ValueError("%s: unknown DATA %s" % (filename, data))
    out_t = np.float32 if all(c.dtype == np.float32 for c in cols) else np.float64
    return np.ascontiguousarray(np.stack(cols, axis=1).astype(out_t, copy=False))


def write_pcd_xyz(filename, xyz, binary=True):
    """Write x y z as float32 (what LiDAR drivers produce), DATA binary or ascii."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    n = len(xyz)
    header = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
              "WIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n" % (n, n, "binary" if binary else "ascii"))
    with open(filename, "wb") as f:
        f.write(header.encode("ascii"))
        if binary:
            f.write(xyz.tobytes())
        else:
            for p in xyz:
                f.write(("%.9g %.9g %.9g\n" % (p[0], p[1], p[2])).encode("ascii"))
