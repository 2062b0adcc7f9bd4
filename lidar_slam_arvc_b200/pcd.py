"""Minimal PCD (Point Cloud Data) reader / writer for the load path.

Stands in for `o3d.io.read_point_cloud(filename)` at keyframemanager/keyframe.py:41-45 of the reference: the
scan-matcher only needs the x/y/z fields.  Supports DATA ascii and DATA binary with float32 or float64
coordinates, and DATA binary_compressed (LZF, decompressed by the native library's host helper).
Like Open3D's reader with default arguments, NaN/inf points are kept (the radius/height filter drops them).
"""
import numpy as np

_NP = {("F", 4): "<f4", ("F", 8): "<f8", ("I", 1): "<i1", ("I", 2): "<i2", ("I", 4): "<i4", ("I", 8): "<i8",
       ("U", 1): "<u1", ("U", 2): "<u2", ("U", 4): "<u4", ("U", 8): "<u8"}


def read_pcd_xyz(filename, alloc=None):
    """Returns an [n,3] array: float32 when the file stores float32 coordinates (the usual case), else float64.
    `alloc(n_rows, dtype)` may supply the output array (e.g. page-locked memory for an asynchronous upload); a binary file
    that holds nothing but float32 x y z is then read straight into it."""
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
        for pos, (fld, sz, tp, cnt) in enumerate(zip(fields, sizes, types, counts)):
            dt = _NP.get((tp.upper(), sz))
            if dt is None:
                raise ValueError("%s: unsupported PCD field type %s of size %d (field %s)" % (filename, tp, sz, fld))
            # PCL writes padding as repeated fields named "_" (e.g. "x y z _ intensity _"): everything but x / y / z gets its
            # position appended, so that the record dtype never sees a duplicate name
            base = fld if fld in ("x", "y", "z") and fld not in names else "%s@%d" % (fld, pos)
            if cnt == 1:
                names.append(base)
                formats.append(dt)
            else:
                for c in range(cnt):
                    names.append("%s_%d" % (base, c))
                    formats.append(dt)
        if data == "binary" and alloc is not None and names == ["x", "y", "z"] and formats == ["<f4"] * 3:
            out = alloc(n, np.float32)
            got = f.readinto(memoryview(out).cast("B")) if n else 0
            if got != 12 * n:
                raise ValueError("%s: truncated PCD payload" % filename)
            return out
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
            # uint32 compressed size, uint32 uncompressed size, LZF stream; the payload is stored field by field (SoA)
            import struct
            from .engine import lzf_decompress
            csize, usize = struct.unpack("<II", f.read(8))
            raw = lzf_decompress(f.read(csize), usize)
            cols, off = {}, 0
            for nme, fmt in zip(names, formats):
                dt = np.dtype(fmt)
                if nme in ("x", "y", "z"):
                    cols[nme] = np.frombuffer(raw, dtype=dt, count=n, offset=off)
                off += dt.itemsize * n
            cols = [cols["x"], cols["y"], cols["z"]]
        else:
            raise ValueError("%s: unknown DATA %s" % (filename, data))
    out_t = np.float32 if all(c.dtype == np.float32 for c in cols) else np.float64
    if alloc is not None:
        out = alloc(len(cols[0]), out_t)
        for k in range(3):
            out[:, k] = cols[k]
        return out
    return np.ascontiguousarray(np.stack(cols, axis=1).astype(out_t, copy=False))


def lzf_compress_literal(data):
    """A valid LZF stream made of literal runs only (no back references): enough to write test fixtures."""
    out = bytearray()
    for i in range(0, len(data), 32):
        chunk = data[i:i + 32]
        out.append(len(chunk) - 1)
        out += chunk
    return bytes(out)


def write_pcd_xyz(filename, xyz, binary=True, compressed=False):
    """Write x y z as float32 (what LiDAR drivers produce), DATA binary, ascii or binary_compressed."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    n = len(xyz)
    kind = "binary_compressed" if compressed else ("binary" if binary else "ascii")
    header = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
              "WIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n" % (n, n, kind))
    with open(filename, "wb") as f:
        f.write(header.encode("ascii"))
        if compressed:
            import struct
            payload = np.ascontiguousarray(xyz.T).tobytes()          # field by field
            comp = lzf_compress_literal(payload)
            f.write(struct.pack("<II", len(comp), len(payload)))
            f.write(comp)
        elif binary:
            f.write(xyz.tobytes())
        else:
            for p in xyz:
                f.write(("%.9g %.9g %.9g\n" % (p[0], p[1], p[2])).encode("ascii"))
