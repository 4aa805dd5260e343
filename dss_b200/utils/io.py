"""PLY point-cloud files, the reference's exchange format for clouds.

``save_ply`` keeps the signature and on-disk layout of DSS/utils/io.py:89-146 (vertex element with x y z [nx ny nz]
[red green blue [alpha]], float32 coordinates, uint8 colours scaled by 255 when given in [0, 1], binary little endian
by default) -- what ``Generator.generate_pointclouds`` exports (DSS/models/point_modeling.py:284-326) and what the
example data ships (example_data/pointclouds/*.ply).  The reference goes through the ``plyfile`` package (absent here);
the format is simple enough to read and write directly with numpy.  Host-side and cold: nothing here touches the GPU.
"""
import os

import numpy as np

__all__ = ["read_ply", "save_ply"]

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
              "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
              "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def read_ply(path):
    """Vertex element of a PLY file (ascii, binary_little_endian or binary_big_endian) ->
    ``dict(points (N,3) f32, normals (N,3) f32 | None, colors (N,3|4) f32 in [0,1] | None, properties {name: array})``.
    Other elements (faces ...) are skipped; list properties inside the vertex element are not supported."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError("%s is not a PLY file" % path)
        fmt, elements = None, []
        while True:
            line = f.readline()
            if not line:
                raise ValueError("%s: unexpected end of header" % path)
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append((tok[1], int(tok[2]), []))
            elif tok[0] == "property":
                if tok[1] == "list":
                    elements[-1][2].append((tok[-1], None))
                else:
                    elements[-1][2].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
            raise ValueError("%s: unsupported PLY format %r" % (path, fmt))
        data = None
        for name, count, props in elements:
            if name != "vertex":
                if data is None:
                    # elements before the vertices would have to be skipped record by record; not produced by anything
                    # on this path
                    raise ValueError("%s: the vertex element must come first" % path)
                break
            if any(t is None for _, t in props):
                raise ValueError("%s: list properties in the vertex element are not supported" % path)
            if fmt == "ascii":
                raw = np.loadtxt(f, max_rows=count, dtype=np.float64, ndmin=2) if count else np.zeros((0, len(props)))
                data = {n: raw[:, i].astype(t) for i, (n, t) in enumerate(props)}
            else:
                end = "<" if fmt == "binary_little_endian" else ">"
                dt = np.dtype([(n, end + t) for n, t in props])
                rec = np.frombuffer(f.read(count * dt.itemsize), dtype=dt, count=count)
                data = {n: np.ascontiguousarray(rec[n]).astype(t) for n, t in props}
    if data is None:
        raise ValueError("%s: no vertex element" % path)

    def stack(names, dtype):
        return np.stack([data[n] for n in names], 1).astype(dtype) if all(n in data for n in names) else None

    pts = stack(("x", "y", "z"), np.float32)
    if pts is None:
        raise ValueError("%s: vertices have no x y z" % path)
    nrm = stack(("nx", "ny", "nz"), np.float32)
    col = stack(("red", "green", "blue", "alpha"), np.float32)
    if col is None:
        col = stack(("red", "green", "blue"), np.float32)
    if col is not None and data["red"].dtype == np.uint8:
        col = col / 255.0
    return {"points": pts, "normals": nrm, "colors": col, "properties": data}


def save_ply(filename, points, colors=None, normals=None, binary=True):
    """DSS/utils/io.py:89-146.  points (N, 2 or 3); colors (N, 3 or 4), scaled by 255 when max <= 1; normals (N, 2 or 3)."""
    points = np.asarray(points)
    if points.ndim != 2:
        raise ValueError("points must be (N, 2 or 3)")
    if points.shape[-1] == 2:
        points = np.concatenate([points, np.zeros_like(points)[:, :1]], axis=-1)
    n = points.shape[0]
    fields = [("x", "<f4"), ("y", "<f4"), ("z", "<f4")]
    cols = [points[:, 0], points[:, 1], points[:, 2]]
    if normals is not None:
        normals = np.asarray(normals)
        if normals.ndim != 2 or len(normals) != n:
            raise ValueError("normals must be (N, 2 or 3) with one row per point")
        if normals.shape[-1] == 2:
            normals = np.concatenate([normals, np.zeros_like(normals)[:, :1]], axis=-1)
        fields += [("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4")]
        cols += [normals[:, 0], normals[:, 1], normals[:, 2]]
    if colors is not None:
        colors = np.asarray(colors)
        if len(colors) != n or colors.ndim != 2 or colors.shape[1] not in (3, 4):
            raise ValueError("colors must be (N, 3 or 4) with one row per point")
        if colors.size and colors.max() <= 1:
            colors = colors * 255
        names = ("red", "green", "blue", "alpha")[:colors.shape[1]]
        fields += [(c, "u1") for c in names]
        cols += [colors[:, i] for i in range(colors.shape[1])]
    rec = np.empty(n, dtype=np.dtype(fields))
    for (name, _), c in zip(fields, cols):
        rec[name] = c
    d = os.path.dirname(filename)
    if d and not os.path.exists(d):
        os.makedirs(d)
    type_names = {"<f4": "float", "u1": "uchar"}
    header = ["ply", "format %s 1.0" % ("binary_little_endian" if binary else "ascii"), "element vertex %d" % n]
    header += ["property %s %s" % (type_names[t], name) for name, t in fields]
    header += ["end_header"]
    with open(filename, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        if binary:
            f.write(rec.tobytes())
        else:
            for row in rec:
                f.write((" ".join(repr(float(v)) if isinstance(v, np.floating) else str(int(v)) for v in row) + "\n").encode("ascii"))
