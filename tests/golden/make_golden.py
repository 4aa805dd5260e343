"""Mint golden vectors from the REFERENCE's own CPU code (oracle/_ref/dss_ref_cpu, compiled from
/root/reference/DSS/csrc/rasterize_points_cpu.cpp) on seeded inputs.  Run in the build container:

    python -m tests.golden.make_golden

The reference ships no fixtures for this path; these files pin the oracle (and
through it the CUDA path) to what the reference computes.  Inputs: (a) random packed screen-space splats,
(b) real point clouds from the reference's example_data projected with a look-at camera.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import build_ref, preprocess_f64          # noqa: E402
from tests.util import random_screen_splats           # noqa: E402
from dss_b200.core.camera import FoVPerspectiveCameras, look_at_view_transform, camera_matrices  # noqa: E402


def read_ply_xyz_normals(path, limit):
    """Minimal ascii/binary-little-endian PLY reader for x y z nx ny nz vertex properties."""
    with open(path, "rb") as f:
        header = []
        while True:
            line = f.readline().decode("ascii", "replace").strip()
            header.append(line)
            if line == "end_header":
                break
        fmt = [h.split()[1] for h in header if h.startswith("format")][0]
        nvert = [int(h.split()[2]) for h in header if h.startswith("element vertex")][0]
        props = []
        in_vertex = False
        for h in header:
            if h.startswith("element"):
                in_vertex = h.startswith("element vertex")
            elif h.startswith("property") and in_vertex:
                props.append((h.split()[1], h.split()[2]))
        names = [p[1] for p in props]
        if fmt == "ascii":
            data = np.loadtxt(f, max_rows=nvert, dtype=np.float64, ndmin=2)
        else:
            dt = np.dtype([(n, {"float": "<f4", "float32": "<f4", "double": "<f8", "uchar": "u1", "uint8": "u1",
                                "int": "<i4", "uint": "<u4"}[t]) for t, n in props])
            raw = np.frombuffer(f.read(nvert * dt.itemsize), dtype=dt, count=nvert)
            data = np.stack([raw[n].astype(np.float64) for n in names], 1)
    col = lambda n: data[:, names.index(n)]
    xyz = np.stack([col("x"), col("y"), col("z")], 1)
    nrm = np.stack([col("nx"), col("ny"), col("nz")], 1)
    idx = np.random.default_rng(0).permutation(nvert)[:limit]
    xyz, nrm = xyz[idx], nrm[idx]
    xyz = xyz - xyz.mean(0)
    xyz = xyz / np.linalg.norm(xyz, axis=1).max()          # unit sphere, create_mvr_data_from_mesh.py:122-126
    nrm = nrm / np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-12)
    return xyz.astype(np.float32), nrm.astype(np.float32)


def reference_outputs(ref, pts, ell, cut, rad, first, num, S, K, grad_occ, radii_s):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    idx, zbuf, q, occ = ref.splat_points_naive_cpu(t(pts), t(ell), t(cut), t(rad), t(first), t(num), 0.05, S, K)
    gb = ref.splat_points_occ_backward_cpu(t(pts), t(rad), t(grad_occ), t(first), t(num), radii_s, 0.05)
    gz = torch.zeros(pts.shape[0], 1)
    gzbuf = t((np.random.default_rng(1).standard_normal(tuple(idx.shape)) * (idx.numpy() >= 0)).astype(np.float32))
    ref.backward_zbuf_cpu(idx, gzbuf, gz)
    return dict(idx=idx.numpy(), zbuf=zbuf.numpy(), qvalue=q.numpy(), occ=occ.numpy(), occ_backward=gb.numpy(),
                grad_zbuf=gzbuf.numpy(), zbuf_backward=gz.numpy()[:, 0])


def main():
    ref = build_ref.ref_cpu()
    assert ref is not None, "needs /root/reference (or a prebuilt oracle/_ref/dss_ref_cpu)"
    # (a) random packed splats, 2 ragged views
    S, K, P, N = 48, 5, 1200, 2
    pts, ell, cut, rad, first, num = random_screen_splats(P, N, S, seed=42)
    g = (np.random.default_rng(2).standard_normal((N, S, S)) * 1e-3).astype(np.float32)
    out = reference_outputs(ref, pts, ell, cut, rad, first, num, S, K, g, 3.0)
    np.savez_compressed(os.path.join(HERE, "random_splats_S48.npz"), points=pts, ellipse=ell, cutoff=cut, radii=rad,
                        first=first, num=num, S=S, K=K, grad_occ=g, radii_s=3.0, **out)
    # (b) reference example clouds through the float64 preprocess
    for name, limit, S in (("teapot_normal_dense", 1500, 64), ("bunny-8000", 1500, 64), ("sphere_2k", 1000, 40)):
        path = "/root/reference/example_data/pointclouds/%s.ply" % name
        if not os.path.exists(path):
            continue
        xyz, nrm = read_ply_xyz_normals(path, limit)
        R, T = look_at_view_transform(dist=2.0, elev=25.0, azim=40.0)
        cams = FoVPerspectiveCameras(znear=0.1, zfar=100.0, R=R, T=T)
        proj, view = camera_matrices(cams)
        h = np.full((1,), 4e-4, np.float32)
        pre = preprocess_f64(proj.numpy(), view.numpy(), xyz, nrm, h, 1.0, 1.0, S)
        f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        pts, ell, rad = f(pre["ndc"]), f(pre["ellipse"]), f(pre["radii"])
        cut = np.ones(len(pts), np.float32)
        first, num = np.zeros(1, np.int64), np.array([len(pts)], np.int64)
        g = (np.random.default_rng(3).standard_normal((1, S, S)) * 1e-3).astype(np.float32)
        out = reference_outputs(ref, pts, ell, cut, rad, first, num, S, 5, g, 2.0)
        np.savez_compressed(os.path.join(HERE, "%s_S%d.npz" % (name.replace("-", "_"), S)), world=xyz, normals=nrm,
                            proj=proj.numpy(), view=view.numpy(), h=h, scaler=f(pre["scaler"]), points=pts,
                            ellipse=ell, cutoff=cut, radii=rad, first=first, num=num, S=S, K=5, grad_occ=g,
                            radii_s=2.0, **out)
    print("golden files:", sorted(x for x in os.listdir(HERE) if x.endswith(".npz")))


if __name__ == "__main__":
    main()
