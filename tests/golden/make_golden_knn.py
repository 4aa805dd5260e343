"""Mint K-NN golden vectors from the REFERENCE's own ground truth, FRNNBruteForceCPU (external/FRNN/frnn/csrc/
bruteforce/bruteforce_cpu.cpp), compiled into oracle/_ref/dss_ref_frnn_cpu.  Run in the build container:

    python -m tests.golden.make_golden_knn

Inputs: the reference's example clouds (normalised like create_mvr_data_from_mesh.py:122-126) and seeded random
clouds, with the query the splat-size rule makes (K = 7, r = 0.2: DSS/core/rasterizer.py:313-326) and a radius small
enough to leave some neighbour lists short."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import build_ref                                   # noqa: E402
from tests.golden.make_golden import read_ply_xyz_normals     # noqa: E402


def main():
    ref = build_ref.ref_frnn_cpu()
    assert ref is not None, "needs /root/reference (or a prebuilt oracle/_ref/dss_ref_frnn_cpu)"
    clouds = {}
    for name, limit in (("teapot_normal_dense", 2500), ("bunny-8000", 2500), ("sphere_2k", 2000)):
        path = "/root/reference/example_data/pointclouds/%s.ply" % name
        if os.path.exists(path):
            clouds[name.replace("-", "_")] = read_ply_xyz_normals(path, limit)[0]
    rng = np.random.default_rng(7)
    clouds["random_cube"] = rng.uniform(-0.5, 0.5, (1500, 3)).astype(np.float32)
    dup = rng.uniform(-0.5, 0.5, (600, 3)).astype(np.float32)
    clouds["random_with_duplicates"] = np.concatenate([dup, dup[:200], dup[:50]]).astype(np.float32)   # exact ties
    out = {}
    for name, xyz in clouds.items():
        p = torch.from_numpy(xyz)[None]
        n = torch.tensor([xyz.shape[0]], dtype=torch.int64)
        for tag, K, r in (("k7_r0.2", 7, 0.2), ("k12_r0.05", 12, 0.05)):
            idxs, dists = ref.frnn_bf_cpu(p, p, n, n, K, r)
            out["%s__%s__dists" % (name, tag)] = dists[0].numpy()
            out["%s__%s__idxs" % (name, tag)] = idxs[0].numpy().astype(np.int32)
        out["%s__points" % name] = xyz
    np.savez_compressed(os.path.join(HERE, "knn_frnn_bruteforce.npz"), **out)
    print("wrote knn_frnn_bruteforce.npz:", sorted(clouds))


if __name__ == "__main__":
    main()
