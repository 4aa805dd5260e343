"""Diagnose idx mismatches between our rasterizer and the reference naive CUDA kernel (C2-like scene)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dss_b200.ops import SplatParams, render_points
from tests.util import scene, packed_offsets
from oracle import build_ref
ref = build_ref.ref_cuda()
dev = torch.device("cuda:0")
P0, N, S, K = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000, 4, 512, 5
pts, nrm, col, proj, view, _ = scene(P0, N, seed=2)
prm = SplatParams(image_size=S, points_per_pixel=K, znear=0.1)
h = torch.full((N,), 2e-4, device=dev)
out = render_points(pts.to(dev), nrm.to(dev), col.to(dev), proj.to(dev), view.to(dev), h, prm, return_fragments=True)
rec = out.records
first, num = (t.to(dev) for t in packed_offsets(N, P0))
ndc, ell, rad = rec[:, :3].contiguous(), rec[:, 5:8].contiguous(), rec[:, 3:5].contiguous()
cut = torch.ones(N * P0, device=dev)
r_idx, r_z, r_q, r_occ = ref.splat_points_naive_cuda(ndc, ell, cut, rad, first, num, 0.05, S, K)
same = (r_idx == out.idx).all(-1)
bad = (~same).nonzero()
print("mismatching pixels:", bad.shape[0], "of", same.numel())
ties = 0
for b in bad[:12].tolist():
    n, y, x = b
    print("pixel", b)
    print("  ours idx", out.idx[n, y, x].tolist(), "z", [("%.9g" % v) for v in out.zbuf[n, y, x].tolist()], "q", [("%.6f" % v) for v in out.qvalue[n, y, x].tolist()])
    print("  ref  idx", r_idx[n, y, x].tolist(), "z", [("%.9g" % v) for v in r_z[n, y, x].tolist()], "q", [("%.6f" % v) for v in r_q[n, y, x].tolist()])
zs, zr = out.zbuf[~same].sort(-1)[0], r_z[~same].sort(-1)[0]
print("same depth multiset on", int((zs == zr).all(-1).sum()), "of", bad.shape[0])
setsame = (out.idx[~same].sort(-1)[0] == r_idx[~same].sort(-1)[0]).all(-1)
print("same id set on", int(setsame.sum()))
