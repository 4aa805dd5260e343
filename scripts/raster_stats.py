import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dss_b200 import _lib
from dss_b200.ops import SplatParams, render_points
from dss_b200.core.camera import camera_matrices
from tests.util import sphere_cloud, random_cameras
dev = torch.device("cuda:0")
P0, V, S = 1_000_000, 8, 512
pts, nrm, col = sphere_cloud(P0)
proj, view = camera_matrices(random_cameras(V))
prm = SplatParams(image_size=S, znear=0.1, clip_pts_grad=0.05)
args = (pts.to(dev), nrm.to(dev), col.repeat(V, 1).to(dev), proj.to(dev), view.to(dev), torch.full((V,), 5e-5, device=dev), prm)
render_points(*args)
_lib.raster_stats(True)
out = render_points(*args)
st = _lib.raster_stats(False)
print(st)
print("visible points per view:", out.visible.view(V, P0).sum(1).tolist())
print("occupied pixels per view:", (out.image[..., 3] > 0).view(V, -1).sum(1).tolist())
print("per splat-view: scanned %.2f survivors %.2f tests %.2f accepted %.3f" % tuple(st[k] / (V * P0) for k in ("entries_scanned", "survivors", "pixel_tests", "accepted")))
from dss_b200 import _C
first = torch.arange(V, device=dev, dtype=torch.int64) * P0
num = torch.full((V,), P0, device=dev, dtype=torch.int64)
radii = out.records[:, 3:5].contiguous()
rs = _C.search_radius(radii, out.visible, first, num, 5.0)
print("search radius (px):", (rs * S / 2).tolist())
vis = out.visible.bool()
print("median radius of visible (px):", float(radii[vis].median() * S / 2), "mean", float(radii[vis].mean() * S / 2), "max", float(radii[vis].max() * S / 2))
