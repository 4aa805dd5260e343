"""Per-tile profile of the forward rasterizer on the bench workload: which tiles make the launch long."""
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
args = (pts.to(dev), nrm.to(dev), col.to(dev), proj.to(dev), view.to(dev), torch.full((V,), 5e-5, device=dev), prm)
render_points(*args)
_lib.raster_stats(True)
render_points(*args)
torch.cuda.synchronize()
prof = _lib.tile_profile()
st = _lib.raster_stats(False)
print(st)
cyc, ln, grp, pos = prof[:, 0].double(), prof[:, 1].double(), prof[:, 2].double(), prof[:, 3]
act = ln > 0
print("tiles %d, non-empty %d, total entries %.0f, max list %d, mean list (non-empty) %.0f" % (len(ln), int(act.sum()), ln.sum(), int(ln.max()), ln[act].mean()))
print("cycles: sum %.3e  max %.0f (%.3f ms at 1.965 GHz)  mean(non-empty) %.0f" % (cyc.sum(), cyc.max(), cyc.max() / 1.965e6, cyc[act].mean()))
print("sum of cycles / (148 SMs x 3 CTAs) = %.3f ms" % (cyc.sum() / (148 * 3) / 1.965e6))
top = cyc.topk(12)
for c, i in zip(top.values.tolist(), top.indices.tolist()):
    print("  tile %5d view %d  cycles %8.0f (%.3f ms)  list %6d  groups %4d  launch pos %d" % (i, i // 1024, c, c / 1.965e6, int(ln[i]), int(grp[i]), int(pos[i])))
import numpy as np
q = np.quantile(cyc[act].numpy(), [0.5, 0.9, 0.99])
print("cycles quantiles (non-empty) 50/90/99%%: %s" % q)
print("cycles per list entry (non-empty tiles): %.1f ; per group: %.0f" % (cyc[act].sum() / ln[act].sum(), cyc[act].sum() / max(grp.sum(), 1)))
