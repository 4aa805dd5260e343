"""GPU tests of the reference-facing Python API (SurfaceSplatting / SurfaceSplattingRenderer /
EllipticalRasterizer): the fused renderer path must agree with the unfused rasterizer + compositor route
(the reference's structure), gradients included, and the documented side effects must happen."""
import numpy as np
import pytest
import torch

from dss_b200.core.camera import FoVPerspectiveCameras, look_at_view_transform
from dss_b200.core.cloud import PointClouds3D, PointCloudsFilters
from dss_b200.core.rasterizer import (EllipticalRasterizer, PointFragments, PointsRasterizationSettings,
                                      SurfaceSplatting, rasterize_elliptical_points)
from dss_b200.core.renderer import NormWeightedCompositor, SurfaceSplattingRenderer
from tests.util import sphere_cloud

pytestmark = pytest.mark.gpu


def _setup(dev, P0=4000, N=3, S=96, **kw):
    pts, nrm, col = sphere_cloud(P0, seed=3)
    R, T = look_at_view_transform(dist=[1.5, 1.8, 2.1][:N], elev=[10, -30, 60][:N], azim=[0, 120, -70][:N])
    cams = FoVPerspectiveCameras(znear=0.1, zfar=100.0, R=R, T=T, device=dev)
    settings = PointsRasterizationSettings(backface_culling=False, cutoff_threshold=1.0, depth_merging_threshold=0.05,
                                           Vrk_invariant=True, Vrk_isotropic=False, radii_backward_scaler=5,
                                           image_size=S, points_per_pixel=5, bin_size=None, clip_pts_grad=0.05, **kw)
    rast = SurfaceSplatting(cameras=cams, raster_settings=settings)
    renderer = SurfaceSplattingRenderer(rast, NormWeightedCompositor())
    return pts.to(dev), nrm.to(dev), col.to(dev), cams, rast, renderer


def test_fused_renderer_matches_unfused_route(cuda_device):
    dev = cuda_device
    pts, nrm, col, cams, rast, renderer = _setup(dev)
    N = len(cams)
    g = torch.randn(N, 96, 96, 4, generator=torch.Generator().manual_seed(1)).to(dev) * 1e-3

    def run(fused):
        p = pts.clone().requires_grad_(True)
        c = col.clone().requires_grad_(True)
        cloud = PointClouds3D([p], normals=[nrm], features=[c])
        img = renderer(cloud, cameras=cams, fused=fused)
        img.backward(g)
        return img.detach(), p.grad.clone(), c.grad.clone()

    img_f, gp_f, gc_f = run(True)
    img_u, gp_u, gc_u = run(False)
    assert img_f.shape == (N, 96, 96, 4)
    assert torch.equal(img_f[..., 3], img_u[..., 3])                     # occupancy identical
    assert float(((img_f - img_u) ** 2).mean()) < 1e-10
    torch.testing.assert_close(gc_f, gc_u, rtol=2e-4, atol=1e-8)
    # position gradients: same occupancy surrogate, chained analytically (fused) vs by autograd (unfused)
    scale = gp_u.abs().max()
    assert scale > 0 and (gp_f - gp_u).abs().max() <= 2e-4 * scale


def test_rasterizer_forward_contract(cuda_device):
    dev = cuda_device
    pts, nrm, col, cams, rast, renderer = _setup(dev)
    cloud = PointClouds3D([pts], normals=[nrm], features=[col])
    filt = PointCloudsFilters(device=dev)
    fragments, cloud_out, info = rast(cloud, point_clouds_filter=filt, cameras=cams, verbose=True)
    assert isinstance(fragments, PointFragments)
    N, S, K, P0 = len(cams), 96, 5, pts.shape[0]
    assert fragments.idx.shape == (N, S, S, K) and fragments.idx.dtype == torch.int32
    assert fragments.occupancy.shape == (N, S, S)
    assert len(cloud_out) == N and cloud_out.points_packed().shape == (N * P0, 3)
    assert set(info) == {"radii", "ellipse_params", "cutoff_threshold", "scaler"}
    # visibility side effect (rasterizer.py:643-653): padded (N, P_max) bool, true exactly for rendered points
    vis = filt.visibility
    assert vis.shape == (N, P0) and vis.dtype == torch.bool
    ids = fragments.idx[fragments.idx >= 0].unique().long()
    want = torch.zeros(N * P0, dtype=torch.bool, device=dev)
    want[ids] = True
    assert torch.equal(vis.reshape(-1), want)
    # scaler fragments are zero exactly where idx < 0 (utils/__init__.py:172-185)
    assert (fragments.scaler[fragments.idx < 0] == 0).all() and (fragments.scaler[fragments.idx >= 0] > 0).all()
    # z-buffer is ascending and within the merge threshold of the first fragment
    z = fragments.zbuf
    ok = fragments.idx[..., 1:] >= 0
    assert (z[..., 1:][ok] >= z[..., :-1][ok]).all()
    assert ((z[..., 1:] - z[..., :1])[ok] <= 0.05 + 1e-6).all()


def test_elliptical_rasterizer_autograd_function(cuda_device):
    """occupancy gradient flows to xy, z-buffer gradient to z, qvalue gradient is dropped
    (rasterizer.py:788-813), clip hook applied by rasterize_elliptical_points (rasterizer.py:735-736)."""
    dev = cuda_device
    pts, nrm, col, cams, rast, renderer = _setup(dev)
    cloud = PointClouds3D([pts], normals=[nrm], features=[col]).extend(len(cams))
    with torch.no_grad():
        info = rast._get_per_point_info(cloud, cameras=cams)
    screen = rast.transform(cloud, cameras=cams).detach().requires_grad_(True)
    first, num = cloud.cloud_to_packed_first_idx(), cloud.num_points_per_cloud()
    idx, zbuf, q, occ = rasterize_elliptical_points((screen, first, num), info["ellipse_params"],
                                                    info["cutoff_threshold"], info["radii"], image_size=96,
                                                    points_per_pixel=5, radii_backward_scaler=5.0, clip_pts_grad=0.05)
    (occ.sum() * 1e-3 + (zbuf * (idx >= 0)).sum() * 1e-4 + q.sum()).backward()
    g = screen.grad
    assert g.shape == screen.shape and torch.isfinite(g).all()
    assert (g.norm(dim=1) <= 0.05 * (1 + 1e-5)).all() and g.norm(dim=1).max() > 0   # clipped per point
    # raw function, no hook: z gradient equals the number of fragments a point owns
    screen2 = screen.detach().clone().requires_grad_(True)
    out = EllipticalRasterizer.apply(screen2, info["ellipse_params"], info["cutoff_threshold"], info["radii"], first,
                                     num, 0.05, 96, 5, 0, 0, 5.0)
    (out[1] * (out[0] >= 0)).sum().backward()
    counts = torch.bincount(out[0][out[0] >= 0].long(), minlength=screen.shape[0]).float()
    torch.testing.assert_close(screen2.grad[:, 2], counts)
    assert (screen2.grad[:, :2] == 0).all()


def test_empty_and_tiny_clouds(cuda_device):
    dev = cuda_device
    pts, nrm, col, cams, rast, renderer = _setup(dev)
    empty = PointClouds3D([pts[:0]], normals=[nrm[:0]], features=[col[:0]])
    assert renderer(empty, cameras=cams) is None                          # renderer.py:41-42
    one = PointClouds3D([pts[:1]], normals=[nrm[:1]], features=[col[:1]])  # "point-one.ply": a single splat
    img = renderer(one, cameras=cams)
    assert img.shape == (len(cams), 96, 96, 4) and torch.isfinite(img).all()
    assert img[..., 3].sum() >= 0


def test_ragged_batch_packed_path(cuda_device):
    """clouds of different sizes (no shared storage) go through the packed first_idx/num_points path."""
    dev = cuda_device
    pts, nrm, col, cams, rast, renderer = _setup(dev, N=2)
    a, b = slice(0, 3000), slice(500, 2000)
    cloud = PointClouds3D([pts[a], pts[b]], normals=[nrm[a], nrm[b]], features=[col[a], col[b]])
    img = renderer(cloud, cameras=cams)
    # each view equals rendering that cloud alone with that camera
    for n, sl in enumerate((a, b)):
        cam_n = FoVPerspectiveCameras(znear=0.1, zfar=100.0, R=cams.R[n:n + 1], T=cams.T[n:n + 1], device=dev)
        single = renderer(PointClouds3D([pts[sl]], normals=[nrm[sl]], features=[col[sl]]), cameras=cam_n)
        assert torch.equal(single[0, ..., 3], img[n, ..., 3])
        torch.testing.assert_close(single[0], img[n], rtol=1e-5, atol=1e-6)


def test_cpu_tensors_fail_loudly():
    """no CPU fallback: the operators refuse CPU tensors instead of silently computing elsewhere."""
    from dss_b200 import _C
    z = torch.zeros(4, 3)
    with pytest.raises(RuntimeError):
        _C.splat_points(z, z, torch.ones(4), torch.ones(4, 2), torch.zeros(1, dtype=torch.int64),
                        torch.full((1,), 4, dtype=torch.int64), 0.05, 16, 5, 0, 0)


def test_filter_renderable_compaction_matches_slot_semantics(cuda_device):
    """SurfaceSplatting.filter_renderable returns the reference's compacted cloud (rasterizer.py:219-254);
    forward(compact_filtered=True) renders it: same fragments as the default slot-keeping path up to the compaction map."""
    dev = cuda_device
    pts, nrm, col, cams, rast, renderer = _setup(dev, P0=6000, N=3, S=96)
    rast.raster_settings.backface_culling = True
    cloud = PointClouds3D([pts], normals=[nrm], features=[col])
    filtered, mask = rast.filter_renderable(cloud, cameras=cams)
    N = len(cams)
    assert mask.shape == (N * 6000,) and len(filtered) == N
    num = filtered.num_points_per_cloud()
    assert torch.equal(num, mask.view(N, -1).sum(1))
    assert 0.3 < mask.float().mean().item() < 0.7
    # same splat size in both modes (left to itself the compact mode sizes the splats from each view's FILTERED cloud,
    # like the reference, the slot mode from the whole cloud: DESIGN.md hazard 12)
    h = torch.full((N,), 3e-4, device=dev)
    frag_slot, _ = rast(cloud, cameras=cams, Vrk_h=h)
    frag_comp, cloud_comp = rast(cloud, cameras=cams, compact_filtered=True, Vrk_h=h)
    assert int(cloud_comp.num_points_per_cloud().sum()) == int(mask.sum())
    remap = (torch.cumsum(mask.long(), 0) - 1).to(torch.int32)
    mapped = torch.where(frag_slot.idx >= 0, remap[frag_slot.idx.clamp(min=0).long()], torch.full_like(frag_slot.idx, -1))
    # the float64-free torch mask and the kernel's fp32 mask agree except for normals within rounding of edge-on
    same = (mapped == frag_comp.idx).all(-1)
    assert same.float().mean().item() > 0.999
    assert torch.equal(frag_slot.zbuf[same], frag_comp.zbuf[same])
    rast.raster_settings.backface_culling = False
