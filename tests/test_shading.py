"""Per-point shading: the torch twin of DSS/core/texture.py / lighting.py on the CPU, and the fused
CUDA route against it (forward colours, and gradients w.r.t. albedo, normals and positions through autograd)."""
import math

import pytest
import torch

from dss_b200.core.camera import camera_matrices
from dss_b200.core.cloud import PointClouds3D
from dss_b200.core.lighting import DirectionalLights, PointLights, pack_lights
from dss_b200.core.texture import LightingTexture, apply_lighting, camera_centres
from tests.util import random_cameras, scene


def _lights(kind, dev="cpu"):
    amb = (((0.3, 0.25, 0.2), (0.1, 0.1, 0.15)),)
    dif = (((0.6, 0.5, 0.4), (0.2, 0.3, 0.5)),)
    spe = (((0.5, 0.5, 0.4), (0.3, 0.2, 0.6)),)
    if kind == "sun":
        return DirectionalLights(ambient_color=amb, diffuse_color=dif, specular_color=spe,
                                 direction=(((0.3, 1.0, 0.4), (-0.8, 0.1, 0.5)),), device=dev)
    return PointLights(ambient_color=amb, diffuse_color=dif, specular_color=spe,
                       location=(((0.7, 1.5, 0.9), (-1.2, 0.3, 1.1)),), device=dev)


def test_lighting_terms_match_hand_computed_values():
    """one point, one directional light: Lambert + Phong by hand (lighting.py:62-69, 139-172)."""
    n = torch.tensor([[0.0, 0.0, 2.0]])                      # not unit length: renormalised
    p = torch.tensor([[0.0, 0.0, 0.0]])
    cam = torch.tensor([[0.0, 1.0, 1.0]])
    lights = DirectionalLights(ambient_color=(((0.1, 0.2, 0.3),),), diffuse_color=(((1.0, 0.5, 0.25),),),
                               specular_color=(((0.5, 0.5, 0.5),),), direction=(((0.0, 3.0, 3.0),),))
    amb, dif, spe = apply_lighting(p, n, lights, cam, shininess=4)
    c = math.cos(math.pi / 4)
    torch.testing.assert_close(amb, torch.tensor([[0.1, 0.2, 0.3]]))
    torch.testing.assert_close(dif, torch.tensor([[1.0, 0.5, 0.25]]) * c)
    # reflect = -d + 2 (n.d) n = (0, -c, c); view = (0, c, c); alpha = v.r = 0 -> no highlight
    torch.testing.assert_close(spe, torch.zeros(1, 3), atol=1e-7, rtol=0)
    # light behind the surface: no diffuse, no specular
    lights2 = DirectionalLights(direction=(((0.0, 0.0, -1.0),),))
    _, dif2, spe2 = apply_lighting(p, n, lights2, cam)
    assert dif2.abs().sum() == 0 and spe2.abs().sum() == 0
    # camera along the mirror direction: alpha = 1 -> specular colour itself
    _, _, spe3 = apply_lighting(p, n, lights, torch.tensor([[0.0, -1.0, 1.0]]), shininess=64)
    torch.testing.assert_close(spe3, torch.tensor([[0.5, 0.5, 0.5]]))


def test_camera_centres_invert_the_view_transform():
    cams = random_cameras(5, seed=3)
    _, view = camera_matrices(cams)
    c = camera_centres(view)
    ch = torch.cat([c, torch.ones(5, 1)], 1)
    in_view = torch.einsum("nk,nkj->nj", ch, view)[:, :3]           # the camera centre maps to the view-space origin
    assert in_view.abs().max() < 1e-5


@pytest.mark.parametrize("kind", ["sun", "point"])
def test_lighting_texture_module_returns_shaded_extended_cloud(kind):
    pts, nrm, col, proj, view, cams = scene(200, 3, seed=4)
    tex = LightingTexture(lights=_lights(kind))
    cloud = PointClouds3D([pts], normals=[nrm], features=[col])
    out = tex(cloud, cameras=cams, shininess=16)
    assert len(out) == 3 and out.features_packed().shape == (600, 3)
    # view 1 by hand
    cam = camera_centres(view)[1].expand(200, 3)
    amb, dif, spe = apply_lighting(pts, nrm, _lights(kind), cam, shininess=16)
    torch.testing.assert_close(out.features_packed()[200:400], col * (amb + dif) + spe)
    rows, ambient, k = pack_lights(_lights(kind))
    assert rows.shape == (2, 9) and k == (0 if kind == "sun" else 1)
    torch.testing.assert_close(ambient, torch.tensor([0.4, 0.35, 0.35]))


def _shade64(p, n, alb, lights, view, shininess):
    """(N*P0,3) float64 shaded colours with autograd -- the unfused route's arithmetic."""
    N, P0 = view.shape[0], p.shape[0]
    l64 = lights.double()
    cam = camera_centres(view.double())
    outs = []
    for v in range(N):
        amb, dif, spe = apply_lighting(p, n, l64, cam[v].expand(P0, 3), shininess=shininess)
        outs.append(alb * (amb + dif) + spe)
    return torch.cat(outs, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["sun", "point"])
def test_fused_shading_matches_the_autograd_route(cuda_device, kind):
    from dss_b200.ops import SplatParams, make_shading, render_points
    d = cuda_device
    P0, N, S, shin = 20000, 3, 96, 24.0
    pts, nrm, col, proj, view, cams = scene(P0, N, seed=12)
    g = torch.Generator().manual_seed(1)
    # normals need not be unit length for the shading (it renormalises); |n| <= 1 keeps the EWA covariance of the
    # rasterizer (J^T (I - n n^T) J) positive semi-definite
    nrm = nrm * (0.85 + 0.15 * torch.rand(P0, 1, generator=g))
    lights = _lights(kind, d)
    prm = SplatParams(image_size=S, znear=0.1, clip_pts_grad=-1.0)
    h = torch.full((N,), 3e-4, device=d)
    gi = (torch.randn(N, S, S, 4, generator=g) * 1e-3).to(d)
    projd, viewd = proj.to(d), view.to(d)
    # ---- fused ----
    p1, n1, a1 = (t.to(d).requires_grad_(True) for t in (pts, nrm, col))
    sh = make_shading(lights, viewd, shininess=shin)
    o1 = render_points(p1, n1, a1, projd, viewd, h, prm, shading=sh)
    o1.image.backward(gi)
    # ---- unfused: float64 torch shading feeding the per-(view, point) colour path ----
    p2, n2, a2 = (t.to(d).double().requires_grad_(True) for t in (pts, nrm, col))
    shaded = _shade64(p2, n2, a2, lights, viewd, shin)
    p2f = p2.float()
    o2 = render_points(p2f, n2.float().detach(), shaded.float(), projd, viewd, h, prm)
    o2.image.backward(gi)
    assert torch.equal(o1.idx, o2.idx)
    assert torch.isfinite(o1.image).all() and torch.isfinite(o2.image).all()
    assert float(((o1.image - o2.image) ** 2).mean().detach()) < 1e-10
    torch.testing.assert_close(o1.image, o2.image, rtol=1e-4, atol=2e-6)
    for name, got, want in (("albedo", a1.grad, a2.grad), ("normals", n1.grad, n2.grad), ("points", p1.grad, p2.grad)):
        scale = want.abs().max().item()
        assert scale > 0, name
        err = (got.double() - want).abs().max().item()
        assert err <= 2e-4 * scale, (name, err, scale)
    # points no view sees receive nothing
    seen = o1.visible.view(N, P0).bool().any(0)
    assert (a1.grad[~seen] == 0).all() and (n1.grad[~seen] == 0).all()


@pytest.mark.gpu
def test_fused_shading_through_grad_sync_equals_single_call(cuda_device):
    from dss_b200.ops import SplatParams, make_shading, render_points
    from dss_b200.parallel import GradSync
    d = cuda_device
    P0, N, S = 12000, 2, 64
    pts, nrm, col, proj, view, cams = scene(P0, N, seed=5)
    prm = SplatParams(image_size=S, znear=0.1, clip_pts_grad=0.05)
    h = torch.full((N,), 3e-4, device=d)
    gi = (torch.randn(N, S, S, 4, generator=torch.Generator().manual_seed(2)) * 1e-3).to(d)
    sh = make_shading(_lights("point", d), view.to(d))
    res = []
    for sync in (None, GradSync()):
        p, n, a = (t.to(d).requires_grad_(True) for t in (pts, nrm, col))
        o = render_points(p, n, a, proj.to(d), view.to(d), h, prm, shading=sh, grad_sync=sync)
        o.image.backward(gi)
        torch.cuda.synchronize()
        res.append((p.grad.clone(), n.grad.clone(), a.grad.clone()))
    torch.testing.assert_close(res[0][0], res[1][0], rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(res[0][1], res[1][1], rtol=1e-4, atol=1e-9)
    torch.testing.assert_close(res[0][2], res[1][2], rtol=1e-4, atol=1e-9)
