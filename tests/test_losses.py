"""Projection / repulsion regularisers against a literal per-point restatement of
DSS/training/losses.py:282-497 in float64 numpy loops; the K-NN comes from a brute-force search on the CPU here and from
the CUDA kernel in the GPU test (both must give the same loss)."""
import numpy as np
import pytest
import torch

from dss_b200.core.cloud import PointClouds3D, PointCloudsFilters
from dss_b200.training.losses import ProjectionLoss, RepulsionLoss
from tests.util import sphere_cloud


def _brute_knn(pts, K):
    d = torch.cdist(pts.double(), pts.double()) ** 2
    dist, idx = torch.sort(d, dim=1)
    # ties by index like the CUDA kernel (stable sort)
    return dist[:, :K].float()[None], idx[:, :K][None], pts[idx[:, :K]][None]


def _reference_loops(pts, nrm, vis, K, sigma, filter_scale, kind):
    """per point, per neighbour, straight from the formulas of losses.py (float64)."""
    P = pts.shape[0]
    p, n = pts.double().numpy(), nrm.double().numpy()
    dist, idx, _ = _brute_knn(pts, K)
    dist, idx = dist[0, :, 1:].double().numpy(), idx[0, :, 1:].numpy()
    h = dist.mean(1, keepdims=True) * 4
    phi = np.clip(1 - dist / h, 0, None) ** 4
    den = (n[idx] * phi[..., None]).sum(1) / np.maximum(phi.sum(1, keepdims=True), 1e-17)
    den[vis] = n[vis]
    unit = lambda a: a / np.maximum(np.linalg.norm(a, axis=-1, keepdims=True), 1e-12)
    nd = unit(den)
    normal_w = np.exp(-((nd[idx] - nd[:, None]) ** 2).sum(-1) / sigma ** 2)
    if kind == "proj":
        vis_w = np.where(vis[idx], 1.0, 0.1)
        w = phi * normal_w * vis_w
        sdf = ((p[idx] - p[:, None]) * den[idx]).sum(-1)
        return (w * sdf ** 2).sum(1) / np.maximum(w.sum(1), 1e-17)
    diag2 = ((p.max(0) - p.min(0)) ** 2).sum()
    spatial = np.exp(-((p[idx] - p[:, None]) ** 2).sum(-1) * (P / diag2) * filter_scale)
    w = spatial * normal_w
    diff = p[:, None] - p[idx]
    proj = diff - (diff * den[idx]).sum(-1, keepdims=True) * den[idx]
    repel = (proj * w[..., None]).sum(1) / np.maximum(w.sum(1, keepdims=True), 1e-17)
    repel = repel * (spatial.sum(1, keepdims=True) + 1.0)
    return np.exp(-np.abs(repel))


@pytest.mark.parametrize("kind", ["proj", "repel"])
def test_losses_match_the_formulas_on_cpu(kind):
    P, K = 400, 12
    pts, nrm, _ = sphere_cloud(P, seed=2)
    g = torch.Generator().manual_seed(0)
    pts = pts + 0.01 * torch.randn(P, 3, generator=g)
    nrm = torch.nn.functional.normalize(nrm + 0.2 * torch.randn(P, 3, generator=g), dim=1)
    vis = torch.rand(P, generator=g) < 0.6
    cloud = PointClouds3D([pts.clone().requires_grad_(True)], normals=[nrm])
    filt = PointCloudsFilters(visibility=vis[None], inmask=torch.ones(1, P, dtype=torch.bool))
    Loss = ProjectionLoss if kind == "proj" else RepulsionLoss
    loss = Loss(reduction="none", filter_scale=2.0, knn_k=K)
    out = loss(cloud, points_filter=filt, rebuild_knn=True, knn_tree=_brute_knn(pts, K))
    want = _reference_loops(pts, nrm, vis.numpy(), K, 0.75, 2.0, kind)
    np.testing.assert_allclose(out.detach().numpy(), want, rtol=2e-4, atol=1e-7)
    # gradients exist and only flow through the query point (neighbours and weights are detached)
    out.sum().backward()
    gpts = cloud.points_list()[0].grad
    assert gpts is not None and torch.isfinite(gpts).all() and gpts.abs().sum() > 0
    mean = Loss(reduction="mean", filter_scale=2.0, knn_k=K)(cloud, points_filter=filt, knn_tree=_brute_knn(pts, K))
    assert abs(float(mean.detach()) - want.mean()) < 1e-4 * max(1.0, abs(want.mean()))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["proj", "repel"])
def test_losses_on_the_cuda_knn_equal_the_bruteforce_tree(cuda_device, kind):
    P, K = 20000, 12
    pts, nrm, _ = sphere_cloud(P, seed=7)
    d = cuda_device
    vis = (torch.rand(P, generator=torch.Generator().manual_seed(1)) < 0.5)
    Loss = ProjectionLoss if kind == "proj" else RepulsionLoss
    outs = []
    for tree in (None, "brute"):
        p = pts.to(d).requires_grad_(True)
        cloud = PointClouds3D([p], normals=[nrm.to(d)])
        filt = PointCloudsFilters(device=d, visibility=vis[None].to(d), inmask=torch.ones(1, P, dtype=torch.bool, device=d))
        kw = {}
        if tree is not None:
            dist = torch.cdist(pts.to(d).double(), pts.to(d).double()) ** 2
            dd, ii = torch.sort(dist, dim=1)
            kw["knn_tree"] = (dd[:, :K].float()[None], ii[:, :K][None], None)
        val = Loss(reduction="mean", filter_scale=2.0, knn_k=K)(cloud, points_filter=filt, rebuild_knn=True, **kw)
        val.backward()
        outs.append((float(val), p.grad.clone()))
    assert abs(outs[0][0] - outs[1][0]) <= 1e-5 * max(1.0, abs(outs[1][0]))
    torch.testing.assert_close(outs[0][1], outs[1][1], rtol=1e-3, atol=1e-7)
