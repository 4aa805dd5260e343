"""GPU tests of the paths round 1 left untested: backface culling, per-splat (isotropic) h, the tile-list overflow
path of the sync-free forward, and the reference-minted golden fixtures fed straight to the CUDA operator."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle
from dss_b200 import _C, _lib
from dss_b200.ops import SplatParams, preprocess_points, render_points
from tests.util import packed_offsets, random_screen_splats, scene

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


def test_backface_culling_equals_rendering_the_filtered_cloud(cuda_device):
    """The reference drops points whose view-space normal has z >= 0 and renders the compacted cloud
    (rasterizer.py:148-181, 219-254).  Ours keeps the slots (z = -1): the image must be the same bit for bit and idx
    must be the same up to the compaction map."""
    d = cuda_device
    P0, N, S = 40000, 3, 128
    pts, nrm, col, proj, view, _ = scene(P0, N, seed=21)
    prm = SplatParams(image_size=S, znear=0.1, zfar=100.0, backface_culling=True)
    h = torch.full((N,), 3e-4, device=d)
    colours = (col.repeat(N, 1) * torch.linspace(0.4, 1.0, N).repeat_interleave(P0)[:, None]).to(d)
    out = render_points(pts.to(d), nrm.to(d), colours, proj.to(d), view.to(d), h, prm, return_fragments=True)
    kept = (out.records[:, 2] >= 0).view(N, P0)
    # the mask is the reference's: n_view.z < 0 and znear <= z_view <= zfar (float64 restatement, rim cases excluded)
    nz = (nrm.double() @ view[:, :3, 2].double().T).T                              # (N,P0)
    ph = torch.cat([pts.double(), torch.ones(P0, 1, dtype=torch.float64)], 1)
    zv = (ph @ view[:, :, 2].double().T).T
    want = (nz < 0) & (zv >= 0.1) & (zv <= 100.0)
    sure = nz.abs() > 1e-6
    assert torch.equal(kept.cpu()[sure], want[sure])
    assert 0.3 < kept.float().mean().item() < 0.7                                   # about half the sphere faces away
    # render the compacted clouds through the packed entry point, culling off
    keep_flat = kept.reshape(-1)
    pts_p = pts.to(d).repeat(N, 1)[keep_flat].contiguous()
    nrm_p = nrm.to(d).repeat(N, 1)[keep_flat].contiguous()
    col_p = colours[keep_flat].contiguous()
    num = kept.sum(1).to(torch.int64)
    first = torch.cumsum(num, 0) - num
    out2 = render_points(pts_p, nrm_p, col_p, proj.to(d), view.to(d), h, prm._replace(backface_culling=False),
                         first_idx=first, num_points=num, shared_cloud=False, return_fragments=True)
    assert torch.equal(out.image, out2.image)
    remap = (torch.cumsum(keep_flat.long(), 0) - 1).to(torch.int32)
    mapped = torch.where(out.idx >= 0, remap[out.idx.clamp(min=0).long()], torch.full_like(out.idx, -1))
    assert torch.equal(mapped, out2.idx)
    assert torch.equal(out.zbuf, out2.zbuf) and torch.equal(out.qvalue, out2.qvalue)
    # culled points are never visible and never receive gradients
    assert (out.visible.view(N, P0)[~kept] == 0).all()


def test_per_splat_h_matches_f64_oracle_and_uniform_h(cuda_device):
    """Vrk_isotropic (the settings default): one h per (view, point) (rasterizer.py:344-402)."""
    d = cuda_device
    P0, N, S = 6000, 2, 128
    pts, nrm, col, proj, view, _ = scene(P0, N, seed=8)
    prm = SplatParams(image_size=S, znear=0.1)
    g = torch.Generator().manual_seed(5)
    h = (torch.rand(N * P0, generator=g) * (1e-2 - 5e-5) + 5e-5)                    # clamp range of :388
    pre = preprocess_points(pts.to(d), nrm.to(d), proj.to(d), view.to(d), h.to(d), prm)
    want = oracle.preprocess_f64(_np(proj), _np(view), _np(pts), _np(nrm), _np(h), 1.0, 1.0, S)
    np.testing.assert_allclose(_np(pre["ndc"]), want["ndc"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(_np(pre["radii"]), want["radii"], rtol=2e-4, atol=1e-7)
    mag = np.abs(want["ellipse"]).max(axis=1, keepdims=True)
    assert (np.abs(_np(pre["ellipse_params"]) - want["ellipse"]) <= 5e-4 * mag).all()
    np.testing.assert_allclose(_np(pre["scaler"]), want["scaler"], rtol=2e-3, atol=2e-3 * want["scaler"].max())
    # a constant per-splat h renders exactly like the same per-view h
    colours = col.repeat(N, 1).to(d)
    a = render_points(pts.to(d), nrm.to(d), colours, proj.to(d), view.to(d), torch.full((N * P0,), 3e-4, device=d), prm)
    b = render_points(pts.to(d), nrm.to(d), colours, proj.to(d), view.to(d), torch.full((N,), 3e-4, device=d), prm)
    assert torch.equal(a.image, b.image) and torch.equal(a.idx, b.idx)
    # and the varying h renders like the oracle rasterizer fed with the CUDA per-point info
    o = render_points(pts.to(d), nrm.to(d), colours, proj.to(d), view.to(d), h.to(d), prm, return_fragments=True)
    first, num = packed_offsets(N, P0)
    widx, wz, wq, wocc = oracle.splat_points_binned(_np(pre["ndc"]), _np(pre["ellipse_params"]),
                                                    np.ones(N * P0, np.float32), _np(pre["radii"]), _np(first),
                                                    _np(num), 0.05, S, 5, 16, fma_mode=1)
    same = (_np(o.idx) == widx).all(-1)
    assert same.mean() > 0.9995
    wimg = oracle.blend_forward(widx, wq, wocc, _np(pre["scaler"]), _np(colours))
    assert float(((_np(o.image) - wimg) ** 2).mean()) < 1e-5


@pytest.mark.parametrize("P0,N,S,K", [(200_000, 3, 256, 5), (60_000, 2, 128, 8)])
def test_tile_list_overflow_path_gives_identical_output(cuda_device, P0, N, S, K):
    """The forward never waits for the size of the tile lists: a list that has outgrown the buffer (sized from the
    previous call) is detected on the device and those tiles take their candidates from the records.  Forced here by
    capping the buffer at a fraction of what the lists need: every output must be bit-identical."""
    d = cuda_device
    pts, nrm, col, proj, view, _ = scene(P0, N, seed=31)
    prm = SplatParams(image_size=S, points_per_pixel=K, znear=0.1)
    h = torch.full((N,), 1e-4, device=d)
    args = (pts.to(d), nrm.to(d), col.to(d), proj.to(d), view.to(d), h, prm)
    ref = render_points(*args, return_fragments=True)
    try:
        for frac in (0.5, 0.05, 0.0):
            _lib.limit_tile_capacity(max(1, int(frac * 1.7 * N * P0)), d)
            for with_stats in (False, True):       # the production kernel and its counting twin
                if with_stats:
                    _lib.raster_stats(True, d)
                out = render_points(*args, return_fragments=True)
                if with_stats:
                    assert _lib.raster_stats(False, d)["overflow_tiles"] > 0
                for a, b in zip((out.image, out.idx, out.zbuf, out.qvalue, out.weights, out.visible),
                                (ref.image, ref.idx, ref.zbuf, ref.qvalue, ref.weights, ref.visible)):
                    assert torch.equal(a, b)
    finally:
        _lib.limit_tile_capacity(0, d)
    again = render_points(*args, return_fragments=True)
    assert torch.equal(again.idx, ref.idx) and torch.equal(again.image, ref.image)


def test_lists_that_outgrow_the_buffer_stay_exact(cuda_device):
    """Lists that grow from call to call (more points) overflow a buffer sized for an earlier call (emulated by a cap
    that fits only the first cloud); the fused path must keep agreeing with the operator-level entry point run without
    the cap."""
    d = cuda_device
    S, K = 128, 5
    prm = SplatParams(image_size=S, points_per_pixel=K, znear=0.1)
    for P0 in (5_000, 40_000, 160_000):
        pts, nrm, col, proj, view, _ = scene(P0, 2, seed=P0)
        h = torch.full((2,), 2e-4, device=d)
        try:
            _lib.limit_tile_capacity(25_000, d)
            out = render_points(pts.to(d), nrm.to(d), col.to(d), proj.to(d), view.to(d), h, prm, return_fragments=True)
        finally:
            _lib.limit_tile_capacity(0, d)
        rec = out.records
        first, num = (t.to(d) for t in packed_offsets(2, P0))
        idx2, z2, q2, _ = _C.splat_points(rec[:, :3].contiguous(), rec[:, 5:8].contiguous(),
                                          torch.ones(2 * P0, device=d), rec[:, 3:5].contiguous(), first, num, 0.05, S, K)
        assert torch.equal(idx2, out.idx) and torch.equal(z2, out.zbuf) and torch.equal(q2, out.qvalue)


GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith("knn_"))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_golden_fixtures_through_the_cuda_operator(cuda_device, path):
    """tests/golden/*.npz hold the outputs of the reference's own RasterizePointsNaiveCpu / ZbufBackwardCpu on seeded
    inputs (tests/golden/make_golden.py).  The CUDA operator must reproduce idx / zbuf / occupancy bit for bit and q to
    the last bits (the x86 build does not contract a*dx*dx + b*dx*dy + c*dy*dy into FMAs, nvcc does: DESIGN 3)."""
    d = cuda_device
    g = np.load(path)
    S, K = int(g["S"]), int(g["K"])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(d)
    idx, zbuf, q, occ = _C.splat_points(t(g["points"]), t(g["ellipse"]), t(g["cutoff"]), t(g["radii"]), t(g["first"]),
                                        t(g["num"]), 0.05, S, K)
    same = (_np(idx) == g["idx"]).all(-1)
    # rim fragments (|q - cutoff| within a few ulp) may flip with the contraction; nothing else may differ
    if not same.all():
        cut = g["cutoff"]
        bad = ~same
        ok = np.zeros_like(bad)
        for (ii, qq) in ((_np(idx), _np(q)), (g["idx"], g["qvalue"])):
            c = np.where(ii >= 0, cut[np.maximum(ii, 0)], np.inf)
            ok |= ((np.abs(qq - c) <= 8 * np.spacing(np.abs(c).astype(np.float32))) & (ii >= 0)).any(-1)
        assert (ok | ~bad).all()
        assert bad.mean() < 2e-3
    assert np.array_equal(_np(zbuf)[same], g["zbuf"][same])
    assert np.array_equal(_np(occ)[same], g["occ"][same])
    np.testing.assert_allclose(_np(q)[same], g["qvalue"][same], rtol=1e-5, atol=1e-6)
    # z-buffer backward on the reference's fragments
    P = g["points"].shape[0]
    zg = torch.zeros(P, 1, device=d)
    _C._backward_zbuf(t(g["idx"]), t(g["grad_zbuf"]), zg)
    np.testing.assert_allclose(_np(zg).reshape(-1), g["zbuf_backward"].reshape(-1), rtol=1e-5, atol=1e-6)
