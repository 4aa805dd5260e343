"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/dss_b200.h declares
(no compute calls without a GPU), the ctypes mirror of dss_render_args matches the C struct, cameras /
clouds / settings / compositor restate the reference conventions, and view sharding + gradient all-reduce
work across 2 gloo processes."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

import oracle
from dss_b200 import _lib
from dss_b200.core.camera import FoVPerspectiveCameras, camera_matrices, look_at_view_transform
from dss_b200.core.cloud import PointClouds3D, PointCloudsFilters
from dss_b200.core.knn import knn_sq_dists
from dss_b200.core.rasterizer import PointsRasterizationSettings, SurfaceSplatting, _splat_params
from dss_b200.core.renderer import NormWeightedCompositor
from dss_b200.parallel import pack_point_grads, shard_views, unpack_point_grads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dss_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"DSS_API[^;(]*?\b(dss_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 20
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), "libdss_b200.so does not export %s" % n
    # and the Python binding declares a prototype for each of them
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)
    assert lib.dss_version() == 1
    assert lib.dss_profile_num_stages() > 5 and lib.dss_profile_stage_name(0)


def test_render_args_struct_layout_matches_c():
    """compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirror."""
    fields = [f[0] for f in _lib.RenderArgs._fields_]
    body = "".join('printf("%s %%zu\\n", offsetof(dss_render_args, %s));\n' % (f, f) for f in fields)
    src = '#include <stddef.h>\n#include <stdio.h>\n#include "dss_b200.h"\nint main(){printf("size %%zu\\n", sizeof(dss_render_args));\n%s return 0;}' % body
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = dict(l.split() for l in subprocess.check_output([exe], text=True).strip().splitlines())
    assert int(out["size"]) == C.sizeof(_lib.RenderArgs)
    for f in fields:
        assert int(out[f]) == getattr(_lib.RenderArgs, f).offset, f


def test_operators_refuse_cpu_tensors_and_missing_gpu():
    from dss_b200 import _C
    z = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        _C.splat_points(z, z, torch.ones(4), torch.ones(4, 2), torch.zeros(1, dtype=torch.int64),
                        torch.full((1,), 4, dtype=torch.int64), 0.05, 16, 5, 0, 0)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            _lib.ctx()


def test_camera_conventions():
    # look-at: camera on +Z looking at the origin sees the origin at NDC (0,0) with view depth = distance
    R, T = look_at_view_transform(dist=2.0, elev=0.0, azim=0.0)
    cams = FoVPerspectiveCameras(znear=0.1, zfar=100.0, R=R, T=T)
    proj, view = camera_matrices(cams)
    o = torch.tensor([[0.0, 0.0, 0.0, 1.0]])
    clip = o @ proj[0]
    assert torch.allclose(clip[0, :2] / clip[0, 3], torch.zeros(2), atol=1e-6)
    assert abs(float((o @ view[0])[0, 2]) - 2.0) < 1e-6
    # +X is LEFT and +Y is UP in NDC: a point to the camera's left has positive ndc x
    cam_c = cams.get_camera_center()[0]
    assert torch.allclose(cam_c, torch.tensor([0.0, 0.0, 2.0]), atol=1e-6)
    up = torch.tensor([[0.0, 0.5, 0.0, 1.0]]) @ proj[0]
    assert float(up[0, 1] / up[0, 3]) > 0
    # fov 60 deg: tan(30deg) at depth 1 maps to |ndc| = 1
    f = 1.0 / np.tan(np.radians(30))
    assert abs(float(proj[0][1, 1]) - f) < 1e-5 or True     # composed with R; check the pure projection instead
    P = cams.get_projection_transform().get_matrix()[0]
    assert abs(float(P[0, 0]) - f) < 1e-5 and abs(float(P[1, 1]) - f) < 1e-5 and float(P[2, 3]) == 1.0
    assert abs(float(P[2, 2]) - 100.0 / 99.9) < 1e-6 and abs(float(P[3, 2]) + 10.0 / 99.9) < 1e-6
    # R is orthonormal, X_view = X_world R + T
    assert torch.allclose(R[0] @ R[0].t(), torch.eye(3), atol=1e-6)
    pts = torch.randn(5, 3)
    vw = cams.get_world_to_view_transform().transform_points(pts)
    assert torch.allclose(vw, pts @ R[0] + T[0], atol=1e-6)


def test_point_cloud_container():
    a, b = torch.randn(5, 3), torch.randn(3, 3)
    pc = PointClouds3D([a, b], normals=[a, b], features=[a, b])
    assert len(pc) == 2 and not pc.isempty() and not pc.equal_sized()
    assert pc.num_points_per_cloud().tolist() == [5, 3] and pc.cloud_to_packed_first_idx().tolist() == [0, 5]
    assert pc.points_packed().shape == (8, 3) and pc.points_padded().shape == (2, 5, 3)
    assert pc.packed_to_cloud_idx().tolist() == [0] * 5 + [1] * 3
    ext = PointClouds3D([a], normals=[a]).extend(4)
    assert len(ext) == 4 and ext.shares_points() and ext.equal_sized()
    assert not pc.shares_points()
    assert PointClouds3D([a[:0]]).isempty()
    # gradients flow through packed accessors back to the leaves
    p = torch.randn(4, 3, requires_grad=True)
    PointClouds3D([p]).extend(3).points_packed().sum().backward()
    assert torch.allclose(p.grad, torch.full((4, 3), 3.0))
    # filters
    filt = PointCloudsFilters(activation=torch.tensor([[True, False, True, True, False]]))
    kept = filt.filter_with(PointClouds3D([a], normals=[a], features=[a]), ("activation",))
    assert kept.points_packed().shape == (3, 3)


def test_settings_defaults_match_reference():
    s = PointsRasterizationSettings()
    # rasterizer.py:73-99
    assert (s.backface_culling, s.cutoff_threshold, s.depth_merging_threshold) == (True, 1, 0.05)
    assert (s.Vrk_invariant, s.Vrk_isotropic, s.radii_backward_scaler) == (False, True, 10)
    assert (s.image_size, s.points_per_pixel, s.bin_size, s.max_points_per_bin) == (256, 8, 0, None)
    assert (s.clip_pts_grad, s.antialiasing_sigma) == (-1, 1.0)
    with pytest.raises(AttributeError):
        s.not_a_setting = 1            # __slots__, like the reference
    cams = FoVPerspectiveCameras(znear=0.1, zfar=50.0)
    prm = _splat_params(s, cams)
    assert prm.znear == pytest.approx(0.1) and prm.zfar == 50.0 and prm.points_per_pixel == 8
    rast = SurfaceSplatting(cameras=cams, raster_settings=s)
    assert rast.cameras is cams and rast.raster_settings is s     # read/mutated live by the trainer (scheduler.py:40-45)
    rast.raster_settings.radii_backward_scaler = 3.0
    assert _splat_params(rast.raster_settings, cams).radii_backward_scaler == 3.0


def test_compositor_matches_oracle_blend():
    rng = np.random.default_rng(0)
    N, S, K, P = 2, 8, 5, 40
    idx = rng.integers(-1, P, (N, S, S, K)).astype(np.int32)
    q = rng.random((N, S, S, K)).astype(np.float32)
    scaler = rng.random(P).astype(np.float32) + 0.1
    col = rng.random((P, 3)).astype(np.float32)
    occ = (idx[..., 0] >= 0).astype(np.float32)
    want = oracle.blend_forward(idx, q, occ, scaler, col)
    t = torch.from_numpy
    frag_scaler = torch.where(t(idx) >= 0, t(scaler)[t(idx).clamp(min=0).long()], torch.zeros(1))
    w = (torch.exp(-0.5 * t(q)) * frag_scaler).permute(0, 3, 1, 2)
    img = NormWeightedCompositor()(t(idx).long().permute(0, 3, 1, 2), w, t(col).t()).permute(0, 2, 3, 1)
    np.testing.assert_allclose(img.numpy(), want[..., :3], rtol=1e-5, atol=1e-6)


def test_knn_is_cuda_only_and_fails_loudly_on_cpu():
    """the splat-size K-NN has no CPU path either (tests/test_knn.py covers the kernel and the oracle)."""
    from dss_b200.frnn_grid import frnn_grid_points, knn_points_packed
    pts = torch.rand(300, 3, generator=torch.Generator().manual_seed(0))
    with pytest.raises(RuntimeError, match="CUDA"):
        knn_sq_dists(pts, K=7, radius=0.2)
    with pytest.raises(RuntimeError, match="CUDA"):
        knn_points_packed(pts, torch.zeros(1, dtype=torch.int64), torch.full((1,), 300, dtype=torch.int64), 7, 0.2)
    with pytest.raises(RuntimeError):
        frnn_grid_points(pts[None], pts[None], K=7, r=0.2)


def test_shard_views_partitions_contiguously():
    for n, w in ((64, 8), (10, 4), (3, 8), (256, 8)):
        spans = [shard_views(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_views(4, 2, 2)
    buf = pack_point_grads(torch.ones(5, 3), None, 2 * torch.ones(5, 3), 5)
    gp, gn, gc = unpack_point_grads(buf)
    assert buf.shape == (5, 9) and gp.eq(1).all() and gn.eq(0).all() and gc.eq(2).all()


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from dss_b200.parallel import (shard_views, pack_point_grads, allreduce_point_grads, allreduce_visibility, GradSync,
                               assign_views)
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank, P0, V = dist.get_rank(), 50, 5
g = torch.Generator().manual_seed(7)
per_view = torch.randn(V, P0, 3, generator=g)            # identical on both ranks: per-view point gradients
a, b = shard_views(V, 2, rank)
buf = pack_point_grads(per_view[a:b].sum(0), None, per_view[a:b].sum(0) * 2, P0)
allreduce_point_grads(buf)
want = pack_point_grads(per_view.sum(0), None, per_view.sum(0) * 2, P0)
assert torch.allclose(buf, want, atol=1e-5), (buf - want).abs().max()
vis = torch.zeros(P0, dtype=torch.bool); vis[rank::7] = True
both = allreduce_visibility(vis)
exp = torch.zeros(P0, dtype=torch.bool); exp[0::7] = True; exp[1::7] = True
assert torch.equal(both, exp)
# the overlapped exchange of a view-sharded step: views dealt by cost, colour-side gradients reduced early, position
# gradients late -- every rank ends up with the sum over ALL views
costs = [5.0, 1.0, 3.0, 2.0, 4.0]
mine = assign_views(costs, 2)[rank]
sync = GradSync()
assert sync.world_size == 2
gc, gp = per_view[mine].sum(0) * 2, per_view[mine].sum(0)
sync.reduce_early(gc, None)
sync.reduce_late(gp)
assert torch.allclose(gp, per_view.sum(0), atol=1e-5) and torch.allclose(gc, 2 * per_view.sum(0), atol=1e-5)
assert sync.calls == 1 and sync.bytes_reduced == 2 * P0 * 3 * 4
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_assign_views_balances_count_and_cost():
    from dss_b200.parallel import assign_views, view_costs_from_cameras
    costs = [9.0, 1.0, 1.0, 8.0, 2.0, 7.0, 3.0, 6.0, 4.0, 5.0, 5.0, 4.0, 6.0, 3.0, 7.0, 2.0]
    for world in (1, 2, 4, 8):
        parts = assign_views(costs, world)
        assert sorted(v for p in parts for v in p) == list(range(len(costs)))
        assert {len(p) for p in parts} == {len(costs) // world}
        sums = [sum(costs[v] for v in p) for p in parts]
        assert max(sums) - min(sums) <= max(costs)
        # contiguous slices of the same views are worse (or equal) for the slowest rank
        per = len(costs) // world
        contiguous = max(sum(costs[r * per:(r + 1) * per]) for r in range(world))
        assert max(sums) <= contiguous + 1e-9
    assert assign_views([1.0, 2.0, 3.0], 2) == [[2], [0, 1]]            # uneven count: 3 | 1 + 2
    view = torch.eye(4).repeat(3, 1, 1)
    view[:, 3, :3] = torch.tensor([[0.0, 0.0, 1.0], [0.0, 0.0, 2.0], [0.0, 2.0, 2.0]])
    c = view_costs_from_cameras(view)
    assert c[0] > c[1] > c[2]


def test_view_sharded_allreduce_two_gloo_processes():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    with tempfile.TemporaryDirectory() as d:
        w = os.path.join(d, "worker.py")
        open(w, "w").write(_WORKER % {"root": ROOT, "port": port})
        procs = [subprocess.Popen([sys.executable, w, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                  text=True) for r in range(2)]
        outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "rank %d ok" % r in o, o


class _ForeignClouds:
    """Only the accessor set of pytorch3d.structures.Pointclouds that the path touches (no shares_points/equal_sized);
    like pytorch3d's extend(), every cloud holds its own CLONE of the data."""

    def __init__(self, pts, nrm, n):
        self._p = [pts.clone() for _ in range(n)]
        self._n = [nrm.clone() for _ in range(n)]
        self.equisized = True

    def points_list(self):
        return self._p

    def normals_list(self):
        return self._n

    def __len__(self):
        return len(self._p)


def test_foreign_cloud_containers_are_duck_typed():
    from dss_b200.core.cloud import PointClouds3D, clouds_equal_sized, clouds_share_points
    g = torch.Generator().manual_seed(0)
    pts, nrm = torch.randn(50, 3, generator=g), torch.randn(50, 3, generator=g)
    foreign = _ForeignClouds(pts, nrm, 3)
    assert clouds_equal_sized(foreign) and clouds_share_points(foreign)          # cloned, same content: shared
    foreign._p[2] = foreign._p[2] + 1e-3
    assert not clouds_share_points(foreign)
    foreign._p[2] = torch.randn(40, 3, generator=g)
    foreign.equisized = False
    assert not clouds_share_points(foreign) and not clouds_equal_sized(foreign)
    ours = PointClouds3D([pts], normals=[nrm]).extend(3)
    assert clouds_share_points(ours) and clouds_equal_sized(ours)
