"""Fused surface-splatting render operator (autograd.Function) over the C ABI.

forward  = dss_render_forward : per-(point,view) EWA preprocess -> tile binning (on-device prefix sum) ->
           per-tile top-K rasterization + normalised blend              (rasterizer.py:584-664, renderer.py:36-82)
backward = dss_render_backward: visibility / median radius -> occupancy gather -> colour scatter ->
           z scatter -> per-point clip -> chain to world space           (rasterizer.py:788-977, 667-673)

Gradients follow the reference exactly: positions receive the occupancy surrogate gradient (and the
z-buffer gradient when one is supplied), colours receive the compositor gradient, the per-point ellipse
parameters are constants (computed under no_grad and detached: rasterizer.py:606-608, 562-565).
"""
import ctypes as C
from typing import NamedTuple, Optional

import torch

from . import _lib

__all__ = ["SplatParams", "Shading", "make_shading", "RenderOutput", "render_points", "preprocess_points"]


class SplatParams(NamedTuple):
    image_size: int = 256
    points_per_pixel: int = 5
    cutoff_threshold: float = 1.0
    depth_merging_threshold: float = 0.05
    antialiasing_sigma: float = 1.0
    radii_backward_scaler: float = 5.0
    clip_pts_grad: float = -1.0
    backface_culling: bool = False
    znear: float = 1.0
    zfar: float = 100.0


class Shading(NamedTuple):
    """Fused per-point shading: pass as ``shading=`` with the per-point albedo as ``colours``.
    Build it from the light objects with :func:`make_shading`."""
    lights: torch.Tensor         # (L,9) {direction | location, diffuse rgb, specular rgb}
    ambient: torch.Tensor        # (3,)
    cam_centres: torch.Tensor    # (N,3) camera centres in world space
    light_type: int = 0          # 0 directional, 1 point
    shininess: float = 64.0


class RenderOutput(NamedTuple):
    image: torch.Tensor      # (N,S,S,4) rgb + occupancy
    idx: torch.Tensor        # (N,S,S,K) int32
    zbuf: Optional[torch.Tensor]
    qvalue: Optional[torch.Tensor]
    weights: torch.Tensor    # (N,S,S,K) normalised blend weights
    visible: torch.Tensor    # (P,) uint8
    records: torch.Tensor    # (P,8) {x,y,z,rx, ry,a,b,c}
    scaler: torch.Tensor     # (P,)


def _fill_shading(a, sh, albedo):
    a.shade = 1
    a.n_lights = int(sh.lights.shape[0])
    a.light_type = int(sh.light_type)
    a.shininess = float(sh.shininess)
    a.albedo, a.lights, a.ambient, a.cam_centres = _lib.ptr(albedo), _lib.ptr(sh.lights), _lib.ptr(sh.ambient), \
        _lib.ptr(sh.cam_centres)


def _check_shading(sh, N, dev):
    if sh.lights.dim() != 2 or sh.lights.shape[1] != 9 or not (1 <= sh.lights.shape[0] <= _lib.MAX_LIGHTS):
        raise RuntimeError("shading.lights must be (L,9) with 1 <= L <= %d" % _lib.MAX_LIGHTS)
    if tuple(sh.ambient.shape) != (3,) or tuple(sh.cam_centres.shape) != (N, 3):
        raise RuntimeError("shading.ambient must be (3,) and shading.cam_centres (N,3)")
    _lib.require_cuda(sh.lights, sh.ambient, sh.cam_centres)
    return Shading(_lib.as_f32(sh.lights.detach(), "lights"), _lib.as_f32(sh.ambient.detach(), "ambient"),
                   _lib.as_f32(sh.cam_centres.detach(), "cam_centres"), int(sh.light_type), float(sh.shininess))


def _fill_common(a, points, normals, colours, proj, view, h, first_idx, num_points, shared, N, P0, P, prm):
    a.points_world = _lib.ptr(points)
    a.normals_world = _lib.ptr(normals)
    a.colours = _lib.ptr(colours)
    a.proj = _lib.ptr(proj)
    a.view = _lib.ptr(view)
    a.h = _lib.ptr(h)
    a.first_idx = _lib.ptr(first_idx)
    a.num_points = _lib.ptr(num_points)
    a.n_views = N
    a.shared_cloud = int(shared)
    a.P0 = P0
    a.P = P
    a.h_per_splat = int(h is not None and h.numel() == P and h.numel() != N)
    a.image_size = int(prm.image_size)
    a.points_per_pixel = int(prm.points_per_pixel)
    a.backface_culling = int(bool(prm.backface_culling))
    a.cutoff_threshold = float(prm.cutoff_threshold)
    a.depth_merging_threshold = float(prm.depth_merging_threshold)
    a.antialiasing_sigma = float(prm.antialiasing_sigma)
    a.znear = float(prm.znear)
    a.zfar = float(prm.zfar)
    a.radii_backward_scaler = float(prm.radii_backward_scaler)
    a.clip_pts_grad = float(prm.clip_pts_grad)


def _layout(points, proj, first_idx, num_points, shared):
    N = proj.shape[0]
    if shared:
        if N > _lib.MAX_SHARED_VIEWS:   # the backward's chain kernel holds the view matrices in shared memory
            raise RuntimeError("a shared cloud takes at most %d views per call, got %d" % (_lib.MAX_SHARED_VIEWS, N))
        P0 = points.shape[0]
        return N, P0, N * P0
    if first_idx is None or num_points is None:
        raise RuntimeError("packed clouds need cloud_to_packed_first_idx and num_points_per_cloud")
    P = points.shape[0]
    # the kernels read both arrays as int64[N] and index the packed arrays with them: check before any launch
    # (their VALUES stay on the device -- reading them back would put a host sync into every forward; rows that no
    #  range covers get zero gradients, see backward)
    for name, t in (("cloud_to_packed_first_idx", first_idx), ("num_points_per_cloud", num_points)):
        if t.dtype != torch.int64 or tuple(t.shape) != (N,):
            raise RuntimeError("%s must be an int64 tensor of shape (%d,), got %s %s"
                               % (name, N, t.dtype, tuple(t.shape)))
    return N, P, P   # P0 is only an upper bound on the points of one view in packed mode


class _RenderFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, normals, colours, proj, view, h, first_idx, num_points, shared, prm, want_frags,
                grad_sync=None, shading=None):
        dev = _lib.require_cuda(points, normals, colours, proj, view, h, first_idx, num_points)
        points_c = _lib.as_f32(points.detach(), "points")
        normals_c = _lib.as_f32(normals.detach(), "normals")
        colours_c = _lib.as_f32(colours.detach(), "colours")
        proj_c = _lib.as_f32(proj.detach(), "proj")
        view_c = _lib.as_f32(view.detach(), "view")
        h_c = _lib.as_f32(h.detach().reshape(-1), "h")
        N, P0, P = _layout(points_c, proj_c, first_idx, num_points, shared)
        if shading is not None:
            if not shared or tuple(colours_c.shape) != (P0, 3):
                raise RuntimeError("fused shading takes a shared cloud and the per-point albedo (P0,3) as colours")
            shading = _check_shading(shading, N, dev)
        shared_col = bool(shared) and (N > 1 or shading is not None) and tuple(colours_c.shape) == (P0, 3)
        if tuple(colours_c.shape) != (P, 3) and not shared_col:
            raise RuntimeError("colours must have shape (%d, 3)%s, got %s"
                               % (P, " or (%d, 3)" % P0 if shared else "", tuple(colours_c.shape)))
        if tuple(proj_c.shape) != (N, 4, 4) or tuple(view_c.shape) != (N, 4, 4):
            raise RuntimeError("proj and view must have shape (N,4,4)")
        if h_c.numel() not in (N, P):
            raise RuntimeError("h must have N or P elements")
        S, K = int(prm.image_size), int(prm.points_per_pixel)
        if K > _lib.MAX_POINTS_PER_PIXEL:
            raise RuntimeError("Must have points_per_pixel <= %d" % _lib.MAX_POINTS_PER_PIXEL)
        f32 = dict(dtype=torch.float32, device=dev)
        records = torch.empty((P, 8), **f32)
        scaler = torch.empty((P,), **f32)
        image = torch.empty((N, S, S, 4), **f32)
        idx = torch.empty((N, S, S, K), dtype=torch.int32, device=dev)
        weights = torch.empty((N, S, S, K), **f32)
        # (capacity rounded up to 4 bytes: the blend epilogue sets the bytes with word atomics when it also counts the
        #  visible splats per backward-binning cell, see dss_render_args.cell_counts)
        visible = torch.empty(((P + 3) // 4 * 4,), dtype=torch.uint8, device=dev)[:P]
        OB = (S + 31) // 32
        cell_counts = torch.empty((N * OB * OB * 1024,), dtype=torch.int32, device=dev) if K <= 8 else None
        zbuf = torch.empty((N, S, S, K), **f32) if want_frags else None
        qvalue = torch.empty((N, S, S, K), **f32) if want_frags else None
        fi = first_idx.contiguous() if first_idx is not None else None
        npts = num_points.contiguous() if num_points is not None else None
        a = _lib.RenderArgs()
        _fill_common(a, points_c, normals_c, colours_c, proj_c, view_c, h_c, fi, npts, shared, N, P0, P, prm)
        a.shared_colours = int(shared_col)
        shaded = None
        if shading is not None:
            shaded = torch.empty((P, 3), **f32)
            _fill_shading(a, shading, colours_c)
            a.shaded = _lib.ptr(shaded)
        a.records, a.scaler, a.image, a.idx = _lib.ptr(records), _lib.ptr(scaler), _lib.ptr(image), _lib.ptr(idx)
        a.weights, a.visible, a.zbuf, a.qvalue = _lib.ptr(weights), _lib.ptr(visible), _lib.ptr(zbuf), _lib.ptr(qvalue)
        a.cell_counts = _lib.ptr(cell_counts)
        with torch.cuda.device(dev):
            rc = _lib.load().dss_render_forward(_lib.ctx(dev), C.byref(a), _lib.stream_ptr(dev))
        _lib.check(rc, "dss_render_forward")
        # no zero-filled gradients for the outputs nobody differentiates (idx, weights, records ... would cost five
        # fill kernels over ~400 MB per backward); backward() handles None
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(points_c, proj_c, view_c, records, idx, weights, visible, fi, npts)
        ctx.shading = (shading, normals_c, colours_c) if shading is not None else None
        ctx.cell_counts = cell_counts
        ctx.meta = (shared, prm, N, P0, P, want_frags, shared_col, grad_sync)
        outs = (image, idx, weights, visible, records, scaler)
        if want_frags:
            ctx.mark_non_differentiable(idx, weights, visible, records, scaler, qvalue)
            return outs + (zbuf, qvalue)
        ctx.mark_non_differentiable(idx, weights, visible, records, scaler)
        return outs

    @staticmethod
    def backward(ctx, grad_image, *rest):
        points_c, proj_c, view_c, records, idx, weights, visible, fi, npts = ctx.saved_tensors
        shared, prm, N, P0, P, want_frags, shared_col, grad_sync = ctx.meta
        dev = points_c.device
        grad_zbuf = rest[5] if (want_frags and len(rest) > 5) else None
        if grad_image is None:
            grad_image = torch.zeros((N, prm.image_size, prm.image_size, 4), dtype=torch.float32, device=dev)
        grad_image = _lib.as_f32(grad_image, "grad_image")
        if grad_zbuf is not None:
            grad_zbuf = _lib.as_f32(grad_zbuf, "grad_zbuf")
        shade = ctx.shading
        # with fused shading the colour scatter fills a (P,3) scratch (d L / d shaded colour) that the shading backward
        # turns into d albedo / d normal / d position
        grad_colours = torch.empty((P if (shade is not None or not shared_col) else P0, 3), dtype=torch.float32, device=dev)
        grad_normals = None
        # packed clouds: rows no view range covers are never written by the kernels
        grad_points = torch.empty_like(points_c) if shared else torch.zeros_like(points_c)
        a = _lib.RenderArgs()
        _fill_common(a, points_c, None, None, proj_c, view_c, None, fi, npts, shared, N, P0, P, prm)
        a.records, a.idx, a.weights, a.visible = _lib.ptr(records), _lib.ptr(idx), _lib.ptr(weights), _lib.ptr(visible)
        a.cell_counts = _lib.ptr(ctx.cell_counts)
        a.shared_colours = int(shared_col)
        a.grad_image, a.grad_zbuf = _lib.ptr(grad_image), _lib.ptr(grad_zbuf)
        a.grad_colours, a.grad_points_world = _lib.ptr(grad_colours), _lib.ptr(grad_points)
        grad_albedo = grad_pshade = None
        if shade is not None:
            sh, normals_c, albedo_c = shade
            grad_albedo, grad_normals, grad_pshade = (torch.empty((P0, 3), dtype=torch.float32, device=dev) for _ in range(3))
            _fill_shading(a, sh, albedo_c)
            a.normals_world = _lib.ptr(normals_c)
            a.grad_albedo, a.grad_normals_world = _lib.ptr(grad_albedo), _lib.ptr(grad_normals)
            a.grad_points_shading = _lib.ptr(grad_pshade)
        lib = _lib.load()
        with torch.cuda.device(dev):
            if grad_sync is None:
                rc = lib.dss_render_backward(_lib.ctx(dev), C.byref(a), _lib.stream_ptr(dev))
                _lib.check(rc, "dss_render_backward")
            else:
                # data-parallel step (dss_b200/parallel.py): the colour gradients are final after the colour scatter --
                # reduce them over the ranks on a side stream while the occupancy path runs here
                main, side = torch.cuda.current_stream(dev), grad_sync.side_stream(dev)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    rc = lib.dss_colour_backward(_lib.ctx(dev), C.byref(a), _lib.stream_ptr(dev))
                    _lib.check(rc, "dss_colour_backward")
                    if shade is not None:
                        shade_done = torch.cuda.Event()
                        shade_done.record(side)
                        grad_sync.reduce_early(grad_albedo, grad_normals)
                    else:
                        grad_sync.reduce_early(grad_colours)
                a.grad_colours = _lib.ptr(None)
                rc = lib.dss_render_backward(_lib.ctx(dev), C.byref(a), _lib.stream_ptr(dev))
                _lib.check(rc, "dss_render_backward")
                if shade is not None:
                    main.wait_event(shade_done)
                    grad_points += grad_pshade           # position gradient through the shading
                grad_sync.reduce_late(grad_points)
                grad_sync.join(dev)
        if shade is not None:
            return (grad_points, grad_normals, grad_albedo) + (None,) * 10
        return (grad_points, None, grad_colours) + (None,) * 10


def render_points(points, normals, colours, proj, view, h, params: SplatParams, first_idx=None, num_points=None,
                  shared_cloud=True, return_fragments=False, grad_sync=None, shading=None) -> RenderOutput:
    """Render ``N`` views of an oriented point cloud to RGBA.

    points, normals : (P0,3) when ``shared_cloud`` (one cloud seen from N cameras) else packed (P,3)
    colours         : (N*P0,3) / (P,3) per-(view,point) features (e.g. shaded rgb), or (P0,3) per-point
                      features used by every view of a shared cloud (gradient then summed over the views)
    proj, view      : (N,4,4) full-projection and world-to-view matrices, row-vector convention
    h               : (N,) per-view or (P,) per-splat variance scale (rasterizer.py:293-402)
    shading         : optional Shading -- `colours` is then the per-point albedo (P0,3) and the colour of every
                      (view, point) is computed in the preprocess kernel: albedo * (ambient + diffuse) + specular
                      (DSS/core/texture.py:74-127); gradients flow to the albedo, the NORMALS and the positions
    grad_sync       : optional dss_b200.parallel.GradSync -- the backward then returns gradients already summed over the
                      ranks of a view-sharded step, the collectives overlapped with the backward kernels
    """
    outs = _RenderFunction.apply(points, normals, colours, proj, view, h, first_idx, num_points,
                                 bool(shared_cloud), params, bool(return_fragments), grad_sync, shading)
    image, idx, weights, visible, records, scaler = outs[:6]
    zbuf, qvalue = (outs[6], outs[7]) if return_fragments else (None, None)
    return RenderOutput(image, idx, zbuf, qvalue, weights, visible, records, scaler)


def make_shading(lights, view_matrices, shininess=64.0) -> Shading:
    """Shading record for the fused route from a DirectionalLights / PointLights object (dss_b200.core.lighting) and the
    (N,4,4) world-to-view matrices of the step's cameras."""
    from .core.lighting import pack_lights
    from .core.texture import camera_centres
    rows, ambient, kind = pack_lights(lights)
    dev = view_matrices.device
    return Shading(rows.to(dev), ambient.to(dev), camera_centres(view_matrices).contiguous().float(), kind, float(shininess))


def preprocess_points(points, normals, proj, view, h, params: SplatParams, first_idx=None, num_points=None,
                      shared_cloud=True):
    """Per-(point,view) screen-space info only (rasterizer.py:525-565 + the transform of :614).

    Returns dict(ndc (P,3), ellipse_params (P,3), radii (P,2), scaler (P,), cutoff_threshold (P,))."""
    dev = _lib.require_cuda(points, normals, proj, view, h, first_idx, num_points)
    points_c, normals_c = _lib.as_f32(points.detach(), "points"), _lib.as_f32(normals.detach(), "normals")
    proj_c, view_c = _lib.as_f32(proj.detach(), "proj"), _lib.as_f32(view.detach(), "view")
    h_c = _lib.as_f32(h.detach().reshape(-1), "h")
    N, P0, P = _layout(points_c, proj_c, first_idx, num_points, shared_cloud)
    f32 = dict(dtype=torch.float32, device=dev)
    ndc, ell = torch.empty((P, 3), **f32), torch.empty((P, 3), **f32)
    radii, scaler = torch.empty((P, 2), **f32), torch.empty((P,), **f32)
    fi = first_idx.contiguous() if first_idx is not None else None
    npts = num_points.contiguous() if num_points is not None else None
    a = _lib.RenderArgs()
    _fill_common(a, points_c, normals_c, None, proj_c, view_c, h_c, fi, npts, shared_cloud, N, P0, P, params)
    a.ndc, a.ellipse, a.radii, a.scaler = _lib.ptr(ndc), _lib.ptr(ell), _lib.ptr(radii), _lib.ptr(scaler)
    with torch.cuda.device(dev):
        rc = _lib.load().dss_preprocess(_lib.ctx(dev), C.byref(a), _lib.stream_ptr(dev))
    _lib.check(rc, "dss_preprocess")
    return {"ndc": ndc, "ellipse_params": ell, "radii": radii, "scaler": scaler,
            "cutoff_threshold": torch.full((P,), float(params.cutoff_threshold), **f32)}
