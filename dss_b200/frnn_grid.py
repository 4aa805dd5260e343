"""Drop-in for the two FRNN entry points the rasterizer backward used (``frnn._C.insert_points_cuda``,
``frnn._C.counting_sort_cuda``; external/FRNN/frnn/csrc/grid/grid.h:43-50, counting_sort.h:4-11), D = 2.

The B200 backward does not need them (it gathers per point, see _C.occ_backward); they are kept so code
written against the reference's radius-binning primitives keeps working."""
import torch

from . import _lib

__all__ = ["insert_points_cuda", "counting_sort_cuda"]


def insert_points_cuda(points, lengths, params, grid_cnt, grid_cell, grid_idx, G):
    dev = _lib.require_cuda(points, lengths, params, grid_cnt, grid_cell, grid_idx)
    if points.dim() != 3 or points.shape[2] != 2:
        raise RuntimeError("for now only 2D is supported by the B200 build (got D=%s)" % (points.shape[-1],))
    if not (grid_cnt.dtype == grid_cell.dtype == grid_idx.dtype == torch.int32):
        raise RuntimeError("grid_cnt, grid_cell, grid_idx must be int32")
    for t in (grid_cnt, grid_cell, grid_idx):
        if not t.is_contiguous():
            raise RuntimeError("output tensors must be contiguous")
    N, P = points.shape[0], points.shape[1]
    with torch.cuda.device(dev):
        rc = _lib.load().dss_grid_insert_points_2d(
            _lib.ctx(dev), _lib.ptr(_lib.as_f32(points, "points")), _lib.ptr(lengths.contiguous()),
            _lib.ptr(_lib.as_f32(params, "params")), _lib.ptr(grid_cnt), _lib.ptr(grid_cell), _lib.ptr(grid_idx),
            N, P, int(G), _lib.stream_ptr(dev))
    _lib.check(rc, "dss_grid_insert_points_2d")


def counting_sort_cuda(points, lengths, grid_cell, grid_idx, grid_off, sorted_points, sorted_points_idxs):
    dev = _lib.require_cuda(points, lengths, grid_cell, grid_idx, grid_off, sorted_points, sorted_points_idxs)
    if points.dim() != 3 or points.shape[2] != 2:
        raise RuntimeError("for now only 2D is supported by the B200 build")
    for t in (sorted_points, sorted_points_idxs):
        if not t.is_contiguous():
            raise RuntimeError("output tensors must be contiguous")
    N, P, G = points.shape[0], points.shape[1], grid_off.shape[1]
    with torch.cuda.device(dev):
        rc = _lib.load().dss_grid_counting_sort_2d(
            _lib.ctx(dev), _lib.ptr(_lib.as_f32(points, "points")), _lib.ptr(lengths.contiguous()),
            _lib.ptr(grid_cell.contiguous()), _lib.ptr(grid_idx.contiguous()), _lib.ptr(grid_off.contiguous()),
            _lib.ptr(sorted_points), _lib.ptr(sorted_points_idxs), N, P, G, _lib.stream_ptr(dev))
    _lib.check(rc, "dss_grid_counting_sort_2d")
