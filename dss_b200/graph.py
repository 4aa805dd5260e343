"""CUDA-graph replay of a whole render step (forward + backward) for loops with fixed shapes.

A step is ~30 kernel launches and two ctypes calls; at the headline size (1.7 ms of kernels) the launches hide behind the
GPU work, at the smaller BASELINE clouds (100k points: 0.6 ms) they do not.  The C library never synchronises or
allocates in steady state (DESIGN.md "Host side"), so the whole step can be captured once and replayed:

    step = GraphedRenderStep(points, normals, colours, proj, view, h, params, grad_image)
    step.replay()                    # image in step.image, gradients in step.grad_points / step.grad_colours
    step.points.data.add_(...)       # update the step's OWN static leaves in place (optimizer step), replay again

The step owns its differentiated inputs (`step.points`, `step.colours`, `step.normals`: fresh leaf copies of what was passed
in): a leaf that has already been through an eager backward on the default stream carries an AccumulateGrad node bound
to that stream, and synchronising with the legacy default stream is not capturable.

Sizes decided on the host at capture time are frozen into the graph: the forward's tile-list capacity and the staged
window of the occupancy gather.  Both have on-device fallbacks (tiles whose list outgrew the buffer are rasterized from
the records; views whose window does not fit take the direct gather), so a replay is always CORRECT; `stale()` tells when
re-capturing would make it faster again.
"""
import torch

from . import _lib
from .ops import render_points

__all__ = ["GraphedRenderStep"]


class GraphedRenderStep:
    def __init__(self, points, normals, colours, proj, view, h, params, grad_image, shading=None, warmup=3,
                 grad_sync=None):
        dev = _lib.require_cuda(points, normals, colours, proj, view, h, grad_image)
        self.device = dev
        leaf = lambda t: t.detach().clone().requires_grad_(True)
        self.points, self.colours = leaf(points), leaf(colours)
        self.normals = leaf(normals) if shading is not None else normals.detach().clone()
        self._args = (self.points, self.normals, self.colours, proj, view, h, params)
        self._shading = shading
        if grad_sync is not None and grad_sync.world_size > 1:
            # measured on 2 x B200: capturing the two overlapped NCCL all-reduces (issued from two streams inside the
            # backward) deadlocks at replay -- a view-sharded step is launched eagerly
            raise NotImplementedError("GraphedRenderStep does not capture the multi-GPU gradient exchange; "
                                      "call render_points(..., grad_sync=...) eagerly")
        self._sync = grad_sync
        self.grad_image = grad_image
        # warm-up on a side stream (sizes the library's scratch and the caching allocator), then capture
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(max(1, warmup)):
                self._eager()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self._capacity_at_capture = self._tile_total()
        self._clear_grads()
        self.graph = torch.cuda.CUDAGraph()
        timing = getattr(grad_sync, "timing", False)
        if grad_sync is not None:
            grad_sync.timing = False    # no timing events inside a capture
        try:
            with torch.cuda.graph(self.graph):
                out = self._eager()
        finally:
            if grad_sync is not None:
                grad_sync.timing = timing
        self.image, self.visible = out.image, out.visible
        self.grad_points, self.grad_colours = self.points.grad, self.colours.grad
        self.grad_normals = self.normals.grad if shading is not None else None

    def _clear_grads(self):
        for t in self._args[:3]:
            if t.requires_grad:
                t.grad = None

    def _eager(self):
        self._clear_grads()
        points, normals, colours, proj, view, h, params = self._args
        out = render_points(points, normals, colours, proj, view, h, params, shading=self._shading, grad_sync=self._sync)
        out.image.backward(self.grad_image)
        return out

    def _tile_total(self):
        """the tile-list size the device published last (mapped pinned word; no synchronisation)"""
        return int(_lib.load().dss_debug_tile_total(_lib.ctx(self.device)))

    def replay(self):
        self.graph.replay()
        return self.image

    def stale(self, slack=1.2) -> bool:
        """True when the tile lists have outgrown what they were at capture time by more than `slack` (the replay is
        still correct -- overflowing tiles are rasterized from the records -- but a fresh capture will be faster)."""
        now = self._tile_total()
        return self._capacity_at_capture > 0 and now > slack * 1.25 * self._capacity_at_capture
