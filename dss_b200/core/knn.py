"""K nearest neighbours within a radius, for the splat-size rule h_k (rasterizer.py:313-326, 369-388).

Status: "next" row (f)1 of SURVEY.md section 8 -- the reference calls ``frnn.frnn_grid_points(K=7, r=0.2)``
(a 3-D grid search).  Until the 3-D grid search kernel lands this is a chunked exact search written with
torch ops on whatever device the points live on; it is NOT part of the timed hot path (bench.py feeds
``h`` directly, as the metric's timed region starts at the per-point preprocess: SURVEY.md section 8d).
"""
import torch

__all__ = ["knn_sq_dists"]


def knn_sq_dists(points: torch.Tensor, K: int = 7, radius: float = 0.2, chunk: int = 4096) -> torch.Tensor:
    """(P, K) ascending squared distances from every point to its K nearest points of the same cloud (self
    included, distance 0).  Neighbours farther than ``radius`` (when radius > 0) are reported as -1 like frnn
    does for missing neighbours (frnn.py:176-301); the callers take ``max`` over the row, so -1 never wins
    as long as one neighbour exists."""
    P = points.shape[0]
    out = torch.empty((P, K), dtype=points.dtype, device=points.device)
    sq = (points * points).sum(-1)
    for s in range(0, P, chunk):
        q = points[s:s + chunk]
        d2 = (sq[s:s + chunk, None] + sq[None, :] - 2.0 * (q @ points.t())).clamp_(min=0)
        vals = torch.topk(d2, k=min(K, P), dim=1, largest=False)[0]
        if vals.shape[1] < K:
            vals = torch.cat([vals, vals.new_full((vals.shape[0], K - vals.shape[1]), -1.0)], 1)
        if radius is not None and radius > 0:
            vals = torch.where(vals > radius * radius, torch.full_like(vals, -1.0), vals)
        out[s:s + chunk] = vals
    return out
