"""Per-point shading -- the twin of DSS/core/texture.py:25-127 (`apply_lighting`, `LightingTexture`).

Two routes, same numbers:
  * `LightingTexture` (torch ops, autograd): features of the returned cloud = rgb * (ambient + diffuse) + specular,
    the module the reference's model calls before the renderer;
  * the fused route: `dss_b200.ops.render_points(..., shading=Shading(...))` evaluates the same formulas inside the
    preprocess kernel and back-propagates to albedo, normals and positions in one kernel.
"""
import torch
import torch.nn as nn

from .camera import camera_matrices
from .lighting import DirectionalLights

__all__ = ["apply_lighting", "camera_centres", "LightingTexture"]


def camera_centres(view_matrices: torch.Tensor) -> torch.Tensor:
    """(N,4,4) world-to-view matrices, row-vector convention [p 1] V  ->  (N,3) camera centres: c R + T = 0."""
    R, T = view_matrices[:, :3, :3], view_matrices[:, 3, :3]
    return -(T.unsqueeze(1) @ R.transpose(1, 2)).squeeze(1)


def apply_lighting(points, normals, lights, camera_position, shininess=64, view_idx=None):
    """-> (ambient (..,3), diffuse, specular) of DSS/core/texture.py:25-62.  Packed (P,3) inputs with the per-point
    view index `view_idx` (P,) and per-point `camera_position` (P,3), or a single view."""
    amb = lights.ambient_color
    amb = amb.sum(1) if amb.dim() == 3 else amb                        # texture.py:52-55
    amb = amb if (view_idx is None or amb.shape[0] == 1) else amb[view_idx]
    dif = lights.diffuse(normals=normals, points=points, view_idx=view_idx)
    spe = lights.specular(normals=normals, points=points, camera_position=camera_position, shininess=shininess,
                          view_idx=view_idx)
    return amb, dif, spe


class LightingTexture(nn.Module):
    """texture.py:65-127: ``forward(pointclouds, shininess=64, lights=, cameras=, points_rgb=)`` returns a clone of the
    (extended) cloud whose features are the shaded colours."""

    def __init__(self, device="cpu", cameras=None, lights=None, materials=None, specular=True):
        super().__init__()
        self.lights = lights if lights is not None else DirectionalLights(device=device)
        self.cameras = cameras
        self.specular = specular

    def forward(self, pointclouds, shininess=64, **kwargs):
        if pointclouds.isempty():
            return pointclouds
        lights = kwargs.get("lights", self.lights).to(pointclouds.device)
        cameras = kwargs.get("cameras", self.cameras)
        _, view = camera_matrices(cameras)
        view = view.to(pointclouds.device)
        if len(pointclouds) == 1 and view.shape[0] != 1:
            pointclouds = pointclouds.extend(view.shape[0])              # texture.py:90-91
        points, normals = pointclouds.points_packed(), pointclouds.normals_packed()
        if normals is None:
            raise ValueError("point normals are required for shading")
        points_rgb = kwargs.get("points_rgb", None)
        if points_rgb is None:
            feats = pointclouds.features_packed()
            points_rgb = feats[:, :3] if feats is not None else torch.ones_like(points)
        if points_rgb.shape[-1] != 3:
            raise ValueError("Expected points_rgb to be 3-channel, got %s" % (tuple(points_rgb.shape),))
        idx = pointclouds.packed_to_cloud_idx()
        cam = camera_centres(view)[idx]
        amb, dif, spe = apply_lighting(points, normals, lights, cam, shininess=shininess, view_idx=idx)
        if not self.specular:
            spe = torch.zeros_like(spe)
        shaded = points_rgb * (amb + dif) + spe                          # texture.py:120
        # same geometry (the clouds keep aliasing one tensor, so the renderer still takes the shared-cloud path), new
        # per-(view, point) features
        return pointclouds.update_features(list(shaded.split([int(n) for n in pointclouds.num_points_per_cloud()])))
