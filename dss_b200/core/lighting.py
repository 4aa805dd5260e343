"""Lights and the per-point lighting terms -- the torch twin of DSS/core/lighting.py (which derives from
pytorch3d.renderer.lighting, absent here) for the unfused route and as the autograd reference of the fused CUDA
shading (dss_render_args.shade, csrc/render.cu: shade_point).

Conventions of the reference: colours / directions / locations are (N, L, 3) or (1, L, 3) tensors (N views, L lights);
`diffuse` is Lambert's law with the renormalised normal and light direction summed over the lights
(lighting.py:10-69), `specular` the reflected ray against the view direction, zero where the light is behind the
surface (lighting.py:72-172).  Packed inputs (P,3) take per-point properties gathered by the caller.
"""
import torch
import torch.nn.functional as F

__all__ = ["diffuse", "specular", "DirectionalLights", "PointLights", "pack_lights"]


def _as_nl3(x, device=None):
    t = torch.as_tensor(x, dtype=torch.float32, device=device)
    if t.dim() == 1:
        t = t.view(1, 1, 3)
    elif t.dim() == 2:
        t = t.unsqueeze(0)
    if t.dim() != 3 or t.shape[-1] != 3:
        raise ValueError("expected (N,L,3), got %s" % (tuple(t.shape),))
    return t


def diffuse(normals, color, direction) -> torch.Tensor:
    """normals (..., 3); color, direction (..., L, 3) broadcastable against normals[..., None, :] -> (..., 3)
    (lighting.py:62-69)."""
    n = F.normalize(normals, p=2, dim=-1, eps=1e-6).unsqueeze(-2)
    d = F.normalize(direction, p=2, dim=-1, eps=1e-6)
    angle = F.relu((n * d).sum(-1))
    return (color * angle[..., None]).sum(-2)


def specular(points, normals, direction, color, camera_position, shininess) -> torch.Tensor:
    """points, normals, camera_position (..., 3); direction, color (..., L, 3) -> (..., 3)  (lighting.py:139-172)."""
    n = F.normalize(normals, p=2, dim=-1, eps=1e-6).unsqueeze(-2)
    d = F.normalize(direction, p=2, dim=-1, eps=1e-6)
    cos_angle = (n * d).sum(-1)
    mask = (cos_angle > 0).to(points.dtype)
    view_direction = F.normalize(camera_position - points, p=2, dim=-1, eps=1e-6).unsqueeze(-2)
    reflect_direction = -d + 2 * (cos_angle[..., None] * n)
    alpha = F.relu((view_direction * reflect_direction).sum(-1)) * mask
    return (color * torch.pow(alpha, shininess)[..., None]).sum(-2)


class _Lights(torch.nn.Module):
    """ambient / diffuse / specular colours (N|1, L, 3) + a direction or a location per light."""
    kind = None

    def __init__(self, ambient_color=(((0.5, 0.5, 0.5),),), diffuse_color=(((0.3, 0.3, 0.3),),),
                 specular_color=(((0.2, 0.2, 0.2),),), vec=(((0, 1, 0),),), device="cpu"):
        super().__init__()
        self.register_buffer("ambient_color", _as_nl3(ambient_color, device))
        self.register_buffer("diffuse_color", _as_nl3(diffuse_color, device))
        self.register_buffer("specular_color", _as_nl3(specular_color, device))
        self.register_buffer("_vec", _as_nl3(vec, device))
        for prop in ("diffuse_color", "specular_color", "_vec"):
            if getattr(self, prop).dim() != 3:
                raise ValueError("%s must be an (N,L,3) tensor" % prop)
        if not (self.diffuse_color.shape[1] == self.specular_color.shape[1] == self._vec.shape[1]):
            raise ValueError("diffuse_color, specular_color and the direction/location need the same number of lights")

    @property
    def device(self):
        return self._vec.device

    def _per_point(self, t, view_idx):
        """(N|1, L, 3) -> (P, L, 3) for packed points of views `view_idx` (P,) or broadcast (1, L, 3)."""
        return t if (view_idx is None or t.shape[0] == 1) else t[view_idx]


class DirectionalLights(_Lights):
    """DSS/core/lighting.py:175-225."""
    kind = 0

    def __init__(self, ambient_color=(((0.5, 0.5, 0.5),),), diffuse_color=(((0.3, 0.3, 0.3),),),
                 specular_color=(((0.2, 0.2, 0.2),),), direction=(((0, 1, 0),),), device="cpu", **kwargs):
        super().__init__(ambient_color, diffuse_color, specular_color, direction, device)

    @property
    def direction(self):
        return self._vec

    def diffuse(self, normals, points=None, view_idx=None):
        return diffuse(normals, self._per_point(self.diffuse_color, view_idx), self._per_point(self._vec, view_idx))

    def specular(self, normals, points, camera_position, shininess, view_idx=None):
        return specular(points, normals, self._per_point(self._vec, view_idx),
                        self._per_point(self.specular_color, view_idx), camera_position, shininess)


class PointLights(_Lights):
    """DSS/core/lighting.py:228-300."""
    kind = 1

    def __init__(self, ambient_color=(((0.5, 0.5, 0.5),),), diffuse_color=(((0.3, 0.3, 0.3),),),
                 specular_color=(((0.2, 0.2, 0.2),),), location=(((0, 1, 0),),), device="cpu", **kwargs):
        super().__init__(ambient_color, diffuse_color, specular_color, location, device)

    @property
    def location(self):
        return self._vec

    def diffuse(self, normals, points, view_idx=None):
        direction = self._per_point(self._vec, view_idx) - points.unsqueeze(-2)
        return diffuse(normals, self._per_point(self.diffuse_color, view_idx), direction)

    def specular(self, normals, points, camera_position, shininess, view_idx=None):
        direction = self._per_point(self._vec, view_idx) - points.unsqueeze(-2)
        return specular(points, normals, direction, self._per_point(self.specular_color, view_idx), camera_position,
                        shininess)


def pack_lights(lights):
    """-> (rows (L,9) {direction|location, diffuse rgb, specular rgb}, ambient (3,), light_type) for the fused path.
    The fused path shares one set of lights between the views ("Currently only supports the same lights for all
    batches", config.py:105)."""
    for t in (lights.diffuse_color, lights.specular_color, lights._vec, lights.ambient_color):
        if t.shape[0] != 1:
            raise ValueError("the fused shading takes one set of lights for all views ((1,L,3) tensors)")
    rows = torch.cat([lights._vec[0], lights.diffuse_color[0], lights.specular_color[0]], dim=1).contiguous().float()
    ambient = lights.ambient_color[0].sum(0).contiguous().float()          # texture.py:52-55: ambient colours are summed
    return rows, ambient, int(lights.kind)
