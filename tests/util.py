"""Seeded synthetic scenes shared by tests, smoke() and bench.py."""
import numpy as np
import torch

from dss_b200.core.camera import FoVPerspectiveCameras, look_at_view_transform, camera_matrices


def sphere_cloud(P0, seed=0, radius=0.5):
    """P0 points uniform on a sphere of radius 0.5 (the training init, config.py:177-182),
    outward normals, random colours."""
    g = torch.Generator().manual_seed(1234 + seed)
    p = torch.nn.functional.normalize(torch.randn(P0, 3, generator=g), dim=1)
    colours = torch.rand(P0, 3, generator=g)
    return (radius * p).contiguous(), p.clone().contiguous(), colours.contiguous()


def random_cameras(N, seed=0, dist=(1.2, 2.2), znear=0.1, zfar=100.0):
    """Look-at cameras as DSS/core/camera.py:42-51 samples them (fov 60, znear 0.1, zfar 100)."""
    g = torch.Generator().manual_seed(4321 + seed)
    d = torch.rand(N, generator=g) * (dist[1] - dist[0]) + dist[0]
    azim = torch.rand(N, generator=g) * 360 - 180
    elev = torch.rand(N, generator=g) * 180 - 90
    at = torch.rand(N, 3, generator=g) * 0.1 - 0.05
    R, T = look_at_view_transform(d, elev, azim, at=at)
    return FoVPerspectiveCameras(znear=znear, zfar=zfar, R=R, T=T)


def scene(P0, N, seed=0):
    pts, nrm, col = sphere_cloud(P0, seed)
    cams = random_cameras(N, seed)
    proj, view = camera_matrices(cams)
    return pts, nrm, col, proj, view, cams


def packed_offsets(N, P0):
    first = torch.arange(N, dtype=torch.int64) * P0
    num = torch.full((N,), P0, dtype=torch.int64)
    return first, num


def random_screen_splats(P, N, S, seed=0, rad_px=(0.8, 4.0), z=(0.5, 3.0), behind_frac=0.05, ragged=True):
    """Random packed screen-space splats for the _C-level operators: points (P,3), ellipse (P,3),
    cutoff (P,), radii (P,2), first_idx, num_points.  Radii are the exact bbox of the conic at cutoff."""
    rng = np.random.default_rng(100 + seed)
    pts = np.concatenate([rng.uniform(-1.05, 1.05, (P, 2)), rng.uniform(z[0], z[1], (P, 1))], 1).astype(np.float32)
    behind = rng.random(P) < behind_frac
    pts[behind, 2] = -pts[behind, 2]
    pix = 2.0 / S
    # random SPD 2x2 variance -> conic = inverse
    s1 = rng.uniform(rad_px[0] * pix, rad_px[1] * pix, P)
    s2 = rng.uniform(rad_px[0] * pix, rad_px[1] * pix, P)
    th = rng.uniform(0, np.pi, P)
    c, s = np.cos(th), np.sin(th)
    G00 = (c * s1) ** 2 + (s * s2) ** 2
    G11 = (s * s1) ** 2 + (c * s2) ** 2
    G01 = c * s * (s1 ** 2 - s2 ** 2)
    det = G00 * G11 - G01 ** 2
    cutoff = rng.uniform(0.5, 2.0, P).astype(np.float32)
    ell = np.stack([G11 / det, -2 * G01 / det, G00 / det], 1).astype(np.float32)
    rad = np.stack([np.sqrt(cutoff * G00), np.sqrt(cutoff * G11)], 1).astype(np.float32)
    if ragged and N > 1:
        cuts = np.sort(rng.choice(np.arange(1, P), N - 1, replace=False))
        bounds = np.concatenate([[0], cuts, [P]])
    else:
        bounds = np.linspace(0, P, N + 1).astype(np.int64)
    first = bounds[:-1].astype(np.int64)
    num = (bounds[1:] - bounds[:-1]).astype(np.int64)
    return pts, ell, cutoff, rad, first, num
