import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dss_b200.ops import SplatParams, make_shading, render_points
from dss_b200.core.lighting import DirectionalLights, PointLights
from dss_b200.core.texture import apply_lighting, camera_centres
from tests.util import scene
d = torch.device("cuda:0")
P0, N, S, shin = 20000, 3, 96, 24.0
pts, nrm, col, proj, view, cams = scene(P0, N, seed=12)
prm = SplatParams(image_size=S, znear=0.1, clip_pts_grad=-1.0)
h = torch.full((N,), 3e-4, device=d)
projd, viewd = proj.to(d), view.to(d)
def run(name, lights):
    sh = make_shading(lights, viewd, shininess=shin)
    o1 = render_points(pts.to(d), nrm.to(d), col.to(d), projd, viewd, h, prm, shading=sh)
    cam = camera_centres(viewd)
    outs = []
    for v in range(N):
        amb, dif, spe = apply_lighting(pts.to(d), nrm.to(d), lights, cam[v].expand(P0, 3), shininess=shin)
        outs.append(col.to(d) * (amb + dif) + spe)
    shaded = torch.cat(outs, 0)
    o2 = render_points(pts.to(d), nrm.to(d), shaded, projd, viewd, h, prm)
    diff = (o1.image - o2.image).abs()
    print(name, "idx equal", bool(torch.equal(o1.idx, o2.idx)), "max diff", float(diff.max()), "mse", float((diff ** 2).mean()),
          "image max", float(o2.image[..., :3].max()), "per-view max diff", [float(diff[v].max()) for v in range(N)])
z = (((0.0, 0.0, 0.0),),)
run("ambient only", DirectionalLights(ambient_color=(((0.3, 0.25, 0.2),),), diffuse_color=z, specular_color=z, direction=(((0.3, 1.0, 0.4),),), device=d))
run("ambient+diffuse", DirectionalLights(ambient_color=(((0.3, 0.25, 0.2),),), diffuse_color=(((0.6, 0.5, 0.4),),), specular_color=z, direction=(((0.3, 1.0, 0.4),),), device=d))
run("full sun", DirectionalLights(ambient_color=(((0.3, 0.25, 0.2),),), diffuse_color=(((0.6, 0.5, 0.4),),), specular_color=(((0.5, 0.5, 0.4),),), direction=(((0.3, 1.0, 0.4),),), device=d))
run("full point", PointLights(ambient_color=(((0.3, 0.25, 0.2),),), diffuse_color=(((0.6, 0.5, 0.4),),), specular_color=(((0.5, 0.5, 0.4),),), location=(((0.7, 1.5, 0.9),),), device=d))
run("two suns", DirectionalLights(ambient_color=(((0.3, 0.25, 0.2), (0.1, 0.1, 0.15)),), diffuse_color=(((0.6, 0.5, 0.4), (0.2, 0.3, 0.5)),), specular_color=(((0.5, 0.5, 0.4), (0.3, 0.2, 0.6)),), direction=(((0.3, 1.0, 0.4), (-0.8, 0.1, 0.5)),), device=d))
