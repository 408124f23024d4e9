"""Scenes with closed-form radiance, shared by tests/test_oracle_transport.py (oracle) and tools/gpu_analytic_check.py (CUDA path)."""
import numpy as np


def corner_form_factor(a, b, c):
    """Differential element to a parallel a x b rectangle whose corner lies on the element's normal, distance c."""
    X, Y = a / c, b / c
    return (X / np.sqrt(1 + X * X) * np.arctan(Y / np.sqrt(1 + X * X)) + Y / np.sqrt(1 + Y * Y) * np.arctan(X / np.sqrt(1 + Y * Y))) / (2 * np.pi)


def form_factor_scene(pkg, light="rectangle"):
    """A diffuse floor point under a rectangular Lambertian light, seen through a very narrow pixel.
    Returns (builder, centre pixel index, expected radiance rho * L * F for maxDepth = 2)."""
    S = pkg.scenes
    rho, L, h, hx, hz = 0.6, (5.0, 3.0, 1.0), 1.5, 0.8, 0.5
    sb = S.SceneBuilder(9, 9, spp=1)
    X = (1, 0, 0)
    sb.rectangle([S.scale(50, 50, 1), S.rotate(X, -90.0)], bsdf=sb.diffuse((rho, rho, rho)))                      # floor, y = 0, +y
    if light == "rectangle":   # Rectangle::samplePosition (rectangle.cpp:210-216)
        sb.rectangle([S.scale(hx, hz, 1), S.rotate(X, 90.0), S.translate(0.3, h, -0.2)], bsdf=-1, radiance=L)     # light, -y
    else:                      # the same light as two triangles: TriMesh::samplePosition over the area cdf (trimesh.cpp:412-423)
        P = [[0.3 - hx, h, -0.2 - hz], [0.3 + hx, h, -0.2 - hz], [0.3 + hx, h, -0.2 + hz], [0.3 - hx, h, -0.2 + hz]]
        sb.trimesh(P=P, T=[[0, 1, 2], [0, 2, 3]], bsdf=-1, radiance=L)                                            # winding: normal -y
    target = np.array([0.5, 0.0, 0.1])
    sb.set_camera((3.0, 1.0, 2.5), tuple(target), (0, 1, 0), 0.05)
    dx0, dx1 = target[0] - (0.3 - hx), (0.3 + hx) - target[0]
    dz0, dz1 = target[2] - (-0.2 - hz), (-0.2 + hz) - target[2]
    assert min(dx0, dx1, dz0, dz1) > 0
    F = sum(corner_form_factor(a, b, h) for a in (dx0, dx1) for b in (dz0, dz1))
    return sb, 4 * 9 + 4, rho * np.array(L) * F


def furnace_scene(pkg, res=16, glass=False, medium=None):
    """The furnace: the six faces of [-1, 1]^3 point inwards, emit L = 1 and reflect rho = 0.5, so the radiance is
    L / (1 - rho) = 2 everywhere. Optional content that must not change that: a glass cube; a purely scattering
    heterogeneous medium medium = (phase, g, method) inside an index-matched cube. Returns (builder, expected value)."""
    S = pkg.scenes
    rho, L = 0.5, 1.0
    sb = S.SceneBuilder(res, res, spp=4)
    mat = sb.diffuse((rho, rho, rho))
    X, Y = (1, 0, 0), (0, 1, 0)
    for ops in ([S.translate(0, 0, -1)],                                   # z = -1, normal +z
                [S.rotate(Y, 180.0), S.translate(0, 0, 1)],                # z = +1, normal -z
                [S.rotate(Y, 90.0), S.translate(-1, 0, 0)],                # x = -1, normal +x
                [S.rotate(Y, -90.0), S.translate(1, 0, 0)],                # x = +1, normal -x
                [S.rotate(X, -90.0), S.translate(0, -1, 0)],               # y = -1, normal +y
                [S.rotate(X, 90.0), S.translate(0, 1, 0)]):                # y = +1, normal -y
        sb.rectangle(ops, bsdf=mat, radiance=(L, L, L))
    if glass:
        sb.cube([S.scale(0.4, 0.3, 0.35), S.rotate(Y, 25.0), S.translate(0.1, -0.1, 0.0)], bsdf=sb.dielectric())
    if medium is not None:
        phase, g, method = medium
        n = 12
        t = (np.arange(n) + 0.5) / n - 0.5
        Z3, Y3, X3 = np.meshgrid(t, t, t, indexing="ij")
        dens = np.clip(1.0 - 3.0 * (X3 * X3 + Y3 * Y3 + Z3 * Z3), 0.05, 1.0).astype(np.float32)   # densities must stay <= 1
        med = sb.medium(dens, (-0.5, -0.5, -0.5), (0.5, 0.5, 0.5), scale_=4.0, albedo=(1.0, 1.0, 1.0), phase=phase, g=g, method=method)
        sb.cube([S.scale(0.5, 0.5, 0.5)], bsdf=-1, interior=med)
    if glass or medium is not None:
        sb.set_camera((0.8, 0.7, 0.9), (0.0, 0.0, 0.0), (0, 1, 0), 50.0)   # outside the cube, looking at it
    else:
        sb.set_camera((0.1, -0.2, 0.3), (0.9, 0.4, -1.0), (0, 1, 0), 70.0)
    return sb, L / (1 - rho)
