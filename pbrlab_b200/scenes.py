"""Input files for the configurations BASELINE.json names: the bundled Cornell scene and the procedurally generated
ones (downloaded hair models / large meshes are unavailable offline).  Everything is written in the
file formats the reference's CLI reads (.obj + .mtl, CyHair .hair), so the same files feed the compiled reference
(oracle/_ref) and this backend.  Plain numpy; nothing here is on the hot path."""
import gzip
import os
import shutil
import struct

import numpy as np

_ROOT = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(_ROOT)
DATA = os.path.join(REPO, "data")
CACHE = os.path.join(DATA, "_generated")


def _cache(name):
    os.makedirs(CACHE, exist_ok=True)
    return os.path.join(CACHE, name)


def cornell():
    """data/cornellbox_suzanne_lucy.obj of the reference (kept gzip-compressed in the repo), unpacked once."""
    obj = _cache("cornellbox_suzanne_lucy.obj")
    mtl = _cache("cornellbox_suzanne_lucy.mtl")
    if not os.path.exists(obj) or os.path.getsize(obj) == 0:
        tmp = obj + ".tmp%d" % os.getpid()
        with gzip.open(os.path.join(DATA, "cornellbox_suzanne_lucy.obj.gz"), "rb") as src, open(tmp, "wb") as dst:
            shutil.copyfileobj(src, dst, 1 << 22)
        os.replace(tmp, obj)
    if not os.path.exists(mtl):
        shutil.copyfile(os.path.join(DATA, "cornellbox_suzanne_lucy.mtl"), mtl)
    return obj


def write_cyhair(path, n_strands=50000, n_points=21, center=(-2.5, 6.0, 0.0), radius=1.2, length=2.5,
                 thickness=0.008, seed=1234, per_point_thickness=False):
    """CyHair file: 128-byte header ("HAIR", strands, points, flags, default segments/thickness/transparency/colour),
    then the point array.  Roots on a sphere, strands grow outwards and droop quadratically under "gravity"
    (SURVEY §8(d) C3).  Every strand has n_points >= 3 vertices (the reference rejects the file otherwise)."""
    assert n_points >= 3
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(n_strands, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v[:, 1] = np.abs(v[:, 1]) * 0.8 + 0.2 * v[:, 1]          # bias roots to the upper hemisphere
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    root = np.asarray(center, np.float64) + radius * v
    s = np.linspace(0.0, 1.0, n_points)[None, :, None]
    jitter = 1.0 + 0.15 * rng.standard_normal((n_strands, 1, 1))
    out = v[:, None, :] * (length * jitter) * s
    droop = np.zeros_like(out)
    droop[..., 1] = -1.6 * length * (s[..., 0] ** 2) * jitter[..., 0]
    wave = 0.03 * length * np.sin(8.0 * s + rng.uniform(0, 6.28, (n_strands, 1, 1))) * rng.normal(size=(n_strands, 1, 3))
    pts = (root[:, None, :] + out + droop + wave).astype(np.float32)
    flags = 0x2 | (0x4 if per_point_thickness else 0)
    header = struct.pack("<4sIIIIff3f88s", b"HAIR", n_strands, n_strands * n_points, flags, n_points - 1,
                         float(thickness), 1.0, 0.5, 0.5, 0.5, b"pbrlab_b200 synthetic hair")
    assert len(header) == 128
    with open(path, "wb") as f:
        f.write(header)
        f.write(pts.tobytes())
        if per_point_thickness:
            taper = (thickness * (1.0 - 0.7 * s[..., 0]) * np.ones((n_strands, 1))).astype(np.float32)
            f.write(taper.tobytes())
    return path


def cyhair(n_strands=50000, n_points=21, **kw):
    path = _cache("hair_%d_%d_%s.hair" % (n_strands, n_points, "_".join("%s%s" % (k, v) for k, v in sorted(kw.items()))))
    if not os.path.exists(path):
        write_cyhair(path + ".tmp", n_strands, n_points, **kw)
        os.replace(path + ".tmp", path)
    return path


def write_light_stage_obj(path, size=6.0, light_y=9.0):
    """A floor, a back wall and a `light...` quad: the smallest OBJ that lights a hair-only scene (lights come only
    from shapes whose name starts with "light", reference pc/pc-common.cc:172-186)."""
    mtl = os.path.splitext(path)[0] + ".mtl"
    with open(mtl, "w") as f:
        f.write("newmtl Floor\nbase_color 0.7 0.7 0.7\nspecular 0.0\n\nnewmtl Light\nbase_color 0.0 0.0 0.0\nspecular 0.0\n")
    h = size
    with open(path, "w") as f:
        f.write("mtllib %s\n" % os.path.basename(mtl))
        f.write("o floor\n")
        for p in [(-h - 2.5, 0, -h), (h - 2.5, 0, -h), (h - 2.5, 0, h), (-h - 2.5, 0, h)]:
            f.write("v %f %f %f\n" % p)
        f.write("usemtl Floor\nf 1 3 2\nf 1 4 3\n")
        f.write("o back\n")
        for p in [(-h - 2.5, 0, -h), (h - 2.5, 0, -h), (h - 2.5, 2 * h, -h), (-h - 2.5, 2 * h, -h)]:
            f.write("v %f %f %f\n" % p)
        f.write("usemtl Floor\nf 5 6 7\nf 5 7 8\n")
        f.write("o light_quad\n")
        for p in [(-4.5, light_y, -2.0), (-0.5, light_y, -2.0), (-0.5, light_y, 2.0), (-4.5, light_y, 2.0)]:
            f.write("v %f %f %f\n" % p)
        f.write("usemtl Light\nf 9 10 11\nf 9 11 12\n")
    return path


def light_stage():
    path = _cache("light_stage.obj")
    if not os.path.exists(path):
        write_light_stage_obj(path)
    return path


def write_displaced_obj(path, n_tris=20_000_000, seed=7, blobs=8):
    """C5: closed, displaced, tessellated surfaces (uv-spheres with smooth noise displacement and per-vertex normals)
    inside an open-front box with a `light...` quad; two materials: GGX (specular 1 / roughness 0.2) and SSS
    (subsurface 1 / subsurface_radius).  Triangles only; total close to n_tris."""
    rng = np.random.default_rng(seed)
    mtl = os.path.splitext(path)[0] + ".mtl"
    with open(mtl, "w") as f:
        f.write("newmtl Wall\nbase_color 0.75 0.75 0.75\nspecular 0.0\n\n"
                "newmtl Light\nbase_color 0.0 0.0 0.0\nspecular 0.0\n\n"
                "newmtl Ggx\nbase_color 0.8 0.6 0.3\nspecular 1.0\nroughness 0.2\n\n"
                "newmtl Sss\nbase_color 0.9 0.7 0.7\nsubsurface 1.0\nsubsurface_radius 1.0 0.3 0.15\n"
                "subsurface_color 0.9 0.7 0.7\nspecular 1.0\nroughness 0.3\n")
    per = max(8, n_tris // blobs)
    # uv sphere with nu x nv quads -> 2*nu*nv triangles (poles are degenerate-free: rows 1..nv-1 + two fans)
    nv = max(4, int(np.sqrt(per / 4.0)))
    nu = 2 * nv
    with open(path, "w", buffering=1 << 24) as f:
        f.write("mtllib %s\n" % os.path.basename(mtl))
        base = 1
        S = 10.0
        f.write("o walls\n")
        box = [(-S, 0, -S), (S, 0, -S), (S, 0, S), (-S, 0, S), (-S, 2 * S, -S), (S, 2 * S, -S), (S, 2 * S, S), (-S, 2 * S, S)]
        for p in box:
            f.write("v %f %f %f\n" % p)
        f.write("usemtl Wall\n")
        for a, b, c, d in [(1, 4, 3, 2), (5, 6, 7, 8), (1, 2, 6, 5), (1, 5, 8, 4), (2, 3, 7, 6)]:
            f.write("f %d %d %d\nf %d %d %d\n" % (a, b, c, a, c, d))
        base += 8
        f.write("o light_panel\n")
        for p in [(-4, 2 * S - 0.05, -4), (4, 2 * S - 0.05, -4), (4, 2 * S - 0.05, 4), (-4, 2 * S - 0.05, 4)]:
            f.write("v %f %f %f\n" % p)
        f.write("usemtl Light\nf %d %d %d\nf %d %d %d\n" % (base, base + 1, base + 2, base, base + 2, base + 3))
        base += 4
        nbase = 1
        for b in range(blobs):
            c = np.array([rng.uniform(-6, 6), rng.uniform(3, 14), rng.uniform(-6, 4)])
            r0 = rng.uniform(1.6, 2.8)
            theta = np.linspace(0, np.pi, nv + 1)[:, None]
            phi = np.linspace(0, 2 * np.pi, nu, endpoint=False)[None, :]
            d = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta) * np.ones_like(phi), np.sin(theta) * np.sin(phi)], -1)
            k = rng.integers(2, 9, size=(4, 3)).astype(np.float64)
            ph = rng.uniform(0, 6.28, size=(4,))
            amp = np.array([0.12, 0.06, 0.03, 0.015]) * r0
            disp = sum(a * np.sin(d @ kk + p) for a, kk, p in zip(amp, k, ph))
            P = c + d * (r0 + disp)[..., None]
            # normals by central differences on the grid (periodic in phi), poles = radial
            dth = np.gradient(P, axis=0)
            dph = (np.roll(P, -1, axis=1) - np.roll(P, 1, axis=1)) * 0.5
            N = np.cross(dph, dth)
            nl = np.linalg.norm(N, axis=-1, keepdims=True)
            N = np.where(nl > 1e-12, N / np.maximum(nl, 1e-12), d)
            N[0] = d[0]; N[-1] = d[-1]
            f.write("o blob%d\n" % b)
            np.savetxt(f, P.reshape(-1, 3), fmt="v %.6f %.6f %.6f")
            np.savetxt(f, N.reshape(-1, 3), fmt="vn %.5f %.5f %.5f")
            f.write("usemtl %s\n" % ("Sss" if b % 2 else "Ggx"))
            i = np.arange(nv)[:, None] * nu
            j = np.arange(nu)[None, :]
            j1 = (j + 1) % nu
            a = (i + j).ravel(); bq = (i + j1).ravel(); cq = (i + nu + j1).ravel(); dq = (i + nu + j).ravel()
            tri = np.concatenate([np.stack([a, cq, bq], 1), np.stack([a, dq, cq], 1)], 0)
            # drop the zero-area triangles at the two poles
            row = tri[:, 0] // nu
            keep = ~(((row == 0) & (np.arange(len(tri)) < len(a))) | ((row == nv - 1) & (np.arange(len(tri)) >= len(a))))
            tri = tri[keep]
            v = tri + base
            n = tri + nbase
            np.savetxt(f, np.stack([v[:, 0], n[:, 0], v[:, 1], n[:, 1], v[:, 2], n[:, 2]], 1), fmt="f %d//%d %d//%d %d//%d")
            base += (nv + 1) * nu
            nbase += (nv + 1) * nu
    return path


def displaced(n_tris=20_000_000, seed=7):
    path = _cache("displaced_%d_%d.obj" % (n_tris, seed))
    done = path + ".done"
    if not os.path.exists(done):
        write_displaced_obj(path, n_tris, seed)
        open(done, "w").close()
    return path
