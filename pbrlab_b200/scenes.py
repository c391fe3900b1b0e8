"""Input files for the configurations BASELINE.json names: the bundled Cornell scene and the procedurally generated
ones (downloaded hair models / large meshes are unavailable offline).  Everything is written in the
file formats the reference's CLI reads (.obj + .mtl, CyHair .hair), so the same files feed the compiled reference
(oracle/_ref) and this backend.  Plain numpy; nothing here is on the hot path."""
import contextlib
import fcntl
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


@contextlib.contextmanager
def _generating(name):
    """One process at a time checks for / writes the files of one generated scene: bench.py under torchrun starts
    N ranks that all ask for the same files at the same moment."""
    os.makedirs(CACHE, exist_ok=True)
    with open(os.path.join(CACHE, "." + name + ".lock"), "w") as f:
        fcntl.flock(f, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(f, fcntl.LOCK_UN)


def cornell():
    """data/cornellbox_suzanne_lucy.obj of the reference (kept gzip-compressed in the repo), unpacked once."""
    obj = _cache("cornellbox_suzanne_lucy.obj")
    mtl = _cache("cornellbox_suzanne_lucy.mtl")
    with _generating("cornell"):
        if not os.path.exists(mtl):
            tmp = mtl + ".tmp%d" % os.getpid()
            shutil.copyfile(os.path.join(DATA, "cornellbox_suzanne_lucy.mtl"), tmp)
            os.replace(tmp, mtl)
        if not os.path.exists(obj) or os.path.getsize(obj) == 0:
            tmp = obj + ".tmp%d" % os.getpid()
            with gzip.open(os.path.join(DATA, "cornellbox_suzanne_lucy.obj.gz"), "rb") as src, open(tmp, "wb") as dst:
                shutil.copyfileobj(src, dst, 1 << 22)
            os.replace(tmp, obj)
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
    with _generating(os.path.basename(path)):
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
    with _generating("light_stage"):
        if not os.path.exists(path + ".done"):
            write_light_stage_obj(path)
            open(path + ".done", "w").close()
    return path


def _append_rows(f, path, kind, data):
    """rows of `v` / `vn` / `f` lines: formatted by the host library's threads when it is built (same text as
    numpy.savetxt, 10x faster), else by savetxt"""
    fmt = ["v %.6f %.6f %.6f", "vn %.5f %.5f %.5f", "f %d//%d %d//%d %d//%d"][kind]
    lib = None
    try:
        import ctypes as C
        so = os.path.join(_ROOT, "lib", "libpbrlab_host.so")
        if os.path.exists(so):
            C.CDLL(os.path.join(_ROOT, "lib", "libpbrgpu.so"), mode=C.RTLD_GLOBAL)
            lib = C.CDLL(so)
            lib.pbrhost_append_rows.restype = C.c_int
    except OSError:
        lib = None
    if lib is None:
        np.savetxt(f, data, fmt=fmt)
        return
    f.flush()
    if kind == 2:
        a = np.ascontiguousarray(data, np.int64)
        ok = lib.pbrhost_append_rows(path.encode(), 2, None, a.ctypes.data_as(C.c_void_p), C.c_uint64(len(a)))
    else:
        a = np.ascontiguousarray(data, np.float64)
        ok = lib.pbrhost_append_rows(path.encode(), kind, a.ctypes.data_as(C.c_void_p), None, C.c_uint64(len(a)))
    if not ok:
        raise RuntimeError("pbrhost_append_rows failed for " + path)
    f.seek(0, os.SEEK_END)


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
            _append_rows(f, path, 0, P.reshape(-1, 3))
            _append_rows(f, path, 1, N.reshape(-1, 3))
            f.write("usemtl %s\n" % ("Sss" if b % 2 else "Ggx"))
            i = np.arange(nv)[:, None] * nu
            j = np.arange(nu)[None, :]
            j1 = (j + 1) % nu
            a = (i + j).ravel(); bq = (i + j1).ravel(); cq = (i + nu + j1).ravel(); dq = (i + nu + j).ravel()
            tri = np.concatenate([np.stack([a, bq, cq], 1), np.stack([a, cq, dq], 1)], 0)   # (v1-v0)x(v2-v0) points outwards, like vn
            # drop the zero-area triangles at the two poles
            row = tri[:, 0] // nu
            keep = ~(((row == 0) & (np.arange(len(tri)) < len(a))) | ((row == nv - 1) & (np.arange(len(tri)) >= len(a))))
            tri = tri[keep]
            v = tri + base
            n = tri + nbase
            _append_rows(f, path, 2, np.stack([v[:, 0], n[:, 0], v[:, 1], n[:, 1], v[:, 2], n[:, 2]], 1))
            base += (nv + 1) * nu
            nbase += (nv + 1) * nu
    return path


def displaced(n_tris=20_000_000, seed=7):
    path = _cache("displaced_v2_%d_%d.obj" % (n_tris, seed))
    done = path + ".done"
    with _generating(os.path.basename(path)):
        if not os.path.exists(done):
            write_displaced_obj(path, n_tris, seed)
            open(done, "w").close()
    return path


# ---------------------------------------------------------------- textured scene
def _png_chunk(tag, data):
    import zlib
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def write_png(path, img, palette=None, trns=None):
    """Minimal PNG encoder for the test textures: img (h, w[, c]) uint8 or uint16; rows cycle through the five PNG
    filter types so that a decoder's un-filtering is exercised.  palette (n,3) uint8 -> colour type 3."""
    import zlib
    img = np.asarray(img)
    if img.ndim == 2:
        img = img[..., None]
    h, w, c = img.shape
    depth = 16 if img.dtype == np.uint16 else 8
    ctype = 3 if palette is not None else {1: 0, 2: 4, 3: 2, 4: 6}[c]
    raw = img.astype(">u2").tobytes() if depth == 16 else img.astype(np.uint8).tobytes()
    bpp = c * depth // 8
    stride = w * bpp
    rows = [np.frombuffer(raw[y * stride:(y + 1) * stride], np.uint8).astype(np.int32) for y in range(h)]
    out = bytearray()
    prev = np.zeros(stride, np.int32)
    for y, cur in enumerate(rows):
        ft = y % 5
        a = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])
        cc = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])
        if ft == 0: f = cur
        elif ft == 1: f = cur - a
        elif ft == 2: f = cur - prev
        elif ft == 3: f = cur - ((a + prev) >> 1)
        else:
            p = a + prev - cc
            pa, pb, pc = np.abs(p - a), np.abs(p - prev), np.abs(p - cc)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, prev, cc))
            f = cur - pred
        out.append(ft)
        out += (f & 0xFF).astype(np.uint8).tobytes()
        prev = cur
    data = b"\x89PNG\r\n\x1a\n" + _png_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0))
    if palette is not None:
        data += _png_chunk(b"PLTE", np.asarray(palette, np.uint8).tobytes())
    if trns is not None:
        data += _png_chunk(b"tRNS", bytes(trns))
    data += _png_chunk(b"IDAT", zlib.compress(bytes(out), 6)) + _png_chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(data)
    return path


def write_textured_obj(path, seed=5):
    """A small lit box whose materials use map_base_color / map_subsurface_color (the two texture slots the
    reference's Principled path samples, cycles-principled-shader.cc:281-301): floor with an sRGB RGB8 checker and
    texcoords running outside [0,1] (clamp addressing), back wall with an RGBA8 gradient declared `-colorspace linear`,
    a pedestal with a palette PNG on a glossy material, a box with a binary PPM, and a subsurface sphere whose
    subsurface colour comes from a 1-channel PNG (channels a texture lacks read as 0)."""
    rng = np.random.default_rng(seed)
    d = os.path.dirname(path)
    checker = np.zeros((48, 64, 3), np.uint8)
    cols = rng.integers(40, 255, size=(6, 8, 3))
    for j in range(6):
        for i in range(8):
            checker[j * 8:(j + 1) * 8, i * 8:(i + 1) * 8] = cols[j, i]
    checker = (checker * (0.6 + 0.4 * np.linspace(0, 1, 64)[None, :, None])).astype(np.uint8)
    write_png(os.path.join(d, "tex_checker.png"), checker)
    gy, gx = np.mgrid[0:32, 0:32]
    grad = np.stack([gx * 8, gy * 8, 255 - gx * 4 - gy * 4, 128 + gx * 0], -1).clip(0, 255).astype(np.uint8)
    write_png(os.path.join(d, "tex grad rgba.png"), grad)          # file name with spaces (tinyobj reads to line end)
    pal = rng.integers(0, 255, size=(16, 3)).astype(np.uint8)
    write_png(os.path.join(d, "tex_palette.png"), rng.integers(0, 16, size=(16, 16)).astype(np.uint8), palette=pal)
    skin = (128 + 100 * np.sin(gx / 5.0) * np.cos(gy / 7.0)).astype(np.uint8)
    write_png(os.path.join(d, "tex_skin_gray.png"), skin)
    wood = (rng.integers(60, 200, size=(20, 24, 1)) * np.array([1.0, 0.7, 0.4])).astype(np.uint8)
    with open(os.path.join(d, "tex_wood.ppm"), "wb") as f:
        f.write(b"P6\n# test texture\n24 20\n255\n" + wood.tobytes())
    deep = (rng.integers(0, 65535, size=(8, 8, 3))).astype(np.uint16)
    write_png(os.path.join(d, "tex_deep16.png"), deep)
    mtl = os.path.splitext(path)[0] + ".mtl"
    with open(mtl, "w") as f:
        f.write("newmtl Floor\nbase_color 0.5 0.5 0.5\nmap_base_color tex_checker.png\nspecular 0.0\n\n"
                "newmtl Back\nbase_color 0.5 0.5 0.5\nmap_base_color -colorspace linear -clamp on tex grad rgba.png\nspecular 0.0\n\n"
                "newmtl Glossy\nbase_color 0.8 0.8 0.8\nmap_base_color -bm 1.0 tex_palette.png\nspecular 1.0\nroughness 0.25\n\n"
                "newmtl Wood\nbase_color 0.2 0.2 0.2\nmap_base_color tex_wood.ppm\nspecular 0.3\nroughness 0.5\n\n"
                "newmtl Skin\nbase_color 0.8 0.6 0.5\nsubsurface 1.0\nsubsurface_radius 0.3 0.2 0.1\n"
                "subsurface_color 0.9 0.6 0.5\nmap_subsurface_color tex_skin_gray.png\nspecular 0.5\nroughness 0.3\n\n"
                "newmtl Deep\nbase_color 0.3 0.3 0.3\nmap_base_color tex_deep16.png\nspecular 0.0\n\n"
                "newmtl Missing\nbase_color 0.6 0.2 0.2\nmap_base_color no_such_file.png\nspecular 0.0\n\n"
                "newmtl Light\nbase_color 0.0 0.0 0.0\nspecular 0.0\n")
    with open(path, "w") as f:
        f.write("mtllib %s\n" % os.path.basename(mtl))
        nv = [0]; nt = [0]; nn = [0]

        def quad(name, mat, P, T=None):
            f.write("o %s\n" % name)
            for p in P:
                f.write("v %f %f %f\n" % tuple(p))
            if T is not None:
                for t in T:
                    f.write("vt %f %f\n" % tuple(t))
            f.write("usemtl %s\n" % mat)
            b, tb = nv[0] + 1, nt[0] + 1
            if T is not None:
                f.write("f %d/%d %d/%d %d/%d\nf %d/%d %d/%d %d/%d\n" % (b, tb, b + 1, tb + 1, b + 2, tb + 2, b, tb, b + 2, tb + 2, b + 3, tb + 3))
                nt[0] += 4
            else:
                f.write("f %d %d %d\nf %d %d %d\n" % (b, b + 1, b + 2, b, b + 2, b + 3))
            nv[0] += 4

        S = 4.0
        quad("floor", "Floor", [(-S, 0, S), (S, 0, S), (S, 0, -S), (-S, 0, -S)], [(-0.2, -0.1), (1.3, -0.1), (1.3, 1.2), (-0.2, 1.2)])
        quad("back", "Back", [(-S, 0, -S), (S, 0, -S), (S, 2 * S, -S), (-S, 2 * S, -S)], [(0, 0), (1, 0), (1, 1), (0, 1)])
        quad("left", "Deep", [(-S, 0, S), (-S, 0, -S), (-S, 2 * S, -S), (-S, 2 * S, S)], [(0, 0), (1, 0), (1, 1), (0, 1)])
        quad("right", "Missing", [(S, 0, -S), (S, 0, S), (S, 2 * S, S), (S, 2 * S, -S)], [(0, 0), (1, 0), (1, 1), (0, 1)])
        quad("pedestal", "Glossy", [(-3, 1.0, 1), (-1, 1.0, 1), (-1, 1.0, -1), (-3, 1.0, -1)], [(0, 0), (1, 0), (1, 1), (0, 1)])
        quad("plank", "Wood", [(1, 0.6, 2), (3, 0.9, 2), (3, 0.9, 0.5), (1, 0.6, 0.5)])      # no texcoords: barycentrics
        quad("light_top", "Light", [(-1.5, 2 * S - 0.01, -1.5), (1.5, 2 * S - 0.01, -1.5), (1.5, 2 * S - 0.01, 1.5), (-1.5, 2 * S - 0.01, 1.5)])
        # uv sphere with normals and texcoords
        nu, nvv = 48, 24
        th = np.linspace(0, np.pi, nvv + 1)[:, None]; ph = np.linspace(0, 2 * np.pi, nu + 1)[None, :]
        dirs = np.stack([np.sin(th) * np.cos(ph), np.cos(th) * np.ones_like(ph), np.sin(th) * np.sin(ph)], -1)
        P = np.array([0.8, 3.0, -0.5]) + 1.3 * dirs
        f.write("o ball\n")
        np.savetxt(f, P.reshape(-1, 3), fmt="v %.6f %.6f %.6f")
        np.savetxt(f, dirs.reshape(-1, 3), fmt="vn %.6f %.6f %.6f")
        uv = np.stack([np.broadcast_to(ph / (2 * np.pi), dirs.shape[:2]), np.broadcast_to(1 - th / np.pi, dirs.shape[:2])], -1)
        np.savetxt(f, uv.reshape(-1, 2), fmt="vt %.6f %.6f")
        f.write("usemtl Skin\n")
        b, tb = nv[0] + 1, nt[0] + 1
        for j in range(nvv):
            for i in range(nu):
                a0 = j * (nu + 1) + i; a1 = a0 + 1; a2 = a0 + nu + 1; a3 = a2 + 1
                if j > 0:
                    f.write("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % (b + a0, tb + a0, a0 + 1, b + a1, tb + a1, a1 + 1, b + a3, tb + a3, a3 + 1))
                if j < nvv - 1:
                    f.write("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % (b + a0, tb + a0, a0 + 1, b + a3, tb + a3, a3 + 1, b + a2, tb + a2, a2 + 1))
    return path


def textured():
    path = _cache("textured_box.obj")
    with _generating("textured_box"):
        write_textured_obj(path)
    return path
