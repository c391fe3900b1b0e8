"""ctypes binding of oracle/_ref/libpbrlab_ref.so — the UNMODIFIED reference (pbrlab + Embree) compiled by
oracle/Makefile.  TEST INFRASTRUCTURE ONLY: imported by tests/, tests/golden/make_golden.py and bench.py's
cpu_baseline / --impl reference legs; never by the product package."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "..", "oracle", "_ref")


def ref_path(count=False):
    return os.path.join(_REF_DIR, "libpbrlab_ref_count.so" if count else "libpbrlab_ref.so")


def available(count=False):
    return os.path.exists(ref_path(count))


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


class RefLib:
    def __init__(self, count=False):
        self.lib = C.CDLL(ref_path(count))
        L = self.lib
        L.ref_scene_create.restype = C.c_void_p
        L.ref_scene_create.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        L.ref_obj_load.restype = C.c_void_p
        L.ref_obj_load.argtypes = [C.c_char_p]
        L.ref_render.restype = C.c_double
        L.ref_obj_shape_faces.restype = C.c_uint32
        L.ref_obj_num_vertices.restype = C.c_uint32
        self.count = count

    # ---- scene
    def scene(self, files):
        arr = (C.c_char_p * len(files))(*[f.encode() for f in files])
        h = self.lib.ref_scene_create(len(files), arr)
        if not h:
            raise RuntimeError("reference CreateScene failed for %r" % (files,))
        return RefScene(self, h)

    def num_threads(self):
        return int(self.lib.ref_num_threads())

    # ---- closures / math KATs
    def rng_draws(self, initstate, initseq, n):
        out = np.empty(n, np.float32)
        self.lib.ref_rng_draws(C.c_uint64(initstate), C.c_uint64(initseq), C.c_uint64(n), _fp(out))
        return out

    def fastmath(self, op, x, y=None):
        x = np.ascontiguousarray(x, np.float32)
        y = np.ascontiguousarray(x if y is None else y, np.float32)
        out = np.empty_like(x)
        self.lib.ref_fastmath(C.c_int(op), _fp(x), _fp(y), C.c_uint64(x.size), _fp(out))
        return out

    def _map(self, fn, n, ncol, *arrays):
        out = np.empty((n, ncol), np.float32)
        fn(*arrays, C.c_uint64(n), _fp(out))
        return out

    def cosine_hemisphere(self, u):
        u = np.ascontiguousarray(u, np.float32)
        return self._map(self.lib.ref_cosine_hemisphere, len(u), 3, _fp(u))

    def uniform_sphere(self, u):
        u = np.ascontiguousarray(u, np.float32)
        return self._map(self.lib.ref_uniform_sphere, len(u), 3, _fp(u))

    def power_heuristic(self, a, b):
        a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
        return self._map(self.lib.ref_power_heuristic, len(a), 1, _fp(a), _fp(b))[:, 0]

    def fresnel_dielectric_cos(self, c, eta):
        c = np.ascontiguousarray(c, np.float32); eta = np.ascontiguousarray(eta, np.float32)
        return self._map(self.lib.ref_fresnel_dielectric_cos, len(c), 1, _fp(c), _fp(eta))[:, 0]

    def ggx_eval(self, wi, wo, ax, ay, distrib):
        wi = np.ascontiguousarray(wi, np.float32); wo = np.ascontiguousarray(wo, np.float32)
        out = np.empty((len(wi), 2), np.float32)
        self.lib.ref_ggx_eval(_fp(wi), _fp(wo), C.c_float(ax), C.c_float(ay), C.c_int(distrib),
                              C.c_uint64(len(wi)), _fp(out))
        return out

    def ggx_sample(self, wo, ax, ay, u, distrib):
        wo = np.ascontiguousarray(wo, np.float32); u = np.ascontiguousarray(u, np.float32)
        out = np.empty((len(wo), 5), np.float32)
        self.lib.ref_ggx_sample(_fp(wo), C.c_float(ax), C.c_float(ay), _fp(u), C.c_int(distrib),
                                C.c_uint64(len(wo)), _fp(out))
        return out

    def principled_eval(self, p23, wi, wo):
        p = np.ascontiguousarray(p23, np.float32)
        wi = np.ascontiguousarray(wi, np.float32); wo = np.ascontiguousarray(wo, np.float32)
        out = np.empty((len(wi), 4), np.float32)
        bsdf = np.empty(40, np.float32)
        self.lib.ref_principled_eval(_fp(p), _fp(wi), _fp(wo), C.c_uint64(len(wi)), _fp(out), _fp(bsdf))
        return out, bsdf[:36]

    def principled_weights(self, p23, wo):
        p = np.ascontiguousarray(p23, np.float32); wo = np.ascontiguousarray(wo, np.float32)
        out = np.empty((len(wo), 4), np.float32)
        self.lib.ref_principled_weights(_fp(p), _fp(wo), C.c_uint64(len(wo)), _fp(out))
        return out

    def hair_eval(self, p20, h, wi, wo):
        p = np.ascontiguousarray(p20, np.float32); h = np.ascontiguousarray(h, np.float32)
        wi = np.ascontiguousarray(wi, np.float32); wo = np.ascontiguousarray(wo, np.float32)
        out = np.empty((len(wi), 4), np.float32)
        self.lib.ref_hair_eval(_fp(p), _fp(h), _fp(wi), _fp(wo), C.c_uint64(len(wi)), _fp(out))
        return out

    def hair_sample(self, p20, h, wo, us):
        p = np.ascontiguousarray(p20, np.float32); h = np.ascontiguousarray(h, np.float32)
        wo = np.ascontiguousarray(wo, np.float32); us = np.ascontiguousarray(us, np.float32)
        out = np.empty((len(wo), 7), np.float32)
        self.lib.ref_hair_sample(_fp(p), _fp(h), _fp(wo), _fp(us), C.c_uint64(len(wo)), _fp(out))
        return out

    def hair_setup(self, p20):
        p = np.ascontiguousarray(p20, np.float32)
        out = np.empty(9, np.float32)
        self.lib.ref_hair_setup(_fp(p), _fp(out))
        return out

    def sss_coefficients(self, albedo, radius, weight):
        i = np.ascontiguousarray(np.concatenate([albedo, radius, weight]), np.float32)
        out = np.empty(9, np.float32)
        self.lib.ref_sss_coefficients(_fp(i), _fp(out))
        return out

    def sss_sample_distance(self, thr, sigma_s, sigma_t, u):
        i = np.ascontiguousarray(np.concatenate([thr, sigma_s, sigma_t, u]), np.float32)
        out = np.empty(4, np.float32)
        self.lib.ref_sss_sample_distance(_fp(i), _fp(out))
        return out

    # ---- loaders
    def output_stage(self, rgba_sums, count, directory, name="rgba.png"):
        """the reference CLI's output statements (pc/pbrlab-cli.cc:47-57) on a RenderLayer: writes directory/name"""
        rgba_sums = np.ascontiguousarray(rgba_sums, np.float32); count = np.ascontiguousarray(count, np.uint32)
        h, w = count.shape
        ok = self.lib.ref_output_stage(_fp(rgba_sums), count.ctypes.data_as(C.c_void_p), C.c_uint32(w), C.c_uint32(h),
                                       name.encode(), directory.encode())
        return bool(ok)

    def obj_load(self, path):
        h = self.lib.ref_obj_load(path.encode())
        if not h:
            raise RuntimeError("reference LoadTriangleMeshFromObj failed: " + path)
        return RefObj(self, h)

    def hair_load(self, path):
        nf = C.c_uint64(0); ni = C.c_uint64(0)
        ok = self.lib.ref_hair_load(path.encode(), None, C.byref(nf), None, C.byref(ni))
        v = np.empty(nf.value, np.float32); idx = np.empty(ni.value, np.uint32)
        ok = self.lib.ref_hair_load(path.encode(), _fp(v), C.byref(nf), _fp(idx), C.byref(ni))
        return bool(ok), v.reshape(-1, 4), idx

    # ---- ray counters (count flavour only)
    def ray_counts(self):
        out = np.zeros(2, np.uint64)
        self.lib.ref_ray_counts(_fp(out))
        return int(out[0]), int(out[1])

    def ray_counts_reset(self):
        self.lib.ref_ray_counts_reset()


class RefObj:
    def __init__(self, ref, h):
        self.ref, self.h = ref, C.c_void_p(h)

    def close(self):
        if self.h:
            self.ref.lib.ref_obj_free(self.h); self.h = None

    def num_shapes(self):
        return int(self.ref.lib.ref_obj_num_shapes(self.h))

    def num_materials(self):
        return int(self.ref.lib.ref_obj_num_materials(self.h))

    def vertices(self):
        n = int(self.ref.lib.ref_obj_num_vertices(self.h))
        v = np.empty((n, 4), np.float32)
        self.ref.lib.ref_obj_vertices(self.h, _fp(v))
        return v

    def shape(self, i):
        name = C.create_string_buffer(256)
        nf = int(self.ref.lib.ref_obj_shape_faces(self.h, C.c_int(i), name, C.c_int(256)))
        vid = np.empty((nf, 3), np.uint32); mid = np.empty(nf, np.uint32)
        self.ref.lib.ref_obj_shape_ids(self.h, C.c_int(i), _fp(vid), _fp(mid))
        return name.value.decode(), vid, mid

    def shading_normal(self, i, prim, uv):
        prim = np.ascontiguousarray(prim, np.uint32); uv = np.ascontiguousarray(uv, np.float32)
        out = np.empty((len(prim), 3), np.float32)
        self.ref.lib.ref_obj_shading_normal(self.h, C.c_int(i), _fp(prim), _fp(uv), C.c_uint64(len(prim)), _fp(out))
        return out

    def material(self, i):
        p = np.empty(23, np.float32); tex = np.empty(2, np.uint32)
        name = C.create_string_buffer(256)
        kind = int(self.ref.lib.ref_obj_material(self.h, C.c_int(i), _fp(p), _fp(tex), name, C.c_int(256)))
        return kind, p, tex, name.value.decode()


    def num_textures(self):
        return int(self.ref.lib.ref_obj_num_textures(self.h))

    def texture(self, i):
        """(h, w, c) float pixels of texture i as the reference's loader left them (de-gammaed where it does)"""
        whc = np.zeros(3, np.uint32)
        self.ref.lib.ref_obj_texture(self.h, C.c_int(i), _fp(whc), None)
        px = np.zeros((int(whc[1]), int(whc[0]), int(whc[2])), np.float32)
        self.ref.lib.ref_obj_texture(self.h, C.c_int(i), _fp(whc), _fp(px))
        return px

    def texture_fetch3(self, i, uv):
        uv = np.ascontiguousarray(uv, np.float32)
        out = np.zeros((len(uv), 3), np.float32)
        self.ref.lib.ref_obj_texture_fetch3(self.h, C.c_int(i), _fp(uv), C.c_uint64(len(uv)), _fp(out))
        return out


class RefScene:
    def __init__(self, ref, h):
        self.ref, self.h = ref, C.c_void_p(h)

    def close(self):
        if self.h:
            self.ref.lib.ref_scene_destroy(self.h); self.h = None

    def aabb(self):
        bmin = np.empty(3, np.float32); bmax = np.empty(3, np.float32)
        self.ref.lib.ref_scene_aabb(self.h, _fp(bmin), _fp(bmax))
        return bmin, bmax

    def camera(self, w, h):
        cam = np.empty(8, np.float32)
        self.ref.lib.ref_camera(self.h, C.c_uint32(w), C.c_uint32(h), _fp(cam))
        return cam

    def trace(self, rays):
        rays = np.ascontiguousarray(rays, np.float32)
        n = len(rays)
        f = np.empty((n, 6), np.float32); ids = np.empty((n, 3), np.uint32)
        self.ref.lib.ref_trace(self.h, _fp(rays), C.c_uint64(n), _fp(f), _fp(ids))
        return f, ids

    def occluded(self, rays):
        rays = np.ascontiguousarray(rays, np.float32)
        out = np.empty(len(rays), np.uint8)
        self.ref.lib.ref_occluded(self.h, _fp(rays), C.c_uint64(len(rays)), _fp(out))
        return out

    def radiance(self, rays, seeds):
        rays = np.ascontiguousarray(rays, np.float32); seeds = np.ascontiguousarray(seeds, np.uint64)
        out = np.empty((len(rays), 3), np.float32)
        self.ref.lib.ref_radiance(self.h, _fp(rays), _fp(seeds), C.c_uint64(len(rays)), _fp(out))
        return out

    def shade(self, rays, seeds):
        rays = np.ascontiguousarray(rays, np.float32); seeds = np.ascontiguousarray(seeds, np.uint64)
        out = np.empty((len(rays), 16), np.float32)
        self.ref.lib.ref_shade(self.h, _fp(rays), _fp(seeds), C.c_uint64(len(rays)), _fp(out))
        return out

    def surface(self, rays):
        rays = np.ascontiguousarray(rays, np.float32)
        out = np.empty((len(rays), 12), np.float32)
        self.ref.lib.ref_surface(self.h, _fp(rays), C.c_uint64(len(rays)), _fp(out))
        return out

    def sample_light(self, seeds):
        seeds = np.ascontiguousarray(seeds, np.uint64)
        out = np.empty((len(seeds), 10), np.float32)
        t = self.ref.lib.ref_sample_light(self.h, _fp(seeds), C.c_uint64(len(seeds)), _fp(out))
        return out, int(t)

    def implicit_light(self, inst, geom, prim):
        out = np.empty(4, np.float32)
        has = self.ref.lib.ref_implicit_light(self.h, C.c_uint32(inst), C.c_uint32(geom), C.c_uint32(prim), _fp(out))
        return bool(has), out

    def render(self, w, h, spp):
        rgba = np.empty((h, w, 4), np.float32); count = np.empty((h, w), np.uint32)
        sec = self.ref.lib.ref_render(self.h, C.c_uint32(w), C.c_uint32(h), C.c_uint32(spp), _fp(rgba), _fp(count))
        return rgba, count, float(sec)
