"""pbrlab_b200 — B200 (sm_100a) backend for pbrlab's path-tracing hot path.

Python is only plumbing here (tests, bench, multi-process launch): the product is
  * lib/libpbrgpu.so       — the C ABI of include/pbrgpu.h (hand-written CUDA kernels), and
  * lib/libpbrlab_host.so  — the C++ mirror of the reference's Scene / Render() / loader API on top of it.
This module binds both with ctypes.  There is no CPU fallback: if the CUDA library is missing or no GPU is present,
creating a context raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(_ROOT)
LIB_DIR = os.path.join(_ROOT, "lib")
GPU_LIB = os.path.join(LIB_DIR, "libpbrgpu.so")
HOST_LIB = os.path.join(LIB_DIR, "libpbrlab_host.so")

INVALID = 0xFFFFFFFF


def build(targets=("gpu", "host")):
    """Compile the native libraries in-tree (nvcc -gencode arch=compute_100a,code=sm_100a; see the Makefile)."""
    subprocess.run(["make", "-C", REPO, "-j8"] + list(targets), check=True)


class Material(C.Structure):
    _fields_ = [("type", C.c_uint32), ("tex_id", C.c_uint32 * 2), ("reserved", C.c_uint32), ("p", C.c_float * 24)]


class LightTables(C.Structure):
    _fields_ = [("num_lights", C.c_uint32), ("light_probability", C.c_void_p), ("light_cdf", C.c_void_p),
                ("light_prim_offset", C.c_void_p), ("num_light_prims", C.c_uint32), ("prim_probability", C.c_void_p),
                ("prim_cdf", C.c_void_p), ("prim_area_pdf", C.c_void_p), ("prim_emission", C.c_void_p),
                ("prim_is_emissive", C.c_void_p), ("prim_triangle", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("sss_rays", C.c_uint64), ("kernel_launches", C.c_uint64), ("seconds", C.c_double),
                ("trace_closest_ms", C.c_double), ("trace_any_ms", C.c_double), ("shade_ms", C.c_double),
                ("sss_ms", C.c_double), ("nodes_visited", C.c_uint64), ("prims_tested", C.c_uint64),
                ("regen_ms", C.c_double), ("device_ms", C.c_double), ("trace_closest_launches", C.c_uint64), ("sss_skipped", C.c_uint64),
                ("shade_vertices", C.c_uint64), ("iterations", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class CommitInfo(C.Structure):
    _fields_ = [("commit_s", C.c_double), ("bvh_s", C.c_double), ("clearance_s", C.c_double), ("upload_s", C.c_double),
                ("tri_builder", C.c_uint32), ("tri_nodes", C.c_uint32), ("curve_nodes", C.c_uint32), ("tri_depth", C.c_uint32)]


class Flat(C.Structure):
    _fields_ = [("verts", C.c_void_p), ("nverts", C.c_uint32), ("normals", C.c_void_p), ("nnormals", C.c_uint32),
                ("texcoords", C.c_void_p), ("ntexcoords", C.c_uint32),
                ("vidx", C.c_void_p), ("nidx", C.c_void_p), ("tidx", C.c_void_p), ("tri_material", C.c_void_p),
                ("tri_instance", C.c_void_p), ("tri_geom", C.c_void_p), ("tri_prim", C.c_void_p), ("ntris", C.c_uint64),
                ("curve_verts", C.c_void_p), ("ncurve_verts", C.c_uint32),
                ("curve_first", C.c_void_p), ("curve_material", C.c_void_p), ("curve_instance", C.c_void_p),
                ("curve_geom", C.c_void_p), ("curve_prim", C.c_void_p), ("nsegs", C.c_uint64),
                ("materials", C.c_void_p), ("nmaterials", C.c_uint32),
                ("lights", LightTables), ("bmin", C.c_float * 3), ("bmax", C.c_float * 3),
                ("tex_pixels", C.c_void_p), ("ntex_floats", C.c_uint64), ("tex_desc", C.c_void_p),
                ("ntextures", C.c_uint32)]


class Texture(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("channels", C.c_uint32),
                ("reserved", C.c_uint32)]


RAY_DTYPE = np.dtype([("org", np.float32, 3), ("tmin", np.float32), ("dir", np.float32, 3), ("tmax", np.float32)])
HIT_DTYPE = np.dtype([("normal_g", np.float32, 3), ("t", np.float32), ("u", np.float32), ("v", np.float32),
                      ("instance_id", np.uint32), ("geom_id", np.uint32), ("prim_id", np.uint32)])


def make_rays(org, direction, tmin=0.0, tmax=1.844e18):
    org = np.asarray(org, np.float32)
    direction = np.asarray(direction, np.float32)
    n = max(org.reshape(-1, 3).shape[0], direction.reshape(-1, 3).shape[0])
    r = np.zeros(n, RAY_DTYPE)
    r["org"] = org
    r["dir"] = direction
    r["tmin"] = tmin
    r["tmax"] = tmax
    return r


def rays_to_f8(rays):
    """(n,8) float view: org.xyz, tmin, dir.xyz, tmax — the layout of oracle/ref_harness.cc."""
    return np.ascontiguousarray(rays).view(np.float32).reshape(-1, 8)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _np_view(ptr, n, dtype, cols=None):
    if not ptr or n == 0:
        return np.zeros((0, cols) if cols else (0,), dtype)
    count = n * (cols or 1)
    buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
    a = np.frombuffer(buf, dtype=dtype, count=count).copy()
    return a.reshape(n, cols) if cols else a


class FlatScene:
    """Host copy of pbrlab::Scene::Flat(): what CommitScene() hands to pbrgpu_set_*."""

    def __init__(self, f):
        self.verts = _np_view(f.verts, f.nverts, np.float32, 4)
        self.normals = _np_view(f.normals, f.nnormals, np.float32, 4)
        self.texcoords = _np_view(f.texcoords, f.ntexcoords, np.float32, 2)
        nt = int(f.ntris)
        self.vidx = _np_view(f.vidx, nt, np.uint32, 3)
        self.nidx = _np_view(f.nidx, nt, np.uint32, 3)
        self.tidx = _np_view(f.tidx, nt, np.uint32, 3)
        self.tri_material = _np_view(f.tri_material, nt, np.uint32)
        self.tri_instance = _np_view(f.tri_instance, nt, np.uint32)
        self.tri_geom = _np_view(f.tri_geom, nt, np.uint32)
        self.tri_prim = _np_view(f.tri_prim, nt, np.uint32)
        ns = int(f.nsegs)
        self.curve_verts = _np_view(f.curve_verts, f.ncurve_verts, np.float32, 4)
        self.curve_first = _np_view(f.curve_first, ns, np.uint32)
        self.curve_material = _np_view(f.curve_material, ns, np.uint32)
        self.curve_instance = _np_view(f.curve_instance, ns, np.uint32)
        self.curve_geom = _np_view(f.curve_geom, ns, np.uint32)
        self.curve_prim = _np_view(f.curve_prim, ns, np.uint32)
        self.materials = _np_view(f.materials, f.nmaterials, np.uint32, 28)   # raw words of pbrgpu_material
        L = f.lights
        nl, npr = int(L.num_lights), int(L.num_light_prims)
        self.light_probability = _np_view(L.light_probability, nl, np.float32)
        self.light_cdf = _np_view(L.light_cdf, nl, np.float32)
        self.light_prim_offset = _np_view(L.light_prim_offset, nl + 1 if nl else 0, np.uint32)
        self.prim_probability = _np_view(L.prim_probability, npr, np.float32)
        self.prim_cdf = _np_view(L.prim_cdf, npr, np.float32)
        self.prim_area_pdf = _np_view(L.prim_area_pdf, npr, np.float32)
        self.prim_emission = _np_view(L.prim_emission, npr, np.float32, 3)
        self.prim_is_emissive = _np_view(L.prim_is_emissive, npr, np.uint32)
        self.prim_triangle = _np_view(L.prim_triangle, npr, np.uint32)
        self.bmin = np.array(list(f.bmin), np.float32)
        self.bmax = np.array(list(f.bmax), np.float32)
        self.tex_pixels = _np_view(f.tex_pixels, int(f.ntex_floats), np.float32)
        self.tex_desc = _np_view(f.tex_desc, int(f.ntextures), np.uint32, 4)   # offset, width, height, channels

    def textures(self):
        """ctypes array of pbrgpu_texture over self.tex_pixels (keep `self` alive while it is in use)."""
        arr = (Texture * max(1, len(self.tex_desc)))()
        for i, (off, w, h, c) in enumerate(self.tex_desc):
            arr[i].pixels = self.tex_pixels.ctypes.data + 4 * int(off)
            arr[i].width, arr[i].height, arr[i].channels = int(w), int(h), int(c)
        return arr

    def light_tables(self):
        t = LightTables()
        t.num_lights = len(self.light_probability)
        t.light_probability = _p(self.light_probability); t.light_cdf = _p(self.light_cdf)
        t.light_prim_offset = _p(self.light_prim_offset)
        t.num_light_prims = len(self.prim_probability)
        t.prim_probability = _p(self.prim_probability); t.prim_cdf = _p(self.prim_cdf)
        t.prim_area_pdf = _p(self.prim_area_pdf); t.prim_emission = _p(self.prim_emission)
        t.prim_is_emissive = _p(self.prim_is_emissive); t.prim_triangle = _p(self.prim_triangle)
        return t

    def upload(self, lib, h, prefix):
        """Feed the flat scene to an object exposing <prefix>set_* / commit (libpbrgpu or the test emulation)."""
        g = lambda name: getattr(lib, prefix + name)
        rc = g("set_textures")(h, self.textures(), C.c_uint32(len(self.tex_desc)))
        if rc: return rc
        rc = g("set_materials")(h, _p(self.materials), C.c_uint32(len(self.materials)))
        if rc: return rc
        rc = g("set_triangles")(h, _p(self.verts), C.c_uint32(len(self.verts)), _p(self.vidx), _p(self.normals),
                                C.c_uint32(len(self.normals)), _p(self.nidx), _p(self.texcoords),
                                C.c_uint32(len(self.texcoords)), _p(self.tidx), _p(self.tri_material),
                                _p(self.tri_instance), _p(self.tri_geom), _p(self.tri_prim),
                                C.c_uint64(len(self.tri_prim)))
        if rc: return rc
        rc = g("set_curves")(h, _p(self.curve_verts), C.c_uint32(len(self.curve_verts)), _p(self.curve_first),
                             _p(self.curve_material), _p(self.curve_instance), _p(self.curve_geom),
                             _p(self.curve_prim), C.c_uint64(len(self.curve_prim)))
        if rc: return rc
        lt = self.light_tables()
        rc = g("set_lights")(h, C.byref(lt))
        if rc: return rc
        return g("commit")(h, _p(self.bmin), _p(self.bmax))


_gpu = None
_host = None


def gpu_lib():
    """libpbrgpu.so.  Raises if it has not been built — never falls back to anything else."""
    global _gpu
    if _gpu is None:
        if not os.path.exists(GPU_LIB):
            raise RuntimeError("%s is missing: run `make gpu` (pbrlab_b200.build()); there is no CPU fallback" % GPU_LIB)
        L = C.CDLL(GPU_LIB, mode=C.RTLD_GLOBAL)
        L.pbrgpu_create.restype = C.c_void_p
        L.pbrgpu_create.argtypes = [C.c_void_p, C.c_int]
        L.pbrgpu_last_error.restype = C.c_char_p
        L.pbrgpu_last_error.argtypes = [C.c_void_p]
        L.pbrgpu_destroy.argtypes = [C.c_void_p]
        for name in ("pbrgpu_set_triangles", "pbrgpu_set_curves", "pbrgpu_set_materials", "pbrgpu_set_lights", "pbrgpu_set_textures",
                     "pbrgpu_resolve_srgb8", "pbrgpu_resolve_srgb8_device",
                     "pbrgpu_commit", "pbrgpu_scene_bounds", "pbrgpu_render", "pbrgpu_render_device",
                     "pbrgpu_get_stats", "pbrgpu_set_wave_spp", "pbrgpu_set_profiling", "pbrgpu_trace", "pbrgpu_occluded",
                     "pbrgpu_trace_device", "pbrgpu_occluded_device", "pbrgpu_radiance", "pbrgpu_radiance_mega",
                     "pbrgpu_shade", "pbrgpu_eval_closure", "pbrgpu_measure_gather", "pbrgpu_nccl_unique_id",
                     "pbrgpu_nccl_init", "pbrgpu_job_rank", "pbrgpu_get_commit_info"):
            getattr(L, name).restype = C.c_int
        _gpu = L
    return _gpu


def host_lib():
    global _host
    if _host is None:
        gpu_lib()
        if not os.path.exists(HOST_LIB):
            raise RuntimeError("%s is missing: run `make host`" % HOST_LIB)
        L = C.CDLL(HOST_LIB)
        L.pbrhost_scene_create.restype = C.c_void_p
        L.pbrhost_scene_create.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int]
        L.pbrhost_scene_create_on.restype = C.c_void_p
        L.pbrhost_scene_create_on.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_void_p, C.c_int]
        L.pbrhost_scene_destroy.argtypes = [C.c_void_p]
        L.pbrhost_scene_flat.argtypes = [C.c_void_p, C.POINTER(Flat)]
        L.pbrhost_scene_ctx.restype = C.c_void_p
        L.pbrhost_scene_ctx.argtypes = [C.c_void_p]
        L.pbrhost_last_error.restype = C.c_char_p
        L.pbrhost_render.restype = C.c_double
        L.pbrhost_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]
        L.pbrhost_render_layer.restype = C.c_double
        L.pbrhost_render_layer.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64,
                                           C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.POINTER(C.c_uint32))]
        L.pbrhost_render_cancel.restype = C.c_double
        L.pbrhost_render_cancel.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32,
                                            C.c_void_p, C.c_void_p, C.c_void_p]
        _host = L
    return _host


class Context:
    """Thin object wrapper of a pbrgpu_ctx* (owned unless borrowed from a Scene)."""

    def __init__(self, handle=None, device_ids=None, owner=None):
        self.lib = gpu_lib()
        self._owner = owner
        if handle is None:
            ids = None
            n = 0
            if device_ids:
                ids = (C.c_int * len(device_ids))(*device_ids)
                n = len(device_ids)
            handle = self.lib.pbrgpu_create(ids, n)
            if not handle:
                raise RuntimeError(self.lib.pbrgpu_last_error(None).decode())
            self._owned = True
        else:
            self._owned = False
        self.h = C.c_void_p(handle)

    def close(self):
        if self._owned and self.h:
            self.lib.pbrgpu_destroy(self.h)
        self.h = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("pbrgpu error %d: %s" % (rc, self.lib.pbrgpu_last_error(self.h).decode()))

    def upload(self, flat):
        self._check(flat.upload(self.lib, self.h, "pbrgpu_"))

    def set_materials(self, words):
        words = np.ascontiguousarray(words, np.uint32)
        self._check(self.lib.pbrgpu_set_materials(self.h, _p(words), C.c_uint32(len(words))))

    def nccl_init(self, id_bytes, rank, world):
        """pbrgpu_nccl_init: join the multi-process job (collective); id_bytes from nccl_unique_id() of rank 0."""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(id_bytes))
        self._check(self.lib.pbrgpu_nccl_init(self.h, buf, C.c_int(rank), C.c_int(world)))

    def set_wave_spp(self, n):
        self._check(self.lib.pbrgpu_set_wave_spp(self.h, C.c_uint32(n)))

    def set_profiling(self, on):
        self._check(self.lib.pbrgpu_set_profiling(self.h, C.c_int(1 if on else 0)))

    def measure_gather(self, working_set_bytes, records_per_thread=4096, chains=8):
        """GB/s of node-sized (80 B) random gathers over a working set: the L2 / HBM gather roofline of the traversal"""
        out = C.c_double(0.0)
        self._check(self.lib.pbrgpu_measure_gather(self.h, C.c_uint64(int(working_set_bytes)),
                                                   C.c_uint32(records_per_thread), C.c_uint32(chains), C.byref(out)))
        return out.value

    def bounds(self):
        a = np.zeros(3, np.float32); b = np.zeros(3, np.float32)
        self._check(self.lib.pbrgpu_scene_bounds(self.h, _p(a), _p(b)))
        return a, b

    def commit_info(self):
        """pbrgpu_get_commit_info: seconds of the last commit and of its parts, and which builder made the triangle BVH"""
        ci = CommitInfo()
        self._check(self.lib.pbrgpu_get_commit_info(self.h, C.byref(ci)))
        d = {k: getattr(ci, k) for k, _ in ci._fields_}
        d["tri_builder"] = ["sah (host)", "ploc (host)", "ploc (device)"][min(2, d["tri_builder"])]
        return d

    def stats(self):
        s = Stats()
        self._check(self.lib.pbrgpu_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def render(self, width, height, spp, seed=1234567890, sample_offset=0, sample_stride=1, cancel=None):
        """pbrgpu_render with HOST output buffers (sums, like RenderLayer).  cancel: optional ctypes c_int the call
        polls (raise it from another thread)."""
        rgba = np.empty((height, width, 4), np.float32)
        count = np.empty((height, width), np.uint32)
        self._check(self.lib.pbrgpu_render(self.h, C.c_uint32(width), C.c_uint32(height), C.c_uint32(spp),
                                           C.c_uint64(seed), C.c_uint32(sample_offset), C.c_uint32(sample_stride),
                                           C.byref(cancel) if cancel is not None else None, _p(rgba), _p(count), None))
        return rgba, count

    def render_device(self, width, height, spp, d_rgba_ptr, d_count_ptr, seed=1234567890, sample_offset=0,
                      sample_stride=1):
        """pbrgpu_render_device: results stay in the given DEVICE buffers (e.g. torch tensors' data_ptr())."""
        self._check(self.lib.pbrgpu_render_device(self.h, C.c_uint32(width), C.c_uint32(height), C.c_uint32(spp),
                                                  C.c_uint64(seed), C.c_uint32(sample_offset),
                                                  C.c_uint32(sample_stride), None, C.c_void_p(d_rgba_ptr),
                                                  C.c_void_p(d_count_ptr), None))

    def resolve_srgb8_device(self, d_rgba_ptr, d_count_ptr, width, height, d_out_ptr):
        """pbrgpu_resolve_srgb8_device: the output stage on caller-supplied DEVICE accumulators"""
        self._check(self.lib.pbrgpu_resolve_srgb8_device(self.h, C.c_void_p(d_rgba_ptr), C.c_void_p(d_count_ptr),
                                                         C.c_uint32(width), C.c_uint32(height), C.c_void_p(d_out_ptr)))

    def resolve_srgb8(self, width, height):
        """pbrgpu_resolve_srgb8: the CLI's output stage (mean -> sRGB -> 8 bit) of the last rendered frame."""
        out = np.empty((height, width, 4), np.uint8)
        self._check(self.lib.pbrgpu_resolve_srgb8(self.h, C.c_uint32(width), C.c_uint32(height), _p(out)))
        return out

    def trace(self, rays):
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(len(rays), HIT_DTYPE)
        self._check(self.lib.pbrgpu_trace(self.h, _p(rays), C.c_uint64(len(rays)), _p(hits)))
        return hits

    def occluded(self, rays):
        rays = np.ascontiguousarray(rays)
        out = np.zeros(len(rays), np.uint8)
        self._check(self.lib.pbrgpu_occluded(self.h, _p(rays), C.c_uint64(len(rays)), _p(out)))
        return out

    def trace_device(self, d_rays_ptr, n, d_hits_ptr=0, collect_stats=0):
        self._check(self.lib.pbrgpu_trace_device(self.h, C.c_void_p(d_rays_ptr), C.c_uint64(n),
                                                 C.c_void_p(d_hits_ptr) if d_hits_ptr else None, C.c_int(collect_stats)))

    def occluded_device(self, d_rays_ptr, n, d_out_ptr=0):
        self._check(self.lib.pbrgpu_occluded_device(self.h, C.c_void_p(d_rays_ptr), C.c_uint64(n),
                                                    C.c_void_p(d_out_ptr) if d_out_ptr else None))

    def radiance(self, rays, seeds, mega=False):
        rays = np.ascontiguousarray(rays); seeds = np.ascontiguousarray(seeds, np.uint64)
        out = np.zeros((len(rays), 3), np.float32)
        fn = self.lib.pbrgpu_radiance_mega if mega else self.lib.pbrgpu_radiance
        self._check(fn(self.h, _p(rays), _p(seeds), C.c_uint64(len(rays)), _p(out)))
        return out

    def shade(self, rays, seeds):
        rays = np.ascontiguousarray(rays); seeds = np.ascontiguousarray(seeds, np.uint64)
        out = np.zeros((len(rays), 16), np.float32)
        self._check(self.lib.pbrgpu_shade(self.h, _p(rays), _p(seeds), C.c_uint64(len(rays)), _p(out)))
        return out

    def eval_closure(self, op, params, inputs, out_stride):
        prm = np.zeros(32, np.float32)
        if params is not None:
            params = np.asarray(params, np.float32).ravel()
            prm[:len(params)] = params
        inputs = np.ascontiguousarray(inputs, np.float32)
        if inputs.ndim == 1:
            inputs = inputs.reshape(-1, 1)
        n, stride = inputs.shape
        out = np.zeros((n, out_stride), np.float32)
        self._check(self.lib.pbrgpu_eval_closure(self.h, C.c_int(op), _p(prm), _p(inputs), C.c_uint32(stride),
                                                 C.c_uint64(n), _p(out), C.c_uint32(out_stride)))
        return out


def nccl_unique_id():
    """pbrgpu_nccl_unique_id: 128 bytes rank 0 hands to every rank of a multi-process job."""
    buf = (C.c_uint8 * 128)()
    rc = gpu_lib().pbrgpu_nccl_unique_id(buf)
    if rc != 0:
        raise RuntimeError("pbrgpu_nccl_unique_id: " + gpu_lib().pbrgpu_last_error(None).decode())
    return bytes(buf)


class Scene:
    """pbrlab::Scene built by CreateScene() from .obj / .hair files (the reference CLI's path)."""

    def __init__(self, files, commit_to_device=True, device_ids=None):
        self.lib = host_lib()
        arr = (C.c_char_p * len(files))(*[f.encode() for f in files])
        if device_ids:
            ids = (C.c_int * len(device_ids))(*device_ids)
            h = self.lib.pbrhost_scene_create_on(len(files), arr, 1 if commit_to_device else 0, ids, len(device_ids))
        else:
            h = self.lib.pbrhost_scene_create(len(files), arr, 1 if commit_to_device else 0)
        if not h:
            raise RuntimeError("CreateScene failed: " + self.lib.pbrhost_last_error().decode())
        self.h = C.c_void_p(h)
        self._flat = None

    def close(self):
        if self.h:
            self.lib.pbrhost_scene_destroy(self.h)
            self.h = None

    def flat(self):
        if self._flat is None:
            f = Flat()
            self.lib.pbrhost_scene_flat(self.h, C.byref(f))
            self._flat = FlatScene(f)
        return self._flat

    def context(self):
        h = self.lib.pbrhost_scene_ctx(self.h)
        if not h:
            raise RuntimeError("scene is not committed to a device")
        return Context(handle=h, owner=self)

    def render(self, width, height, spp, seed=1234567890):
        """pbrlab::Render() through the C++ entry point; returns (rgba sums, count, seconds)."""
        rgba = np.empty((height, width, 4), np.float32)
        count = np.empty((height, width), np.uint32)
        sec = self.lib.pbrhost_render(self.h, width, height, spp, seed, _p(rgba), _p(count))
        if sec < 0:
            raise RuntimeError("Render failed: " + self.lib.pbrhost_last_error().decode())
        return rgba, count, sec

    def render_layer(self, width, height, spp, seed=1234567890):
        """pbrlab::Render() into the RenderLayer the scene keeps across frames (as the reference's GUI / CLI hold one):
        returns (rgba sums, count, seconds) as VIEWS of the layer's host buffers, valid until the next call."""
        pr = C.POINTER(C.c_float)()
        pc = C.POINTER(C.c_uint32)()
        sec = self.lib.pbrhost_render_layer(self.h, width, height, spp, seed, C.byref(pr), C.byref(pc))
        if sec < 0:
            raise RuntimeError("Render failed: " + self.lib.pbrhost_last_error().decode())
        rgba = np.ctypeslib.as_array(pr, shape=(height, width, 4))
        count = np.ctypeslib.as_array(pc, shape=(height, width))
        return rgba, count, sec

    def render_cancelled(self, width, height, spp, cancel_at_pass, seed=1234567890):
        """pbrlab::Render() with the cancel flag raised by a second thread once finish_pass >= cancel_at_pass.
        Returns (rgba sums, count, info) with info = dict(returned, finish_pass, progress_violation, raised_at, seconds)."""
        rgba = np.empty((height, width, 4), np.float32)
        count = np.empty((height, width), np.uint32)
        out = np.zeros(4, np.uint64)
        sec = self.lib.pbrhost_render_cancel(self.h, width, height, spp, seed, cancel_at_pass, _p(rgba), _p(count), _p(out))
        if sec < 0:
            raise RuntimeError("Render threw: " + self.lib.pbrhost_last_error().decode())
        return rgba, count, {"returned": bool(out[0]), "finish_pass": int(out[1]), "progress_violation": bool(out[2]),
                             "raised_at": int(out[3]), "seconds": sec}
