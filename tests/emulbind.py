"""ctypes binding of tests/host_emul/libpbr_emul.so: the device headers compiled by g++ (TEST ONLY)."""
import ctypes as C
import os

import numpy as np

import pbrlab_b200 as pb

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emul", "libpbr_emul.so")


def available():
    return os.path.exists(_PATH)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Emul:
    def __init__(self, flat=None):
        self.lib = C.CDLL(_PATH)
        self.lib.emul_create.restype = C.c_void_p
        self.lib.emul_last_error.restype = C.c_char_p
        self.lib.emul_last_error.argtypes = [C.c_void_p]
        self.h = C.c_void_p(self.lib.emul_create())
        if flat is not None:
            rc = flat.upload(self.lib, self.h, "emul_")
            if rc:
                raise RuntimeError("emul upload failed: " + self.lib.emul_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.emul_destroy(self.h); self.h = None

    def bounds(self):
        a = np.zeros(3, np.float32); b = np.zeros(3, np.float32)
        self.lib.emul_scene_bounds(self.h, _p(a), _p(b))
        return a, b

    def texture_fetch3(self, tex, uv):
        uv = np.ascontiguousarray(uv, np.float32)
        out = np.zeros((len(uv), 3), np.float32)
        rc = self.lib.emul_texture_fetch3(self.h, C.c_uint32(tex), _p(uv), C.c_uint64(len(uv)), _p(out))
        assert rc == 0
        return out

    def material_classes(self):
        out = np.zeros(4096, np.uint32)
        n = self.lib.emul_material_classes(self.h, _p(out))
        return out[:n]

    def bvh_info(self):
        o = np.zeros(6, np.float64)
        self.lib.emul_bvh_info(self.h, _p(o))
        return o

    def trace(self, rays, stats=False):
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(len(rays), pb.HIT_DTYPE)
        st = np.zeros(2, np.uint64)
        self.lib.emul_trace(self.h, _p(rays), C.c_uint64(len(rays)), _p(hits), _p(st))
        return (hits, st) if stats else hits

    def occluded(self, rays):
        rays = np.ascontiguousarray(rays)
        out = np.zeros(len(rays), np.uint8)
        self.lib.emul_occluded(self.h, _p(rays), C.c_uint64(len(rays)), _p(out))
        return out

    def radiance(self, rays, seeds, counts=False):
        rays = np.ascontiguousarray(rays); seeds = np.ascontiguousarray(seeds, np.uint64)
        out = np.zeros((len(rays), 3), np.float32)
        c = np.zeros(3, np.uint64)
        self.lib.emul_radiance(self.h, _p(rays), _p(seeds), C.c_uint64(len(rays)), _p(out), _p(c))
        return (out, c) if counts else out

    def curve_probe(self):
        """(curve leaf tests, tests that passed CurveMayHit) since the last call"""
        o = np.zeros(2, np.uint64)
        self.lib.emul_curve_probe(_p(o))
        return [int(x) for x in o]

    def clearance_field(self):
        """(bytes [dz, dy, dx], origin xyz, 1 / cell, quantum) of the committed scene's clearance field, or None"""
        geo = np.zeros(7, np.float32); dims = np.zeros(3, np.uint32)
        self.lib.emul_clearance_field.restype = C.c_uint64
        n = self.lib.emul_clearance_field(self.h, _p(geo), _p(dims), None)
        if not n:
            return None
        b = np.zeros(int(n), np.uint8)
        self.lib.emul_clearance_field(self.h, _p(geo), _p(dims), _p(b))
        return b.reshape(int(dims[2]), int(dims[1]), int(dims[0])), geo[:3].copy(), float(geo[3]), float(geo[4])

    def clearance_probe(self):
        """(walk segments, segments the clearance field skips, skipped segments that hit) since the last call"""
        o = np.zeros(3, np.uint64)
        self.lib.emul_clearance_probe(_p(o))
        return [int(x) for x in o]

    def shade(self, rays, seeds):
        rays = np.ascontiguousarray(rays); seeds = np.ascontiguousarray(seeds, np.uint64)
        out = np.zeros((len(rays), 16), np.float32)
        self.lib.emul_shade(self.h, _p(rays), _p(seeds), C.c_uint64(len(rays)), _p(out))
        return out

    def surface(self, rays):
        rays = np.ascontiguousarray(rays)
        out = np.zeros((len(rays), 12), np.float32)
        self.lib.emul_surface(self.h, _p(rays), C.c_uint64(len(rays)), _p(out))
        return out

    def sample_light(self, seeds):
        seeds = np.ascontiguousarray(seeds, np.uint64)
        out = np.zeros((len(seeds), 10), np.float32)
        self.lib.emul_sample_light(self.h, _p(seeds), C.c_uint64(len(seeds)), _p(out))
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
        self.lib.emul_eval_closure(C.c_int(op), _p(prm), _p(inputs), C.c_uint32(stride), C.c_uint64(n), _p(out),
                                   C.c_uint32(out_stride))
        return out

    def job_share(self, rank, world, spp, job_offset=0, job_stride=1, device=0, num_devices=1):
        """(offset, stride, count): the samples pbrgpu_render gives this worker (csrc/job_split.h)"""
        out = np.zeros(3, np.uint32)
        ok = self.lib.emul_job_share(C.c_uint32(job_offset), C.c_uint32(job_stride), C.c_uint32(rank), C.c_uint32(world),
                                     C.c_uint32(device), C.c_uint32(num_devices), C.c_uint32(spp), _p(out))
        assert ok
        return int(out[0]), int(out[1]), int(out[2])

    def sample_order(self, order, npix, probe, rest, block):
        """(pixel, local sample) of every camera sample id of a worker's frame (csrc/job_split.h: SampleOfId)"""
        n = npix * (probe + rest)
        pix = np.zeros(n, np.uint32); smp = np.zeros(n, np.uint32)
        o = None if order is None else np.ascontiguousarray(order, np.uint32)
        self.lib.emul_sample_order(_p(o), C.c_uint32(npix), C.c_uint32(probe), C.c_uint32(rest), C.c_uint32(block),
                                   _p(pix), _p(smp))
        return pix, smp

    def render(self, width, height, spp, seed=1234567890, sample_offset=0, sample_stride=1):
        rgba = np.zeros((height, width, 4), np.float32); count = np.zeros((height, width), np.uint32)
        c = np.zeros(3, np.uint64)
        self.lib.emul_render(self.h, C.c_uint32(width), C.c_uint32(height), C.c_uint32(spp), C.c_uint64(seed),
                             C.c_uint32(sample_offset), C.c_uint32(sample_stride), _p(rgba), _p(count), _p(c))
        return rgba, count, c
