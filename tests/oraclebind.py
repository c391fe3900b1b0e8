"""ctypes binding of oracle/libpbr_oracle.so, the CPU restatement of the reference path (TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs, never by pbrlab_b200/)."""
import ctypes as C
import os

import numpy as np

import pbrlab_b200 as pb

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "libpbr_oracle.so")


def available():
    return os.path.exists(_PATH)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """Same duck type as pbrlab_b200.Context / emulbind.Emul for trace / occluded / radiance / render."""

    def __init__(self, flat):
        self.lib = C.CDLL(_PATH)
        self.lib.pbo_create.restype = C.c_void_p
        self.lib.pbo_last_error.restype = C.c_char_p
        self.lib.pbo_last_error.argtypes = [C.c_void_p]
        self.lib.pbo_render.restype = C.c_double
        self.h = C.c_void_p(self.lib.pbo_create())
        rc = flat.upload(self.lib, self.h, "pbo_")
        if rc:
            raise RuntimeError("oracle upload failed: " + self.lib.pbo_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.pbo_destroy(self.h)
            self.h = None

    def trace(self, rays):
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(len(rays), pb.HIT_DTYPE)
        self.lib.pbo_trace(self.h, _p(rays), C.c_uint64(len(rays)), _p(hits))
        return hits

    def occluded(self, rays):
        rays = np.ascontiguousarray(rays)
        out = np.zeros(len(rays), np.uint8)
        self.lib.pbo_occluded(self.h, _p(rays), C.c_uint64(len(rays)), _p(out))
        return out

    def radiance(self, rays, seeds, counts=False):
        rays = np.ascontiguousarray(rays)
        seeds = np.ascontiguousarray(seeds, np.uint64)
        out = np.zeros((len(rays), 3), np.float32)
        c = np.zeros(3, np.uint64)
        self.lib.pbo_radiance(self.h, _p(rays), _p(seeds), C.c_uint64(len(rays)), _p(out), _p(c))
        return (out, c) if counts else out

    def render(self, width, height, spp, seed=1234567890, threads=0):
        rgba = np.zeros((height, width, 4), np.float32)
        count = np.zeros((height, width), np.uint32)
        c = np.zeros(3, np.uint64)
        sec = self.lib.pbo_render(self.h, C.c_uint32(width), C.c_uint32(height), C.c_uint32(spp), C.c_uint64(seed),
                                  _p(rgba), _p(count), C.c_int(threads), _p(c))
        if sec < 0:
            raise RuntimeError("oracle render failed")
        return rgba, count, sec, c
