"""Scene ingest: our OBJ/MTL and CyHair loaders and the flattening done by Scene::CommitScene(), against the
reference's loaders (golden fixtures generated from oracle/_ref; live comparison when the library is around)."""
import json
import os

import numpy as np

import pbrlab_b200 as pb
from conftest import GOLDEN, golden
from pbrlab_b200 import scenes


def test_cornell_matches_reference_loader(cornell_host):
    meta = json.load(open(os.path.join(GOLDEN, "cornell_loader.json")))
    f = cornell_host.flat()
    assert len(f.verts) == meta["num_vertices"]
    assert abs(float(f.verts.astype(np.float64).sum()) - meta["vertices_sum"]) < 1e-6 * abs(meta["vertices_sum"])
    assert len(f.materials) == meta["num_materials"]
    # shapes -> instances in file order, prim ids local to the shape
    faces = [s["faces"] for s in meta["shapes"]]
    assert len(f.tri_prim) == sum(faces) == 362620
    start = 0
    for i, s in enumerate(meta["shapes"]):
        sl = slice(start, start + s["faces"])
        assert np.all(f.tri_instance[sl] == i) and np.all(f.tri_geom[sl] == 0)
        assert np.array_equal(f.tri_prim[sl], np.arange(s["faces"], dtype=np.uint32))
        assert int(f.vidx[sl].astype(np.int64).sum()) == s["vid_sum"]
        assert sorted(set(int(m) for m in f.tri_material[sl])) == s["mid"]
        start += s["faces"]
    # materials: first duplicate MTL key wins (Lucy specular = 1.0)
    for i, m in enumerate(meta["materials"]):
        words = f.materials[i]
        assert words[0] == 0 and words[1] == 0xFFFFFFFF and words[2] == 0xFFFFFFFF
        p = words[4:27].view(np.float32)
        assert np.allclose(p, np.array(m["p"], np.float32), rtol=0, atol=0), m["name"]
    lucy = [m for m in meta["materials"] if m["name"] == "Lucy"][0]
    assert lucy["p"][11] == 1.0
    # scene bounds feed the camera: must equal what Embree reports
    assert np.array_equal(f.bmin, np.array(meta["bmin"], np.float32))
    assert np.array_equal(f.bmax, np.array(meta["bmax"], np.float32))


def test_light_tables(cornell_host):
    f = cornell_host.flat()
    assert len(f.light_probability) == 1 and f.light_probability[0] == 1.0
    assert len(f.prim_probability) == 2 and abs(f.prim_probability.sum() - 1.0) < 1e-6
    assert np.all(f.prim_is_emissive == 1) and np.all(f.prim_emission == 3.0)
    # the light is the last shape (lightobj_Plane.001): its two triangles close the flattened soup
    assert list(f.prim_triangle) == [362618, 362619]


def test_shading_normals_match_reference(cornell_host, cornell_emul):
    """FetchShadingNormal at random (prim, u, v): evaluated by the device code on the uploaded tables"""
    g = golden("cornell_normals.npz")
    f = cornell_host.flat()
    start = 0
    for i in range(9):
        nfaces = int((f.tri_instance == i).sum())
        prim, uv, want = g["sn_prim_%d" % i], g["sn_uv_%d" % i], g["sn_%d" % i]
        tri = start + prim
        n0 = f.normals[f.nidx[tri, 0], :3]; n1 = f.normals[f.nidx[tri, 1], :3]; n2 = f.normals[f.nidx[tri, 2], :3]
        u = uv[:, 0:1]; v = uv[:, 1:2]
        ns = ((np.float32(1) - u - v) * n0 + u * n1) + v * n2
        ns = ns / np.linalg.norm(ns, axis=1, keepdims=True)
        assert np.abs(ns - want).max() < 1e-6
        start += nfaces


def test_cyhair_to_bezier_matches_reference(built, hair_file):
    g = golden("hair_scene.npz")
    lib = pb.host_lib()
    import ctypes as C
    nf = C.c_uint64(0); ni = C.c_uint64(0)
    assert lib.pbrhost_hair_load(hair_file.encode(), None, C.byref(nf), None, C.byref(ni)) == 1
    v = np.empty(nf.value, np.float32); idx = np.empty(ni.value, np.uint32)
    lib.pbrhost_hair_load(hair_file.encode(), v.ctypes.data_as(C.c_void_p), C.byref(nf), idx.ctypes.data_as(C.c_void_p), C.byref(ni))
    assert np.array_equal(idx, g["bezier_idx"])
    assert np.array_equal(v.reshape(-1, 4), g["bezier"])     # same float operations: bit exact
    assert len(idx) == 400 * 8                                # N points -> N-1 segments per strand


def test_hair_scene_bounds_match_embree(hair_host, built, hair_file):
    g = golden("hair_scene.npz")
    f = hair_host.flat()
    assert np.allclose(f.bmin, g["bmin"], rtol=0, atol=1e-6) and np.allclose(f.bmax, g["bmax"], rtol=0, atol=1e-6)
    only = pb.Scene([hair_file], commit_to_device=False).flat()
    b = golden("hair_bounds.npz")
    assert np.allclose(only.bmin, b["bmin"], rtol=0, atol=1e-6) and np.allclose(only.bmax, b["bmax"], rtol=0, atol=1e-6)


def test_rejects_short_strands(built, tmp_path):
    """every strand needs >= 3 vertices or the whole file yields no curves (reference curve-util.cc:104-106)"""
    path = str(tmp_path / "short.hair")
    scenes.write_cyhair(path, n_strands=10, n_points=3, seed=1)
    ok = pb.Scene([scenes.light_stage(), path], commit_to_device=False).flat()
    assert len(ok.curve_prim) == 20
    import struct
    raw = bytearray(open(path, "rb").read())
    raw[16:20] = struct.pack("<I", 1)     # default_segments = 1 -> 2 points per strand
    raw[8:12] = struct.pack("<I", 20)
    open(path, "wb").write(bytes(raw[:128 + 20 * 12]))
    bad = pb.Scene([scenes.light_stage(), path], commit_to_device=False).flat()
    assert len(bad.curve_prim) == 0


def test_quad_and_polygon_triangulation(built, tmp_path):
    path = str(tmp_path / "quad.obj")
    with open(path, "w") as f:
        f.write("o light_q\nv 0 0 0\nv 2 0 0\nv 2 1 0\nv 0 1 0\nv 1 2 0\nf 1 2 3 4\nf 1 2 3 5 4\n")
    fl = pb.Scene([path], commit_to_device=False).flat()
    assert len(fl.tri_prim) == 2 + 3
    # both diagonals have the same length here -> the reference picks [0,1,3],[1,2,3] (sqr02 < sqr13 is false)
    assert fl.vidx[:2].tolist() == [[0, 1, 3], [1, 2, 3]]
    assert np.all(fl.tri_material == 0xFFFFFFFF)          # no usemtl -> no material -> paths are absorbed
    assert len(fl.light_probability) == 1                 # name starts with "light"


def test_binary_scene_cache_round_trip(built, tmp_path, monkeypatch):
    """PBRLAB_SCENE_CACHE=1: the second load of an OBJ comes from `<file>.pbrcache` and yields the same
    flat scene, array for array; touching the OBJ invalidates the cache"""
    import shutil, os, time
    src = scenes.cornell()
    obj = str(tmp_path / "cornell.obj")
    shutil.copy(src, obj)
    shutil.copy(os.path.splitext(src)[0] + ".mtl", str(tmp_path / os.path.basename(os.path.splitext(src)[0] + ".mtl")))
    plain = pb.Scene([obj], commit_to_device=False).flat()
    assert not os.path.exists(obj + ".pbrcache")
    monkeypatch.setenv("PBRLAB_SCENE_CACHE", "1")
    first = pb.Scene([obj], commit_to_device=False).flat()
    assert os.path.exists(obj + ".pbrcache")
    t = time.time()
    cached = pb.Scene([obj], commit_to_device=False).flat()
    for name in ("verts", "normals", "texcoords", "vidx", "nidx", "tidx", "tri_material", "tri_instance", "tri_prim",
                 "materials", "prim_cdf", "light_cdf", "bmin", "bmax"):
        a, b, c = getattr(plain, name), getattr(first, name), getattr(cached, name)
        assert np.array_equal(a, b) and np.array_equal(a, c), name
    # a newer OBJ is parsed again (and the cache rewritten)
    stamp = os.path.getmtime(obj + ".pbrcache")
    os.utime(obj, (time.time() + 5, time.time() + 5))
    again = pb.Scene([obj], commit_to_device=False).flat()
    assert np.array_equal(again.vidx, plain.vidx)
    assert os.path.getmtime(obj + ".pbrcache") >= stamp
