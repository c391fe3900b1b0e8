#!/usr/bin/env python
"""Benchmark of the path-tracing hot path (BASELINE.json: "Msamples/s and Mrays/s at 1/2/4/8 B200 vs Embree CPU").

    python bench.py --gpus N --steps K --warmup W [--workload c1|c2|c3|c4|c5] [--spp S] [--impl reference]
    (N > 1: launched by torchrun, one rank per GPU)

A step is one Render() of the workload: every camera sample of the frame traced to termination.
  value  = Msamples/s of the whole job with the scene and the accumulators resident in HBM (pbrgpu_render_device;
           for N > 1 each rank renders its interleaved share of the samples and the accumulators are summed by ONE
           NCCL reduce), timed with CUDA events, max over ranks.
  e2e    = the same through the reference-facing call with HOST buffers: pbrlab::Render() (N = 1) /
           pbrgpu_render_device + reduce + device->host read (N > 1); the material table goes host->device every step
           (Render() re-uploads it, as the reference reads materials live) and the RenderLayer sums come back.
  roofline: the dominant kernel is the closest-hit traversal; achieved = algorithmic bytes per ray (SURVEY §8(d):
           64 + 80*ceil(log8(N/4)) + 4*S) x rays per launch / mean launch time (CUDA events on the launching stream,
           profiling mode of the library), against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline: the unmodified reference (oracle/_ref: pbrlab + Embree) timed on this box's host cores on a bounded
           sample of the same workload.
--impl reference runs only that CPU arm.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (description, width, height, spp per GPU, scene builder)
    "c1": ("cornellbox_suzanne_lucy 512x512 64spp PrincipledBSDF + area light", 512, 512, 64),
    "c2": ("cornellbox_suzanne_lucy 1920x1080 1024spp, Lucy random-walk SSS", 1920, 1080, 1024),
    "c3": ("synthetic CyHair 50k strands (1M segments) + light stage, 1920x1080 256spp, Principled Hair", 1920, 1080, 256),
    "c4": ("cornellbox_suzanne_lucy + synthetic CyHair 50k strands, 3840x2160 512spp", 3840, 2160, 512),
    "c5": ("synthetic displaced 20M-triangle OBJ, GGX + SSS materials, 3840x2160 1024spp", 3840, 2160, 1024),
}


def scene_files(workload):
    from pbrlab_b200 import scenes
    if workload in ("c1", "c2"):
        return [scenes.cornell()], 362620, 0
    if workload == "c3":
        return [scenes.light_stage(), scenes.cyhair(50000, 21, center=(-2.5, 3.5, 0.0), radius=1.2, length=2.5,
                                                    thickness=0.008)], 6, 1000000
    if workload == "c4":
        return [scenes.cornell(), scenes.cyhair(50000, 21, center=(-2.5, 6.0, 0.0), radius=1.2, length=2.5,
                                                thickness=0.008)], 362620, 1000000
    if workload == "c5":
        return [scenes.displaced(20_000_000)], 19_920_012, 0
    raise SystemExit("unknown workload " + workload)


def algorithmic_bytes_per_ray(ntris, nsegs):
    """ray in (32) + hit out (32) + 80 B per level of an 8-wide tree with 4-primitive leaves + one leaf"""
    if nsegs and not ntris > 1000:
        return 64 + 80 * math.ceil(math.log(nsegs / 4.0, 8)) + 4 * 64
    b_tri = 64 + 80 * math.ceil(math.log(max(ntris, 8) / 4.0, 8)) + 4 * 48
    if not nsegs:
        return b_tri
    b_cur = 64 + 80 * math.ceil(math.log(nsegs / 4.0, 8)) + 4 * 64
    return (b_tri + b_cur) // 2


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.rows = []
        self.stop = False
        self.index = index

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def cpu_render(files, w, h, spp, warm_spp=1):
    """pbrlab::Render() on the host cores: the compiled unmodified reference (oracle/_ref) when it travelled with the
    snapshot, else the restated oracle (oracle/libpbr_oracle.so).  Returns (seconds, cores, kind)."""
    import refbind
    if refbind.available():
        R = refbind.RefLib()
        S = R.scene(files)
        if warm_spp:
            S.render(w, h, warm_spp)
        _, _, sec = S.render(w, h, spp)
        return sec, R.num_threads(), "reference"
    import oraclebind
    import pbrlab_b200 as pb
    if not oraclebind.available():
        return None, 0, "unavailable"
    host = pb.Scene(files, commit_to_device=False)
    O = oraclebind.Oracle(host.flat())
    if warm_spp:
        O.render(w, h, warm_spp)
    _, _, sec, _ = O.render(w, h, spp)
    return sec, os.cpu_count() or 1, "port"


def reference_arm(args, rank):
    """times the reference's own CPU implementation of the path on the host cores (all hardware threads, as
    pbrlab::Render() always does) on a bounded sample of the workload"""
    desc, w, h, spp = WORKLOADS[args.workload]
    if rank != 0:
        return
    files, _, _ = scene_files(args.workload)
    sample_spp = args.ref_spp
    secs = []
    cores, kind = 0, "unavailable"
    for i in range(args.warmup + args.steps):
        sec, cores, kind = cpu_render(files, w, h, 1 if i < args.warmup else sample_spp, warm_spp=0)
        if sec is None:
            emit({"impl": "reference", "unavailable": "neither oracle/_ref nor oracle/libpbr_oracle.so is present"})
            return
        if i >= args.warmup:
            secs.append(sec)
    t = sum(secs) / len(secs)
    v = w * h * sample_spp / t * 1e-6
    sample = "%dx%d, %d of %d spp per step (throughput is spp-independent)" % (w, h, sample_spp, spp)
    line = {"metric": "Msamples/s", "value": v, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": desc, "sample": sample},
            "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """The host C++ code mirrors the reference's progress prints on stdout ("Load obj file ...", "finish pass N").
    The contract is ONE JSON line on stdout: everything else written to fd 1, by Python or by native code, goes to
    stderr from here on; emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel per GPU")
    ap.add_argument("--ref-spp", type=int, default=16, help="spp of the bounded CPU-reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    import pbrlab_b200 as pb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    desc, w, h, spp_per_gpu = WORKLOADS[args.workload]
    if args.spp:
        spp_per_gpu = args.spp
    files, ntris, nsegs = scene_files(args.workload)
    t0 = time.time()
    scene = pb.Scene(files, device_ids=[local_rank])
    ctx = scene.context()
    commit_s = time.time() - t0
    npix = w * h
    spp_total = spp_per_gpu * world          # weak scaling: every GPU renders spp_per_gpu samples of every pixel
    d_rgba = torch.zeros(npix * 4, dtype=torch.float32, device="cuda")
    d_count = torch.zeros(npix, dtype=torch.int32, device="cuda")
    h_rgba = torch.empty(npix * 4, dtype=torch.float32).pin_memory()
    h_count = torch.empty(npix, dtype=torch.int32).pin_memory()
    flat = scene.flat()
    mat_words = flat.materials
    ntris, nsegs = int(len(flat.vidx)), int(len(flat.curve_first))      # what was actually loaded
    seed = 1234567890

    def step_device():
        """HBM-resident step: render this rank's samples into device buffers, one NCCL reduce to rank 0"""
        ctx.render_device(w, h, spp_total, d_rgba.data_ptr(), d_count.data_ptr(), seed=seed, sample_offset=rank,
                          sample_stride=world)
        if world > 1:
            dist.reduce(d_rgba, 0, op=dist.ReduceOp.SUM)
            dist.reduce(d_count, 0, op=dist.ReduceOp.SUM)

    def step_e2e():
        """host-to-host step"""
        if world == 1:
            return scene.render(w, h, spp_total, seed=seed)      # pbrlab::Render(): materials H2D, RenderLayer D2H
        ctx.set_materials(mat_words)
        step_device()
        if rank == 0:
            h_rgba.copy_(d_rgba, non_blocking=True)
            h_count.copy_(d_count, non_blocking=True)
        torch.cuda.synchronize()
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """CUDA events on the current stream around `steps` calls; max over ranks"""
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        launches = 0
        for _ in range(steps):
            fn()
            launches += ctx.stats()["kernel_launches"]
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item()), launches

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, launches = timed(step_device, args.steps)
    st = ctx.stats()
    rays_step = st["closest_rays"] + st["shadow_rays"] + st["sss_rays"]       # this rank, last step
    rays_t = torch.tensor([float(rays_step)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(rays_t)
    ms_e2e, _ = timed(step_e2e, args.steps)
    sampler.stop = True

    # roofline of the dominant kernel, measured live: one more step in profiling mode (CUDA events per kernel family)
    ctx.set_profiling(True)
    step_device()
    ps = ctx.stats()
    ctx.set_profiling(False)
    a_ray = algorithmic_bytes_per_ray(ntris, nsegs)
    fam = {"trace_closest": ps["trace_closest_ms"], "sss_walk": ps["sss_ms"], "shade": ps["shade_ms"],
           "trace_any": ps["trace_any_ms"], "regenerate": ps["regen_ms"]}
    n_launch = max(1, ps["trace_closest_launches"])
    dur_ms = ps["trace_closest_ms"] / n_launch
    rays_per_launch = ps["closest_rays"] / n_launch
    achieved = a_ray * rays_per_launch / (dur_ms * 1e-3) / 1e9 if dur_ms > 0 else 0.0
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = json.load(open(peaks_path))["hbm_gbs"]; peak_src = "measured (MEASURED_PEAKS.json)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
    # the L2-resident scenes are bounded by L2 gather bandwidth, which MEASURED_PEAKS.json does not hold:
    # node-sized (80 B) dependent random gathers over a 32 MB (L2) and a 4 GB (HBM) working set, 8 chains per thread
    gather_l2 = ctx.measure_gather(32 << 20, 4096, 8) if rank == 0 else None
    gather_hbm = ctx.measure_gather(4 << 30, 2048, 8) if rank == 0 else None
    # nodes visited / primitives tested per ray, counted by the traversal kernel itself on a batch of
    # 1 Mi camera rays + 1 Mi secondary rays leaving from what they hit, and the bytes those steps touch
    trav = None
    if rank == 0:
        import pbrlab_b200 as _pb
        rng = np.random.default_rng(7)
        bmin, bmax = ctx.bounds()
        hs = bmax[0] - bmin[0]; vs = bmax[1] - bmin[1]
        if hs > vs: vs = hs * h / w
        else: hs = vs * w / h
        eye = np.array([(bmax[0] + bmin[0]) * 0.5, (bmax[1] + bmin[1]) * 0.5, bmax[2] + hs * 0.5 * np.sqrt(3.0)], np.float32)
        nr = 1 << 20
        px = rng.random((nr, 2)).astype(np.float32)
        tgt = np.stack([eye[0] - hs * 0.5 + hs * px[:, 0], eye[1] + vs * 0.5 - vs * px[:, 1], np.full(nr, bmax[2], np.float32)], 1)
        d = tgt - eye; d /= np.linalg.norm(d, axis=1, keepdims=True)
        cam_rays = _pb.make_rays(np.tile(eye, (nr, 1)), d.astype(np.float32))
        hits = ctx.trace(cam_rays); st1 = ctx.stats()
        hit = hits["instance_id"] != 0xFFFFFFFF
        P = cam_rays["org"][hit] + hits["t"][hit, None] * cam_rays["dir"][hit]
        d2 = rng.normal(size=P.shape).astype(np.float32); d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
        ctx.trace(_pb.make_rays(P, d2, tmin=1e-3)); st2 = ctx.stats()
        prim_bytes = 64 if (nsegs and not ntris > 1000) else 48
        per = lambda st, n: {"nodes_per_ray": st["nodes_visited"] / max(n, 1), "prims_per_ray": st["prims_tested"] / max(n, 1),
                             "touched_bytes_per_ray": 64 + 80 * st["nodes_visited"] / max(n, 1) + prim_bytes * st["prims_tested"] / max(n, 1)}
        trav = {"camera_rays": per(st1, nr), "secondary_rays": per(st2, int(hit.sum()))}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload, {}).get("TraceClosestKernel_bytes_per_launch")

    if rank == 0:
        samples_step = npix * spp_total
        value = samples_step * args.steps / (ms * 1e-3) * 1e-6
        e2e_value = samples_step * args.steps / (ms_e2e * 1e-3) * 1e-6
        mrays = float(rays_t.item()) * args.steps / (ms * 1e-3) * 1e-6
        line = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "width": w, "height": h, "spp_per_gpu": spp_per_gpu, "spp_total": spp_total,
                       "split": "interleaved samples, scene replicated, one NCCL reduce per frame" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2: the path pool (~2 GB of slot lines) streams through every iteration between two visits of a slot",
                       "triangles": ntris, "curve_segments": nsegs,
                       "scene_commit_s": commit_s, "seed": seed},
            "Mrays_per_s": mrays, "rays_per_sample": float(rays_t.item()) / samples_step,
            "e2e": {"value": e2e_value, "unit": "Msamples/s",
                    "h2d_bytes_per_step": int(mat_words.nbytes), "d2h_bytes_per_step": int(npix * 20)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "TraceClosestKernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_ray": a_ray, "rays_per_launch": rays_per_launch,
                         "launch_ms": dur_ms, "kernel_family_ms_per_step": fam,
                         "gather_80B_l2_gbs": gather_l2, "gather_80B_hbm_gbs": gather_hbm,
                         "frac_of_l2_gather": (achieved / gather_l2) if gather_l2 else None,
                         "traversal_counters": trav},
            "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline:
            sec, cores, kind = cpu_render(files, w, h, args.ref_spp)
            sample = "%dx%d at %d spp (of %d), pbrlab::Render() of %s" % (
                w, h, args.ref_spp, spp_total, "oracle/_ref" if kind == "reference" else "oracle/pbr_oracle.cc")
            line["cpu_baseline"] = {"value": (w * h * args.ref_spp / sec * 1e-6) if sec else None, "unit": "Msamples/s",
                                    "cores": cores, "kind": kind, "sample": sample}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
