"""Per-kernel-family figures of an ncu capture, for bench.py's roofline_families:

    python scripts/ncu_families.py REP.ncu-rep WORKLOAD [NOTE]   ->  merges into profiles/ncu_families.json

One entry per family (trace_closest, trace_any, sss_walk, shade) with the metrics north_star names: achieved DRAM
bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), L2 hit rate, lanes per instruction (warp execution
efficiency), issue-slot utilisation, FMA pipe utilisation.  When a family was captured several times (or has several
kernels, like shade = general + diffuse-only + hair + SSS exit) durations and bytes are summed per wavefront iteration
and the ratios are weighted by executed instructions.  ncu replays kernels cold and serialised: shares, not absolutes.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILY = [("TraceClosestKernel", "trace_closest"), ("TraceAnyKernel", "trace_any"), ("SssWalkKernel", "sss_walk"),
          ("ShadeSurfaceKernel", "shade"), ("ShadeHairKernel", "shade"), ("SssExitKernel", "shade")]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def to_ms(v, unit):
    return v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "s": 1e3, "second": 1e3, "nsecond": 1e-6}.get(unit, 1)


def main():
    rep, workload = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {k: i for i, k in enumerate(hdr)}
    ki = col["Kernel Name"]
    acc = {}
    for d in data:
        fam = next((f for key, f in FAMILY if key in d[ki]), None)
        if fam is None:
            continue
        g = lambda k: num(d[col[k]]) if k in col else 0.0
        u = lambda k: units[col[k]] if k in col else ""
        inst = g("smsp__inst_executed.sum")
        a = acc.setdefault(fam, {"kernels": {}, "inst": 0.0, "lanes": 0.0, "issue": 0.0, "fma": 0.0, "l2hit": 0.0,
                                 "warps": 0.0, "dur_w": 0.0})
        name = d[ki].split("(")[0].replace("void ", "").replace("pbr::", "") + ("<1>" if "(bool)1" in d[ki] else "<0>" if "(bool)0" in d[ki] else "")
        k = a["kernels"].setdefault(name, {"launches": 0, "ms": 0.0, "dram": 0.0})
        k["launches"] += 1
        k["ms"] += to_ms(g("gpu__time_duration.sum"), u("gpu__time_duration.sum"))
        k["dram"] += to_bytes(g("dram__bytes_read.sum"), u("dram__bytes_read.sum")) + to_bytes(g("dram__bytes_write.sum"), u("dram__bytes_write.sum"))
        dur = to_ms(g("gpu__time_duration.sum"), u("gpu__time_duration.sum"))
        a["inst"] += inst
        a["lanes"] += inst * g("smsp__thread_inst_executed_per_inst_executed.ratio")
        a["dur_w"] += dur
        a["issue"] += dur * g("smsp__issue_active.avg.pct_of_peak_sustained_active")
        a["fma"] += dur * g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")
        a["l2hit"] += dur * g("lts__t_sector_hit_rate.pct")
        a["warps"] += dur * g("sm__warps_active.avg.pct_of_peak_sustained_active")
    out = {}
    for fam, a in acc.items():
        ms = sum(k["ms"] / k["launches"] for k in a["kernels"].values())       # per wavefront iteration
        dram = sum(k["dram"] / k["launches"] for k in a["kernels"].values())
        out[fam] = {"dram_bytes_per_launch": dram, "ncu_launch_ms": ms,
                    "ncu_dram_gbs": dram / (ms * 1e-3) / 1e9 if ms else None,
                    "lanes_per_instruction": a["lanes"] / a["inst"] if a["inst"] else None,
                    "issue_slots_busy_pct": a["issue"] / a["dur_w"] if a["dur_w"] else None,
                    "fma_pipe_pct": a["fma"] / a["dur_w"] if a["dur_w"] else None,
                    "l2_hit_pct": a["l2hit"] / a["dur_w"] if a["dur_w"] else None,
                    "warps_active_pct": a["warps"] / a["dur_w"] if a["dur_w"] else None,
                    "ncu_kernels": sorted(a["kernels"])}
    path = os.path.join(ROOT, "profiles", "ncu_families.json")
    allf = json.load(open(path)) if os.path.exists(path) else {}
    allf[workload] = out
    src = allf.get("_source", {}) if isinstance(allf.get("_source"), dict) else {}
    src[workload] = (os.path.basename(rep) + (": " + note if note else ""))
    allf["_source"] = src
    json.dump(allf, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
