"""Attribute the executed instructions and stall samples of one kernel in an .ncu-rep to CUDA source lines.

    python scripts/ncu_lines.py REP KERNEL_REGEX [LIB.so] [--top N] [--outer]

ncu's CSV source page carries per-SASS-instruction counters but no line numbers; nvdisasm -g prints the line (and the
inlining chain) of every SASS instruction of the cubin inside the .so.  Both list the kernel's instructions in the same
order, so the two are joined by position.  The .so must be the build that was profiled.  Output: per source line (the
innermost inlined location by default, the outermost frame inside the kernel's file with --outer), share of executed
warp instructions, mean active lanes, share of stall samples.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_rows(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                          "regex:" + kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    kernels = []
    cur = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r and r[0].startswith("0x"):
            cur["rows"].append(r)
    return kernels


def disasm(lib):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    text = ""
    for f in sorted(os.listdir(tmp)):
        if f.endswith(".cubin"):
            text += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    return text


def function_lines(text, mangled_part):
    """list of (line location chain) per instruction of the first function whose section name contains mangled_part"""
    funcs = {}
    cur = None
    loc = None
    chain = []
    for ln in text.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            cur = m.group(1)
            if funcs.get(cur):       # a second (empty) cubin of the same .so repeats the section names
                cur = None
                continue
            funcs[cur] = []
            loc = None
            chain = []
            continue
        if cur is None:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            inl = "inlined at" in m.group(3)
            entry = (os.path.basename(m.group(1)), int(m.group(2)))
            if inl:
                chain = [entry]
                mm = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
                for f_, l_ in mm:
                    chain.append((os.path.basename(f_), int(l_)))
            else:
                chain = [entry]
            loc = list(chain)
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if re.match(r"\s*/\*[0-9a-f]{4}\*/", ln):
            funcs[cur].append(loc)
    for k, v in funcs.items():
        if mangled_part in k:
            return k, v
    return None, None


def main():
    argv = sys.argv[1:]
    top = 40
    if "--top" in argv:
        i = argv.index("--top")
        top = int(argv[i + 1])
        del argv[i:i + 2]
    args = [a for a in argv if not a.startswith("--")]
    rep, kernel = args[0], args[1]
    lib = args[2] if len(args) > 2 else "pbrlab_b200/lib/libpbrgpu.so"
    outer = "--outer" in sys.argv
    text = disasm(lib)
    for k in sass_rows(rep, kernel):
        hdr = k["hdr"]
        ie, te, sm = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        short = k["name"].replace("void ", "").split("(pbr::")[0].replace("(bool)", "")
        base = re.sub(r"<.*", "", short).split("::")[-1]
        tmpl = re.search(r"<\(bool\)(\d)", k["name"])
        # mangled: kernel base name + template bool args
        cands = [base]
        name, locs = None, None
        secs = re.findall(r"\.section\s+\.text\.(\S+?),", text)
        for s in secs:
            if base in s:
                if tmpl and ("ILb%sE" % tmpl.group(1)) not in s and "ILb" in s:
                    continue
                name, locs = function_lines(text, s)
                if locs is not None and len(locs) == len(k["rows"]):
                    break
        if locs is None or len(locs) != len(k["rows"]):
            print("## %s: cannot align (%s SASS rows in report, %s in cubin)" % (short, len(k["rows"]), None if locs is None else len(locs)))
            continue
        agg = defaultdict(lambda: [0, 0, 0])
        tot_i = tot_s = 0
        for r, loc in zip(k["rows"], locs):
            i, t, s = int(r[ie] or 0), int(r[te] or 0), int(r[sm] or 0)
            if loc is None:
                key = ("?", 0)
            else:
                key = loc[-1] if outer else loc[0]
            a = agg[key]
            a[0] += i; a[1] += t; a[2] += s
            tot_i += i; tot_s += s
        print("## %s — %d SASS instructions, %d warp instructions executed, %d samples" % (short, len(locs), tot_i, tot_s))
        print("| file:line | % of warp instructions | active lanes | % of samples |")
        print("|---|---|---|---|")
        for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            print("| %s:%d | %.1f | %.1f | %.1f |" % (key[0], key[1], 100.0 * a[0] / max(tot_i, 1), a[1] / max(a[0], 1),
                                                   100.0 * a[2] / max(tot_s, 1)))
        # coarse groups: traversal pieces by line range, everything else by file
        def group(key):
            f, l = key
            if f == "traverse.cuh":
                if 250 <= l <= 340: return "node test (NodeIntersect)"
                if 30 <= l <= 56: return "triangle test"
                if 57 <= l <= 249: return "curve tests"
                return "traverse.cuh other"
            if f == "trav_engine.cuh":
                if 62 <= l <= 101: return "node step (fetch, stack)"
                if 116 <= l <= 132: return "triangle step"
                if 133 <= l <= 211: return "curve steps"
                if 226 <= l: return "engine loop (ballots, phase control)"
                return "TravBegin/Advance"
            return f
        g = defaultdict(lambda: [0, 0, 0])
        for key, a in agg.items():
            x = g[group(key)]
            x[0] += a[0]; x[1] += a[1]; x[2] += a[2]
        print("| group | % of warp instructions | active lanes | % of samples |")
        print("|---|---|---|---|")
        for key, a in sorted(g.items(), key=lambda kv: -kv[1][0]):
            print("| %s | %.1f | %.1f | %.1f |" % (key, 100.0 * a[0] / max(tot_i, 1), a[1] / max(a[0], 1),
                                                100.0 * a[2] / max(tot_s, 1)))
        print()


if __name__ == "__main__":
    main()
