"""Summarise an .ncu-rep (read here, no GPU needed) into markdown: python scripts/ncu_summary.py REP [title]
Per captured kernel: duration, registers, occupancy, issue / pipe utilisation, threads per instruction, DRAM bytes,
cache hit rates, top stall reasons; plus the hottest SASS segments of every kernel (instruction share, active lanes)."""
import csv, io, subprocess, sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
M = [("gpu__time_duration.sum", "duration"), ("launch__registers_per_thread", "regs/thread"),
     ("launch__grid_size", "grid"), ("launch__block_size", "block"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
     ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction (of 32)"),
     ("smsp__inst_executed.sum", "warp instructions"),
     ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
     ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
     ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
     ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
     ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
     ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
     ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
     ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
     ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
     ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
     ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle")]
ki = hdr.index("Kernel Name")
names = [d[ki].split("(")[0].replace("void ", "").replace("pbr::", "") for d in data]
print("# %s\n" % title)
print("`ncu --set full --clock-control none --import-source on`, one launch per kernel out of the middle of a frame "
      "(cold-cache, serialised: compare shares, not absolutes).\n")
print("| metric | " + " | ".join(names) + " |")
print("|---|" + "---|" * len(names))
for key, label in M:
    if key not in hdr:
        continue
    i = hdr.index(key)
    vals = []
    for d in data:
        v = d[i].replace(",", "")
        try:
            f = float(v)
            v = ("%.0f" % f) if f >= 1000 else ("%.2f" % f)
        except ValueError:
            pass
        vals.append(v + (" " + units[i] if units[i] and units[i] not in ("%", "inst") else ""))
    print("| %s | %s |" % (label, " | ".join(vals)))

print("\n## Hot SASS segments (runs of instructions executed equally often)\n")
for d, name in zip(data, names):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + name.split("<")[0]],
                         capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(src)))
    if len(r) < 3 or "Instructions Executed" not in r[1]:
        continue
    h = r[1]
    ie, at, sm = h.index("Instructions Executed"), h.index("Avg. Threads Executed"), h.index("# Samples")
    rows_ = [x for x in r[2:] if len(x) > at and x[ie].isdigit()]
    if len(rows_) > 2 and rows_[0][1] == rows_[len(rows_) // 2][1]:
        rows_ = rows_[:len(rows_) // 2]
    tot = sum(int(x[ie]) for x in rows_) or 1
    tots = sum(int(x[sm]) for x in rows_) or 1
    seg = []
    for k, x in enumerate(rows_):
        n, a, s = int(x[ie]), float(x[at]), int(x[sm])
        if seg and seg[-1][1] == n and abs(seg[-1][2] - a) < 1e-6:
            seg[-1][3] += 1; seg[-1][4] += s; seg[-1][6] = x[1].strip()
        else:
            seg.append([k, n, a, 1, s, x[1].strip(), x[1].strip()])
    print("### %s (%d SASS instructions)\n" % (name, len(rows_)))
    print("| first..last instruction | length | executed (warp level) | share of instructions | active lanes | share of samples |")
    print("|---|---|---|---|---|---|")
    for k, n, a, ln, s, first, last in seg:
        if n * ln > tot * 0.02 or s > tots * 0.03:
            print("| `%s` .. `%s` | %d | %d | %.1f %% | %.1f | %.1f %% |" % (first[:32], last[:32], ln, n, 100.0 * n * ln / tot, a, 100.0 * s / tots))
    print()
