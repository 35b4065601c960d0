"""Summarise an .ncu-rep here (no GPU needed): key raw metrics per kernel + hottest SASS regions.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring] [--sass]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    if filt and filt not in r[ki]:
        continue
    print("==", r[ki][:70])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"   {w:75s} {r[i]:>16s} {units[i]}")
if "--sass" in sys.argv:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["-k", f"regex:{filt}"] if filt else []),
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
    h = rows[hi]
    si, ii, sm, ti = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
    data = []
    for r in rows[hi + 1:]:
        if len(r) <= ti or not r[ii].isdigit():
            break
        data.append(r)
    tot = sum(int(r[ii]) for r in data) or 1
    tots = sum(int(r[sm]) for r in data) or 1
    print("SASS instructions:", len(data), "executed warp-inst", tot, "samples", tots)
    blk = int(sys.argv[sys.argv.index("--sass") + 1]) if len(sys.argv) > sys.argv.index("--sass") + 1 else 64
    for s in range(0, len(data), blk):
        seg = data[s:s + blk]
        n = sum(int(r[ii]) for r in seg); ss = sum(int(r[sm]) for r in seg); th = sum(int(r[ti]) for r in seg)
        ops = {}
        for r in seg:
            t = r[si].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
        if n / tot > 0.004 or ss / tots > 0.004:
            print(f"{s:5d} inst {n / tot * 100:5.1f}% samp {ss / tots * 100:5.1f}% lanes {th / max(n, 1):4.1f} {top}")
