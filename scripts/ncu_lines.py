"""Developer tool: per-CUDA-source-line instruction counts / stall samples of one kernel from an .ncu-rep
(`ncu --set full --import-source on`), read here without a GPU.
usage: python scripts/ncu_lines.py gpurun_out/x.ncu-rep [top N] [kernel regex]"""
import csv, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
kern = ["-k", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"] + kern, capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname, hdr, agg = None, None, {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 10 and r[0] == "Line No":
        hdr = r
        ii, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif len(r) > 10 and r[0] != "" and hdr:
        try:
            key = (fname, int(r[0]), r[1].strip()[:100])
            agg[key] = (agg.get(key, (0, 0))[0] + int(r[ii]), agg.get(key, (0, 0))[1] + int(r[ss]))
        except ValueError:
            pass
tot_i, tot_s = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print(f"total warp instructions {tot_i}, samples {tot_s}")
for (f, ln, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*n/tot_i:5.1f}% inst {100*s/max(tot_s,1):5.1f}% smp  {f}:{ln:<4} {src}")
