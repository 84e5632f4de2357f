"""Summarise an ncu report (raw metrics + SASS hot spots) into a markdown file under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.md "title" [steps_per_launch]
"""
import csv
import io
import subprocess
import sys

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
steps = float(sys.argv[4]) if len(sys.argv) > 4 else None

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
STALLS = "smsp__average_warps_issue_stalled_"


def ncu(page):
    r = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True)
    return list(csv.reader(io.StringIO(r.stdout)))


raw = ncu("raw")
hdr, units, vals = raw[0], raw[1], raw[2]
lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none, one launch)", "", "## raw metrics", "",
         "| metric | unit | value |", "|---|---|---|"]
name_i = hdr.index("Kernel Name")
lines.insert(2, f"kernel: `{vals[name_i]}`")
for i, h in enumerate(hdr):
    if h in KEYS or (h.startswith(STALLS) and h.endswith("_per_issue_active.ratio")):
        lines.append(f"| {h} | {units[i]} | {vals[i]} |")

src = ncu("source")
sh, data = src[1], src[2:]
isrc, isamp, iex = sh.index("Source"), sh.index("# Samples"), sh.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(sh) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp] or 0) for r in data)
totex = sum(int(r[iex] or 0) for r in data)
lines += ["", "## SASS hot spots (>= 0.8 % of stall samples)", "",
          f"total samples {tot}, warp instructions executed {totex}" + (f" = {totex / steps:.1f} per warp step" if steps else ""), "",
          "| # | samples % | exec/step | SASS | top stall |", "|---|---|---|---|---|"]
agg = {}
for k, r in enumerate(data):
    s, ex = int(r[isamp] or 0), int(r[iex] or 0)
    for i in stall_cols:
        agg[sh[i]] = agg.get(sh[i], 0) + int(r[i] or 0)
    if s >= 0.008 * tot:
        top = max(((int(r[i] or 0), sh[i][6:]) for i in stall_cols))
        per = f"{ex / steps:.2f}" if steps else str(ex)
        lines.append(f"| {k} | {100 * s / tot:.1f} | {per} | `{r[isrc].strip()[:70]}` | {top[1]} |")
lines += ["", "## stall samples by reason", "", "| reason | % |", "|---|---|"]
for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    lines.append(f"| {h[6:]} | {100 * v / max(tot, 1):.1f} |")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
