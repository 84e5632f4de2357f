"""Instruction mix of a profiled kernel from the SASS source page of an ncu report (CSV written on the GPU box by
tools/gpu_profile_r2.sh): executed warp instructions per 32-event warp step by opcode, the Philox share, the
split between the ALU pipe (two issue cycles per warp instruction) and the rest, and the pipe utilisations.

    python tools/ncu_mix.py gpurun_out/NAME_sass.csv gpurun_out/NAME_raw.csv EVENTS_PER_LAUNCH >> profiles/...md
"""
import collections
import csv
import re
import sys

sass, raw, events = sys.argv[1], sys.argv[2], float(sys.argv[3])
rows = list(csv.reader(open(sass)))
hdr, data = rows[1], rows[2:]
isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
steps = events / 32.0
tot = sum(int(r[iex] or 0) for r in data)
ALU = {"LOP3", "SHF", "ISETP", "LEA", "SEL", "IADD3", "VIADD", "PLOP3", "PRMT", "FLO", "BREV", "POPC", "VIMNMX", "MOV",
       "VIADDMNMX", "IABS", "LOP", "IADD", "VOTE", "SGXT", "BMSK", "R2P", "P2R"}
ops = collections.Counter()
philox = other_alu = 0.0
for r in data:
    e = int(r[iex] or 0) / steps
    if e <= 0:
        continue
    s = r[isrc].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", s)
    op = m.group(2) if m else s
    base = op.split(".")[0]
    key = ".".join(op.split(".")[:2]) if base in ("IMAD", "BAR", "ATOMS", "ATOMG", "LDS", "LDG", "LDC", "LDCU", "SYNCS", "RED") else base
    ops[key] += e
    if ("-0x2daee0ad" in s or "-0x326172a9" in s) or (base == "LOP3" and "0x96" in s and "UR" in s):
        philox += e
    elif base in ALU:
        other_alu += e
print(f"\n## instruction mix ({tot / steps:.1f} warp instructions per 32-event step; {tot:.4g} per launch, {events:.4g} events)\n")
print("| opcode | executed per step | pipe |")
print("|---|---|---|")
for op, c in ops.most_common(28):
    pipe = "ALU (2 issue cycles)" if op.split(".")[0] in ALU else ("FMA / IMAD" if op.startswith("IMAD") else "")
    print(f"| {op} | {c:.1f} | {pipe} |")
alu = sum(c for o, c in ops.items() if o.split(".")[0] in ALU)
print(f"\nALU-pipe instructions: {alu:.1f} per step ({100 * alu * steps / tot:.0f} %); Philox4x32-10 (IMAD.WIDE by the round constants + "
      f"three-input XORs with the round keys): {philox:.1f} per step ({100 * philox * steps / tot:.0f} %).")
r = list(csv.reader(open(raw)))
h, v = r[0], r[2]
want = ["sm__inst_executed.avg.per_cycle_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
print("\n| pipe / unit metric | value |")
print("|---|---|")
for w in want:
    if w in h:
        print(f"| {w} | {v[h.index(w)]} {r[1][h.index(w)]} |")
