"""Development aid: print the SASS between two mbarrier arrives of the RING kernel (one unrolled step copy),
with the branch structure, to count the instructions of the common path by eye.
    python tools/sass_path.py /tmp/h.sass [copy index]"""
import re
import sys

lines = [l for l in open(sys.argv[1]) if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]
ins = []
for l in lines:
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?)\s*;", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
arr = [i for i, (a, t) in enumerate(ins) if "SYNCS.ARRIVE" in t]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lo, hi = arr[k], arr[k + 1]
for a, t in ins[lo:hi + 1]:
    print(f"{a:06x}  {t}")
print(len(ins[lo:hi + 1]), "static instructions between arrives", file=sys.stderr)
