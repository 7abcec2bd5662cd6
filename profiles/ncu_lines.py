"""Per-CUDA-line summary of an ncu report captured with --import-source on (not a product path):
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv ; python profiles/ncu_lines.py x.csv [top]
Sums 'Instructions Executed' and '# Samples' (warp stall samples) of the SASS rows under each source line and prints the
heaviest lines with their dominant stall reasons."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr, fname = None, ""
inst, samp, text = defaultdict(int), defaultdict(int), {}
stalls = defaultdict(lambda: defaultdict(int))
cur = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
        st_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0].strip():
        cur = (fname, int(r[0]))
        text[cur] = r[1]
    if cur is None or not r[ci].isdigit():
        continue
    inst[cur] += int(r[ci])
    samp[cur] += int(r[cs]) if r[cs].isdigit() else 0
    for i, h in st_cols:
        if r[i].isdigit():
            stalls[cur][h] += int(r[i])
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total warp instructions {ti}, stall samples {ts}")
for k in sorted(inst, key=lambda k: -samp[k])[:top]:
    s = sorted(stalls[k].items(), key=lambda kv: -kv[1])[:3]
    print(f"{k[0][:22]:22s}:{k[1]:4d} {100 * inst[k] / max(ti, 1):5.1f}% inst {100 * samp[k] / max(ts, 1):5.1f}% smp  "
          f"{' '.join(f'{a[6:]}={b}' for a, b in s if b):40s} | {text[k].strip()[:100]}")
