"""Top source lines of each kernel in an `ncu --page source --csv --print-source cuda,sass` dump, by stall samples."""
import csv, sys
from collections import defaultdict
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 22; pat = sys.argv[3] if len(sys.argv) > 3 else ""
rows = list(csv.reader(open(path)))
cur = None; hdr = None; idx = None
data = defaultdict(lambda: defaultdict(lambda: [0, 0, ""]))
for r in rows:
    if not r: continue
    if r[0] == "File Path": f = r[1].split("/")[-1]
    elif r[0] == "Function Name": cur = r[1][:60]
    elif r[0] == "Line No": hdr = r; idx = {k: j for j, k in enumerate(hdr)}
    elif cur and hdr and len(r) >= len(hdr) - 2:
        try:
            ln = r[idx["Line No"]]; inst = int(r[idx["Instructions Executed"]] or 0); samp = int(r[idx["# Samples"]] or 0)
        except Exception: continue
        if not ln.strip(): continue
        e = data[cur][(f, ln)]; e[0] += inst; e[1] += samp; e[2] = e[2] or r[1][:100]
for k, d in data.items():
    if pat and pat not in k: continue
    tot = sum(v[1] for v in d.values()); ti = sum(v[0] for v in d.values())
    print("=====", k, "samples", tot, "inst", ti)
    for (f, ln), v in sorted(d.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{f}:{ln:>4} samp {v[1]:6d} {100*v[1]/tot:5.1f}%  inst {v[0]:9d} {100*v[0]/ti:5.1f}% | {v[2]}")
