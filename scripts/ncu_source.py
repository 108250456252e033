"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line: instructions executed
and stall samples per line, per kernel."""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    i = 0
    kernels = []
    cur = None
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1]
        elif r[0] == "Function Name":
            cur = {"name": r[1], "file": cur_file, "lines": defaultdict(lambda: [0, 0, 0, ""]), "total": [0, 0]}
            kernels.append(cur)
        elif r[0] == "Line No":
            hdr = r
            idx = {k: j for j, k in enumerate(hdr)}
        elif cur is not None and hdr is not None and len(r) >= len(hdr) - 2:
            try:
                ln = r[idx["Line No"]]
                inst = int(r[idx["Instructions Executed"]] or 0)
                samp = int(r[idx["# Samples"]] or 0)
                thr = int(r[idx["Thread Instructions Executed"]] or 0)
            except (ValueError, KeyError):
                continue
            if not ln.strip():
                continue  # SASS rows of the combined view: already counted in their CUDA line
            key = (cur["file"].split("/")[-1], ln)
            e = cur["lines"][key]
            e[0] += inst; e[1] += samp; e[2] += thr
            if not e[3]:
                e[3] = r[1][:110]
            cur["total"][0] += inst; cur["total"][1] += samp
    # merge sections of the same kernel (one per file)
    merged = {}
    for k in kernels:
        m = merged.setdefault(k["name"], {"lines": {}, "total": [0, 0]})
        m["lines"].update(k["lines"])
        m["total"][0] += k["total"][0]; m["total"][1] += k["total"][1]
    for name, m in merged.items():
        print("=" * 100)
        print(name[:150], " instr:", m["total"][0], " samples:", m["total"][1])
        for (f, ln), e in sorted(m["lines"].items(), key=lambda x: -x[1][0])[:top]:
            if e[0] == 0 and e[1] == 0:
                continue
            print(f"{f}:{ln:>5s} inst {e[0]:>12d} ({100*e[0]/max(m['total'][0],1):5.1f}%) samples {e[1]:>7d} ({100*e[1]/max(m['total'][1],1):5.1f}%) thr/inst {e[2]/max(e[0],1):5.1f} | {e[3]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
