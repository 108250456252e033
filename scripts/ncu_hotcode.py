"""Instruction footprint of a kernel from an `ncu --page source --csv --print-source cuda,sass` dump: how many SASS
instructions exist, how many of them are HOT (executed at least `per` times, e.g. once per tile), and the source lines that
own most of the hot ones.  The SM's instruction cache holds 32 KB = 2048 instructions; a kernel whose warps loop over more
than that is bound by instruction fetch.

    python scripts/ncu_hotcode.py <dump.csv> <per> [top]
"""
import csv
import sys
from collections import defaultdict


def main(path, per, top=25):
    hdr = idx = cur_file = cur_line = None
    ins = {}
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr, idx = r, {k: j for j, k in enumerate(r)}
        elif hdr is None:
            continue
        elif r[0].strip():
            try:
                cur_line = int(r[0])
            except ValueError:
                pass
        elif len(r) > 3 and r[2].startswith("0x"):
            a = int(r[2], 16)
            try:
                ex = int(r[idx["Instructions Executed"]] or 0)
            except (ValueError, KeyError):
                ex = 0
            ins.setdefault(a, (ex, cur_file, cur_line, r[3].strip()))
    base = min(ins)
    hot = [a for a in ins if ins[a][0] >= per]
    lines = {(a - base) // 128 for a in hot}
    print(f"SASS instructions {len(ins)} ({len(ins) * 16 / 1024:.1f} KB); executed >= {per:.0f} times: {len(hot)} "
          f"({len(hot) * 16 / 1024:.1f} KB, in {len(lines)} 128-byte lines = {len(lines) / 8:.1f} KB)")
    by = defaultdict(lambda: [0, 0, 0])
    for a, (ex, f, l, _) in ins.items():
        e = by[(f, l)]
        e[0] += 1
        e[1] += 1 if ex >= per else 0
        e[2] += ex
    for (f, l), e in sorted(by.items(), key=lambda x: -x[1][1])[:top]:
        print(f"  {f}:{l}  instructions {e[0]:4d}  hot {e[1]:4d}  executed {e[2] / 1e6:9.1f} M")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 25)
