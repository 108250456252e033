"""Per CUDA source line of an `ncu --page source --csv --print-source cuda,sass` dump: L2 sectors requested by global
accesses, L1 tag requests, stall samples -- which lines of a kernel generate its memory traffic."""
import csv
import sys
from collections import defaultdict


def main(path, top=30):
    cur_file, hdr, idx = None, None, None
    agg = defaultdict(lambda: [0, 0, 0, 0, ""])
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            idx = {k: j for j, k in enumerate(hdr)}
        elif hdr and len(r) >= len(hdr) - 2 and r[0].strip():
            try:
                e = agg[(cur_file, int(r[0]))]
                e[0] += int(r[idx["L2 Theoretical Sectors Global"]] or 0)
                e[1] += int(r[idx["L1 Tag Requests Global"]] or 0)
                e[2] += int(r[idx["# Samples"]] or 0)
                e[3] += int(r[idx["Instructions Executed"]] or 0)
                e[4] = r[1][:100]
            except (ValueError, KeyError):
                pass
    tot = [sum(e[k] for e in agg.values()) for k in range(4)]
    print(f"total: L2 sectors {tot[0]:,}  L1 tag requests {tot[1]:,}  samples {tot[2]:,}  instructions {tot[3]:,}")
    for (f, ln), e in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        if e[0]:
            print(f"{f}:{ln:<5d} L2sect {e[0]:>14,} ({100 * e[0] / max(tot[0], 1):5.1f}%) tagreq {e[1]:>13,} samples {e[2]:>7,} ({100 * e[2] / max(tot[2], 1):4.1f}%) | {e[4]}")
    print("-- by stall samples")
    for (f, ln), e in sorted(agg.items(), key=lambda x: -x[1][2])[:top]:
        print(f"{f}:{ln:<5d} samples {e[2]:>7,} ({100 * e[2] / max(tot[2], 1):4.1f}%) inst {e[3]:>13,} | {e[4]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
