"""ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch) of a bench.py run with two
builds -> profiles/traffic_r2.json: DRAM bytes and time per kernel of ONE build (the second), what bench.py reports as
`roofline.traffic`.  Only the library's own kernels between two launches of read_stats / repack (one build) are counted.

    python scripts/ncu_traffic.py gpurun_out/launches_<tag>.csv <workload> <scale> <n_gpus> > profiles/traffic_r2.json
"""
import collections
import csv
import re
import json
import sys

path, workload, scale, n_gpus = sys.argv[1], sys.argv[2], float(sys.argv[3]), int(sys.argv[4])
rows = list(csv.DictReader([l for l in open(path) if not l.startswith("==")]))
launches = collections.OrderedDict()  # launch id -> {name, metrics}
for r in rows:
    e = launches.setdefault(r["ID"], {"name": re.sub(r"^(void )?(alga::(<unnamed>::)?|CUB_[0-9]+_SM_[0-9]+::)?", "", r["Kernel Name"]), "m": {}})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        v = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0) * v  # -> ms
    else:
        v = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0) * v  # -> bytes
    e["m"][r["Metric Name"]] = v
seq = list(launches.values())
ours = ("repack_reads", "build_index", "seed_records", "fill_buckets", "chain_overflow", "csr_records", "csr_rows", "DeviceRadixSort",
        "phase1_", "phase2_", "rebuild_rows", "over_to_csr", "rows_to_csr", "scan_", "end_cursor",
        "scatter_", "split_pairs", "sort_rows", "sort_big_rows", "count_sources", "peek", "read_stats")
# a build starts with the seed records (sorted index: they come before the repack) or, without them, with the repack
first = "seed_records" if any(e["name"].startswith("seed_records") for e in seq) else "repack_reads"
starts = [i for i, e in enumerate(seq) if e["name"].startswith(first)]
if len(starts) < 2:
    raise SystemExit(f"expected two builds (two {first} launches) in the list")
build = [e for e in seq[starts[-1]:] if any(e["name"].startswith(o) for o in ours)]
per = collections.OrderedDict()
for e in build:
    key = e["name"].split("(")[0].split("<")[0]
    a = per.setdefault(key, {"launches": 0, "ms": 0.0, "dram_bytes": 0.0})
    a["launches"] += 1
    a["ms"] += e["m"].get("gpu__time_duration.sum", 0.0)
    a["dram_bytes"] += e["m"].get("dram__bytes_read.sum", 0.0) + e["m"].get("dram__bytes_write.sum", 0.0)
tot_b = sum(a["dram_bytes"] for a in per.values())
tot_ms = sum(a["ms"] for a in per.values())
out = {"workload": workload, "scale": scale, "n_gpus": n_gpus, "source": path.split("/")[-1],
       "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; second of two builds",
       "pipeline_dram_bytes_per_step": tot_b, "pipeline_kernel_ms_serialised": tot_ms,
       "phase1": sum(a["dram_bytes"] for k, a in per.items() if k.startswith("phase1_")),
       "phase2": sum(a["dram_bytes"] for k, a in per.items() if k.startswith("phase2_")),
       "kernels": {k: {"launches": a["launches"], "ms": round(a["ms"], 4), "dram_bytes": a["dram_bytes"],
                       "share_of_time": round(a["ms"] / tot_ms, 4) if tot_ms else None} for k, a in per.items()}}
print(json.dumps(out, indent=1))
