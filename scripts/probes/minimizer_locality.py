"""CPU-only groundwork for a seed index with locality: how many DISTINCT buckets do the probes of one
read touch if the bucket is chosen by the minimizer of the K-window instead of by the window itself, and how full do those
buckets get?  Synthetic reads of BASELINE config 2's shape (2 x 150 bp trimmed to 144, 50x, both strands), K = 32.

    python scripts/probes/minimizer_locality.py [genome_size] [m ...]
"""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/scripts/", 1)[0])
from alga_b200 import synth  # noqa: E402

K, LEN, LMIN, RS = 32, 144, 82, 116


def mmer_hash(codes: np.ndarray, m: int) -> np.ndarray:
    """(n, LEN) codes -> (n, LEN - m + 1) scrambled m-mer values."""
    n, ln = codes.shape
    v = np.zeros((n, ln - m + 1), np.uint64)
    for j in range(m):
        v = (v << np.uint64(2)) | codes[:, j : ln - m + 1 + j].astype(np.uint64)
    v = v * np.uint64(0x9E3779B97F4A7C15)
    return v ^ (v >> np.uint64(29))


def window_minimizer(h: np.ndarray, w: int) -> np.ndarray:
    """sliding minimum over w consecutive m-mers: (n, P) -> (n, P - w + 1)"""
    out = h[:, : h.shape[1] - w + 1].copy()
    for j in range(1, w):
        np.minimum(out, h[:, j : h.shape[1] - w + 1 + j], out=out)
    return out


def distinct_per_row(a: np.ndarray) -> np.ndarray:
    s = np.sort(a, axis=1)
    return 1 + (s[:, 1:] != s[:, :-1]).sum(axis=1)


def main():
    genome = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
    ms = [int(x) for x in sys.argv[2:]] or [12, 16, 20]
    rng = np.random.default_rng(2)
    g = synth.make_genome(genome, rng)
    m1, m2 = synth.sample_paired_end(g, 150, 50, rng, 0.0)
    nodes = synth.remove_duplicate_nodes(synth.strand_nodes(m1, m2))  # both strands, trimmed, duplicates removed
    n = nodes.shape[0]
    print(f"genome {genome}, {n} strand-reads of {LEN} nt, K = {K}")
    print("today: bucket = hash(window): phase 2 probes 29 windows -> 29 buckets, phase 1 probes ~11 of 34 -> ~11 buckets, "
          "table entries per bucket: mean 2 of 8 slots by construction")
    for m in ms:
        w = K - m + 1
        h = mmer_hash(nodes, m)
        mins = window_minimizer(h, w)  # minimizer of the K-window that starts at each position 0 .. LEN-K
        # phase 2 (per target): windows that end at L = RS .. LEN, i.e. start at L - K
        p2 = mins[:, RS - K : LEN - K + 1]
        # phase 1 (per source): windows that start at LEN - L for L = RS-1 down to LMIN; the kernel stops after 3 hits (~11 lengths)
        p1_all = mins[:, LEN - (RS - 1) : LEN - LMIN + 1]
        p1_11 = p1_all[:, :11]
        d2, d1a, d1 = distinct_per_row(p2), distinct_per_row(p1_all), distinct_per_row(p1_11)
        # table entries: suffix table = minimizer of the last K window, prefix table = of the first K window
        for name, col in (("suffix", mins[:, LEN - K]), ("prefix", mins[:, 0])):
            _, cnt = np.unique(col, return_counts=True)
            over8 = (cnt > 8).sum() / cnt.shape[0]
            print(f"  m = {m:2d} (w = {w:2d})  {name} table: {cnt.shape[0]} distinct minimizers for {n} entries, entries per minimizer "
                  f"mean {cnt.mean():.2f}  p99 {np.percentile(cnt, 99):.0f}  max {cnt.max()}  > 8 slots: {100 * over8:.2f} % of the runs")
            # the table as the library sizes it (n / 2 buckets of 8 slots, open addressing): what a probe meets in its FIRST bucket
            for load in (2, 1):
                nb = max(64, n // load)
                bucket = ((col * np.uint64(0xC2B2AE3D27D4EB4F)) >> np.uint64(32)) * np.uint64(nb) >> np.uint64(32)
                fill = np.bincount(bucket.astype(np.int64), minlength=nb)
                full = (fill >= 8)
                # entries that do not fit their home bucket spill into the chain; probes whose home bucket is full walk it
                entries_in_full = fill[full].sum() / n
                print(f"      {nb} buckets ({load} entries per bucket on average): home buckets with >= 8 entries {100 * full.mean():.2f} %, "
                      f"holding {100 * entries_in_full:.1f} % of the entries (window-hash buckets at the same size: "
                      f"{100 * (np.random.default_rng(0).poisson(load, nb) >= 8).mean():.2f} %)")
        print(f"  m = {m:2d}: distinct minimizers among the probes of a read: phase 2 (29 windows) mean {d2.mean():.2f} p99 "
              f"{np.percentile(d2, 99):.0f};  phase 1 first 11 windows mean {d1.mean():.2f}, all 34 mean {d1a.mean():.2f}")


if __name__ == "__main__":
    main()
