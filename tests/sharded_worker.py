"""Worker of tests/test_sharded_multi_gpu.py (launched with torch.distributed.run, one rank per GPU): the sharded
build of alga_b200.distributed.ShardedPrefSuf on real peer memory, gathered on rank 0 and compared with the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from alga_b200 import readset, synth  # noqa: E402
from alga_b200.distributed import ShardedPrefSuf  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
    w = synth.make_config("cfg2", scale=scale)
    rs = w.reads
    n = rs.n  # not a multiple of the world size in general: the last rank owns fewer reads
    W = int(rs.word_off[1] - rs.word_off[0])
    n_shard = ((n + world - 1) // world + 1) & ~1
    lo, hi = min(rank * n_shard, n), min((rank + 1) * n_shard, n)
    words = torch.from_numpy(rs.words.view(np.int32).reshape(n, W))
    sp = ShardedPrefSuf(w.params.min_overlap, w.params.rs_min_overlap, 0, w.params.max_len_cap, dev, rank, world,
                        len_nt=int(rs.len_nt[0]), n_shard=n_shard, words_per_read=W, n_total=n)
    sp.load_shard(words[lo:hi].to(dev))
    ok = True
    for _ in range(2):  # the workspaces are reused across builds
        sp.run()
        e = sp.plan.result_host().edges()
        e[:, 0] += lo
        parts = [None] * world
        dist.all_gather_object(parts, e)
        if rank == 0:
            from oracle import oracle

            got = np.concatenate(parts)
            want = oracle.prefsuf(rs, w.params.min_overlap, w.params.rs_min_overlap, 0)
            ok = ok and got.shape == want.shape and np.array_equal(got, want)
            print(f"sharded world={world} nodes={n} edges={got.shape[0]} match={ok} stages={sp.stats()['stage_ms']}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
