import ctypes as C, os, sys
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libprobe.so"))
lib.probe_scatter.restype = C.c_float
lib.probe_scatter.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_int, C.c_int]
n_t, cap, n_ops = 2_600_000, 16, 7_800_000
indeg = symm.empty(n_t, dtype=torch.int32, device=dev); rows = symm.empty(n_t * cap * 4, dtype=torch.int32, device=dev)
indeg.zero_(); rows.zero_()
hi = symm.rendezvous(indeg, dist.group.WORLD); hr = symm.rendezvous(rows, dist.group.WORLD)
hi.barrier(); torch.cuda.synchronize()
peer = (rank + 1) % world
for blocks in (148 * 4, 148 * 16):
    loc = lib.probe_scatter(hi.buffer_ptrs[rank], hr.buffer_ptrs[rank], n_t, n_ops, cap, blocks, 5)
    dist.barrier()
    rem = lib.probe_scatter(hi.buffer_ptrs[peer], hr.buffer_ptrs[peer], n_t, n_ops, cap, blocks, 5)
    dist.barrier()
    print(f"[r{rank}] blocks={blocks} local {loc:.3f} ms ({n_ops/loc/1e6:.1f} G/s)  remote {rem:.3f} ms ({n_ops/rem/1e6:.1f} G/s)", flush=True)
dist.barrier(); dist.destroy_process_group()
