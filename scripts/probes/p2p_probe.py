"""2-GPU probe: which peer-memory plumbing works on this box (torch symmetric memory, CUDA IPC)?"""
import ctypes as C, os, sys, time, traceback
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def log(*a):
    print(f"[r{rank}]", *a, flush=True)
log("can_device_access_peer", [torch.cuda.can_device_access_peer(local, p) for p in range(world) if p != local])
# --- 1. symmetric memory
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.int32, device=dev)
    t.fill_(rank + 1)
    h = symm.rendezvous(t, dist.group.WORLD)
    log("symm ok; buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "signal_pad_ptrs", len(h.signal_pad_ptrs))
    h.barrier()
    peer = (rank + 1) % world
    v = h.get_buffer(peer, (1 << 20,), torch.int32)
    torch.cuda.synchronize()
    log("symm peer read", int(v[12345].item()), "expected", peer + 1)
    x = torch.empty(1 << 20, dtype=torch.int32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    big = symm.empty(64 << 20, dtype=torch.int32, device=dev); hb = symm.rendezvous(big, dist.group.WORLD); hb.barrier()
    vb = hb.get_buffer(peer, (64 << 20,), torch.int32); y = torch.empty_like(big)
    y.copy_(vb); torch.cuda.synchronize(); e0.record(); y.copy_(vb); e1.record(); torch.cuda.synchronize()
    log("symm peer copy 256 MiB GB/s", 0.268 / (e0.elapsed_time(e1) / 1e3))
    h.barrier()
except Exception:
    log("symm FAILED"); traceback.print_exc()
# --- 2. CUDA IPC through libalga_gpu
try:
    from alga_b200 import _lib
    lib = C.CDLL(_lib.LIB_PATH)
    buf = torch.full((1 << 20,), rank + 100, dtype=torch.int32, device=dev)  # torch caching allocator block
    raw = C.c_void_p()
    cudart = None
    hnd = (C.c_ubyte * 64)()
    rc = lib.alga_gpu_ipc_export(C.c_void_p(buf.data_ptr()), hnd)
    log("ipc export rc", rc, lib.alga_gpu_last_error and C.c_char_p(lib.alga_gpu_last_error()).value if rc else "")
    hs = [None] * world
    dist.all_gather_object(hs, bytes(hnd))
    peer = (rank + 1) % world
    p = C.c_void_p()
    hb = (C.c_ubyte * 64).from_buffer_copy(hs[peer])
    rc = lib.alga_gpu_ipc_open(hb, C.byref(p))
    log("ipc open rc", rc, hex(p.value or 0))
    if rc == 0:
        from alga_b200.plan import _as_tensor
        v = _as_tensor(p.value, (1 << 20,), "<i4", torch.int32, dev)
        torch.cuda.synchronize()
        log("ipc peer read", int(v[777].item()), "expected", peer + 100)
        dist.barrier()
        lib.alga_gpu_ipc_close(p)
except Exception:
    log("ipc FAILED"); traceback.print_exc()
dist.barrier(); dist.destroy_process_group()
