"""Bandwidth of the ways a raw batch can travel from rank 0 to the other ranks of a node:
ncclBroadcast, copy-engine copies into CUDA-IPC mapped peer buffers, copy-engine copies into
symmetric-memory peer buffers.  torchrun --nproc-per-node N microbench/peer_copy.py [MB]"""
import os, sys, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = mb << 20
src = torch.randint(0, 255, (n,), dtype=torch.uint8, device=dev)
reps = 10


def timed(fn, label):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = e0.elapsed_time(e1) / reps
    if rank == 0:
        print(f'{label:44s} {ms:8.3f} ms per {mb} MB to {world - 1} peer(s) = {n * (world - 1) / ms / 1e6:8.1f} GB/s egress', flush=True)


buf = torch.empty(n, dtype=torch.uint8, device=dev)
timed(lambda: dist.broadcast(buf if rank else src, src=0), 'ncclBroadcast')

# ---- CUDA IPC (torch.multiprocessing reductions)
try:
    from torch.multiprocessing.reductions import reduce_tensor
    mine = torch.empty(n, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    got = [None] * world
    dist.all_gather_object(got, reduce_tensor(mine) if rank else None)
    peers = []
    if rank == 0:
        peers = [fn(*a) for (fn, a) in [g for g in got if g is not None]]
        lanes = [torch.cuda.Stream() for _ in peers]

    def ipc_copy():
        if rank == 0:
            cur = torch.cuda.current_stream()
            for lane, p in zip(lanes, peers):
                lane.wait_stream(cur)
                with torch.cuda.stream(lane):
                    p.copy_(src, non_blocking=True)
                cur.wait_stream(lane)
    timed(ipc_copy, 'IPC peer buffers, Tensor.copy_')
    if rank == 0:
        print('   peer tensor devices:', [str(p.device) for p in peers], flush=True)
    dist.barrier()
    ok = bool((mine == src.to(dev)).all().item()) if False else None
    chk = src.clone() if rank == 0 else mine
    lst = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    if rank == 0:
        print('   IPC data correct:', all(bool((t == src).all().item()) for t in lst), flush=True)
except Exception as ex:
    print(f'rank {rank}: IPC path failed: {ex!r}', flush=True)

# ---- symmetric memory
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(n, dtype=torch.uint8, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD)
    views = [hdl.get_buffer(r, (n,), torch.uint8) for r in range(world) if r != 0] if rank == 0 else []
    lanes2 = [torch.cuda.Stream() for _ in views]

    def sm_copy():
        if rank == 0:
            cur = torch.cuda.current_stream()
            for lane, v in zip(lanes2, views):
                lane.wait_stream(cur)
                with torch.cuda.stream(lane):
                    v.copy_(src, non_blocking=True)
                cur.wait_stream(lane)
    timed(sm_copy, 'symmetric memory peer views, Tensor.copy_')
    if rank == 0:
        print('   view devices:', [str(v.device) for v in views], flush=True)
    chk = src.clone() if rank == 0 else t
    lst = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    if rank == 0:
        print('   symmetric-memory data correct:', all(bool((x == src).all().item()) for x in lst), flush=True)

    # receivers pull instead
    view0 = hdl.get_buffer(0, (n,), torch.uint8)
    if rank == 0:
        t.copy_(src)
    torch.cuda.synchronize()
    dist.barrier()
    dst = torch.empty(n, dtype=torch.uint8, device=dev)

    def sm_pull():
        if rank != 0:
            dst.copy_(view0, non_blocking=True)
    for _ in range(2):
        sm_pull()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sm_pull()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f'{"symmetric memory, receivers pull":44s} {ms.item():8.3f} ms = {n * (world - 1) / ms.item() / 1e6:8.1f} GB/s egress', flush=True)
except Exception as ex:
    print(f'rank {rank}: symmetric-memory path failed: {ex!r}', flush=True)
dist.barrier()
dist.destroy_process_group()
