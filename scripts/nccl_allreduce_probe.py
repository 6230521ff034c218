import os, time, torch, torch.distributed as dist
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev=torch.device("cuda",rank)
dist.init_process_group("nccl", device_id=dev)
n=110_000_000
def bench(t,label):
    for _ in range(3): dist.all_reduce(t)
    torch.cuda.synchronize(); dist.barrier()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): dist.all_reduce(t)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    if rank==0: print(f"{label}: {ms:.3f} ms  algbw {t.numel()*t.element_size()/ms/1e6:.0f} GB/s busbw {t.numel()*t.element_size()*2*(world-1)/world/ms/1e6:.0f} GB/s", flush=True)
a=torch.zeros(n,device=dev)
bench(a,"plain fp32 440MB")
b=torch.zeros(n,device=dev,dtype=torch.bfloat16)
bench(b,"plain bf16 220MB")
try:
    backend = dist.group.WORLD._get_backend(dev)
    pool = torch.cuda.MemPool(backend.mem_allocator)
    with torch.cuda.use_mem_pool(pool):
        c = torch.zeros(n, device=dev)
    backend.register_mem_pool(pool)
    bench(c,"registered (ncclMemAlloc pool) fp32 440MB")
except Exception as e:
    if rank==0: print("registered pool failed:", repr(e)[:300], flush=True)
dist.destroy_process_group()
