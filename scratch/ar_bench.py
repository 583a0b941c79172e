import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for n in (138241, 2 * 138241):
    t = torch.randn(n, device="cuda")
    for _ in range(20): dist.all_reduce(t, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): dist.all_reduce(t, op=dist.ReduceOp.AVG)
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f"ALGO={os.environ.get('NCCL_ALGO','-')} PROTO={os.environ.get('NCCL_PROTO','-')} n={n}: {e0.elapsed_time(e1)/200*1000:.1f} us", flush=True)
dist.destroy_process_group()
