"""all_reduce of the cfg2 gradient buffer (14.7 MB fp32) on N GPUs: time per call under the NCCL settings in the env."""
import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for nbytes in (14_714_112, 2_000_000):
    x = torch.ones(nbytes // 4, device="cuda")
    for _ in range(10):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"[{os.environ.get('NCCL_ALGO', 'default')}] all_reduce {nbytes / 1e6:.1f} MB x{world}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us", flush=True)
dist.destroy_process_group()
