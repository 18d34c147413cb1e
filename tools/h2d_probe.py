"""Host -> device bandwidth of this box with 1..N ranks copying at once (pinned memory, 2 GiB per copy).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py"""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30  # int16 elements = 2 GiB
host = torch.empty(n, dtype=torch.int16, pin_memory=True)
host.zero_()
dev = torch.empty(n, dtype=torch.int16, device="cuda")
def run(active):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if active:
        for _ in range(3):
            dev.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return 3 * n * 2 / dt / 1e9 if active else 0.0
run(True)
alone = run(rank == 0)
if world > 1:
    dist.barrier()
together = run(True)
print(f"rank {rank}/{world}: alone(rank0 only) {alone:.1f} GB/s, all ranks at once {together:.1f} GB/s per rank", flush=True)
if world > 1:
    dist.destroy_process_group()
