"""Diagnostic: latency of the tiny all-gather the sharded search uses."""
import os, time, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
x = torch.zeros(2 * 256 * 10, dtype=torch.int32, device="cuda")
out = torch.empty(world * x.numel(), dtype=torch.int32, device="cuda")
for _ in range(10):
    dist.all_gather_into_tensor(out, x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(200):
    dist.all_gather_into_tensor(out, x)
e1.record(); torch.cuda.synchronize()
t1 = time.perf_counter()
if rank == 0:
    print(f"all_gather 20KB: device {e0.elapsed_time(e1)/200*1e3:.1f} us/call, wall {(t1-t0)/200*1e6:.1f} us/call, "
          f"peer access 0->1: {torch.cuda.can_device_access_peer(0, 1)}")
dist.destroy_process_group()
