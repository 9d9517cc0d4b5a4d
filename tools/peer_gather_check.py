"""Run under torchrun on >= 2 GPUs: PeerGather against the NCCL all-gather - equality and time for gradient-sized rows."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfdm_b200.distributed import PeerGather, gather_designs

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rows, cols = 256, 132612
pg = PeerGather(rows * world, (cols,), device=dev)
pg1 = PeerGather(rows * world, (), device=dev)
g = torch.Generator(device=dev).manual_seed(100 + rank)
ok = True
for trial in range(3):
    x = torch.randn(rows, cols, dtype=torch.float64, device=dev, generator=g)
    y = torch.randn(rows, dtype=torch.float64, device=dev, generator=g)
    a = pg.gather(x, sync=False); b = pg1.gather(y, sync=False); pg.fence()
    ok &= bool(torch.equal(a, gather_designs(x, rows * world))) and bool(torch.equal(b, gather_designs(y, rows * world)))
def timed(fn, n=10):
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
x = torch.randn(rows, cols, dtype=torch.float64, device=dev, generator=g)
t_peer = timed(lambda: pg.gather(x))
t_nccl = timed(lambda: gather_designs(x, rows * world))
if rank == 0:
    gb = x.numel() * 8 * (world - 1) / 1e9
    print("world %d  equal %s  peer %.3f ms (%.0f GB/s out per rank)  nccl %.3f ms (%.0f GB/s)" % (world, ok, t_peer, gb / t_peer * 1e3, t_nccl, gb / t_nccl * 1e3))
pg.close(); pg1.close()
dist.destroy_process_group()
