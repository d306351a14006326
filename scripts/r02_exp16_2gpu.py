"""2-GPU breakdown of the sharded exp scan (torchrun --nproc-per-node 2)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, ".")
import bench
from numbagg_b200 import decorators as D, distributed as nd
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
n = 1_000_000_000
lo, hi = bench.shard_bounds(n, rank, world, bench.GEN_BLOCK)
shard = bench.gen_flat(torch, dev, torch.float64, lo, hi, 0.3, seed=3).view(1, -1)
def t(name, fn, steps=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"{name:40s} {ms.item():8.3f} ms", flush=True)
t("kernel out+agg", lambda: D.run_move_exp("move_exp_nanmean", [shard], 0.1, 0.0, -1, None, True, True))
def ag():
    out, agg = D.run_move_exp("move_exp_nanmean", [shard], 0.1, 0.0, -1, None, True, True)
    return nd._all_gather(agg, None)
t("kernel + all_gather", ag)
t("all_gather only", lambda: nd._all_gather(torch.zeros((1, 11), dtype=torch.float64, device=dev), None))
t("move_exp_sharded", lambda: nd.move_exp_sharded("move_exp_nanmean", shard, alpha=0.1))
t("fill_sharded", lambda: nd.fill_sharded("ffill", shard))
t("kernel ffill out+agg", lambda: D.run_fill("ffill", shard, n, -1, None, True, True))
dist.destroy_process_group()
