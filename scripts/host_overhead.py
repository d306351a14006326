"""Host-side cost of one device-resident call (cfg1 is launch-bound): cProfile of run_move."""
import cProfile, pstats, sys, time, torch
sys.path.insert(0, ".")
from numbagg_b200.decorators import run_move, run_reduce
a = torch.rand((100, 10000), device="cuda", dtype=torch.float64)
for _ in range(20): run_move("move_mean", [a], 20, 1, -1)
torch.cuda.synchronize()
N = 2000
t0 = time.perf_counter()
for _ in range(N): run_move("move_mean", [a], 20, 1, -1)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"run_move: host {1e6*(t1-t0)/N:.1f} us/call, incl. drain {1e6*(t2-t0)/N:.1f} us/call")
t0 = time.perf_counter()
for _ in range(N): run_reduce("nansum", a, (1,))
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"run_reduce: host {1e6*(t1-t0)/N:.1f} us/call, incl. drain {1e6*(t2-t0)/N:.1f} us/call")
pr = cProfile.Profile(); pr.enable()
for _ in range(N): run_move("move_mean", [a], 20, 1, -1)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
