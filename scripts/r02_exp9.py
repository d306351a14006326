"""Where does the time of the sharded exp scan go?  (1 GPU, pieces timed separately)"""
import sys, torch, time
sys.path.insert(0, ".")
from numbagg_b200 import decorators as D
import bench
dev = torch.device("cuda", 0)
n = 500_000_000
x = bench.gen_flat(torch, dev, torch.float64, 0, n, 0.3, seed=3).view(1, n)
def t(fn, steps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
print("out only        ", t(lambda: D.run_move_exp("move_exp_nanmean", [x], 0.1, 0.0, -1)))
print("out + agg       ", t(lambda: D.run_move_exp("move_exp_nanmean", [x], 0.1, 0.0, -1, None, True, True)))
print("agg only        ", t(lambda: D.run_move_exp("move_exp_nanmean", [x], 0.1, 0.0, -1, None, True, False)))
c = D.run_move_exp("move_exp_nanmean", [x], 0.1, 0.0, -1, None, True, False)[1]
print("out with carry  ", t(lambda: D.run_move_exp("move_exp_nanmean", [x], 0.1, 0.0, -1, c, False, True)))
h = x.narrow(1, 0, 7277)
print("head only       ", t(lambda: D.run_move_exp("move_exp_nanmean", [h], 0.1, 0.0, -1, c, False, True)))
print("ffill out+agg   ", t(lambda: D.run_fill("ffill", x, n, -1, None, True, True)))
print("ffill out       ", t(lambda: D.run_fill("ffill", x, n, -1)))
