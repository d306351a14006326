import sys, numpy as np, torch
sys.path.insert(0, ".")
import numbagg_b200 as nb
from numbagg_b200.decorators import run_group
from oracle import oracle
torch.cuda.set_device(0)
def ev(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    ts=[]
    for _ in range(reps):
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return np.median(ts), min(ts)
ROWS = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
for dt, rows, n, K in [(torch.float32, ROWS, 1_000_000, 1000), (torch.float64, ROWS//2, 1_000_000, 1000), (torch.float32, 100, 100_000, 12), (torch.float32, 1, 10_000_000, 12), (torch.float64, 1, 100_000_000, 1_000_000)]:
    g=torch.Generator(device="cuda").manual_seed(0)
    a=torch.rand((rows,n), generator=g, device="cuda", dtype=dt); a[a<=0.1]=float("nan")
    lab_np=np.random.RandomState(0).randint(0,K,size=n).astype(np.int64)
    lab=torch.from_numpy(lab_np).cuda()
    sub=a[:11].cpu().numpy()
    funcs = ["group_nansum","group_nanmean","group_nanstd","group_nancount","group_nansum_of_squares","group_nanmax","group_nanargmax","group_nanfirst"] if rows>1 else ["group_nansum","group_nanvar","group_nanargmax","group_nanfirst"]
    for name in funcs:
        got=run_group(name, a[:11].contiguous(), lab, K, 1).cpu().numpy()
        exp=getattr(oracle,name)(sub, lab_np, num_labels=K, axis=-1)
        exact=np.array_equal(got,exp,equal_nan=True)
        close=np.allclose(got,exp,rtol=1e-5 if dt==torch.float32 else 1e-12,equal_nan=True)
        med,best=ev(lambda: run_group(name, a, lab, K, 1))
        nbytes=a.numel()*a.element_size()+n*8+rows*K*a.element_size()
        print(f"{name:24s} {str(dt):14s} ({rows},{n}) K={K}: {med:.3f} ms {a.numel()/med/1e6:.1f} Gel/s {nbytes/med/1e6:.0f} GB/s ({nbytes/med/1e6/6447.8:.2%}) exact={exact} close={close}", flush=True)
    del a
