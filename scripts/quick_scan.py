import sys, numpy as np, torch
sys.path.insert(0, ".")
import numbagg_b200 as nb
from numbagg_b200.decorators import run_move_exp, run_fill
from oracle import oracle
torch.cuda.set_device(0)
def ev(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    ts=[]
    for _ in range(reps):
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return np.median(ts), min(ts)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
for dt, shape in [(torch.float64,(1,N)),(torch.float64,(2000,100000)),(torch.float32,(1000,1000000))]:
    g=torch.Generator(device="cuda").manual_seed(0)
    a=torch.rand(shape, generator=g, device="cuda", dtype=dt); a[a<=0.3]=float("nan")
    b=a*a+1
    # parity on a prefix / row subset vs the oracle
    sub = a[:2, :3_000_000].cpu().numpy()
    for name in ["ffill","bfill"]:
        got = run_fill(name, a[:2, :3_000_000].contiguous(), 3_000_000, -1)[0].cpu().numpy()
        exp = getattr(oracle, name)(sub)
        assert np.array_equal(got, exp, equal_nan=True), name
        med,best=ev(lambda: run_fill(name, a, a.shape[-1], -1))
        nbytes=a.numel()*a.element_size()*2
        print(f"{name:18s} {str(dt):14s} {shape}: {med:.3f} ms  {a.numel()/med/1e6:.1f} Gel/s  {nbytes/med/1e6:.0f} GB/s ({nbytes/med/1e6/6447.8:.2%})", flush=True)
    for name in ["move_exp_nanmean","move_exp_nansum","move_exp_nancount","move_exp_nanvar","move_exp_nanstd","move_exp_nancov","move_exp_nancorr"]:
        arrs=[a,b] if name in("move_exp_nancov","move_exp_nancorr") else [a]
        subs=[x[:2, :3_000_000].contiguous() for x in arrs]
        got = run_move_exp(name, subs, 0.1, 0.0, -1)[0].cpu().numpy()
        exp = getattr(oracle, name)(*[s.cpu().numpy() for s in subs], alpha=(0.1 if dt==torch.float64 else np.float32(0.1)))
        rt = 1e-12 if dt==torch.float64 else 1e-5
        ok = np.array_equal(np.isnan(got), np.isnan(exp)) and np.allclose(got, exp, rtol=rt, atol=rt*4, equal_nan=True)
        med,best=ev(lambda: run_move_exp(name, arrs, 0.1 if dt==torch.float64 else float(np.float32(0.1)), 0.0, -1))
        nbytes=a.numel()*a.element_size()*(len(arrs)+1)
        print(f"{name:18s} {str(dt):14s} {shape}: {med:.3f} ms  {a.numel()/med/1e6:.1f} Gel/s  {nbytes/med/1e6:.0f} GB/s ({nbytes/med/1e6/6447.8:.2%}) parity={'OK' if ok else 'FAIL'}", flush=True)
    del a,b
