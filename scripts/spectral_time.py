import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
ops = bench.pkg("ops")
def timeit(fn, iters=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
for n, d in [(7050, 64), (18357, 64), (23033, 128)]:
    g = torch.Generator().manual_seed(0)
    img = torch.randn(n, d, generator=g).cuda().requires_grad_(True); txt = torch.randn(n, d, generator=g).cuda().requires_grad_(True)
    ws = [torch.randn(d // 2 + 1, 2, generator=g).cuda().requires_grad_(True) for _ in range(3)]
    outs = ops.spectrum_convolution(img, txt, *ws, True)
    go = [torch.randn_like(o) for o in outs]
    tf = timeit(lambda: ops.spectrum_convolution(img, txt, *ws, True))
    def fb():
        o = ops.spectrum_convolution(img, txt, *ws, True)
        torch.autograd.backward(o, go)
    tfb = timeit(fb)
    print(f"spectral n={n} d={d}: fwd {tf:.1f} us, fwd+bwd {tfb:.1f} us", flush=True)
from torch.profiler import ProfilerActivity, profile
n, d = 7050, 64
for n, d in [(7050, 64), (23033, 128)]:
    g = torch.Generator().manual_seed(0)
    img = torch.randn(n, d, generator=g).cuda().requires_grad_(True); txt = torch.randn(n, d, generator=g).cuda().requires_grad_(True)
    ws = [torch.randn(d // 2 + 1, 2, generator=g).cuda().requires_grad_(True) for _ in range(3)]
    go = [torch.randn(n, d, device="cuda") for _ in range(3)]
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(10):
            o = ops.spectrum_convolution(img, txt, *ws, True)
            torch.autograd.backward(o, go)
        torch.cuda.synchronize()
    for e in prof.key_averages():
        if "spectral" in e.key:
            print(n, d, e.key.split("(")[0][-40:], f"{e.device_time_total / e.count:.1f} us")
