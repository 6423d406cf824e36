import sys, importlib, torch
sys.path.insert(0, ".")
ops = importlib.import_module("recommendar-systems_b200.ops")
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))
for M, N, K in [(7050, 64, 4096), (300, 64, 384), (26495, 64, 64), (129, 128, 100), (64, 8, 36)]:
    gen = torch.Generator().manual_seed(29)
    x, W, b = torch.randn(M, K, generator=gen), torch.randn(N, K, generator=gen) * 0.05, torch.randn(N, generator=gen)
    gy = torch.randn(M, N, generator=gen)
    xg, Wg, bg = (t.cuda().requires_grad_(True) for t in (x, W, b))
    y = ops.linear(xg, Wg, bg); y.backward(gy.cuda())
    xo, Wo, bo = (t.double().requires_grad_(True) for t in (x, W, b))
    yo = torch.nn.functional.linear(xo, Wo, bo); yo.backward(gy.double())
    yc = torch.nn.functional.linear(x.cuda(), W.cuda(), b.cuda())
    print(M, N, K, "y", rel(y, yo), "cublas", rel(yc, yo), "dx", rel(xg.grad, xo.grad), "dW", rel(Wg.grad, Wo.grad), "db", rel(bg.grad, bo.grad))
