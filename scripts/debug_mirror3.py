import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ops = importlib.import_module("recommendar-systems_b200.ops")
DEV = "cuda:0"
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))
gen = torch.Generator().manual_seed(5)
I, F, d = 1500, 384, 64
emb = torch.nn.Embedding.from_pretrained(torch.randn(I, F, generator=gen).to(DEV), freeze=False)
W = torch.nn.Parameter((torch.randn(d, F, generator=gen) * 0.05).to(DEV))
b = torch.nn.Parameter(torch.randn(d, generator=gen).to(DEV))
Gy = torch.randn(I, d, generator=gen).to(DEV)
y = ops.table_project(emb, W, b)
(y * Gy).sum().backward()
variant = sys.argv[1] if len(sys.argv) > 1 else "a"
if variant in ("b", "c"):
    Xr, Wr, br = (t.detach().double().requires_grad_(True) for t in (emb.weight, W, b))
    yr = torch.nn.functional.linear(Xr, Wr, br)
    (yr * Gy.double()).sum().backward()
    print(rel(y, yr), rel(W.grad, Wr.grad), rel(b.grad, br.grad))
lr = emb.weight._mmrec_lowrank
if variant == "c":
    print("dense", rel(lr.dense(), Xr.grad))
coef = torch.tensor([0.37], device=DEV)
dY1, W1 = lr.dY, lr.W.clone()
emb.weight._mmrec_lowrank = None
emb.weight._mmrec_delta = (coef, dY1, W1)
W.grad = b.grad = None
y2 = ops.table_project(emb, W, b)
(y2 * Gy).sum().backward()
X = emb.weight.detach().double()
base = Gy.double().t() @ X
corrd = 0.37 * (Gy.double().t() @ dY1.double()) @ W1.double()
print(variant, "vs corrected", rel(W.grad, base - corrd), "vs uncorrected", rel(W.grad, base), "vs plus", rel(W.grad, base + corrd))
X2 = (emb.weight.detach().double() - 0.37 * dY1.double() @ W1.double())
W2 = W.detach().double().requires_grad_(True)
y2r = torch.nn.functional.linear(X2, W2, b.detach().double())
(y2r * Gy.double()).sum().backward()
print(variant, "test-style", rel(y2, y2r), rel(W.grad, W2.grad), "W2.grad vs base-corrd", rel(W2.grad, base - corrd))
