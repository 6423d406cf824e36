import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ops = importlib.import_module("recommendar-systems_b200.ops")
DEV = "cuda:0"
def rel(a, b): return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max() / b.detach().double().abs().max())
gen = torch.Generator().manual_seed(5)
I, F, d = 1500, 384, 64
emb = torch.nn.Embedding.from_pretrained(torch.randn(I, F, generator=gen).to(DEV), freeze=False)
W = torch.nn.Parameter((torch.randn(d, F, generator=gen) * 0.05).to(DEV))
b = torch.nn.Parameter(torch.randn(d, generator=gen).to(DEV))
Gy = torch.randn(I, d, generator=gen).to(DEV)
y = ops.table_project(emb, W, b)
(y * Gy).sum().backward()
lr = emb.weight._mmrec_lowrank
print("dY is Gy-valued:", rel(lr.dY, Gy), "W same obj:", lr.W.data_ptr() == W.data_ptr())
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
print("vs corrected", rel(W.grad, base - corrd), "vs uncorrected", rel(W.grad, base), "vs plus", rel(W.grad, base + corrd))
print("lowrank dY after 2nd:", rel(emb.weight._mmrec_lowrank.dY, Gy))
