import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ops = importlib.import_module("recommendar-systems_b200.ops")
dev = "cuda:0"
g = torch.Generator().manual_seed(1)
def rel(a, b): return float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max())
for (M, N, K) in [(64, 64, 1500), (64, 64, 7050), (64, 384, 64), (64, 4096, 64), (64, 64, 4096), (64, 64, 384), (1500, 64, 64), (128, 128, 23033), (32, 32, 1000), (32, 384, 32)]:
    A = torch.randn(M, K, generator=g).to(dev); B = torch.randn(N, K, generator=g).to(dev)
    want = A.double() @ B.double().t()
    def t(f):
        try:
            return rel(f(), want)
        except RuntimeError as e:
            return "n/a"
    print(M, N, K, "kk", t(lambda: ops.gemm(A, True, B, True, M, N, K)),
          "k,n", t(lambda: ops.gemm(A, True, B.t().contiguous(), False, M, N, K)),
          "m,n", t(lambda: ops.gemm(A.t().contiguous(), False, B.t().contiguous(), False, M, N, K)), flush=True)
