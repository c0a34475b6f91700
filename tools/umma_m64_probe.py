"""Where does tcgen05.mma with M = 64 (cta_group::1) put the rows of D in TMEM?  (both operands transposed views)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_umma import _run
cuda = torch.device("cuda")
K, N = 32, 32
g = torch.Generator().manual_seed(1)
bf = lambda t: t.to(torch.bfloat16).float()
A2 = bf(torch.randn(K, 128, generator=g)).to(cuda)
B1 = bf(torch.randn(K, N, generator=g)).to(cuda)
D = _run(4, A2, B1, K, N, cuda).cpu()
ref = (A2.T @ B1).cpu()   # [128, N]; rows 0..63 are what M = 64 computes
written = [l for l in range(128) if not torch.all(D[l] == -777.0)]
print("lanes written:", written)
for l in written:
    err = (ref[:64] - D[l]).abs().sum(1)
    r = int(err.argmin())
    print(f"lane {l:3d} <- row {r:3d} (err {float(err[r]):.2e})")
