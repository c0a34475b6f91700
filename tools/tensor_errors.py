"""Prints the tensor-path errors (loss, gradient, worst leaf) against the float64 Taylor twin for several d."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from pde_inverse_problem_b200 import ops, _lib as L
from tensor_v2_model import kfp_0T_schedule
cuda = torch.device("cuda")
for d, n, seed in [(8, 20000, 3), (16, 5000, 7), (4, 5000, 5), (32, 3000, 11), (2, 2000, 9), (8, 1000, 21), (16, 300, 22)]:
    g = torch.Generator().manual_seed(seed)
    dims = [d, 32, 32, 40]
    W = [torch.randn(dims[i], dims[i + 1], generator=g, dtype=torch.float64) * (2.0 / dims[i]) ** 0.5 for i in range(3)]
    b = [0.1 * torch.randn(dims[i + 1], generator=g, dtype=torch.float64) for i in range(3)]
    z = torch.randn(n, 2 * d, generator=g, dtype=torch.float64) * torch.cat([torch.full((d,), 2.0), torch.full((d,), 0.6)]).double()
    r = kfp_0T_schedule(W, b, z[:, :d], z[:, d:], 0.5, 1.0 / n)
    flat = torch.cat([torch.cat([w_.reshape(-1), b_.reshape(-1)]) for w_, b_ in zip(W, b)]).float().to(cuda)
    spec = ops.ModelSpec(L.MODEL_MLP, d, 32, 2)
    acc = ops.ResidualAccumulator(spec, device=cuda).begin()
    acc.accumulate(L.SET_KFP_0T, flat, z.float().to(cuda), 1.0 / n, coef=0.5, true_grad=ops.TrueGrad(L.DRIFT_NONE), path=L.PATH_TENSOR)
    s, gr = acc.finalize()
    s, gr = s.double().cpu(), gr.double().cpu()
    ref = torch.cat([torch.cat([w_.reshape(-1), b_.reshape(-1)]) for w_, b_ in zip(r["dW"], r["db"])])
    off, worst, per_leaf = 0, 0.0, []
    for li, (w_, b_) in enumerate(zip(r["dW"], r["db"])):
        for nm, leaf in (("W", w_), ("b", b_)):
            k = leaf.numel()
            e = ((gr[off:off + k] - leaf.reshape(-1)).abs().max() / leaf.abs().max()).item()
            per_leaf.append(f"{nm}{li} {e:.1e} (|leaf| {leaf.abs().max().item():.1e})")
            worst = max(worst, e)
            off += k
    print(f"d={d:2d} n={n:5d}: loss {abs(float(s[L.SUM_LOSS]) - float(r['loss'])) / abs(float(r['loss'])):.1e} "
          f"grad {((gr - ref).abs().max() / ref.abs().max()).item():.1e} worst-leaf {worst:.1e}   " + "  ".join(per_leaf))
