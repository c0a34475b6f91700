"""Times the production configuration of K1 alone (C3 shape by default)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pde_inverse_problem_b200 import _lib as L, ops
d = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = int(sys.argv[2]) if len(sys.argv) > 2 else 16
path = L.PATH_TENSOR if (len(sys.argv) > 3 and sys.argv[3] == "tensor") else L.PATH_FP32
n = int(sys.argv[4]) if len(sys.argv) > 4 else (1 << 18 if d <= 8 else 75776)
S = 200
blk = len(sys.argv) > 5 and sys.argv[5] == "block"
dev = torch.device('cuda')
z0 = torch.randn(n, 2 * d, device=dev)
if K > 0:
    params, kind = (torch.rand(K, d, device=dev) * 8 - 4), L.DRIFT_GMM
else:
    f = torch.randn(d, d + 1, device=dev); params, kind = (f @ f.T) / d, L.DRIFT_LINEAR
traj = torch.empty(3 * d * S * n, device=dev)
zl = torch.empty_like(z0)
def run():
    ops.kl_integrate(z0, S, 0.01, 0.5, kind, params, n_gaussian=K, seed=1, traj_layout=L.TRAJ_BLOCK128 if blk else L.TRAJ_TIME_SOA, traj_out=traj,
                     z_last_out=zl, emit_drift=True, path=path)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"d={d} K={K} n={n} path={'tensor' if path else 'fp32'} layout={'block128' if blk else 'time_soa'} minb={os.environ.get('PDEIP_ITC_MINB', '-')} status={ops.tensor_path_status()}: {ms:.3f} ms  {n*(S+1)/ms*1e3:.3e} particle-steps/s  {n*S*3*d*4/ms/1e6:.0f} GB/s")
