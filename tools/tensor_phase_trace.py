import sys, ctypes as C, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pde_inverse_problem_b200 import ops, _lib as L
from oracle import model as o_model
cuda = torch.device('cuda')
d=8; n=(1<<18)*200
p = o_model.init_mlp_params(d, 32, 2)
flat = o_model.flatten_params(p).float().to(cuda)
pts = torch.randn(2*d, n, device=cuda)
TG = ops.TrueGrad(L.DRIFT_GMM, (torch.rand(16, d, device=cuda)*8-4), 1.0)
spec = ops.ModelSpec(L.MODEL_MLP, d, 32, 2)
acc = ops.ResidualAccumulator(spec, device=cuda).begin()
import time
for _ in range(2):
    acc.accumulate(L.SET_KFP_0T, flat, pts, 1.0/n, coef=0.5, layout=L.LAYOUT_SOA, path=L.PATH_TENSOR, true_grad=TG)
torch.cuda.synchronize()
lib = C.CDLL(L.LIB_PATH)
lib.pdeip_debug_tensor_trace.argtypes=[C.c_void_p, C.c_int]
buf = (C.c_longlong*64)()
print(lib.pdeip_debug_tensor_trace(buf, 64))
t = list(buf)
t0 = t[0]
print("phase: t_arrive(rel)  arrive->mma_start  mma_start->issued  issued->done   total")
for ph in range(12):
    a, b, c_, dn = t[4*ph:4*ph+4]
    print(ph, a-t0, b-a, c_-b, dn-c_, " total", dn-a)
print("tile start->S0 arrive", t[0]-t[60], " tile total", t[61]-t[60])
ntile = (n + 127)//128
per_cta = ntile/148
print("CTA0 loop cycles", t[63]-t[62], "tiles/CTA", per_cta, "cycles/tile", (t[63]-t[62])/per_cta)
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record(); acc.accumulate(L.SET_KFP_0T, flat, pts, 1.0/n, coef=0.5, layout=L.LAYOUT_SOA, path=L.PATH_TENSOR, true_grad=TG); e1.record(); torch.cuda.synchronize()
print("kernel ms", e0.elapsed_time(e1), "evals/s", n/e0.elapsed_time(e1)*1e3)
