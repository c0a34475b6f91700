"""Per-phase clock trace of the tcgen05 residual kernel (build csrc with EXTRA=-DPDEIP_TC_TRACE first)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pde_inverse_problem_b200 import _lib as L, ops  # noqa: E402
from pde_inverse_problem_b200.core.model import V_hypothesis  # noqa: E402

cuda = torch.device('cuda')
d = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = (1 << 18) * 200
model = V_hypothesis(1, [32, 32], d)
flat = model.flat(model.init(11, torch.zeros(d, device=cuda)))
layout = {"soa": L.LAYOUT_SOA, "blk": L.LAYOUT_BLOCK128}[sys.argv[2] if len(sys.argv) > 2 else "blk"]  # blk = the pipeline's layout
pts = torch.randn(3 * d, n, device=cuda) if layout == L.LAYOUT_SOA else torch.randn(n // 128, 3 * d, 128, device=cuda)
spec = ops.ModelSpec(L.MODEL_MLP, d, 32, 2)
acc = ops.ResidualAccumulator(spec, device=cuda).begin()
tg = ops.TrueGrad(L.DRIFT_IN_POINTS)
for _ in range(2):
    acc.accumulate(L.SET_KFP_0T, flat, pts, 1.0 / n, coef=0.5, layout=layout, path=L.PATH_TENSOR, true_grad=tg)
torch.cuda.synchronize()
lib = C.CDLL(L.LIB_PATH)
lib.pdeip_debug_tensor_trace.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_longlong * 256)()
lib.pdeip_debug_tensor_trace(buf, 256)  # discard the warm-up launches (the read clears the sums)
acc.accumulate(L.SET_KFP_0T, flat, pts, 1.0 / n, coef=0.5, layout=layout, path=L.PATH_TENSOR, true_grad=tg)
torch.cuda.synchronize()
print("rc", lib.pdeip_debug_tensor_trace(buf, 256))
NT = 64  # kTraceTiles of residual_tensor.cu: every stamp is a sum over 64 consecutive tile rounds of CTA 0
t = [v / NT for v in buf]
t0 = t[0]
print("ph s | wait_start wait_end arrive | mma_ready fast_issued all_issued   (mean over 64 tiles, cycles since E0(s0) wait start)")
for ph in range(12):
    for s in range(2):
        r = t[(ph * 2 + s) * 8:(ph * 2 + s) * 8 + 8]
        if r[0] == 0:
            continue
        f = lambda x: f"{x:.0f}"
        if r[2] == 0:  # no hand-off after this phase (kGram: E6 runs into E7)
            print(ph, s, "|", f(r[0] - t0), f(r[1] - t0), "- | - - -    wait", f(r[1] - r[0]))
            continue
        print(ph, s, "|", f(r[0] - t0), f(r[1] - t0), f(r[2] - t0), "|", f(r[4] - t0), f(r[5] - t0) if r[5] else '-', f(r[6] - t0),
              "   wait", f(r[1] - r[0]), "(mbar", f(r[3] - r[0]) if r[3] else '-', ") epi", f(r[2] - r[1]),
              "(arrive", f(r[2] - r[7]) if r[7] else '-', ")  ready->issued", f(r[6] - r[4]), " issued->wait_end(next)")
for s_ in range(2):
    f = t[240 + 8 * s_:240 + 8 * s_ + 8]
    r = t[(4 * 2 + s_) * 8:(4 * 2 + s_) * 8 + 8]
    if f[0]:
        print(f"E4 slot {s_}: wait-end->tmem/lds data {f[0] - r[1]}, math+STS {f[1] - f[0]}, park(st+wait) {f[2] - f[1]}, "
              f"to fence {r[7] - f[2]}, fence.proxy.async {f[3] - r[7]}, tcgen05.fence {f[4] - f[3]}, bar.arrive {r[2] - f[4]}")
e0 = torch.cuda.Event(enable_timing=True)
e1 = torch.cuda.Event(enable_timing=True)
e0.record()
acc.accumulate(L.SET_KFP_0T, flat, pts, 1.0 / n, coef=0.5, layout=layout, path=L.PATH_TENSOR, true_grad=tg)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("kernel ms", ms, "evals/s", n / ms * 1e3, "cycles/tile @1.965GHz", ms * 1e-3 * 1.965e9 / (n / 128 / 148))
