"""TMEM read bandwidth of one SM (tcgen05.ld.32x32b.x16 streams) for 4 / 8 / 16 warps."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pde_inverse_problem_b200 import _lib as L
torch.zeros(1, device='cuda')
lib = C.CDLL(L.LIB_PATH)
out = (C.c_longlong * 4)()
for nw in (1, 4, 8, 16):
    for _ in range(2):
        rc = lib.pdeip_debug_tmem_bw(nw, 2000, out)
    print(f"warps {nw:2d}: rc {rc} cycles {out[0]} bytes {out[1]} -> {out[1] / max(out[0], 1):.1f} B/cycle/SM")
