"""tcgen05.mma cost probe (csrc/umma_selftest.cu: umma_cost_kernel): cycles per 128 x N x 16 bf16 MMA issued back to back."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: F401  (CUDA context)
from pde_inverse_problem_b200 import _lib as L

torch.zeros(1, device="cuda")
lib = C.CDLL(L.LIB_PATH)
lib.pdeip_debug_umma_cost.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
reps = 64
print("variant bits: 1 = A from TMEM, 2 = M 64, 4 = four rotating accumulators, 8 = B MN-major view, 16 = four rotating A tiles, 32 = four rotating B tiles, 64 = (A0 B0)(A0 B1)(A1 B0)(A1 B1)")
for variant in (0, 1, 2, 16, 17, 32, 33, 48, 49, 52, 53, 64, 65, 68, 69, 18, 50):
    row = []
    for N in (16, 32, 48, 64, 96, 128, 256):
        if (variant & 4) and N > 112:
            row.append("   -  ")
            continue
        buf = (C.c_longlong * 2)()
        best = None
        for _ in range(3):
            rc = lib.pdeip_debug_umma_cost(N, reps, variant, buf)
            assert rc == 0
            v = (buf[0] / reps, buf[1] / reps)
            best = v if best is None or v[1] < best[1] else best
        row.append(f"{best[0]:5.1f}/{best[1]:5.1f}")
    print(f"variant {variant:2d}: " + "  ".join(f"N={n}: {r}" for n, r in zip((16, 32, 48, 64, 96, 128, 256), row)), " (issue / complete cycles per MMA)")

# ---- hand-off latency: commit issued -> the waiting warps are past their mbarrier wait (umma_handoff_kernel) ----
lib.pdeip_debug_umma_handoff.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_void_p]
print("\nhand-off: 8 waiting warps, warp 8 issues `reps` MMAs (N = 32) + commit; cycles first issue -> commit issued | commit issued -> wake, per warp")
for ts in (0, 1):
    for reps in (1, 4, 10):
        for kind, hint, name in ((0, 0, "try_wait"), (1, 0, "test_wait"), (2, 20, "try_wait hint 20 ns"), (2, 1000, "try_wait hint 1 us")):
            buf = (C.c_longlong * 9)()
            assert lib.pdeip_debug_umma_handoff(32, reps, ts, kind, hint, buf) == 0
            print(f"ts={ts} reps={reps:2d} {name:22s}: issue {buf[0]:4d} | wake " + " ".join(f"{buf[1 + w]:4d}" for w in range(8)))
