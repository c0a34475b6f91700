"""GPU: tcgen05 building blocks (descriptor / core-matrix layout conventions of csrc/umma.cuh)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(mode, A, B, K, N, cuda):
    from pde_inverse_problem_b200 import _lib as L
    lib = L.load()
    D = torch.full((128, N), float("nan"), device=cuda)
    status = torch.zeros(1, dtype=torch.int32, device=cuda)
    st = lib.pdeip_debug_umma(mode, A.data_ptr(), B.data_ptr() if B is not None else None, D.data_ptr(), K, N,
                              status.data_ptr(), torch.cuda.current_stream().cuda_stream)
    L.check(st, "pdeip_debug_umma")
    torch.cuda.synchronize()
    assert status.item() == 0, "tcgen05 self-test timed out"
    return D


@pytest.mark.parametrize("K,N", [(16, 16), (32, 32), (48, 32), (32, 48), (128, 48)])
def test_umma_layout_conventions(cuda, K, N):
    g = torch.Generator().manual_seed(K * 100 + N)
    bf = lambda t: t.to(torch.bfloat16).float()
    A = bf(torch.randn(128, K, generator=g)).to(cuda)
    B0 = bf(torch.randn(N, K, generator=g)).to(cuda)
    D = _run(0, A, B0, K, N, cuda)
    assert torch.allclose(D, A @ B0.T, rtol=1e-4, atol=1e-4), "K-major x K-major"
    B1 = bf(torch.randn(K, N, generator=g)).to(cuda)
    D = _run(1, A, B1, K, N, cuda)
    assert torch.allclose(D, A @ B1, rtol=1e-4, atol=1e-4), "K-major x transposed view"
    A2 = bf(torch.randn(K, 128, generator=g)).to(cuda)
    D = _run(2, A2, B1, K, N, cuda)
    assert torch.allclose(D, A2.T @ B1, rtol=1e-4, atol=1e-4), "transposed view x transposed view"


def test_tmem_store_load_roundtrip(cuda):
    A = torch.randn(128, 32, device=cuda)
    D = _run(3, A, None, 32, 32, cuda)
    assert torch.equal(D, A)


def test_m64_instruction_shape_lane_mapping(cuda):
    """tcgen05.mma with M = 64 (cta_group::1): row i of D is written to TMEM lane (i % 16) + 32 (i / 16); the other
    lanes keep their contents.  The residual kernel's dW GEMMs rely on this mapping."""
    K, N = 32, 32
    g = torch.Generator().manual_seed(64)
    bf = lambda t: t.to(torch.bfloat16).float()
    A2 = bf(torch.randn(K, 128, generator=g)).to(cuda)
    B1 = bf(torch.randn(K, N, generator=g)).to(cuda)
    D = _run(4, A2, B1, K, N, cuda)
    ref = A2.T @ B1
    for i in range(64):
        lane = (i % 16) + 32 * (i // 16)
        assert torch.allclose(D[lane], ref[i], rtol=1e-4, atol=1e-4), i
    untouched = [l for l in range(128) if (l % 32) >= 16]
    assert torch.all(D[untouched] == -777.0)


@pytest.mark.parametrize("K,N", [(16, 32), (32, 32), (64, 32), (48, 48)])
def test_a_operand_from_tensor_memory(cuda, K, N):
    """tcgen05.mma with the A operand in TMEM (thread = row = lane stores its row as packed bf16 pairs, 32-bit column c =
    elements k = 2c, 2c + 1; 8 columns per K = 16 step), B K-major in shared memory: D = A B^T."""
    g = torch.Generator().manual_seed(K * 7 + N)
    bf = lambda t: t.to(torch.bfloat16).float()
    A = bf(torch.randn(128, K, generator=g)).to(cuda)
    B0 = bf(torch.randn(N, K, generator=g)).to(cuda)
    D = _run(5, A, B0, K, N, cuda)
    assert torch.allclose(D, A @ B0.T, rtol=1e-4, atol=1e-4)
