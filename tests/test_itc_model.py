"""The split-operand formulation of the tensor-core GMM drift (tests/itc_model.py, the CPU twin of
csrc/integrator_tc.cu) against the float64 closed form of the reference (oracle/potential.py <- core/potential.py:32-61)."""
import pytest
import torch

from conftest import relmax
from itc_model import gmm_grad_tensor_model
from oracle import potential as o_pot


@pytest.mark.parametrize("d,K,k_pad", [(8, 16, 16), (32, 64, 64), (16, 33, 48), (32, 7, 16), (8, 3, 16)])
def test_split_operand_gmm_gradient_matches_closed_form(d, K, k_pad):
    """C3 / C5 shaped ensembles (x0 ~ N(0, 4 I), centres ~ U[-4, 4]^d, GMM.py:17-42): two bf16 terms per operand keep
    the tensor-core formulation within 2e-4 (per-tensor max-norm) of the float64 gradient; a single bf16 term does not."""
    g = torch.Generator().manual_seed(d * 100 + K)
    x = torch.randn(4096, d, generator=g, dtype=torch.float64) * 2.0
    mus = torch.rand(K, d, generator=g, dtype=torch.float64) * 8 - 4
    ref = o_pot.vg_gmm_V(x, mus, 1.0)
    got = gmm_grad_tensor_model(x, mus, 1.0, k_pad=k_pad)
    assert torch.isfinite(got).all()
    assert relmax(got, ref) < 2e-4
    # what the split buys: plain bf16 operands (no lo terms) are orders of magnitude worse
    xb = x.to(torch.bfloat16).double()
    mb = mus.to(torch.bfloat16).double()
    assert relmax(o_pot.vg_gmm_V(xb, mb, 1.0), ref) > 10 * relmax(got, ref)


def test_split_operand_gmm_gradient_far_from_all_centres():
    """Particles tens of sigma away from every centre: the Gaussian weights underflow in fp32, the max-subtracted
    logits do not (the row constant -c |x|^2 / 2 is never formed)."""
    g = torch.Generator().manual_seed(5)
    mus = torch.rand(64, 32, generator=g, dtype=torch.float64) * 8 - 4
    x = torch.randn(512, 32, generator=g, dtype=torch.float64) * 12.0
    got = gmm_grad_tensor_model(x, mus, 1.0, k_pad=64)
    assert torch.isfinite(got).all()
    assert relmax(got, o_pot.vg_gmm_V(x, mus, 1.0)) < 5e-4


def test_sigma_scaling():
    g = torch.Generator().manual_seed(6)
    mus = torch.rand(16, 8, generator=g, dtype=torch.float64) * 8 - 4
    x = torch.randn(1000, 8, generator=g, dtype=torch.float64) * 2.0
    for sigma in (0.7, 1.0, 2.5):
        assert relmax(gmm_grad_tensor_model(x, mus, sigma, k_pad=16), o_pot.vg_gmm_V(x, mus, sigma)) < 3e-4
