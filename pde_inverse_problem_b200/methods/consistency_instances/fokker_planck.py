"""Mirror of methods/consistency_instances/fokker_planck.py (value_and_grad_fn :33-63, test_fn :66-85,
create_model_fn :89-100) on the fused CUDA residual (exact Laplacian from d tangent streams)."""
from __future__ import annotations

import torch

from ... import _lib as L
from ...core.model import get_model, model_of
from ...utils import rng as jrandom
from . import common


def value_and_grad_fn(forward_fn, params, data, rng, pde_instance, path=None):
    """loss = mean|gV|^2 - 2 mean Laplacian V + mean|gV_true|^2 + (2 mean V(x_T) - 2 mean V(x_0)) / T
    (fokker_planck.py:47-53); data[k] are [n, d] CUDA tensors."""
    model = model_of(forward_fn)
    flat = model.flat(params)
    T = float(pde_instance.total_evolving_time)
    acc = common.accumulator_for(model, flat.device).begin()
    acc.accumulate(L.SET_FP_0T, flat, data["0T"], 1.0 / data["0T"].shape[0], true_grad=pde_instance.true_grad_spec())
    acc.accumulate(L.SET_FP_BOUNDARY, flat, data["terminal"], 1.0 / data["terminal"].shape[0], coef=2.0 / T)
    acc.accumulate(L.SET_FP_BOUNDARY, flat, data["initial"], 1.0 / data["initial"].shape[0], coef=-2.0 / T)
    sums, grad = acc.finalize()
    return common.result_dict(model, params, sums, grad)


def test_fn(forward_fn, pde_instance, rng):
    """fokker_planck.py:66-85: relative L2 error of grad V on 10k initial / terminal samples.
    `forward_fn` is functools.partial(net.apply, params) (methods/consistency.py:28)."""
    model = model_of(forward_fn)
    params = forward_fn.args[0]
    rng_initial, rng_terminal = jrandom.split(rng, 2)
    out = {}
    for name, dist_, key in (("initial", pde_instance.distribution_initial, rng_initial),
                             ("terminal", pde_instance.distribution_terminal, rng_terminal)):
        data = dist_.sample(10000, key)
        pred = model.gradient(params, data)
        true = pde_instance.potential.gradient(data)
        # two scalar reductions of [10000, d] tensors (test-time metric, not the training hot path)
        err = torch.sqrt(torch.mean(torch.sum((pred - true) ** 2, -1)) / torch.mean(torch.sum(true ** 2, -1)))
        out[f"relative error of gradient estimation {name}"] = err
    return out


def create_model_fn(pde_instance):
    """fokker_planck.py:89-100.  (The reference calls get_model(cfg, DEBUG=False) without pde_instance, which
    only works for estimation_mode == "non-parametric": defect D3 of SURVEY.md §2.3.)"""
    net = get_model(pde_instance.cfg, DEBUG=False, pde_instance=pde_instance)
    if net is None:
        raise NotImplementedError("FokkerPlanck has no parametric model (reference defect D3): "
                                  "use estimation_mode='non-parametric'")
    x = pde_instance.distribution_initial.sample(1, jrandom.PRNGKey(1))[0]
    params = net.init(jrandom.PRNGKey(11), x)
    return net, params
