"""Mirror of methods/consistency_instances/kinetic_fokker_planck.py (value_and_grad_fn :11-69, test_fn :72-92,
create_model_fn :96-104) on the fused CUDA residual."""
from __future__ import annotations

import torch

from ... import _lib as L
from ...core.model import get_model, model_of
from ...utils import rng as jrandom
from . import common


def value_and_grad_fn(forward_fn, params, data, rng, pde_instance, path=None):
    """loss = mean|gV|^2 - 2 mean v'Hv + 2 gamma mean gV.v + mean|gV_true|^2
              + (2 mean gV(x_T).v_T - 2 mean gV(x_0).v_0) / T          (kinetic_fokker_planck.py:33-50)
    `data[k]` are [n, 2d] CUDA tensors (x first d, v last d, :13-15).  Returns the reference's dict (:64-69)."""
    model = model_of(forward_fn)
    flat = model.flat(params)
    path = common.DEFAULT_PATH["path"] if path is None else path
    gamma = float(pde_instance.initial_configuration["gamma_friction"])
    T = float(pde_instance.total_evolving_time)
    acc = common.accumulator_for(model, flat.device).begin()
    acc.accumulate(L.SET_KFP_0T, flat, data["0T"], 1.0 / data["0T"].shape[0], coef=gamma,
                   true_grad=pde_instance.true_grad_spec(), path=path)
    acc.accumulate(L.SET_KFP_BOUNDARY, flat, data["terminal"], 1.0 / data["terminal"].shape[0], coef=2.0 / T, path=path)
    acc.accumulate(L.SET_KFP_BOUNDARY, flat, data["initial"], 1.0 / data["initial"].shape[0], coef=-2.0 / T, path=path)
    sums, grad = acc.finalize()
    return common.result_dict(model, params, sums, grad)


def test_fn(forward_fn, pde_instance, rng):
    """kinetic_fokker_planck.py:72-92: the reference's body is commented out and returns {}."""
    return {}


def create_model_fn(pde_instance):
    """kinetic_fokker_planck.py:96-104."""
    net = get_model(pde_instance.cfg, DEBUG=False, pde_instance=pde_instance)
    z = pde_instance.distribution_initial.sample(1, jrandom.PRNGKey(1))[0]
    x = z[: pde_instance.dim]
    params = net.init(jrandom.PRNGKey(11), x)
    return net, params
