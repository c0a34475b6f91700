"""GPU parity against the committed golden vectors (tests/golden/*.npz, float64 oracle outputs)."""
import os

import numpy as np
import pytest
import torch

from conftest import relmax

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-5


def _ops():
    from pde_inverse_problem_b200 import ops, _lib
    return ops, _lib


def _load(name, cuda):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return z, (lambda k: torch.as_tensor(z[k], dtype=torch.float32, device=cuda).contiguous())


def test_golden_integrator(cuda):
    ops, L = _ops()
    z, T = _load("integrator_gmm", cuda)
    zl, tr, tau = ops.kl_integrate(T("z0"), int(z["S"]), float(z["dt"]), float(z["gamma"]), L.DRIFT_GMM, T("mus"),
                                   n_gaussian=z["mus"].shape[0], noise=T("noise"), tau0=T("tau0"), want_tau=True)
    assert relmax(zl, z["last"]) < TOL and relmax(tr, z["traj"]) < TOL and relmax(tau, z["tau"]) < 1e-6
    val, grd = ops.gmm_value_grad(T("z0")[:, :4].contiguous(), T("mus"), 1.0, want_value=True)
    assert relmax(grd, z["gmm_grad"]) < TOL and relmax(val, z["gmm_value"]) < TOL
    z, T = _load("integrator_ou", cuda)
    zl, tr, _ = ops.kl_integrate(T("z0"), int(z["S"]), float(z["dt"]), float(z["gamma"]), L.DRIFT_LINEAR, T("F"),
                                 noise=T("noise"), tau0=T("tau0"))
    assert relmax(zl, z["last"]) < TOL and relmax(tr, z["traj"]) < TOL


def test_golden_model_eval_and_residuals(cuda):
    ops, L = _ops()
    z, T = _load("model_eval", cuda)
    spec = ops.ModelSpec(L.MODEL_MLP, 4, 32, 2)
    out = ops.model_eval(spec, T("params"), T("x"), T("v"), want=("value", "grad", "vHv", "laplacian"))
    for k, gk in (("value", "value"), ("grad", "grad"), ("vHv", "vHv"), ("laplacian", "lap")):
        assert relmax(out[k], z[gk]) < TOL, k
    z, T = _load("kfp_residual", cuda)
    acc = ops.ResidualAccumulator(spec, device=cuda).begin()
    tg = ops.TrueGrad(L.DRIFT_LINEAR, T("tilde_F"))
    acc.accumulate(L.SET_KFP_0T, T("params"), T("data_0T"), 1.0 / z["data_0T"].shape[0], coef=float(z["gamma"]), true_grad=tg)
    acc.accumulate(L.SET_KFP_BOUNDARY, T("params"), T("data_terminal"), 1.0 / z["data_terminal"].shape[0], coef=2.0 / float(z["T"]))
    acc.accumulate(L.SET_KFP_BOUNDARY, T("params"), T("data_initial"), 1.0 / z["data_initial"].shape[0], coef=-2.0 / float(z["T"]))
    s, g = acc.finalize()
    assert relmax(s[L.SUM_LOSS], z["loss"]) < TOL and relmax(g, z["grad"]) < TOL
    assert relmax(s[L.SUM_GRADNORM], z["grad_norm"]) < TOL and relmax(s[L.SUM_GT], z["loss_gt"]) < TOL
    z, T = _load("fp_residual", cuda)
    acc = ops.ResidualAccumulator(spec, device=cuda).begin()
    tg = ops.TrueGrad(L.DRIFT_LINEAR, T("F"))
    acc.accumulate(L.SET_FP_0T, T("params"), T("data_0T"), 1.0 / z["data_0T"].shape[0], true_grad=tg)
    acc.accumulate(L.SET_FP_BOUNDARY, T("params"), T("data_terminal"), 1.0 / z["data_terminal"].shape[0], coef=2.0 / float(z["T"]))
    acc.accumulate(L.SET_FP_BOUNDARY, T("params"), T("data_initial"), 1.0 / z["data_initial"].shape[0], coef=-2.0 / float(z["T"]))
    s, g = acc.finalize()
    assert relmax(s[L.SUM_LOSS], z["loss"]) < TOL and relmax(g, z["grad"]) < TOL


def test_golden_adam_and_philox(cuda):
    ops, L = _ops()
    from oracle import optim as o_optim
    z, T = _load("adam", cuda)
    p = T("params")[0].clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    sched = o_optim.cosine_decay_schedule(1e-2)
    for i in range(3):
        ops.adam_l2_step(p, T("grads")[i].contiguous(), m, v, count=i + 1, lr=sched(i))
        assert relmax(p, z["params"][i + 1]) < 1e-6
    z, _ = _load("philox", cuda)
    ids = z["ids"]
    dev = ops.philox_normals(len(ids), 3, 6, int(z["seed"]), particle_offset=int(ids[0])).cpu().double().numpy()
    assert np.abs(dev - z["normals"]).max() < 2e-5
    u = ops.philox_uniforms(len(ids), int(z["seed"]), particle_offset=int(ids[0])).cpu().double().numpy()
    assert np.array_equal(u, z["uniforms"])
