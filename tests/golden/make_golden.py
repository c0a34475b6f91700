"""Generates tests/golden/*.npz — small seeded inputs and the float64 oracle outputs for them.

The reference cannot run in this image (no jax/flax/optax), so these are ORACLE outputs ("parity unpinned",
see oracle/__init__.py); they pin the oracle against accidental change and give the GPU tests fixed vectors
that do not depend on torch's CPU RNG stream.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch
from torch.func import grad, vmap

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import integrator as o_int, model as o_model, optim as o_optim, philox as o_philox  # noqa: E402
from oracle import potential as o_pot, problems as o_prob, residuals as o_res, moments as o_mom  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def npy(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def main():
    g = torch.Generator().manual_seed(20261018)
    # ---- integrator, GMM drift ----
    d, K, N, S, T, gamma = 4, 3, 64, 10, 2.0, 0.5
    dt = T / S
    z0 = torch.randn(N, 2 * d, generator=g, dtype=torch.float64)
    noise = torch.randn(N, S + 1, d, generator=g, dtype=torch.float64)
    tau0 = torch.rand(N, generator=g, dtype=torch.float64) * dt
    mus = torch.rand(K, d, generator=g, dtype=torch.float64) * 8 - 4
    last, traj, tau = o_int.underdamped_langevin_dynamics_scan(z0, S, dt, noise, tau0,
                                                               o_pot.GMMPotential(mus, 1.0).gradient, gamma)
    np.savez(os.path.join(OUT, "integrator_gmm.npz"), z0=npy(z0), noise=npy(noise), tau0=npy(tau0), mus=npy(mus),
             last=npy(last), traj=npy(traj), tau=npy(tau), S=S, dt=dt, gamma=gamma,
             gmm_grad=npy(o_pot.vg_gmm_V(z0[:, :d], mus, 1.0)),
             gmm_value=npy(o_pot.GMMPotential(mus, 1.0).value(z0[:, :d])))
    # ---- integrator, linear (OU) drift ----
    cfg = o_mom.kinetic_ou_configuration(d)
    F = torch.as_tensor(cfg["tilde_F"]) / d
    last, traj, tau = o_int.underdamped_langevin_dynamics_scan(z0, S, dt, noise, tau0,
                                                               o_pot.LinearDrift(F).gradient, 1.0)
    np.savez(os.path.join(OUT, "integrator_ou.npz"), z0=npy(z0), noise=npy(noise), tau0=npy(tau0), F=npy(F),
             last=npy(last), traj=npy(traj), S=S, dt=dt, gamma=1.0)
    # ---- model evaluation + KFP / FP residuals ----
    params = o_model.init_mlp_params(d, 32, 2, seed=11)
    for k in params["params"]:
        b = params["params"][k]["bias"]
        params["params"][k]["bias"] = 0.1 * torch.randn(b.shape, generator=g, dtype=torch.float64)
    flat = o_model.flatten_params(params)
    x = torch.randn(96, d, generator=g, dtype=torch.float64)
    v = torch.randn(96, d, generator=g, dtype=torch.float64)
    V = lambda xx: o_model.mlp_apply(params, xx)[0]
    np.savez(os.path.join(OUT, "model_eval.npz"), params=npy(flat), x=npy(x), v=npy(v), value=npy(vmap(V)(x)),
             grad=npy(vmap(grad(V))(x)),
             vHv=npy(vmap(lambda a, b: torch.dot(b, o_res.hessian_vector_product(V, a, b)))(x, v)),
             lap=npy(vmap(lambda a: torch.diagonal(torch.func.jacfwd(grad(V))(a)).sum())(x)))
    pde = o_prob.KineticOUProblem(d, T=2.0)
    data = {k: torch.randn(n, 2 * d, generator=g, dtype=torch.float64) for k, n in
            (("initial", 80), ("terminal", 70), ("0T", 150))}
    ref = o_res.kfp_value_and_grad_fn(o_model.mlp_apply, params, data, pde)
    np.savez(os.path.join(OUT, "kfp_residual.npz"), params=npy(flat), tilde_F=npy(pde.initial_configuration["tilde_F"]),
             gamma=1.0, T=2.0, **{f"data_{k}": npy(v_) for k, v_ in data.items()}, loss=npy(ref["loss"]),
             grad=npy(o_model.flatten_params(ref["grad"])), grad_norm=npy(ref["grad_norm"]),
             loss_gt=npy(ref["loss ground truth"]))
    pde2 = o_prob.OverdampedOUProblem(d, T=5.0)
    data2 = {k: torch.randn(n, d, generator=g, dtype=torch.float64) for k, n in
             (("initial", 80), ("terminal", 70), ("0T", 150))}
    ref = o_res.fp_value_and_grad_fn(o_model.mlp_apply, params, data2, pde2)
    np.savez(os.path.join(OUT, "fp_residual.npz"), params=npy(flat), F=npy(pde2.initial_configuration["F"]), T=5.0,
             **{f"data_{k}": npy(v_) for k, v_ in data2.items()}, loss=npy(ref["loss"]),
             grad=npy(o_model.flatten_params(ref["grad"])), grad_norm=npy(ref["grad_norm"]),
             loss_gt=npy(ref["loss ground truth"]))
    # ---- optimizer: 3 steps of add_decayed_weights + adam with the cosine schedule ----
    p = torch.randn(50, generator=g, dtype=torch.float64)
    st = o_optim.AdamL2State(p)
    sched = o_optim.cosine_decay_schedule(1e-2)
    grads, ps = [], [npy(p)]
    for _ in range(3):
        gr = torch.randn(50, generator=g, dtype=torch.float64)
        p = o_optim.adam_l2_step(p, gr, st, sched)
        grads.append(npy(gr))
        ps.append(npy(p))
    np.savez(os.path.join(OUT, "adam.npz"), grads=np.stack(grads), params=np.stack(ps))
    # ---- Philox normals ----
    np.savez(os.path.join(OUT, "philox.npz"), seed=np.uint64(0x1234567890ABCDEF), ids=np.arange(5, 37),
             normals=np.stack([o_philox.normals(0x1234567890ABCDEF, np.arange(5, 37), s, 6) for s in range(3)], 1),
             uniforms=o_philox.uniform01(0x1234567890ABCDEF, np.arange(5, 37)))
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
