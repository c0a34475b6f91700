"""CPU: pin the oracle against analytic known-answer tests and against the
committed golden vectors.  The reference ships no golden vectors for this path, so these KATs are the pins."""
import math
import os

import numpy as np
import pytest
import scipy.integrate
import torch
from torch.func import grad, vmap

from conftest import relmax
from oracle import integrator as o_int, model as o_model, moments as o_mom, optim as o_optim, philox as o_philox
from oracle import potential as o_pot, problems as o_prob, residuals as o_res, sampler as o_s, taylor as o_tay

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_kat1_lyapunov_matches_ode_integration():
    """m' = Fm, P' = FP + PF^T + L (OU.py:78-84): Van Loan expm vs scipy's RK45 on the same ODE."""
    cfg = o_mom.kinetic_ou_configuration(3)
    n = 6

    def rhs(t, y):
        m, P = y[:n], y[n:].reshape(n, n)
        return np.concatenate([cfg["F"] @ m, (cfg["F"] @ P + P @ cfg["F"].T + cfg["L"]).reshape(-1)])

    cfg["m_0"] = np.arange(n) * 0.1
    sol = scipy.integrate.solve_ivp(rhs, (0, 1.3), np.concatenate([cfg["m_0"], cfg["P_0"].reshape(-1)]),
                                    rtol=1e-11, atol=1e-13)
    m, P = o_mom.lyapunov_mean_cov(1.3, cfg)
    assert np.abs(m - sol.y[:n, -1]).max() < 1e-8
    assert np.abs(P - sol.y[n:, -1].reshape(n, n)).max() < 1e-8


def test_kat1_overdamped_closed_form_matches_ode():
    """fokker_planck_example.py:48-55 vs its own disabled check test_OU (:101-116)."""
    cfg = o_mom.overdamped_ou_configuration(4)
    n = 4

    def rhs(t, y):
        m, P = y[:n], y[n:].reshape(n, n)
        return np.concatenate([-cfg["F"] @ m, (-cfg["F"] @ P - P @ cfg["F"] + cfg["L"]).reshape(-1)])

    sol = scipy.integrate.solve_ivp(rhs, (0, 0.7), np.concatenate([cfg["m_0"], cfg["P_0"].reshape(-1)]),
                                    rtol=1e-11, atol=1e-13)
    m, P = o_mom.overdamped_ou_mean_cov(0.7, cfg)
    assert np.abs(m - sol.y[:n, -1]).max() < 1e-7
    assert np.abs(P - sol.y[n:, -1].reshape(n, n)).max() < 1e-7


def test_kat2_discrete_recursion_matches_oracle_ensemble():
    """Exact moments of the reference scheme for linear drift vs an oracle ensemble (Monte-Carlo error)."""
    d, S, T, N = 2, 25, 1.0, 200_000
    dt = T / S
    cfg = o_mom.kinetic_ou_configuration(d)
    g = torch.Generator().manual_seed(0)
    z0 = torch.randn(N, 2 * d, generator=g, dtype=torch.float64)
    noise = torch.randn(N, S, d, generator=g, dtype=torch.float64)
    F = torch.as_tensor(cfg["tilde_F"])
    zT = o_int.fixed_step_scan(z0, S, dt, noise, o_pot.LinearDrift(F).gradient, cfg["gamma_friction"])
    m_d, P_d = o_mom.discrete_mean_cov(S, dt, cfg, tau0=0.0)
    cov = np.cov(zT.numpy().T)
    assert np.linalg.norm(cov - P_d) / np.linalg.norm(P_d) < 1.5e-2
    assert np.abs(zT.mean(0).numpy() - m_d).max() < 1.5e-2
    # and the scheme carries an O(dt) bias to the continuous law
    _, P_c = o_mom.lyapunov_mean_cov(T, cfg)
    assert np.linalg.norm(P_d - P_c) / np.linalg.norm(P_c) > 1e-3


def test_kat2_tau0_schedule_equals_step_composition():
    """step(tau0), (S-1) x step(dt), step(dt - tau0) for one particle vs the matrix recursion with xi = 0."""
    d, S, dt = 2, 5, 0.1
    cfg = o_mom.kinetic_ou_configuration(d)
    z0 = torch.tensor([[0.3, -0.2, 0.5, 0.1]], dtype=torch.float64)
    noise = torch.zeros(1, S + 1, d, dtype=torch.float64)
    tau0 = torch.tensor([0.037], dtype=torch.float64)
    last, traj, tau = o_int.underdamped_langevin_dynamics_scan(
        z0, S, dt, noise, tau0, o_pot.LinearDrift(torch.as_tensor(cfg["tilde_F"])).gradient, 1.0)
    cfg["m_0"], cfg["P_0"] = z0[0].numpy(), np.zeros((4, 4))
    m, _ = o_mom.discrete_mean_cov(S, dt, cfg, tau0=0.037)
    assert np.abs(last[0].numpy() - m).max() < 1e-12
    assert traj.shape == (1, S, 4) and tau.shape == (1, S)
    assert abs(tau[0, 3].item() - (0.037 + 3 * dt)) < 1e-15


def test_kat3_gmm_closed_forms():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(50, 5, generator=g, dtype=torch.float64) * 2
    mus = torch.randn(7, 5, generator=g, dtype=torch.float64) * 2
    assert relmax(o_pot.gmm_gradient_closed_form(x, mus), o_pot.vg_gmm_V(x, mus, 1.0)) < 1e-13
    # K = 1 reduces to x - mu
    assert relmax(o_pot.vg_gmm_V(x, mus[:1], 1.0), x - mus[:1]) < 1e-13
    # v^T grad^2 U v = |v|^2 - Var_w(c), c_k = (x - mu_k).v
    v = torch.randn(50, 5, generator=g, dtype=torch.float64)
    U = lambda xx: o_pot.gmm_V(xx, mus, 1.0)
    vhv = vmap(lambda a, b: torch.dot(b, o_res.hessian_vector_product(U, a, b)))(x, v)
    r = x[:, None] - mus[None]
    w = torch.softmax(-0.5 * (r * r).sum(-1), 1)
    c = (r * v[:, None]).sum(-1)
    closed = (v * v).sum(-1) - ((w * c * c).sum(1) - (w * c).sum(1) ** 2)
    assert relmax(closed, vhv) < 1e-12


def test_kat4_quadratic_parametric_model():
    d = 3
    pde = o_prob.KineticOUProblem(d)
    F = pde.initial_configuration["tilde_F"]
    params = {"params": {"tilde_F": {"kernel": F / 2, "bias": torch.zeros(d, dtype=torch.float64)}}}
    g = torch.Generator().manual_seed(2)
    data = {k: torch.randn(40, 2 * d, generator=g, dtype=torch.float64) for k in ("initial", "terminal", "0T")}
    out = o_res.kfp_value_and_grad_fn(o_model.quadratic_parametric_apply, params, data, pde)
    assert abs(out["loss ground truth"].item()) < 1e-20  # W = F/2, b = 0 is the truth
    # closed-form twin vs autodiff at a generic point
    W = torch.randn(d, d, generator=g, dtype=torch.float64)
    b = torch.randn(d, generator=g, dtype=torch.float64)
    params = {"params": {"tilde_F": {"kernel": W, "bias": b}}}
    ref = o_res.kfp_value_and_grad_fn(o_model.quadratic_parametric_apply, params, data, pde)
    dW = db = 0
    for name, al, be, cg in (("0T", -2.0, 2.0, 1.0), ("terminal", 0.0, 1.0, 0.0), ("initial", 0.0, -1.0, 0.0)):
        z = data[name]
        _, a_, b_, *_ = o_tay.quad_param_point_set(W, b, z[:, :d], z[:, d:], al, be, cg, 1.0 / z.shape[0])
        dW, db = dW + a_, db + b_
    assert relmax(dW, ref["grad"]["params"]["tilde_F"]["kernel"]) < 1e-12
    assert relmax(db, ref["grad"]["params"]["tilde_F"]["bias"]) < 1e-12


def test_kat4_gmm_parametric_twin_matches_autodiff():
    d, K = 3, 4
    pde = o_prob.KineticGMMProblem(d, K)
    g = torch.Generator().manual_seed(3)
    mus = torch.randn(K, d, generator=g, dtype=torch.float64)
    data = {k: torch.randn(60, 2 * d, generator=g, dtype=torch.float64) for k in ("initial", "terminal", "0T")}
    ref = o_res.kfp_value_and_grad_fn(o_model.gmm_parametric_apply, {"params": {"mus": mus}}, data, pde)
    tot = dm = 0
    for name, al, be, cg in (("0T", -2.0, 1.0, 1.0), ("terminal", 0.0, 1.0, 0.0), ("initial", 0.0, -1.0, 0.0)):
        z = data[name]
        val, dmus, *_ = o_tay.gmm_param_point_set(mus, z[:, :d], z[:, d:], al, be, cg, 1.0 / z.shape[0])
        tot, dm = tot + val, dm + dmus
    gt = vmap(grad(pde.V_true_fn))(data["0T"][:, :d])
    assert relmax(tot + (gt ** 2).sum(-1).mean(), ref["loss"]) < 1e-12
    assert relmax(dm, ref["grad"]["params"]["mus"]) < 1e-12


def test_kat5_loss_is_unbiased_estimator_of_ground_truth():
    """With exact kinetic-OU samples, E[loss] = E["loss ground truth"]."""
    d, T, n = 2, 1.0, 60_000
    pde = o_prob.KineticOUProblem(d, T=T)
    g = torch.Generator().manual_seed(4)
    params = o_model.init_mlp_params(d, 32, 2)
    W, b = o_tay.unpack(params)
    n_t = 60
    ts = (torch.rand(n_t, generator=g, dtype=torch.float64) * T).tolist()
    zs = []
    for t in ts:
        m, P = pde.get_mean_cov(t)
        zs.append(o_s.gaussian_sample(m, P, torch.randn(n // n_t, 2 * d, generator=g, dtype=torch.float64)))
    mT, PT = pde.get_mean_cov(T)
    data = {"0T": torch.cat(zs),
            "initial": torch.randn(n, 2 * d, generator=g, dtype=torch.float64),
            "terminal": o_s.gaussian_sample(mT, PT, torch.randn(n, 2 * d, generator=g, dtype=torch.float64))}
    gt = data["0T"][:, :d] @ pde.initial_configuration["tilde_F"].T
    out = o_tay.kfp_value_and_grad(params, data, 1.0, T, gt)
    scale = max(abs(out["loss ground truth"].item()), 1.0)
    assert abs(out["loss"].item() - out["loss ground truth"].item()) / scale < 0.1


def test_kat6_finite_differences_of_model_derivatives():
    d = 3
    params = o_model.init_mlp_params(d, 32, 2)
    V = lambda xx: o_model.mlp_apply(params, xx)[0]
    x = torch.tensor([0.3, -0.7, 0.2], dtype=torch.float64)
    v = torch.tensor([0.5, 0.1, -0.4], dtype=torch.float64)
    eps = 1e-5
    g_fd = torch.stack([(V(x + eps * e) - V(x - eps * e)) / (2 * eps) for e in torch.eye(d, dtype=torch.float64)])
    assert relmax(grad(V)(x), g_fd) < 1e-8
    hv_fd = (grad(V)(x + eps * v) - grad(V)(x - eps * v)) / (2 * eps)
    assert relmax(o_res.hessian_vector_product(V, x, v), hv_fd) < 1e-7
    # parameter gradient of the KFP loss by central differences on a few entries
    pde = o_prob.KineticOUProblem(d)
    g = torch.Generator().manual_seed(5)
    data = {k: torch.randn(20, 2 * d, generator=g, dtype=torch.float64) for k in ("initial", "terminal", "0T")}
    ref = o_res.kfp_value_and_grad_fn(o_model.mlp_apply, params, data, pde)
    for (layer, leaf, idx) in (("layers_0", "kernel", (1, 5)), ("layers_1", "kernel", (7, 3)), ("layers_2", "bias", (9,))):
        t = params["params"][layer][leaf]
        old = t[idx].item()
        t[idx] = old + 1e-6
        lp = o_res.kfp_value_and_grad_fn(o_model.mlp_apply, params, data, pde)["loss"].item()
        t[idx] = old - 1e-6
        lm = o_res.kfp_value_and_grad_fn(o_model.mlp_apply, params, data, pde)["loss"].item()
        t[idx] = old
        fd = (lp - lm) / 2e-6
        assert abs(fd - ref["grad"]["params"][layer][leaf][idx].item()) < 1e-5 * max(1.0, abs(fd))


def test_kat7_partial_s_log_density_finite_differences():
    """Restates test_partial_s_log_density.py:257-311: analytic d_s log rho and d_s^2 log rho vs central FD."""
    pde = o_prob.KineticOUProblem(4)
    g = torch.Generator().manual_seed(0)
    xs = torch.rand(3, 4, generator=g, dtype=torch.float64)
    s = 0.1
    for x in xs:
        d1 = pde.partial_s_log_density_fn(s, x).item()
        fd1 = (pde.log_density_x(s + 1e-5, x) - pde.log_density_x(s - 1e-5, x)).item() / 2e-5
        assert abs(d1 - fd1) < 1e-6 * max(1.0, abs(fd1))
        d2 = pde.partial_s2_log_density_fn(s, x).item()
        fd2 = (pde.partial_s_log_density_fn(s + 1e-5, x) - pde.partial_s_log_density_fn(s - 1e-5, x)).item() / 2e-5
        assert abs(d2 - fd2) < 1e-5 * max(1.0, abs(fd2))


def test_kat8_adam_l2_cosine_hand_calculation():
    p = torch.tensor([1.0, -2.0], dtype=torch.float64)
    gr = torch.tensor([0.5, 0.25], dtype=torch.float64)
    st = o_optim.AdamL2State(p)
    sched = o_optim.cosine_decay_schedule(1e-2, 20000, 0.001)
    assert sched(0) == pytest.approx(1e-2)
    assert sched(20000) == pytest.approx(1e-5) and sched(50000) == pytest.approx(1e-5)
    p1 = o_optim.adam_l2_step(p, gr, st, sched)
    gd = gr + 1e-3 * p                       # L2 added before Adam
    m, v = 0.1 * gd, 0.001 * gd * gd
    u = (m / 0.1) / (torch.sqrt(v / 0.001) + 1e-4)   # bias correction with count = 1; eps outside the sqrt
    assert torch.allclose(p1, p - 1e-2 * u, rtol=0, atol=1e-15)
    e = o_optim.ema_update(p, p1)
    assert torch.allclose(e, 0.999 * p + 0.001 * p1)


def test_philox_published_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = o_philox.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == want
    n = o_philox.normals(7, np.arange(100_000), 0, 4)
    assert abs(n.mean()) < 0.01 and abs(n.std() - 1) < 0.01
    u = o_philox.uniform01(7, np.arange(100_000))
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.01


def test_taylor_twin_matches_autodiff_oracle():
    for d in (2, 5):
        pde = o_prob.KineticOUProblem(d)
        params = o_model.init_mlp_params(d, 32, 2)
        g = torch.Generator().manual_seed(d)
        for k in params["params"]:
            b = params["params"][k]["bias"]
            params["params"][k]["bias"] = 0.1 * torch.randn(b.shape, generator=g, dtype=torch.float64)
        data = {k: torch.randn(40, 2 * d, generator=g, dtype=torch.float64) for k in ("initial", "terminal", "0T")}
        ref = o_res.kfp_value_and_grad_fn(o_model.mlp_apply, params, data, pde)
        gt = data["0T"][:, :d] @ pde.initial_configuration["tilde_F"].T
        tw = o_tay.kfp_value_and_grad(params, data, 1.0, 2.0, gt)
        assert relmax(tw["loss"], ref["loss"]) < 1e-12
        assert relmax(o_model.flatten_params(tw["grad"]), o_model.flatten_params(ref["grad"])) < 1e-12
        pde2 = o_prob.OverdampedOUProblem(d)
        data2 = {k: torch.randn(40, d, generator=g, dtype=torch.float64) for k in ("initial", "terminal", "0T")}
        ref = o_res.fp_value_and_grad_fn(o_model.mlp_apply, params, data2, pde2)
        tw = o_tay.fp_value_and_grad(params, data2, 5.0, data2["0T"] @ pde2.initial_configuration["F"].T)
        assert relmax(tw["loss"], ref["loss"]) < 1e-12
        assert relmax(o_model.flatten_params(tw["grad"]), o_model.flatten_params(ref["grad"])) < 1e-12


def test_offline_subsampler_indexing():
    ds = torch.arange(20 * 15 * 2, dtype=torch.float32).reshape(20, 15, 2)
    perm = torch.tensor([3, 0, 19, 7, 5, 1, 2, 4, 6, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18])
    out = o_s.offline_subsample_0T(ds, 2, perm)
    assert out.shape == (4 * 3, 2)
    assert torch.equal(out[0], ds[3, 2]) and torch.equal(out[1], ds[3, 7]) and torch.equal(out[3], ds[0, 2])


@pytest.mark.parametrize("name", ["integrator_gmm", "integrator_ou", "model_eval", "kfp_residual", "fp_residual",
                                  "adam", "philox"])
def test_oracle_reproduces_golden_vectors(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    T = lambda k: torch.as_tensor(z[k])
    if name.startswith("integrator"):
        drift = (o_pot.GMMPotential(T("mus"), 1.0) if name.endswith("gmm") else o_pot.LinearDrift(T("F"))).gradient
        last, traj, _ = o_int.underdamped_langevin_dynamics_scan(T("z0"), int(z["S"]), float(z["dt"]), T("noise"),
                                                                 T("tau0"), drift, float(z["gamma"]))
        assert relmax(last, z["last"]) < 1e-13 and relmax(traj, z["traj"]) < 1e-13
    elif name == "philox":
        n = np.stack([o_philox.normals(int(z["seed"]), z["ids"], s, 6) for s in range(3)], 1)
        assert np.array_equal(n, z["normals"])
    elif name == "adam":
        p = T("params")[0]
        st = o_optim.AdamL2State(p)
        for i in range(3):
            p = o_optim.adam_l2_step(p, T("grads")[i], st, o_optim.cosine_decay_schedule(1e-2))
            assert relmax(p, z["params"][i + 1]) < 1e-14
    elif name == "model_eval":
        params = _unflatten(T("params"), 4)
        V = lambda xx: o_model.mlp_apply(params, xx)[0]
        assert relmax(vmap(grad(V))(T("x")), z["grad"]) < 1e-13
    else:
        params = _unflatten(T("params"), 4)
        data = {k: T("data_" + k) for k in ("initial", "terminal", "0T")}
        if name == "kfp_residual":
            pde = o_prob.KineticOUProblem(4, T=2.0)
            out = o_res.kfp_value_and_grad_fn(o_model.mlp_apply, params, data, pde)
        else:
            pde = o_prob.OverdampedOUProblem(4, T=5.0)
            out = o_res.fp_value_and_grad_fn(o_model.mlp_apply, params, data, pde)
        assert relmax(out["loss"], z["loss"]) < 1e-13
        assert relmax(o_model.flatten_params(out["grad"]), z["grad"]) < 1e-12


def _unflatten(flat, d, hidden=32, layers=2):
    dims = [d] + [hidden] * layers + [40]
    tree, off = {}, 0
    for i in range(len(dims) - 1):
        nw = dims[i] * dims[i + 1]
        tree[f"layers_{i}"] = {"kernel": flat[off:off + nw].reshape(dims[i], dims[i + 1]),
                               "bias": flat[off + nw:off + nw + dims[i + 1]]}
        off += nw + dims[i + 1]
    return {"params": tree}


# ---------------------------------------------------------------------------------------------------------
# phase schedule of the tcgen05 residual kernel (tests/tensor_v2_model.py) against the hand-derived Taylor twin
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d", [2, 8, 16, 32])
def test_tensor_kernel_schedule_matches_taylor_oracle(d):
    """The kernel's rescaled streams (a2^ = -a2/2, g^ = g/2, za^ = 4 za), the adjoints taken from the input-gradient
    chain, the pz terms and the merged band c = a2^ + ag^ are algebraically identical to SURVEY §9 (float64)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from tensor_v2_model import kfp_0T_schedule
    from oracle import taylor as o_tay, model as o_model
    g = torch.Generator().manual_seed(40 + d)
    p = o_model.init_mlp_params(d, 32, 2, seed=d)
    for k in p["params"]:
        b = p["params"][k]["bias"]
        p["params"][k]["bias"] = 0.1 * torch.randn(b.shape, generator=g, dtype=torch.float64)
    W, b = o_tay.unpack(p)
    n, gamma = 300, 0.7
    z = torch.randn(n, 2 * d, generator=g, dtype=torch.float64)
    val, dW, db, _, gvec = o_tay.point_set(W, b, z[:, :d], [(z[:, d:], -2.0, 2.0 * gamma)], 0.0, 1.0, 1.0 / n)
    r = kfp_0T_schedule(W, b, z[:, :d], z[:, d:], gamma, 1.0 / n)
    assert abs(float(val) - float(r["loss"])) <= 1e-12 * abs(float(val))
    assert (gvec - r["g"]).abs().max() <= 1e-12 * gvec.abs().max()
    for l in range(3):
        assert (dW[l] - r["dW"][l]).abs().max() <= 1e-12 * dW[l].abs().max()
        assert (db[l] - r["db"][l]).abs().max() <= 1e-12 * db[l].abs().max()
    # masked tail points contribute nothing
    mask = (torch.arange(n) < n - 37).double()
    rm = kfp_0T_schedule(W, b, z[:, :d], z[:, d:], gamma, 1.0 / n, mask=mask)
    rs = kfp_0T_schedule(W, b, z[: n - 37, :d], z[: n - 37, d:], gamma, 1.0 / n)
    assert abs(float(rm["loss"]) - float(rs["loss"])) <= 1e-12 * abs(float(rs["loss"]))
    for l in range(3):
        assert (rm["dW"][l] - rs["dW"][l]).abs().max() <= 1e-12 * rs["dW"][l].abs().max()
        assert (rm["db"][l] - rs["db"][l]).abs().max() <= 1e-12 * rs["db"][l].abs().max()


def test_tensor_kernel_schedule_bf16_emulation_within_tolerance():
    """With every operand rounded to bf16 where the kernel rounds it, the schedule stays within the 1e-2 tolerance of
    BASELINE.json for the bf16 GEMM path (per-tensor max-norm metric)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from tensor_v2_model import kfp_0T_schedule
    from oracle import taylor as o_tay, model as o_model
    d, n, gamma = 8, 20000, 0.5
    g = torch.Generator().manual_seed(3)
    p = o_model.init_mlp_params(d, 32, 2, seed=3)
    W, b = o_tay.unpack(p)
    scale = torch.cat([torch.full((d,), 2.0), torch.full((d,), 0.6)]).double()
    z = torch.randn(n, 2 * d, generator=g, dtype=torch.float64) * scale
    val, dW, db, _, _ = o_tay.point_set(W, b, z[:, :d], [(z[:, d:], -2.0, 2.0 * gamma)], 0.0, 1.0, 1.0 / n)
    r = kfp_0T_schedule(W, b, z[:, :d], z[:, d:], gamma, 1.0 / n, emulate_bf16=True)
    ref = torch.cat([torch.cat([w_.reshape(-1), b_.reshape(-1)]) for w_, b_ in zip(dW, db)])
    got = torch.cat([torch.cat([w_.reshape(-1), b_.reshape(-1)]) for w_, b_ in zip(r["dW"], r["db"])])
    assert abs(float(val) - float(r["loss"])) < 1e-2 * abs(float(val))
    assert (ref - got).abs().max() < 1e-2 * ref.abs().max()


def test_oracle_matches_reference_golden_when_present():
    """The pin that is one command away (tests/golden/make_golden_from_reference.py): when ref_*.npz files produced by the
    REAL JAX reference are present, every committed golden output (oracle float64) must agree with them."""
    import glob
    import os
    import numpy as np
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    refs = sorted(glob.glob(os.path.join(here, "ref_*.npz")))
    if not refs:
        pytest.skip("no reference-generated golden vectors (jax is not installable in this image): parity unpinned")
    for path in refs:
        ref = np.load(path)
        ours = np.load(os.path.join(here, os.path.basename(path)[4:]))
        for key in ref.files:
            a, b = np.asarray(ours[key], dtype=np.float64), np.asarray(ref[key], dtype=np.float64)
            assert a.shape == b.shape, (path, key)
            assert np.abs(a - b).max() <= 1e-9 * max(1.0, np.abs(b).max()), (path, key)
