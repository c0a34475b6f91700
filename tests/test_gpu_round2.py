"""Round-2 GPU parity tests: mean-field drift (frozen mean and the per-step ensemble-mean table), KMV reference sets /
moment closure / device-side density terms, the FP 0T set on the tcgen05 kernel, zero-padded hidden widths and deep
stacks, the torch custom-op registration, and the NaN poisoning of a timed-out tensor phase."""
import numpy as np
import pytest
import torch

from conftest import relmax
from oracle import integrator as o_int, model as o_model, potential as o_pot, problems as o_prob
from oracle import residuals as o_res, taylor as o_tay

pytestmark = pytest.mark.gpu


def _ops():
    from pde_inverse_problem_b200 import ops, _lib
    return ops, _lib


def _spd(d, seed):
    g = torch.Generator().manual_seed(seed)
    F = torch.randn(d, d + 1, generator=g, dtype=torch.float64)
    return (F @ F.T) / d


# ---------------------------------------------------------------------------------------------------------------
# mean-field drift
# ---------------------------------------------------------------------------------------------------------------
def test_meanfield_frozen_mean_vs_oracle(cuda):
    """PDEIP_DRIFT_MEANFIELD (grad U = A (x - xbar), xbar in the parameter block) against the oracle on injected noise:
    linear contractive drift -> the whole trajectory holds rtol 1e-5."""
    ops, L = _ops()
    d, n, S, T, gamma = 6, 700, 60, 1.5, 1.0
    dt = T / S
    g = torch.Generator().manual_seed(3)
    A = _spd(d, 5)
    xbar = torch.randn(d, generator=g, dtype=torch.float64)
    z0 = torch.randn(n, 2 * d, generator=g, dtype=torch.float64)
    noise = torch.randn(n, S + 1, d, generator=g, dtype=torch.float64)
    tau0 = torch.rand(n, generator=g, dtype=torch.float64) * dt
    last, traj, _ = o_int.underdamped_langevin_dynamics_scan(z0, S, dt, noise, tau0, lambda x: (x - xbar) @ A.T, gamma)
    params = torch.cat([A.reshape(-1), xbar]).float().to(cuda)
    zl, tr, _ = ops.kl_integrate(z0.float().to(cuda), S, dt, gamma, L.DRIFT_MEANFIELD, params,
                                 noise=noise.float().to(cuda), tau0=tau0.float().to(cuda))
    assert relmax(tr, traj) < 1e-5 and relmax(zl, last) < 1e-5


def _oracle_interacting(z0, S, dt, gamma, A, noise):
    return o_int.interacting_langevin_scan(z0, S, dt, noise, A, gamma)


@pytest.mark.parametrize("d,n,S,fast", [(4, 1024, 50, True), (6, 900, 40, False), (16, 2048, 30, True)])
def test_meanfield_table_equals_per_step_empirical_mean(cuda, d, n, S, fast):
    """PDEIP_DRIFT_MEANFIELD_TABLE: the closed recursion of the ensemble mean (one noise pre-pass, no per-step exchange)
    against the oracle that recomputes the empirical mean EVERY step — the mean table, the trajectory and the emitted
    drift.  `fast`: the production configuration (BLOCK128 trajectory with grad U, kl_integrate_fast_kernel)."""
    ops, L = _ops()
    T, gamma, seed = 1.0, 1.0, 41
    dt = T / S
    A = _spd(d, 7 + d)
    g = torch.Generator().manual_seed(d)
    z0 = (torch.randn(n, 2 * d, generator=g, dtype=torch.float64) + 0.7).float()  # non-zero mean: the table matters
    z0c = z0.to(cuda)
    # two "ranks": the sums of both halves add up to the global sums (what the all-reduce does)
    half = n // 2
    sums = ops.meanfield_noise_sums(z0c[:half].contiguous(), S, seed, particle_offset=0)
    ops.meanfield_noise_sums(z0c[half:].contiguous(), S, seed, particle_offset=half, out=sums)
    params, xbar = ops.meanfield_drift_params(sums, n, A.float().to(cuda), S, dt, gamma)
    noise = ops.philox_normals(n, S + 1, d, seed=seed, device=cuda).double().cpu()
    last, traj, xbars, drifts = _oracle_interacting(z0.double(), S, dt, gamma, A, noise)
    assert relmax(xbar, xbars) < 1e-5
    if fast:
        zl, tr, _ = ops.kl_integrate(z0c, S, dt, gamma, L.DRIFT_MEANFIELD_TABLE, params, seed=seed,
                                     traj_layout=L.TRAJ_BLOCK128, emit_drift=True)
        got = tr.permute(1, 3, 0, 2).reshape(n, S, 3 * d)
        assert relmax(got[..., : 2 * d], traj) < 1e-5
        assert relmax(got[..., 2 * d:], drifts) < 1e-5
    else:
        zl, tr, _ = ops.kl_integrate(z0c, S, dt, gamma, L.DRIFT_MEANFIELD_TABLE, params, seed=seed)
        assert relmax(tr, traj) < 1e-5
    assert relmax(zl, last) < 1e-5
    # the ensemble mean of the integrated particles IS the table (closed recursion == empirical mean)
    emp = tr.permute(1, 3, 0, 2).reshape(n, S, 3 * d)[:, -1, :d].mean(0) if fast else tr[:, -1, :d].mean(0)
    assert (emp.double().cpu() - xbars[S].double()).abs().max() < 1e-5 * max(1.0, xbars.abs().max().item())


def test_meanfield_hotpath_step_vs_oracle(cuda):
    """HotPath with the interacting drift (noise pre-pass -> mean table -> integrate -> residual with the emitted
    A (x - xbar_t) as grad V_true) against the float64 oracle that recomputes the empirical mean every step and feeds
    the closed-form residual twin.  Two chunks; fp32 path; linear contractive dynamics: the 1e-5 class holds end to end."""
    ops, L = _ops()
    from pde_inverse_problem_b200.core.model import V_hypothesis
    from pde_inverse_problem_b200.pipeline import HotPath, HotPathConfig
    d, n, S, T, gamma, seed = 8, 512, 24, 1.2, 1.0, 13
    dt = T / S
    A = _spd(d, 2)
    g = torch.Generator().manual_seed(4)
    z0 = (torch.randn(n, 2 * d, generator=g, dtype=torch.float64) + 0.4).float()
    model = V_hypothesis(1, [32, 32], d)
    params = model.init(11, torch.zeros(d, device=cuda))
    params["_flat"].add_(0.05 * torch.randn(params["_flat"].shape, generator=g).to(cuda))
    cfg = HotPathConfig(d=d, n_steps=S, total_time=T, gamma=gamma, drift_kind=L.DRIFT_MEANFIELD_TABLE, chunk=256,
                        path=L.PATH_FP32)
    hp = HotPath(cfg, model, params, A.float().to(cuda).contiguous(), ops.TrueGrad(L.DRIFT_IN_POINTS), device=cuda)
    out = hp.step(z0.to(cuda), seed=seed, apply_optimizer=False)
    noise = ops.philox_normals(n, S + 1, d, seed=seed, device=cuda).double().cpu()
    last, traj, xbars, drifts = _oracle_interacting(z0.double(), S, dt, gamma, A, noise)
    assert relmax(hp.xbar, xbars) < 1e-5
    p64 = o_model.init_mlp_params(d, 32, 2)
    off, flat = 0, params["_flat"].double().cpu()
    for i in range(3):
        for name in ("kernel", "bias"):
            leaf = p64["params"][f"layers_{i}"][name]
            p64["params"][f"layers_{i}"][name] = flat[off:off + leaf.numel()].reshape(leaf.shape).clone()
            off += leaf.numel()
    data = {"initial": z0.double(), "terminal": last, "0T": traj.reshape(-1, 2 * d)}
    ref = o_tay.kfp_value_and_grad(p64, data, gamma, T, grad_true_0T=drifts.reshape(-1, d))
    e = dict(loss=relmax(out["loss"], ref["loss"]), gt=relmax(out["loss ground truth"], ref["loss ground truth"]),
             grad=relmax(out["grad"], o_model.flatten_params(ref["grad"])))
    print("mean-field HotPath.step vs oracle:", {k: f"{v:.2e}" for k, v in e.items()})
    assert e["loss"] < 1e-5 and e["gt"] < 1e-5 and e["grad"] < 1e-5, e


# ---------------------------------------------------------------------------------------------------------------
# KMV: device-side density terms, reference sets, moment closure
# ---------------------------------------------------------------------------------------------------------------
def _kmv_problem(cuda, d):
    from pde_inverse_problem_b200 import registry
    from pde_inverse_problem_b200.config import make_config
    cfg = make_config("kinetic_mckean_vlasov", **{"pde_instance.domain_dim": d, "estimation_mode": "parametric"})
    return registry.get_pde_instance(cfg)(cfg=cfg, rng=1, device=cuda)


@pytest.mark.parametrize("d,nt", [(2, 1), (4, 3)])
def test_kmv_density_terms_on_device_vs_oracle(cuda, d, nt):
    """d_s log rho and d_ss log rho (kinetic_mckean_vlasov_example_quadratic.py:51-69,120-177) evaluated by
    pdeip_kmv_density_terms against the oracle's float64 formulas with the SAME law (tilde_F of the mirror)."""
    ops, L = _ops()
    pde = _kmv_problem(cuda, d)
    pde.np_configuration["m_0"] = np.linspace(-0.5, 0.8, 2 * d)  # non-zero mean: exercise the linear terms too
    opde = o_prob.KineticOUProblem(d, T=2.0)
    cfg = pde.np_configuration
    opde.np_cfg = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in cfg.items()}
    opde.initial_configuration = {k: (torch.as_tensor(v) if isinstance(v, np.ndarray) else v) for k, v in cfg.items()}
    n = 500
    g = torch.Generator().manual_seed(d)
    xv = torch.randn(n, nt, 2 * d, generator=g, dtype=torch.float64)
    tau = [0.3 + 0.5 * t for t in range(nt)]
    c, ps, ps2 = ops.kmv_density_terms(xv.float().to(cuda), pde.density_coefficients(tau, cuda), 1.0, want_parts=True)
    for t in range(nt):
        r1 = opde.partial_s_log_density_fn(torch.tensor(tau[t], dtype=torch.float64), xv[:, t, :d])
        r2 = opde.partial_s2_log_density_fn(torch.tensor(tau[t], dtype=torch.float64), xv[:, t, :d])
        assert relmax(ps[t], r1) < 1e-5 and relmax(ps2[t], r2) < 1e-5
        assert relmax(c[t], r2 + r1 ** 2 + r1) < 1e-5


def _kmv_case(cuda, model_kind, d, nt, n, seed):
    ops, L = _ops()
    pde = o_prob.KineticOUProblem(d, T=2.0)
    g = torch.Generator().manual_seed(seed)
    data0T = torch.randn(n * nt, 2 * d, generator=g, dtype=torch.float64)
    tau = torch.linspace(0.2, 1.4, nt, dtype=torch.float64)
    if model_kind == "mlp":
        params = o_model.init_mlp_params(d, 32, 2)
        apply_fn, spec = o_model.mlp_apply, ops.ModelSpec(L.MODEL_MLP, d, 32, 2)
    else:
        W = torch.randn(d, d, generator=g, dtype=torch.float64)
        b = torch.randn(d, generator=g, dtype=torch.float64)
        params = {"params": {"tilde_F": {"kernel": W, "bias": b}}}
        apply_fn, spec = o_model.quadratic_parametric_apply, ops.ModelSpec(L.MODEL_QUADRATIC, d)
    x3 = data0T[:, :d].reshape(-1, nt, d)
    psl = torch.stack([pde.partial_s_log_density_fn(tau[t], x3[:, t]) for t in range(nt)], 0).reshape(-1, nt)
    ps2l = torch.stack([pde.partial_s2_log_density_fn(tau[t], x3[:, t]) for t in range(nt)], 0).reshape(-1, nt)
    c = ps2l + psl ** 2 + 1.0 * psl

    class M:
        pass
    model = M()
    model.spec = spec
    flat = o_model.flatten_params(params).float().to(cuda)
    xv = data0T.reshape(n, nt, 2 * d).float().to(cuda)

    def run(**kw):
        acc = ops.ResidualAccumulator(spec, device=cuda)
        return ops.kmv_value_and_grad(model, params, flat, xv, c.float().contiguous().to(cuda),
                                      pde.initial_configuration["tilde_F"].float().to(cuda), acc,
                                      lambda m_, p_, s, gr: (s.cpu().double(), gr.cpu().double()), **kw)
    return pde, apply_fn, params, data0T, tau, run


@pytest.mark.parametrize("model_kind,d,nt,m", [("mlp", 2, 1, 11), ("quadratic", 3, 2, 9)])
def test_kmv_subsampled_reference_set_vs_oracle(cuda, model_kind, d, nt, m):
    """Reference set = the first m < n trajectories against the oracle with ref_0T = x_0T[:m]."""
    _, L = _ops()
    n = 41
    pde, apply_fn, params, data0T, tau, run = _kmv_case(cuda, model_kind, d, nt, n, seed=d + nt + m)
    ref = o_res.kmv_value_and_grad_fn(apply_fn, params, {"0T": data0T, "tau_0T": tau}, pde, m=m)
    sums, gflat = run(m=m)
    assert relmax(sums[L.SUM_LOSS], ref["loss"]) < 1e-5
    assert relmax(sums[L.SUM_GT], ref["loss ground truth"]) < 1e-5
    assert relmax(gflat, o_model.flatten_params(ref["grad"])) < 1e-5


@pytest.mark.parametrize("d,nt,n", [(2, 1, 53), (4, 2, 37)])
def test_kmv_moment_closure_equals_full_pair_set(cuda, d, nt, n):
    """Quadratic interaction model: the pair set against the reference MEAN plus the covariance correction (O(n))
    reproduces the full m = n pair set (O(n^2)) — against the oracle's pairwise restatement and against the GPU pair
    kernel."""
    _, L = _ops()
    pde, apply_fn, params, data0T, tau, run = _kmv_case(cuda, "quadratic", d, nt, n, seed=100 + d)
    ref = o_res.kmv_value_and_grad_fn(apply_fn, params, {"0T": data0T, "tau_0T": tau}, pde)
    s_full, g_full = run()
    s_cl, g_cl = run(closure=True)
    for s_x, g_x in ((s_full, g_full), (s_cl, g_cl)):
        assert relmax(s_x[L.SUM_LOSS], ref["loss"]) < 1e-5
        assert relmax(s_x[L.SUM_GT], ref["loss ground truth"]) < 1e-5
        assert relmax(g_x, o_model.flatten_params(ref["grad"])) < 1e-5


def test_kmv_moment_closure_at_scale(cuda):
    """2^18 samples, d = 16 (C4 shape): the closed form runs in O(n); the pair set would be 6.9e10 pairs.  Checked
    against the float64 closed form evaluated on the host."""
    ops, L = _ops()
    d, n, nt = 16, 1 << 18, 1
    g = torch.Generator().manual_seed(1)
    W = torch.randn(d, d, generator=g, dtype=torch.float64) / d
    b = torch.randn(d, generator=g, dtype=torch.float64) * 0.1
    A = _spd(d, 3)
    xv = torch.randn(n, nt, 2 * d, generator=g, dtype=torch.float64)
    c = torch.randn(n, nt, generator=g, dtype=torch.float64) * 0.3
    x, v = xv[:, 0, :d], xv[:, 0, d:]
    rbar = x.mean(0)
    C = (x - rbar).T @ (x - rbar) / n
    y = x - rbar
    Wp = W.clone().requires_grad_(True)
    bp = b.clone().requires_grad_(True)
    G = y @ (Wp + Wp.T).T + bp
    loss = (G ** 2).sum(-1).mean() - 2 * (2 * ((v @ Wp) * v).sum(-1)).mean() \
        + 2 * ((((y @ Wp) * y).sum(-1) + y @ bp + (Wp * C).sum()) * c[:, 0]).mean() + ((y @ A.T) ** 2).sum(-1).mean()
    loss.backward()

    class M:
        pass
    model = M()
    model.spec = ops.ModelSpec(L.MODEL_QUADRATIC, d)
    flat = torch.cat([W.reshape(-1), b]).float().to(cuda)
    acc = ops.ResidualAccumulator(model.spec, device=cuda)
    sums, gflat = ops.kmv_value_and_grad(model, None, flat, xv.float().to(cuda), c.float().to(cuda), A.float().to(cuda),
                                         acc, lambda m_, p_, s, gr: (s.cpu().double(), gr.cpu().double()), closure=True)
    assert relmax(sums[L.SUM_LOSS], loss.detach()) < 1e-4
    assert relmax(gflat, torch.cat([Wp.grad.reshape(-1), bp.grad])) < 1e-4


# ---------------------------------------------------------------------------------------------------------------
# FP 0T set on the tcgen05 kernel
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,n,layout", [(4, 3000, "aos"), (2, 999, "soa"), (8, 1500, "aos"), (16, 257, "aos")])
def test_fp_0T_tensor_path_matches_oracle(cuda, d, n, layout):
    """fokker_planck.py:33-63 (|g|^2 - 2 Laplacian V, d forward-mode tangents) on the tensor path: the d tangent streams
    are rows of the tcgen05 tile; rtol 1e-2 against the float64 closed-form twin, and against the fp32 kernel."""
    ops, L = _ops()
    p = o_model.init_mlp_params(d, 32, 2, seed=5)
    g = torch.Generator().manual_seed(17 + d)
    for k in p["params"]:
        bb = p["params"][k]["bias"]
        p["params"][k]["bias"] = 0.1 * torch.randn(bb.shape, generator=g, dtype=torch.float64)
    x = torch.randn(n, d, generator=g, dtype=torch.float64) * 1.5
    F = _spd(d, 23)
    W, b = o_tay.unpack(p)
    eye = torch.eye(d, dtype=torch.float64)
    dirs = [(eye[i].expand(n, d), -2.0, 0.0) for i in range(d)]
    val, dW, db, _, gvec = o_tay.point_set(W, b, x, dirs, 0.0, 1.0, 1.0 / n)
    gt = x @ F.T
    ref_loss = val + (gt ** 2).sum(-1).mean()
    ref_gt = ((gt - gvec) ** 2).sum(-1).mean()
    ref_grad = torch.cat([torch.cat([w_.reshape(-1), b_.reshape(-1)]) for w_, b_ in zip(dW, db)])
    spec = ops.ModelSpec(L.MODEL_MLP, d, 32, 2)
    flat = o_model.flatten_params(p).float().to(cuda)
    tg = ops.TrueGrad(L.DRIFT_LINEAR, F.float().to(cuda))
    pts = x.float().to(cuda)
    out = {}
    for name, path in (("fp32", L.PATH_FP32), ("tensor", L.PATH_TENSOR)):
        acc = ops.ResidualAccumulator(spec, device=cuda).begin()
        if layout == "soa":
            acc.accumulate(L.SET_FP_0T, flat, pts.t().contiguous(), 1.0 / n, true_grad=tg, path=path, layout=L.LAYOUT_SOA)
        else:
            acc.accumulate(L.SET_FP_0T, flat, pts, 1.0 / n, true_grad=tg, path=path)
        s, gr = acc.finalize()
        out[name] = (s.cpu().double(), gr.cpu().double())
    assert ops.tensor_path_status() == 0
    s32, g32 = out["fp32"]
    assert relmax(s32[L.SUM_LOSS], ref_loss) < 1e-5 and relmax(g32, ref_grad) < 1e-5
    st, gtc = out["tensor"]
    print(f"FP tensor d={d} n={n}: loss {relmax(st[L.SUM_LOSS], ref_loss):.2e} gt {relmax(st[L.SUM_GT], ref_gt):.2e} "
          f"lap {relmax(st[L.SUM_D2], s32[L.SUM_D2]):.2e} grad {relmax(gtc, ref_grad):.2e}")
    assert relmax(st[L.SUM_LOSS], ref_loss) < 1e-2
    assert relmax(st[L.SUM_GT], ref_gt) < 1e-2
    assert relmax(st[L.SUM_D2], s32[L.SUM_D2]) < 1e-2
    assert relmax(st[L.SUM_G2], s32[L.SUM_G2]) < 1e-2
    assert relmax(gtc, ref_grad) < 1e-2


# ---------------------------------------------------------------------------------------------------------------
# model envelope: zero-padded hidden widths, deep stacks
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("hidden,layers,d", [(20, 8, 4), (20, 3, 2), (32, 6, 3), (7, 1, 5)])
def test_padded_hidden_width_and_deep_stack_vs_oracle(cuda, hidden, layers, d):
    """configurations/neural_network/MLP.yaml (hidden_dim = 20, layers = 8) and other shapes inside the envelope: the
    reference-shaped parameter tree lives in a zero-padded 32-wide buffer; loss, gradient (un-padded leaves) and the
    exactly-zero padding gradient against the autodiff oracle on the un-padded network."""
    ops, L = _ops()
    from pde_inverse_problem_b200.core.model import V_hypothesis
    from pde_inverse_problem_b200.methods.consistency_instances import common
    model = V_hypothesis(1, [hidden] * layers, d)
    params = model.init(11, torch.zeros(d, device=cuda))
    g = torch.Generator().manual_seed(hidden + layers)
    p64 = {"params": {}}
    for i in range(layers + 1):
        leaf = params["params"][f"layers_{i}"]
        leaf["bias"].copy_(0.1 * torch.randn(leaf["bias"].shape, generator=g))
        p64["params"][f"layers_{i}"] = {"kernel": leaf["kernel"].double().cpu().clone(),
                                        "bias": leaf["bias"].double().cpu().clone()}
    assert tuple(params["params"]["layers_0"]["kernel"].shape) == (d, hidden)
    pde = o_prob.KineticOUProblem(d, T=2.0)
    data = {k: torch.randn(nn, 2 * d, generator=g, dtype=torch.float64) for k, nn in
            (("initial", 150), ("terminal", 140), ("0T", 400))}
    ref = o_res.kfp_value_and_grad_fn(o_model.mlp_apply, p64, data, pde)
    flat = model.flat(params)
    tg = ops.TrueGrad(L.DRIFT_LINEAR, pde.initial_configuration["tilde_F"].float().to(cuda))
    acc = ops.ResidualAccumulator(model.spec, device=cuda).begin()
    acc.accumulate(L.SET_KFP_0T, flat, data["0T"].float().to(cuda), 1.0 / 400, coef=1.0, true_grad=tg)
    acc.accumulate(L.SET_KFP_BOUNDARY, flat, data["terminal"].float().to(cuda), 1.0 / 140, coef=1.0)
    acc.accumulate(L.SET_KFP_BOUNDARY, flat, data["initial"].float().to(cuda), 1.0 / 150, coef=-1.0)
    sums, grad = acc.finalize()
    out = common.result_dict(model, params, sums, grad)
    assert relmax(out["loss"], ref["loss"]) < 1e-5
    total, kept = grad.double().pow(2).sum().item(), 0.0
    for i in range(layers + 1):
        for name in ("kernel", "bias"):
            got = out["grad"]["params"][f"layers_{i}"][name]
            want = ref["grad"]["params"][f"layers_{i}"][name]
            assert tuple(got.shape) == tuple(want.shape)
            assert relmax(got, want) < 1e-5, (i, name)
            kept += got.double().pow(2).sum().item()
    assert abs(total - kept) <= 1e-12 * max(total, 1e-30)  # every padding entry of the gradient is exactly zero
    # forward value through the reference-shaped API
    xs = data["0T"][:50, :d]
    val = model.apply(params, xs.float().to(cuda))
    want = torch.stack([o_model.mlp_apply(p64, xx)[0] for xx in xs])
    assert relmax(val, want) < 1e-5


def test_hessian_vector_product_reference_contract(cuda):
    """utils/common_utils.py:6-14: (f, x, v) -> H v."""
    from functools import partial
    from pde_inverse_problem_b200.core.model import V_hypothesis
    from pde_inverse_problem_b200.utils.common_utils import hessian_vector_product
    d = 5
    model = V_hypothesis(1, [32, 32], d)
    params = model.init(11, torch.zeros(d, device=cuda))
    p64 = {"params": {k: {"kernel": v["kernel"].double().cpu(), "bias": v["bias"].double().cpu()}
                      for k, v in params["params"].items()}}
    g = torch.Generator().manual_seed(0)
    x = torch.randn(64, d, generator=g, dtype=torch.float64)
    v = torch.randn(64, d, generator=g, dtype=torch.float64)
    hv = hessian_vector_product(partial(model.apply, params), x.float().to(cuda), v.float().to(cuda))
    V = lambda xx: o_model.mlp_apply(p64, xx)[0]
    want = torch.stack([o_res.hessian_vector_product(V, a, b) for a, b in zip(x, v)])
    assert relmax(hv, want) < 1e-4  # polarisation of two fp32 quadratic forms
    one = hessian_vector_product(partial(model.apply, params), x[0].float().to(cuda), v[0].float().to(cuda))
    assert tuple(one.shape) == (d,) and relmax(one, want[0]) < 1e-4


# ---------------------------------------------------------------------------------------------------------------
# boundary: torch custom ops; failure visibility
# ---------------------------------------------------------------------------------------------------------------
def test_entry_points_are_torch_custom_ops(cuda):
    """The C ABI is reached through torch.ops.pdeip.* (torch.library.custom_op): call two ops directly."""
    ops, L = _ops()
    from pde_inverse_problem_b200 import torch_ops
    for name in torch_ops.REGISTERED:
        assert hasattr(torch.ops.pdeip, name)
    d, K, n = 3, 4, 100
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, d, generator=g, dtype=torch.float64)
    mus = torch.randn(K, d, generator=g, dtype=torch.float64)
    out = torch.empty(n, d, device=cuda)
    torch.ops.pdeip.gmm_value_grad(x.float().to(cuda), mus.float().to(cuda), K, 1.0, None, out, n, d)
    assert relmax(out, o_pot.gmm_gradient_closed_form(x, mus, 1.0)) < 1e-5
    p = torch.ones(10, device=cuda)
    gr, m, v = torch.full((10,), 0.5, device=cuda), torch.zeros(10, device=cuda), torch.zeros(10, device=cuda)
    norms = torch.empty(2, device=cuda)
    torch.ops.pdeip.adam_l2_step(p, gr, m, v, None, 1e-2, 0.9, 0.999, 1e-4, 1e-3, 1, 1.0, 0, 0.999, norms)
    # first Adam step: u = g' / (|g'| + eps), g' = g + wd p
    gp = 0.5 + 1e-3
    assert abs(p[0].item() - (1.0 - 1e-2 * gp / (gp + 1e-4))) < 1e-6
    with pytest.raises(Exception):
        torch.ops.pdeip.linear_grad(torch.zeros(2, 2), torch.zeros(2, 2), torch.zeros(2, 2), 2, 2)  # CPU: no kernel


def test_timed_out_tensor_phase_poisons_loss_and_gradient(cuda):
    """a timed-out tcgen05 phase must not feed a garbage gradient into Adam.  The status words live in
    device memory; setting the residual word by hand must turn sums and grad of the next finalize into NaN, and
    begin must clear it again."""
    import ctypes
    ops, L = _ops()
    d, n = 4, 256
    spec = ops.ModelSpec(L.MODEL_MLP, d, 32, 2)
    flat = o_model.flatten_params(o_model.init_mlp_params(d, 32, 2)).float().to(cuda)
    pts = torch.randn(n, 2 * d, device=cuda)
    acc = ops.ResidualAccumulator(spec, device=cuda).begin()
    acc.accumulate(L.SET_KFP_0T, flat, pts, 1.0 / n, coef=0.5, path=L.PATH_TENSOR)
    lib = L.load()
    lib.pdeip_debug_set_status.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.pdeip_debug_set_status.restype = ctypes.c_int
    assert lib.pdeip_debug_set_status(0, 1, torch.cuda.current_stream().cuda_stream) == 0
    sums, grad = acc.finalize()
    assert torch.isnan(sums[L.SUM_LOSS]) and torch.isnan(grad).all()
    assert ops.tensor_path_status() == 1 and ops.tensor_path_status() == 0  # read-and-clear
    lib.pdeip_debug_set_status(1, 1, torch.cuda.current_stream().cuda_stream)
    acc.begin()  # clears both words
    acc.accumulate(L.SET_KFP_0T, flat, pts, 1.0 / n, coef=0.5, path=L.PATH_TENSOR)
    sums, grad = acc.finalize()
    assert torch.isfinite(sums[L.SUM_LOSS]) and torch.isfinite(grad).all()


# ---------------------------------------------------------------------------------------------------------------
# edge cases of the round-2 paths
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 127, 129, 300])
def test_fp_and_boundary_tensor_modes_edge_sizes(cuda, n):
    """Empty, single-point and tile-boundary point counts through the FP 0T direction tiles and the KFP boundary mode of the
    tcgen05 kernel: agreement with the fp32 kernel (n = 0 is a no-op)."""
    ops, L = _ops()
    d = 4
    spec = ops.ModelSpec(L.MODEL_MLP, d, 32, 2)
    g = torch.Generator().manual_seed(n + 5)
    flat = o_model.flatten_params(o_model.init_mlp_params(d, 32, 2, seed=3)).float().to(cuda)
    x = (torch.randn(n, d, generator=g) * 1.5).to(cuda)
    z = (torch.randn(n, 2 * d, generator=g) * 1.2).to(cuda)
    res = {}
    for name, path in (("fp32", L.PATH_FP32), ("tensor", L.PATH_TENSOR)):
        acc = ops.ResidualAccumulator(spec, device=cuda).begin()
        acc.accumulate(L.SET_FP_0T, flat, x, 1.0 / max(n, 1), path=path)
        s1, g1 = (t.cpu().double().clone() for t in acc.finalize())
        acc.begin()
        acc.accumulate(L.SET_KFP_BOUNDARY, flat, z, 1.0 / max(n, 1), coef=-1.0, path=path)
        s2, g2 = (t.cpu().double().clone() for t in acc.finalize())
        res[name] = (s1, g1, s2, g2)
    assert ops.tensor_path_status() == 0
    if n == 0:
        for t in res["tensor"]:
            assert float(t.abs().sum()) == 0.0
        return
    tol = 1e-2 if n >= 128 else 3e-2  # fewer points than a tile: single-point bf16 rounding, nothing averages
    s1, g1, s2, g2 = res["fp32"]
    t1, h1, t2, h2 = res["tensor"]
    # the loss |g|^2 - 2 Laplacian is a cancelling scalar: its error is measured against the size of its two terms
    terms = abs(float(s1[L.SUM_G2])) + 2.0 * abs(float(s1[L.SUM_D2]))
    assert abs(float(t1[L.SUM_LOSS] - s1[L.SUM_LOSS])) < tol * terms
    assert relmax(t1[L.SUM_G2], s1[L.SUM_G2]) < tol and relmax(t1[L.SUM_D2], s1[L.SUM_D2]) < tol
    assert relmax(h1, g1) < tol
    assert relmax(h2, g2) < tol


@pytest.mark.parametrize("d,n,S", [(3, 77, 7), (5, 130, 1), (2, 1, 4)])
def test_meanfield_table_generic_kernel_ragged(cuda, d, n, S):
    """Mean-field table on the generic integrator: d not a multiple of 4, n not a multiple of 128, a single particle
    (the ensemble mean is the particle itself: the drift vanishes) and a single step."""
    ops, L = _ops()
    T, gamma, seed = 0.5, 0.7, 8
    dt = T / S
    A = _spd(d, 31 + d)
    g = torch.Generator().manual_seed(d * 100 + n)
    z0 = (torch.randn(n, 2 * d, generator=g, dtype=torch.float64) - 0.3).float()
    sums = ops.meanfield_noise_sums(z0.to(cuda), S, seed)
    params, xbar = ops.meanfield_drift_params(sums, n, A.float().to(cuda), S, dt, gamma)
    noise = ops.philox_normals(n, S + 1, d, seed=seed, device=cuda).double().cpu()
    last, traj, xbars, _ = _oracle_interacting(z0.double(), S, dt, gamma, A, noise)
    zl, tr, _ = ops.kl_integrate(z0.to(cuda), S, dt, gamma, L.DRIFT_MEANFIELD_TABLE, params, seed=seed)
    assert relmax(xbar, xbars) < 1e-5 and relmax(tr, traj) < 1e-5 and relmax(zl, last) < 1e-5


def test_kmv_reference_set_of_one_and_full_agree_with_oracle(cuda):
    """m = 1 (a single reference trajectory) and m = n through the same entry points."""
    _, L = _ops()
    n, d, nt = 23, 2, 2
    pde, apply_fn, params, data0T, tau, run = _kmv_case(cuda, "mlp", d, nt, n, seed=77)
    for m in (1, n):
        ref = o_res.kmv_value_and_grad_fn(apply_fn, params, {"0T": data0T, "tau_0T": tau}, pde, m=m)
        sums, gflat = run(m=m)
        assert relmax(sums[L.SUM_LOSS], ref["loss"]) < 1e-5
        assert relmax(gflat, o_model.flatten_params(ref["grad"])) < 1e-5
