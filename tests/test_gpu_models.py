"""GPU parity: parametric models (K5), the KMV pairwise residual, exact samplers and the reference-shaped
host API (registry -> method -> value_and_grad_fn / trainer) vs the oracle."""
import numpy as np
import pytest
import torch
from torch.func import grad, vmap

from conftest import relmax
from oracle import model as o_model, moments as o_mom, problems as o_prob, residuals as o_res

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _ops():
    from pde_inverse_problem_b200 import ops, _lib
    return ops, _lib


def _kfp_sets(ops, L, spec, flat, data, tg, gamma, T, cuda):
    acc = ops.ResidualAccumulator(spec, device=cuda).begin()
    acc.accumulate(L.SET_KFP_0T, flat, data["0T"].float().to(cuda), 1.0 / data["0T"].shape[0], coef=gamma, true_grad=tg)
    acc.accumulate(L.SET_KFP_BOUNDARY, flat, data["terminal"].float().to(cuda), 1.0 / data["terminal"].shape[0], coef=2.0 / T)
    acc.accumulate(L.SET_KFP_BOUNDARY, flat, data["initial"].float().to(cuda), 1.0 / data["initial"].shape[0], coef=-2.0 / T)
    s, g = acc.finalize()
    return s.cpu().double(), g.cpu().double()


@pytest.mark.parametrize("d,K", [(4, 3), (8, 16), (2, 1), (16, 40), (32, 64), (8, 1), (4, 9)])
def test_gmm_parametric_residual(cuda, d, K):
    """V_parametric of GMM.py:214-234 under kinetic_fokker_planck.py:11-69."""
    ops, L = _ops()
    pde = o_prob.KineticGMMProblem(d, max(K, 2), T=2.0)
    g = torch.Generator().manual_seed(d * 7 + K)
    mus = torch.randn(K, d, generator=g, dtype=torch.float64) * 1.5
    data = {k: torch.randn(n, 2 * d, generator=g, dtype=torch.float64) for k, n in
            (("initial", 333), ("terminal", 410), ("0T", 777))}
    ref = o_res.kfp_value_and_grad_fn(o_model.gmm_parametric_apply, {"params": {"mus": mus}}, data, pde)
    spec = ops.ModelSpec(L.MODEL_GMM, d, n_gaussian=K)
    tg = ops.TrueGrad(L.DRIFT_GMM, pde.mus.float().to(cuda), 1.0)
    sums, gflat = _kfp_sets(ops, L, spec, mus.reshape(-1).float().to(cuda), data, tg, 0.5, 2.0, cuda)
    assert relmax(sums[L.SUM_LOSS], ref["loss"]) < TOL
    assert relmax(sums[L.SUM_GT], ref["loss ground truth"]) < TOL
    assert relmax(gflat, ref["grad"]["params"]["mus"].reshape(-1)) < TOL
    assert relmax(sums[L.SUM_GRADNORM], ref["grad_norm"]) < TOL
    # model_eval of the same model
    x, v = data["0T"][:, :d], data["0T"][:, d:]
    V = lambda xx: o_model.gmm_parametric_apply({"params": {"mus": mus}}, xx)[0]
    out = ops.model_eval(spec, mus.reshape(-1).float().to(cuda), x.float().to(cuda), v.float().to(cuda),
                         want=("value", "grad", "vHv", "laplacian"))
    assert relmax(out["grad"], vmap(grad(V))(x)) < TOL
    assert relmax(out["vHv"], vmap(lambda a, b: torch.dot(b, o_res.hessian_vector_product(V, a, b)))(x, v)) < TOL
    assert relmax(out["laplacian"], vmap(lambda a: torch.diagonal(torch.func.jacfwd(grad(V))(a)).sum())(x)) < TOL


@pytest.mark.parametrize("d,K,n", [(8, 16, 128 * 70), (32, 64, 128 * 9), (4, 3, 128 * 33)])
def test_gmm_parametric_fast_kernel_layouts_and_generic_twin(cuda, d, K, n):
    """The production GMM-model kernel (parametric_fast.cu: d in {4, 8, 16, 32}) on the three point layouts with grad V_true
    stored in the points (what the pipeline feeds it), against the generic kernel (PDEIP_NO_FAST_PARAMETRIC) on the same
    points; ragged tail (n - 37 points) through the AoS layout."""
    import os
    ops, L = _ops()
    g = torch.Generator().manual_seed(d + K)
    mus = (torch.randn(K, d, generator=g) * 1.5).reshape(-1).to(cuda)
    pts3 = (torch.randn(n, 3 * d, generator=g) * 1.3).to(cuda)
    spec = ops.ModelSpec(L.MODEL_GMM, d, n_gaussian=K)
    tg = ops.TrueGrad(L.DRIFT_IN_POINTS)

    def run(points, layout, n_pts):
        acc = ops.ResidualAccumulator(spec, device=cuda).begin()
        acc.accumulate(L.SET_KFP_0T, mus, points, 1.0 / n_pts, coef=0.5, true_grad=tg, layout=layout)
        acc.accumulate(L.SET_KFP_BOUNDARY, mus, pts3[:1000, : 2 * d].contiguous(), 1e-3, coef=1.0)
        s, gr = acc.finalize()
        return s.cpu().double().clone(), gr.cpu().double().clone()

    s_a, g_a = run(pts3, L.LAYOUT_AOS, n)
    s_s, g_s = run(pts3.t().contiguous(), L.LAYOUT_SOA, n)
    s_b, g_b = run(pts3.view(n // 128, 128, 3 * d).permute(0, 2, 1).contiguous(), L.LAYOUT_BLOCK128, n)
    assert torch.equal(s_a, s_s) and torch.equal(g_a, g_s) and torch.equal(s_a, s_b) and torch.equal(g_a, g_b)
    s_r, g_r = run(pts3[: n - 37].contiguous(), L.LAYOUT_AOS, n - 37)
    os.environ["PDEIP_NO_FAST_PARAMETRIC"] = "1"
    try:
        s_g, g_g = run(pts3, L.LAYOUT_AOS, n)
        s_gr, g_gr = run(pts3[: n - 37].contiguous(), L.LAYOUT_AOS, n - 37)
    finally:
        del os.environ["PDEIP_NO_FAST_PARAMETRIC"]
    for (s1, g1), (s2, g2) in (((s_a, g_a), (s_g, g_g)), ((s_r, g_r), (s_gr, g_gr))):
        assert relmax(s1[L.SUM_LOSS], s2[L.SUM_LOSS]) < TOL and relmax(s1[L.SUM_GT], s2[L.SUM_GT]) < TOL
        assert relmax(g1, g2) < TOL


@pytest.mark.parametrize("d,n", [(4, 128 * 50), (8, 128 * 21), (16, 128 * 30), (32, 128 * 7)])
def test_quadratic_parametric_fast_kernel_layouts_and_generic_twin(cuda, d, n):
    """The production quadratic-model kernel (parametric_fast.cu) on the three point layouts with grad V_true stored in
    the points, against the generic kernel (PDEIP_NO_FAST_PARAMETRIC); ragged tail through the AoS layout."""
    import os
    ops, L = _ops()
    g = torch.Generator().manual_seed(3 * d)
    flat = torch.cat([(torch.randn(d, d, generator=g) / d).reshape(-1), 0.2 * torch.randn(d, generator=g)]).to(cuda)
    pts3 = (torch.randn(n, 3 * d, generator=g) * 1.1).to(cuda)
    spec = ops.ModelSpec(L.MODEL_QUADRATIC, d)
    tg = ops.TrueGrad(L.DRIFT_IN_POINTS)

    def run(points, layout, n_pts):
        acc = ops.ResidualAccumulator(spec, device=cuda).begin()
        acc.accumulate(L.SET_KFP_0T, flat, points, 1.0 / n_pts, coef=0.7, true_grad=tg, layout=layout)
        acc.accumulate(L.SET_KFP_BOUNDARY, flat, pts3[:900, : 2 * d].contiguous(), 1.0 / 900, coef=-1.0)
        s, gr = acc.finalize()
        return s.cpu().double().clone(), gr.cpu().double().clone()

    s_a, g_a = run(pts3, L.LAYOUT_AOS, n)
    s_s, g_s = run(pts3.t().contiguous(), L.LAYOUT_SOA, n)
    s_b, g_b = run(pts3.view(n // 128, 128, 3 * d).permute(0, 2, 1).contiguous(), L.LAYOUT_BLOCK128, n)
    assert torch.equal(s_a, s_s) and torch.equal(g_a, g_s) and torch.equal(s_a, s_b) and torch.equal(g_a, g_b)
    s_r, g_r = run(pts3[: n - 55].contiguous(), L.LAYOUT_AOS, n - 55)
    os.environ["PDEIP_NO_FAST_PARAMETRIC"] = "1"
    try:
        s_g, g_g = run(pts3, L.LAYOUT_AOS, n)
        s_gr, g_gr = run(pts3[: n - 55].contiguous(), L.LAYOUT_AOS, n - 55)
    finally:
        del os.environ["PDEIP_NO_FAST_PARAMETRIC"]
    for (s1, g1), (s2, g2) in (((s_a, g_a), (s_g, g_g)), ((s_r, g_r), (s_gr, g_gr))):
        assert relmax(s1[L.SUM_LOSS], s2[L.SUM_LOSS]) < TOL and relmax(s1[L.SUM_GT], s2[L.SUM_GT]) < TOL
        assert relmax(s1[L.SUM_BOUNDARY], s2[L.SUM_BOUNDARY]) < TOL
        assert relmax(g1, g2) < TOL


@pytest.mark.parametrize("d", [2, 4, 8, 16, 32])
def test_quadratic_parametric_residual(cuda, d):
    """V_parametric of OU.py:209-220 under kinetic_fokker_planck.py:11-69; KAT-4: W = F/2, b = 0 -> ground truth 0."""
    ops, L = _ops()
    pde = o_prob.KineticOUProblem(d, T=2.0)
    g = torch.Generator().manual_seed(d)
    W = torch.randn(d, d, generator=g, dtype=torch.float64)
    b = torch.randn(d, generator=g, dtype=torch.float64)
    data = {k: torch.randn(n, 2 * d, generator=g, dtype=torch.float64) for k, n in
            (("initial", 300), ("terminal", 290), ("0T", 800))}
    params = {"params": {"tilde_F": {"kernel": W, "bias": b}}}
    ref = o_res.kfp_value_and_grad_fn(o_model.quadratic_parametric_apply, params, data, pde)
    spec = ops.ModelSpec(L.MODEL_QUADRATIC, d)
    F = pde.initial_configuration["tilde_F"]
    tg = ops.TrueGrad(L.DRIFT_LINEAR, F.float().to(cuda))
    flat = torch.cat([W.reshape(-1), b]).float().to(cuda)
    sums, gflat = _kfp_sets(ops, L, spec, flat, data, tg, 1.0, 2.0, cuda)
    assert relmax(sums[L.SUM_LOSS], ref["loss"]) < TOL
    assert relmax(sums[L.SUM_GT], ref["loss ground truth"]) < TOL
    assert relmax(gflat[: d * d], ref["grad"]["params"]["tilde_F"]["kernel"].reshape(-1)) < TOL
    assert relmax(gflat[d * d:], ref["grad"]["params"]["tilde_F"]["bias"]) < TOL
    flat_true = torch.cat([(F / 2).reshape(-1), torch.zeros(d, dtype=torch.float64)]).float().to(cuda)
    sums, _ = _kfp_sets(ops, L, spec, flat_true, data, tg, 1.0, 2.0, cuda)
    assert abs(sums[L.SUM_GT].item()) < 1e-8 * max(1.0, abs(sums[L.SUM_GTRUE2].item()))


@pytest.mark.parametrize("model_kind,d,nt", [("mlp", 2, 1), ("quadratic", 2, 1), ("mlp", 3, 2), ("quadratic", 4, 1)])
def test_kmv_pairwise_residual(cuda, model_kind, d, nt):
    """kinetic_mckean_vlasov.py:11-120 with m = n (the batch is its own reference set)."""
    ops, L = _ops()
    from pde_inverse_problem_b200.methods.consistency_instances import common
    n = 37
    pde = o_prob.KineticOUProblem(d, T=2.0)
    g = torch.Generator().manual_seed(d + nt)
    data0T = torch.randn(n * nt, 2 * d, generator=g, dtype=torch.float64)
    tau = torch.linspace(0.2, 1.4, nt, dtype=torch.float64)
    if model_kind == "mlp":
        params = o_model.init_mlp_params(d, 32, 2)
        apply_fn = o_model.mlp_apply
        spec = ops.ModelSpec(L.MODEL_MLP, d, 32, 2)
    else:
        W = torch.randn(d, d, generator=g, dtype=torch.float64)
        b = torch.randn(d, generator=g, dtype=torch.float64)
        params = {"params": {"tilde_F": {"kernel": W, "bias": b}}}
        apply_fn = o_model.quadratic_parametric_apply
        spec = ops.ModelSpec(L.MODEL_QUADRATIC, d)
    ref = o_res.kmv_value_and_grad_fn(apply_fn, params, {"0T": data0T, "tau_0T": tau}, pde)
    # the coefficient c[n, nt] exactly as the reference builds it (vmap over (tau, x[:, t]) then reshape)
    x3 = data0T[:, :d].reshape(-1, nt, d)
    psl = torch.stack([pde.partial_s_log_density_fn(tau[t], x3[:, t]) for t in range(nt)], 0).reshape(-1, nt)
    ps2l = torch.stack([pde.partial_s2_log_density_fn(tau[t], x3[:, t]) for t in range(nt)], 0).reshape(-1, nt)
    c = ps2l + psl ** 2 + 1.0 * psl

    class M:
        pass
    model = M()
    model.spec = spec
    flat = o_model.flatten_params(params).float().to(cuda)
    acc = ops.ResidualAccumulator(spec, device=cuda)
    out = ops.kmv_value_and_grad(model, params, flat, data0T.reshape(n, nt, 2 * d).float().to(cuda),
                                 c.float().contiguous().to(cuda), pde.initial_configuration["tilde_F"].float().to(cuda),
                                 acc, lambda m_, p_, s, gr: (s.cpu().double(), gr.cpu().double()))
    sums, gflat = out
    e = (relmax(sums[L.SUM_LOSS], ref["loss"]), relmax(sums[L.SUM_GT], ref["loss ground truth"]),
         relmax(gflat, o_model.flatten_params(ref["grad"])))
    print(f"KMV {model_kind} d={d} nt={nt}: loss {e[0]:.2e} gt {e[1]:.2e} grad {e[2]:.2e}")
    assert max(e) < TOL, e


def test_exact_samplers_reproduce_analytic_moments(cuda):
    """Gaussian.sample (distribution.py:64-65), grouped per-time sampler (OU.py:140-190) and the overdamped
    per-sample random-time sampler (fokker_planck_example.py:84-96)."""
    ops, L = _ops()
    d2, N = 6, 400_000
    g = torch.Generator().manual_seed(1)
    A = torch.randn(d2, d2, generator=g, dtype=torch.float64)
    cov = A @ A.T + 0.5 * torch.eye(d2, dtype=torch.float64)
    mu = torch.randn(d2, generator=g, dtype=torch.float64)
    half = torch.as_tensor(o_mom.gaussian_cov_half(cov.numpy()))
    z = ops.gaussian_sample(N, d2, mu.float().to(cuda), half.float().to(cuda), seed=3)
    s1, s2 = ops.ensemble_moments(z)
    m = (s1 / N).cpu().double()
    C = (s2 / N).cpu().double() - torch.outer(m, m)
    assert (m - mu).abs().max() < 0.02 and ((C - cov).norm() / cov.norm()).item() < 0.01
    # grouped: two groups with different laws
    mus = torch.stack([mu, -mu]).float().to(cuda)
    halves = torch.stack([half, 2 * half]).float().to(cuda)
    zz = ops.gaussian_sample_grouped(2, N // 2, d2, mus, halves, seed=4)
    for gi, (mm, sc) in enumerate(((mu, 1.0), (-mu, 4.0))):
        s1, s2 = ops.ensemble_moments(zz[gi].contiguous())
        m = (s1 / (N // 2)).cpu().double()
        C = (s2 / (N // 2)).cpu().double() - torch.outer(m, m)
        assert (m - mm).abs().max() < 0.03 and ((C - sc * cov).norm() / (sc * cov).norm()).item() < 0.015
    # overdamped OU at a fixed time (t_min = t_max): closed-form mean / covariance
    cfg = o_mom.overdamped_ou_configuration(4)
    f32 = lambda a: torch.as_tensor(a, dtype=torch.float32, device=cuda).contiguous()
    t = 0.37
    x = ops.ou_exact_sample(N, f32(cfg["U"]), f32(cfg["s"]), f32(cfg["B_0"]), f32(cfg["B"]),
                            f32(cfg["U"].T @ cfg["m_0"]), t, t, seed=5)
    m_t, P_t = o_mom.overdamped_ou_mean_cov(t, cfg)
    s1, s2 = ops.ensemble_moments(x)
    m = (s1 / N).cpu().double().numpy()
    C = (s2 / N).cpu().double().numpy() - np.outer(m, m)
    assert np.abs(m - m_t).max() < 0.02 and np.linalg.norm(C - P_t) / np.linalg.norm(P_t) < 0.01


def _cfg(pde, **kw):
    from pde_inverse_problem_b200.config import make_config
    base = {"neural_network.hidden_dim": 32, "neural_network.layers": 2, "train.number_of_iterations": 6,
            "train.optimizer.learning_rate.initial": 1e-2, "train.optimizer.learning_rate.scheduling": "cosine",
            "test.frequency": 3}
    base.update(kw)
    return make_config(pde, **base)


@pytest.mark.parametrize("name,pde,kw", [
    ("KGMM online SDE, MLP", "kinetic_fokker_planck",
     {"pde_instance.potential": "GMM", "estimation_mode": "non-parametric", "pde_instance.n_steps": 20,
      "solver.train.batch_size_0T": 200}),
    ("KGMM offline, parametric", "kinetic_fokker_planck",
     {"pde_instance.potential": "GMM", "estimation_mode": "parametric", "pde_instance.sample_mode": "offline",
      "pde_instance.sample_initial_size": 4000, "pde_instance.sample_terminal_size": 3000,
      "pde_instance.sample_0T_size": 500, "pde_instance.n_steps_terminal": 40, "pde_instance.n_steps_0T": 40}),
    ("KOU online exact, parametric", "kinetic_fokker_planck",
     {"estimation_mode": "parametric", "solver.train.batch_size_0T": 5000, "solver.train.batch_size_init": 2000,
      "solver.train.batch_size_terminal": 2000}),
    ("OU-FP online exact, MLP", "fokker_planck",
     {"estimation_mode": "non-parametric", "pde_instance.total_evolving_time": 5.0, "solver.train.batch_size_0T": 3000,
      "solver.train.batch_size_init": 3000, "solver.train.batch_size_terminal": 3000}),
    ("KOU online exact, MLP with the reference's default network (hidden_dim 20, layers 8: MLP.yaml)",
     "kinetic_fokker_planck",
     {"estimation_mode": "non-parametric", "neural_network.hidden_dim": 20, "neural_network.layers": 8,
      "solver.train.batch_size_0T": 2000, "solver.train.batch_size_init": 1000, "solver.train.batch_size_terminal": 1000}),
    ("KMV online, parametric, moment closure", "kinetic_mckean_vlasov",
     {"estimation_mode": "parametric", "pde_instance.domain_dim": 2, "solver.train.sample_mode": "grid_time",
      "solver.train.n_time_stamps": 1, "solver.train.sample_per_time": 3000, "solver.train.batch_size_init": 100,
      "solver.train.batch_size_terminal": 100, "pde_instance.kmv": {"moment_closure": True}}),
    ("KMV online, parametric", "kinetic_mckean_vlasov",
     {"estimation_mode": "parametric", "pde_instance.domain_dim": 2, "solver.train.sample_mode": "grid_time",
      "solver.train.n_time_stamps": 1, "solver.train.sample_per_time": 300, "solver.train.batch_size_init": 100,
      "solver.train.batch_size_terminal": 100}),
])
def test_reference_shaped_training_loop_runs(cuda, name, pde, kw):
    """registry -> problem -> ConsistencyBased -> create_model_fn -> JaxTrainer.fit, as main.py:47-66."""
    from pde_inverse_problem_b200 import main as pmain
    cfg = _cfg(pde, **kw)
    logs = []
    params, pde_instance, net = pmain.run(cfg, log_fn=lambda d_, step: logs.append((step, dict(d_))), device=cuda)
    train_logs = [l for _, l in logs if "loss" in l]
    assert len(train_logs) == 6
    for l in train_logs:
        assert set(l) == {"loss", "grad_norm", "loss ground truth", "params_norm"}
        assert all(torch.isfinite(v).item() for v in l.values())
    assert torch.isfinite(params["_flat"]).all()


def test_loss_is_unbiased_estimate_of_ground_truth_on_device(cuda):
    """KAT-5 on the CUDA path: exact kinetic-OU samples, quadratic model away from the truth."""
    from pde_inverse_problem_b200 import registry
    from pde_inverse_problem_b200.utils import rng as R
    cfg = _cfg("kinetic_fokker_planck", **{"estimation_mode": "parametric", "solver.train.batch_size_0T": 400000,
                                           "solver.train.batch_size_init": 200000,
                                           "solver.train.batch_size_terminal": 200000, "pde_instance.domain_dim": 2})
    pde = registry.get_pde_instance(cfg)(cfg=cfg, rng=1, device=cuda)
    method = registry.get_method(cfg)(pde_instance=pde, cfg=cfg, rng=2)
    net, params = method.create_model_fn()
    out = method.value_and_grad_fn(net.apply, params, R.PRNGKey(5))
    loss, gt = out["loss"].item(), out["loss ground truth"].item()
    assert abs(loss - gt) < 0.05 * max(1.0, abs(gt)), (loss, gt)
