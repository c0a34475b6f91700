"""GPU parity: model evaluation, residual value+gradient and the optimizer step vs the autodiff oracle."""
import math

import pytest
import torch
from torch.func import grad, vmap

from conftest import relmax
from oracle import model as o_model
from oracle import optim as o_optim
from oracle import problems as o_prob
from oracle import residuals as o_res

pytestmark = pytest.mark.gpu

TOL = 1e-5  # fp32 path, BASELINE.json: rtol 1e-5 in the per-tensor max-norm metric


def _ops():
    from pde_inverse_problem_b200 import ops, _lib
    return ops, _lib


def _params(d, layers=2, seed=11, bias_scale=0.1):
    p = o_model.init_mlp_params(d, 32, layers, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    for k in p["params"]:
        b = p["params"][k]["bias"]
        p["params"][k]["bias"] = bias_scale * torch.randn(b.shape, generator=g, dtype=torch.float64)
    return p


@pytest.mark.parametrize("d,layers", [(2, 2), (4, 2), (8, 2), (16, 1), (32, 2), (8, 3)])
def test_model_eval_matches_autodiff(cuda, d, layers):
    ops, L = _ops()
    p = _params(d, layers)
    g = torch.Generator().manual_seed(d)
    x = torch.randn(1031, d, generator=g, dtype=torch.float64)
    v = torch.randn(1031, d, generator=g, dtype=torch.float64)
    V = lambda xx: o_model.mlp_apply(p, xx)[0]
    ref_val = vmap(V)(x)
    ref_grad = vmap(grad(V))(x)
    ref_vhv = vmap(lambda xx, vv: torch.dot(vv, o_res.hessian_vector_product(V, xx, vv)))(x, v)
    ref_lap = vmap(lambda xx: torch.diagonal(torch.func.jacfwd(grad(V))(xx)).sum())(x)
    spec = ops.ModelSpec(L.MODEL_MLP, d, 32, layers)
    out = ops.model_eval(spec, o_model.flatten_params(p).float().to(cuda), x.float().to(cuda), v.float().to(cuda),
                         want=("value", "grad", "vHv", "laplacian"))
    assert relmax(out["value"], ref_val) < TOL
    assert relmax(out["grad"], ref_grad) < TOL
    assert relmax(out["vHv"], ref_vhv) < TOL
    assert relmax(out["laplacian"], ref_lap) < TOL


def _run_kfp(ops, L, cuda, p, data, pde, true_grad, gamma, T, layout=None):
    d = pde.dim
    layers = len(p["params"]) - 1
    spec = ops.ModelSpec(L.MODEL_MLP, d, 32, layers)
    flat = o_model.flatten_params(p).float().to(cuda)
    acc = ops.ResidualAccumulator(spec, device=cuda).begin()
    pts = {k: v.float().to(cuda) for k, v in data.items()}
    lay = L.LAYOUT_AOS
    p0 = pts["0T"]
    if layout == "soa":
        lay = L.LAYOUT_SOA
        p0 = p0.t().contiguous()
    acc.accumulate(L.SET_KFP_0T, flat, p0, 1.0 / data["0T"].shape[0], coef=gamma, true_grad=true_grad, layout=lay)
    acc.accumulate(L.SET_KFP_BOUNDARY, flat, pts["terminal"], 1.0 / data["terminal"].shape[0], coef=2.0 / T)
    acc.accumulate(L.SET_KFP_BOUNDARY, flat, pts["initial"], 1.0 / data["initial"].shape[0], coef=-2.0 / T)
    sums, g = acc.finalize()
    return sums.cpu().double(), g.cpu().double()


@pytest.mark.parametrize("d,layers,layout", [(4, 2, "aos"), (8, 2, "soa"), (16, 2, "aos"), (32, 2, "aos"),
                                              (8, 1, "aos"), (4, 3, "aos")])
def test_kfp_value_and_grad_matches_oracle(cuda, d, layers, layout):
    """kinetic_fokker_planck.py:11-69 on identical inputs: loss, grad, grad_norm, loss ground truth."""
    ops, L = _ops()
    pde = o_prob.KineticOUProblem(d, T=2.0)
    p = _params(d, layers)
    g = torch.Generator().manual_seed(100 + d)
    data = {"initial": torch.randn(700, 2 * d, generator=g, dtype=torch.float64),
            "terminal": torch.randn(515, 2 * d, generator=g, dtype=torch.float64),
            "0T": torch.randn(1300, 2 * d, generator=g, dtype=torch.float64)}
    ref = o_res.kfp_value_and_grad_fn(o_model.mlp_apply, p, data, pde)
    tg = ops.TrueGrad(L.DRIFT_LINEAR, pde.initial_configuration["tilde_F"].float().to(cuda))
    sums, grad_flat = _run_kfp(ops, L, cuda, p, data, pde, tg, 1.0, 2.0, layout)
    assert relmax(sums[L.SUM_LOSS], ref["loss"]) < TOL
    assert relmax(sums[L.SUM_GT], ref["loss ground truth"]) < TOL
    assert relmax(sums[L.SUM_GRADNORM], ref["grad_norm"]) < TOL
    assert relmax(grad_flat, o_model.flatten_params(ref["grad"])) < TOL
    # per-leaf check in the same metric
    off = 0
    for name in sorted(ref["grad"]["params"], key=lambda s: (len(s), s)):
        for leaf in ("kernel", "bias"):
            r = ref["grad"]["params"][name][leaf].reshape(-1)
            assert relmax(grad_flat[off:off + r.numel()], r) < 5 * TOL, (name, leaf)
            off += r.numel()


def test_kfp_gmm_true_gradient(cuda):
    ops, L = _ops()
    d, K = 8, 16
    pde = o_prob.KineticGMMProblem(d, K, T=2.0)
    p = _params(d, 2)
    g = torch.Generator().manual_seed(5)
    data = {k: torch.randn(n, 2 * d, generator=g, dtype=torch.float64) * 1.5
            for k, n in (("initial", 300), ("terminal", 300), ("0T", 900))}
    ref = o_res.kfp_value_and_grad_fn(o_model.mlp_apply, p, data, pde)
    tg = ops.TrueGrad(L.DRIFT_GMM, pde.mus.float().to(cuda), sigma=1.0)
    sums, grad_flat = _run_kfp(ops, L, cuda, p, data, pde, tg, 0.5, 2.0)
    assert relmax(sums[L.SUM_LOSS], ref["loss"]) < TOL
    assert relmax(sums[L.SUM_GT], ref["loss ground truth"]) < TOL
    assert relmax(grad_flat, o_model.flatten_params(ref["grad"])) < TOL


@pytest.mark.parametrize("d", [2, 4, 8])
def test_fp_value_and_grad_matches_oracle(cuda, d):
    """fokker_planck.py:33-63 (exact Laplacian from d tangent streams)."""
    ops, L = _ops()
    pde = o_prob.OverdampedOUProblem(d, T=5.0)
    p = _params(d, 2)
    g = torch.Generator().manual_seed(200 + d)
    data = {k: torch.randn(n, d, generator=g, dtype=torch.float64) + 0.3
            for k, n in (("initial", 600), ("terminal", 450), ("0T", 1100))}
    ref = o_res.fp_value_and_grad_fn(o_model.mlp_apply, p, data, pde)
    spec = ops.ModelSpec(L.MODEL_MLP, d, 32, 2)
    flat = o_model.flatten_params(p).float().to(cuda)
    tg = ops.TrueGrad(L.DRIFT_LINEAR, pde.initial_configuration["F"].float().to(cuda))
    acc = ops.ResidualAccumulator(spec, device=cuda).begin()
    acc.accumulate(L.SET_FP_0T, flat, data["0T"].float().to(cuda), 1.0 / 1100, true_grad=tg)
    acc.accumulate(L.SET_FP_BOUNDARY, flat, data["terminal"].float().to(cuda), 1.0 / 450, coef=2.0 / 5.0)
    acc.accumulate(L.SET_FP_BOUNDARY, flat, data["initial"].float().to(cuda), 1.0 / 600, coef=-2.0 / 5.0)
    sums, grad_flat = acc.finalize()
    assert relmax(sums[L.SUM_LOSS], ref["loss"]) < TOL
    assert relmax(sums[L.SUM_GT], ref["loss ground truth"]) < TOL
    assert relmax(grad_flat, o_model.flatten_params(ref["grad"])) < TOL


def test_residual_is_deterministic_and_chunk_invariant(cuda):
    """Same points in one launch or in chunks: the chunked sum stays within fp32 rounding; repeated runs are
    bit-identical (no atomics)."""
    ops, L = _ops()
    d = 8
    p = _params(d, 2)
    flat = o_model.flatten_params(p).float().to(cuda)
    pts = torch.randn(5000, 2 * d, device=cuda)
    spec = ops.ModelSpec(L.MODEL_MLP, d, 32, 2)
    outs = []
    for _ in range(2):
        acc = ops.ResidualAccumulator(spec, device=cuda).begin()
        acc.accumulate(L.SET_KFP_0T, flat, pts, 1.0 / 5000, coef=0.5)
        s, g = acc.finalize()
        outs.append((s.clone(), g.clone()))
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][0], outs[1][0])
    acc = ops.ResidualAccumulator(spec, device=cuda).begin()
    for lo in range(0, 5000, 1250):
        acc.accumulate(L.SET_KFP_0T, flat, pts[lo:lo + 1250].contiguous(), 1.0 / 5000, coef=0.5)
    s, g = acc.finalize()
    assert relmax(g, outs[0][1]) < 1e-5


def test_adam_l2_step_matches_oracle(cuda):
    """KAT-8: add_decayed_weights + adam(b1=.9,b2=.999,eps=1e-4), cosine schedule, EMA; 5 consecutive steps."""
    ops, L = _ops()
    g = torch.Generator().manual_seed(3)
    n = 2664
    p0 = torch.randn(n, generator=g, dtype=torch.float64)
    sched = o_optim.cosine_decay_schedule(1e-2)
    st = o_optim.AdamL2State(p0)
    p_ref = p0.clone()
    p = p0.float().to(cuda)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    ema = p.clone()
    ema_ref = p0.clone()
    for it in range(5):
        grad_ = torch.randn(n, generator=g, dtype=torch.float64)
        p_ref = o_optim.adam_l2_step(p_ref, grad_, st, sched)
        use_ema = it >= 3
        if it == 3:
            ema_ref = p_ref_before.clone()  # trainer.py:97-100: EMA re-seeded with the current params
            ema.copy_(p)
        if use_ema:
            ema_ref = o_optim.ema_update(ema_ref, p_ref)
            p_ref = ema_ref.clone()
        norms = ops.adam_l2_step(p, grad_.float().to(cuda), m, v, count=it + 1, lr=sched(it), ema=ema,
                                 use_ema=use_ema)
        p_ref_before = p_ref.clone()
        assert relmax(p, p_ref) < 1e-6, it
        assert relmax(norms[0], grad_.norm()) < 1e-6
        assert relmax(norms[1], p_ref.norm()) < 1e-6


def test_gather_0T_matches_reference_indexing(cuda):
    from oracle import sampler as o_s
    ops, L = _ops()
    g = torch.Generator().manual_seed(0)
    ds = torch.randn(50, 40, 8, generator=g)
    perm = torch.randperm(50, generator=g)
    ref = o_s.offline_subsample_0T(ds, 3, perm)
    out = ops.gather_0T(ds.to(cuda), perm[:10].to(cuda), 5, 3)
    assert torch.equal(out.cpu(), ref)


def test_wrong_device_and_shape_fail_loudly(cuda):
    ops, L = _ops()
    with pytest.raises(ops.PdeipError):
        ops.gmm_value_grad(torch.randn(4, 4), torch.randn(2, 4))  # CPU tensor: no fallback
    with pytest.raises(ops.PdeipError):
        ops.gmm_value_grad(torch.randn(4, 64, device=cuda), torch.randn(2, 64, device=cuda))  # d > 32
    with pytest.raises(ops.PdeipError):
        ops.ModelSpec(99, 4)
