"""GPU parity of the COMPOSED hot path against the oracle (round-1 verdict: the pieces were oracle-checked, the
composition was not): `pipeline.HotPath.step` (chunks, weights 1 / N_global, emit_drift -> PDEIP_DRIFT_IN_POINTS,
BLOCK128 trajectory, boundary sets from z_last / z0, Adam) and an N-iteration `JaxTrainer.fit`, each against the
float64 oracle fed with the SAME Philox draws / the SAME batches.

Tolerances.  fp32 path: the north-star's 1e-5 class is asserted on the whole composition (loss, "loss ground truth",
gradient, gradient norm), for the linear drift AND for the GMM drift at S = 200: individual float32 trajectories
deviate from float64 by up to 3e-5 after 200 steps, but the loss and the gradient are ensemble means
and come out at 1e-7 .. 1e-6 (measured, printed by the tests).  Tensor path: the bf16-GEMM tolerance 1e-2."""
import pytest
import torch

from conftest import relmax
from oracle import integrator as o_int, model as o_model, optim as o_optim, potential as o_pot, problems as o_prob
from oracle import residuals as o_res, taylor as o_tay

pytestmark = pytest.mark.gpu


def _mods():
    from pde_inverse_problem_b200 import _lib as L
    from pde_inverse_problem_b200 import ops
    from pde_inverse_problem_b200.core.model import V_hypothesis
    from pde_inverse_problem_b200.core.optimizer import AdamL2
    from pde_inverse_problem_b200.pipeline import HotPath, HotPathConfig
    return L, ops, V_hypothesis, AdamL2, HotPath, HotPathConfig


def _oracle_params_from(flat, d):
    """Flax-shaped float64 tree with the values of the flat CUDA buffer (MLP d -> 32 x 2 -> 40)."""
    p = o_model.init_mlp_params(d, 32, 2)
    off = 0
    flat = flat.detach().double().cpu()
    for i in range(3):
        leaf = p["params"][f"layers_{i}"]
        for name in ("kernel", "bias"):
            k = leaf[name].numel()
            leaf[name] = flat[off:off + k].reshape(leaf[name].shape).clone()
            off += k
    return p


def _hotpath_vs_oracle(cuda, d, K, n, S, chunk, path_name, use_autodiff):
    L, ops, V_hypothesis, AdamL2, HotPath, HotPathConfig = _mods()
    T, seed = 2.0, 77
    dt = T / S
    g = torch.Generator().manual_seed(5 + d + S)
    if K > 0:
        pde = o_prob.KineticGMMProblem(d, K, T=T)
        gamma = 0.5
        drift64 = pde.mus
        grad_fn = o_pot.GMMPotential(pde.mus, 1.0).gradient
        drift_kind = L.DRIFT_GMM
        z0 = torch.randn(n, 2 * d, generator=g, dtype=torch.float64) * torch.cat(
            [torch.full((d,), 2.0), torch.full((d,), 0.316)]).double()
    else:
        pde = o_prob.KineticOUProblem(d, T=T)
        gamma = float(pde.initial_configuration["gamma_friction"])
        drift64 = pde.initial_configuration["tilde_F"] / d
        pde.initial_configuration["tilde_F"] = drift64
        grad_fn = o_pot.LinearDrift(drift64).gradient
        drift_kind = L.DRIFT_LINEAR
        z0 = torch.randn(n, 2 * d, generator=g, dtype=torch.float64)
    path = L.PATH_FP32 if path_name == "fp32" else L.PATH_TENSOR
    model = V_hypothesis(1, [32, 32], d)
    params = model.init(11, torch.zeros(d, device=cuda))
    # non-zero biases: exercise every leaf of the gradient
    params["_flat"].add_(0.05 * torch.randn(params["_flat"].shape, generator=g).to(cuda))
    flat0 = params["_flat"].clone()
    drift = drift64.float().to(cuda).contiguous()
    opt = AdamL2(lambda count: 1e-2, 1e-3)
    cfg = HotPathConfig(d=d, n_steps=S, total_time=T, gamma=gamma, drift_kind=drift_kind, n_gaussian=K, chunk=chunk,
                        path=path)
    hp = HotPath(cfg, model, params, drift, ops.TrueGrad(drift_kind, drift, 1.0), optimizer=opt, device=cuda)
    out = hp.step(z0.float().to(cuda), seed=seed)
    torch.cuda.synchronize()
    assert ops.tensor_path_status() == 0

    # ---- oracle: the same Philox draws, float64 ------------------------------------------------------------------
    noise = ops.philox_normals(n, S + 1, d, seed=seed, device=cuda).double().cpu()
    tau0 = ops.philox_uniforms(n, seed=seed, device=cuda).float().mul(dt).double().cpu()
    last, traj, _ = o_int.underdamped_langevin_dynamics_scan(z0.float().double(), S, dt, noise, tau0, grad_fn, gamma)
    data = {"initial": z0.float().double(), "terminal": last, "0T": traj.reshape(-1, 2 * d)}
    p64 = _oracle_params_from(flat0, d)
    pde.initial_configuration["gamma_friction"] = gamma
    if use_autodiff:
        ref = o_res.kfp_value_and_grad_fn(o_model.mlp_apply, p64, data, pde, chunk=4096)
        ref_loss, ref_gt, ref_grad = ref["loss"], ref["loss ground truth"], o_model.flatten_params(ref["grad"])
    else:  # closed-form twin (oracle/taylor.py, checked against the autodiff restatement to 1e-12 in test_oracle_kat.py)
        x0T = data["0T"][:, :d]
        gt = grad_fn(x0T)
        ref = o_tay.kfp_value_and_grad(p64, data, gamma, T, grad_true_0T=gt)
        ref_loss, ref_gt, ref_grad = ref["loss"], ref["loss ground truth"], o_model.flatten_params(ref["grad"])
    # optimizer step of the oracle on the oracle gradient
    st = o_optim.AdamL2State(flat0.double().cpu())
    p_new = o_optim.adam_l2_step(flat0.double().cpu(), ref_grad, st, o_optim.constant_schedule(1e-2))
    errs = dict(loss=relmax(out["loss"], ref_loss), gt=relmax(out["loss ground truth"], ref_gt),
                grad=relmax(out["grad"], ref_grad), gnorm=relmax(out["sums"][L.SUM_GRADNORM], ref_grad.norm()),
                params=relmax(params["_flat"], p_new))
    print(f"HotPath.step vs oracle d={d} K={K} n={n} S={S} chunk={chunk} path={path_name}: "
          + " ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    return errs


@pytest.mark.parametrize("path_name", ["fp32", "tensor"])
def test_hotpath_step_c3_shape_vs_oracle(cuda, path_name):
    """C3 shape (KGMM d = 8, K = 16, S = 200), 2 chunks of 128 particles (BLOCK128 trajectory), both paths, against the
    float64 closed-form twin on the same Philox draws."""
    e = _hotpath_vs_oracle(cuda, d=8, K=16, n=256, S=200, chunk=128, path_name=path_name, use_autodiff=False)
    tol = 1e-5 if path_name == "fp32" else 1e-2
    assert e["loss"] < tol and e["gt"] < tol and e["grad"] < tol and e["gnorm"] < tol, e
    # One Adam step: the first update is lr * g / (|g| + eps) per entry, i.e. it AMPLIFIES the relative error of the small
    # gradient entries (max-norm-small errors flip u between -1 and 1), so the updated parameters are compared on the
    # fp32 path only; the tensor path's gradient itself is checked above.
    if path_name == "fp32":
        assert e["params"] < 1e-5, e


@pytest.mark.parametrize("path_name", ["fp32", "tensor"])
def test_hotpath_step_vs_autodiff_oracle(cuda, path_name):
    """The same composition against the AUTODIFF restatement (oracle/residuals.py), S = 40, n = 320 with chunk = 192: a
    192-particle chunk (not a multiple of 128: time-SoA trajectory layout) and a 128-particle tail (BLOCK128), GMM drift."""
    e = _hotpath_vs_oracle(cuda, d=8, K=16, n=320, S=40, chunk=192, path_name=path_name, use_autodiff=True)
    tol = 1e-5 if path_name == "fp32" else 1e-2
    assert e["loss"] < tol and e["gt"] < tol and e["grad"] < tol, e


def test_hotpath_step_linear_drift_fp32_1e5(cuda):
    """Contractive linear drift (kinetic OU, d = 4, S = 100, C2 shape): the fp32 composition holds the rtol-1e-5 class
    end to end (trajectory + residual + gradient), 3 chunks with a ragged tail (time-SoA trajectory layout)."""
    e = _hotpath_vs_oracle(cuda, d=4, K=0, n=300, S=100, chunk=128, path_name="fp32", use_autodiff=False)
    assert e["loss"] < 1e-5 and e["gt"] < 1e-5 and e["grad"] < 1e-5, e


@pytest.mark.parametrize("ema_start", [None, 6])
def test_trainer_fit_20_iterations_vs_oracle_loop(cuda, ema_start):
    """core/trainer.py:61-107 on the CUDA path against an oracle loop (oracle/residuals.py + oracle/optim.py) on
    IDENTICAL batches: the batches the method samples on the device are captured and fed to the float64 oracle,
    parameters are compared after every one of 20 iterations (cosine schedule, L2-in-Adam).  ema_start = 6: the EMA
    branch of the trainer (trainer.py:87-103: state re-seeded with the parameters at the switch epoch — 40 000 in the
    reference, moved to 6 here —, then every step ema <- 0.999 ema + 0.001 p_new and the parameters are OVERWRITTEN by
    the raw accumulator)."""
    from pde_inverse_problem_b200 import registry
    from pde_inverse_problem_b200.config import make_config
    from pde_inverse_problem_b200.core.optimizer import get_optimizer
    from pde_inverse_problem_b200.core.trainer import JaxTrainer
    from pde_inverse_problem_b200.utils import rng as R
    n_iter, d = 20, 2
    cfg = make_config("kinetic_fokker_planck", **{
        "neural_network.hidden_dim": 32, "neural_network.layers": 2, "train.number_of_iterations": n_iter,
        "train.optimizer.learning_rate.initial": 1e-2, "train.optimizer.learning_rate.scheduling": "cosine",
        "test.frequency": 1000, "estimation_mode": "non-parametric", "pde_instance.domain_dim": d,
        "solver.train.batch_size_0T": 600, "solver.train.batch_size_init": 400, "solver.train.batch_size_terminal": 400,
        "train.optimizer.use_ema": ema_start is not None})
    pde = registry.get_pde_instance(cfg)(cfg=cfg, rng=1, device=cuda)
    method = registry.get_method(cfg)(pde_instance=pde, cfg=cfg, rng=2)
    net, params = method.create_model_fn()
    flat0 = params["_flat"].clone()
    batches, snaps = [], []
    sample_data = method.sample_data

    def capture(rng):
        data = sample_data(rng)
        batches.append({k: v.detach().double().cpu() for k, v in data.items() if k in ("initial", "terminal", "0T")})
        return data

    method.sample_data = capture
    logs = []

    def log_fn(dct, step):
        if "loss" in dct:
            logs.append({k: float(v) for k, v in dct.items()})
            snaps.append(params["_flat"].detach().double().cpu().clone())

    trainer = JaxTrainer(cfg=cfg, method=method, rng=R.PRNGKey(3), optimizer=get_optimizer(cfg.train.optimizer),
                         forward_fn=net.apply, params=params, log_fn=log_fn)
    if ema_start is not None:
        trainer.EMA_START_EPOCH = ema_start
    trainer.fit()
    assert len(batches) == n_iter and len(snaps) == n_iter

    # ---- oracle loop on the captured batches -----------------------------------------------------------------------
    opde = o_prob.KineticOUProblem(d, T=float(pde.total_evolving_time))
    opde.initial_configuration["tilde_F"] = pde.initial_configuration["tilde_F"].double().cpu()
    opde.initial_configuration["gamma_friction"] = float(pde.initial_configuration["gamma_friction"])
    p64 = _oracle_params_from(flat0, d)
    flat = o_model.flatten_params(p64)
    st = o_optim.AdamL2State(flat)
    sched = o_optim.cosine_decay_schedule(1e-2)
    worst = dict(params=0.0, loss=0.0, gnorm=0.0)
    for it in range(n_iter):
        ref = o_res.kfp_value_and_grad_fn(o_model.mlp_apply, _oracle_params_from(flat, d), batches[it], opde)
        gflat = o_model.flatten_params(ref["grad"])
        if ema_start is not None and it == ema_start:
            ema = flat.clone()                                   # EmaState(count=0, ema=params), trainer.py:97-100
        flat = o_optim.adam_l2_step(flat, gflat, st, sched)
        if ema_start is not None and it >= ema_start:
            ema = o_optim.ema_update(ema, flat)                  # trainer.py:67-69: params <- raw ema accumulator
            flat = ema.clone()
        worst["params"] = max(worst["params"], relmax(snaps[it], flat))
        worst["loss"] = max(worst["loss"], relmax(logs[it]["loss"], ref["loss"]))
        worst["gnorm"] = max(worst["gnorm"], relmax(logs[it]["grad_norm"], ref["grad_norm"]))
        assert abs(logs[it]["params_norm"] - float(flat.norm())) < 1e-5 * float(flat.norm())
    print("fit vs oracle loop, worst over 20 iterations:", {k: f"{v:.2e}" for k, v in worst.items()})
    assert worst["params"] < 1e-5 and worst["loss"] < 1e-5 and worst["gnorm"] < 1e-5, worst
