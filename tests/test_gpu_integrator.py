"""GPU parity: K1 integrator and K2 drifts through the C ABI vs the oracle (injected noise)."""
import math

import numpy as np
import pytest
import torch

from conftest import relmax
from oracle import integrator as o_int
from oracle import moments as o_mom
from oracle import philox as o_philox
from oracle import potential as o_pot

pytestmark = pytest.mark.gpu


def _ops():
    from pde_inverse_problem_b200 import ops, _lib
    return ops, _lib


@pytest.mark.parametrize("d,K", [(2, 3), (4, 3), (8, 16), (5, 7), (32, 64)])
def test_gmm_grad_matches_oracle(cuda, d, K):
    ops, L = _ops()
    g = torch.Generator().manual_seed(d * 100 + K)
    x = torch.randn(4099, d, generator=g, dtype=torch.float64) * 2.0
    mus = torch.rand(K, d, generator=g, dtype=torch.float64) * 8 - 4
    ref_g = o_pot.vg_gmm_V(x, mus, 1.0)
    ref_v = o_pot.GMMPotential(mus, 1.0).value(x)
    val, grd = ops.gmm_value_grad(x.float().to(cuda), mus.float().to(cuda), 1.0, want_value=True)
    assert relmax(grd, ref_g) < 1e-5  # fp32 path, BASELINE.json rtol 1e-5 (max-norm metric)
    assert relmax(val, ref_v) < 1e-5


def test_gmm_single_gaussian_is_x_minus_mu(cuda):
    ops, L = _ops()
    x = torch.randn(1000, 8, device=cuda)
    mu = torch.randn(1, 8, device=cuda)
    _, grd = ops.gmm_value_grad(x, mu, 1.0)
    assert relmax(grd, x - mu) < 1e-6  # KAT-3, K = 1


def test_linear_grad(cuda):
    ops, L = _ops()
    x = torch.randn(3001, 16, dtype=torch.float64)
    A = torch.randn(16, 16, dtype=torch.float64)
    out = ops.linear_grad(x.float().to(cuda), A.float().to(cuda))
    assert relmax(out, x @ A.T) < 1e-5


def test_philox_raw_known_answers(cuda):
    """Random123 known-answer vectors for Philox4x32-10 (bit-exact)."""
    ops, L = _ops()
    def i32(v):
        return np.array(v, dtype=np.uint32).view(np.int32)
    cases = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
             ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
             ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
              (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in cases:
        out = ops.philox_raw(torch.from_numpy(i32([ctr])).to(cuda), torch.from_numpy(i32(key)).to(cuda))
        got = out.cpu().numpy().view(np.uint32)[0]
        assert tuple(int(x) for x in got) == want


def test_philox_normals_match_numpy_restatement(cuda):
    ops, L = _ops()
    n, draws, d, seed, off = 513, 3, 8, 0x1234567890ABCDEF, 7
    dev = ops.philox_normals(n, draws, d, seed, particle_offset=off, step_offset=5).cpu().double().numpy()
    for s in range(draws):
        ref = o_philox.normals(seed, np.arange(n) + off, 5 + s, d)
        assert np.abs(dev[:, s] - ref).max() < 2e-5  # fast-math log/sincos on device
    u = ops.philox_uniforms(n, seed, particle_offset=off).cpu().double().numpy()
    assert np.array_equal(u, o_philox.uniform01(seed, np.arange(n) + off))  # exact


def _setup_traj(d, K, N, S, seed=0):
    g = torch.Generator().manual_seed(seed)
    z0 = torch.randn(N, 2 * d, generator=g, dtype=torch.float64)
    noise = torch.randn(N, S + 1, d, generator=g, dtype=torch.float64)
    tau0 = torch.rand(N, generator=g, dtype=torch.float64)
    mus = torch.rand(K, d, generator=g, dtype=torch.float64) * 8 - 4
    return z0, noise, tau0, mus


@pytest.mark.parametrize("d", [2, 4, 8, 16])
def test_single_step_gmm_rtol_1e5(cuda, d):
    """Parity protocol (i): one reference trajectory with S = 1 (two update_steps) from identical states."""
    ops, L = _ops()
    N, S, dt, gamma = 2053, 1, 0.01, 0.5
    z0, noise, tau0, mus = _setup_traj(d, 5, N, S, seed=d)
    tau0 = tau0 * dt
    pot = o_pot.GMMPotential(mus, 1.0)
    last, traj, tau = o_int.underdamped_langevin_dynamics_scan(z0, S, dt, noise, tau0, pot.gradient, gamma)
    zl, tr, ta = ops.kl_integrate(z0.float().to(cuda), S, dt, gamma, L.DRIFT_GMM, mus.float().to(cuda),
                                  n_gaussian=5, sigma=1.0, noise=noise.float().to(cuda),
                                  tau0=tau0.float().to(cuda), want_tau=True)
    assert relmax(zl, last) < 1e-5
    assert relmax(tr, traj) < 1e-5
    assert relmax(ta, tau) < 1e-6


@pytest.mark.parametrize("d", [4, 16])
def test_full_trajectory_linear_drift_rtol_1e5(cuda, d):
    """Parity protocol (iii): contractive linear (OU) drift holds rtol 1e-5 over the whole trajectory."""
    ops, L = _ops()
    N, S, T, gamma = 1500, 100, 2.0, 1.0
    dt = T / S
    z0, noise, tau0, _ = _setup_traj(d, 1, N, S, seed=3)
    tau0 = tau0 * dt
    cfg = o_mom.kinetic_ou_configuration(d)
    F = torch.as_tensor(cfg["tilde_F"]) / d  # keep dt * |F| well inside the stability region
    drift = o_pot.LinearDrift(F)
    last, traj, _ = o_int.underdamped_langevin_dynamics_scan(z0, S, dt, noise, tau0, drift.gradient, gamma)
    for layout in (L.TRAJ_PARTICLE_MAJOR, L.TRAJ_TIME_MAJOR, L.TRAJ_TIME_SOA):
        zl, tr, _ = ops.kl_integrate(z0.float().to(cuda), S, dt, gamma, L.DRIFT_LINEAR, F.float().to(cuda),
                                     noise=noise.float().to(cuda), tau0=tau0.float().to(cuda),
                                     traj_layout=layout)
        if layout == L.TRAJ_TIME_MAJOR:
            tr = tr.permute(1, 0, 2)
        elif layout == L.TRAJ_TIME_SOA:
            tr = tr.permute(2, 1, 0)
        assert relmax(zl, last) < 1e-5
        assert relmax(tr, traj) < 1e-5


def test_multi_step_gmm_within_fp32_twin_band(cuda):
    """Parity protocol (ii): GMM trajectories are locally unstable, so the GPU error vs the float64 oracle is
    compared with the error of the CPU fp32 twin of the same code: within 2x (plus a floor)."""
    ops, L = _ops()
    d, K, N, S, T, gamma = 8, 16, 4096, 200, 2.0, 0.5
    dt = T / S
    z0, noise, tau0, mus = _setup_traj(d, K, N, S, seed=11)
    z0[:, :d] *= 2.0
    z0[:, d:] *= math.sqrt(0.1)
    tau0 = tau0 * dt
    last64, traj64, _ = o_int.underdamped_langevin_dynamics_scan(
        z0, S, dt, noise, tau0, o_pot.GMMPotential(mus, 1.0).gradient, gamma)
    last32, traj32, _ = o_int.underdamped_langevin_dynamics_scan(
        z0.float(), S, dt, noise.float(), tau0.float(), o_pot.GMMPotential(mus.float(), 1.0).gradient, gamma)
    zl, tr, _ = ops.kl_integrate(z0.float().to(cuda), S, dt, gamma, L.DRIFT_GMM, mus.float().to(cuda),
                                 n_gaussian=K, noise=noise.float().to(cuda), tau0=tau0.float().to(cuda))
    err_gpu = (tr.cpu().double() - traj64).abs()
    err_twin = (traj32.double() - traj64).abs()
    scale = traj64.abs().max().item()
    for s in (0, 9, 99, 199):
        g_max, t_max = err_gpu[:, s].max().item() / scale, err_twin[:, s].max().item() / scale
        g_med, t_med = err_gpu[:, s].median().item(), err_twin[:, s].median().item()
        assert g_max <= 2.0 * t_max + 1e-6, (s, g_max, t_max)
        assert g_med <= 2.0 * t_med + 1e-7, (s, g_med, t_med)
    assert relmax(tr[:, 0], traj64[:, 0]) < 1e-5  # first sample: strict


def test_emit_every_and_soa_state(cuda):
    ops, L = _ops()
    d, N, S, dt, gamma = 4, 777, 20, 0.05, 1.0
    z0, noise, tau0, _ = _setup_traj(d, 1, N, S, seed=5)
    tau0 = tau0 * dt
    F = torch.eye(d, dtype=torch.float64) * 0.7
    last, traj, _ = o_int.underdamped_langevin_dynamics_scan(z0, S, dt, noise, tau0, o_pot.LinearDrift(F).gradient, gamma)
    zl, tr, _ = ops.kl_integrate(z0.float().t().contiguous().to(cuda), S, dt, gamma, L.DRIFT_LINEAR, F.float().to(cuda),
                                 noise=noise.float().to(cuda), tau0=tau0.float().to(cuda),
                                 state_layout=L.LAYOUT_SOA, traj_layout=L.TRAJ_TIME_SOA, emit_every=5, emit_offset=2)
    assert relmax(zl.t(), last) < 1e-5
    assert tr.shape == (2 * d, 4, N)
    assert relmax(tr.permute(2, 1, 0), traj[:, 2::5]) < 1e-5


def test_philox_mode_equals_injected_device_noise(cuda):
    """The in-register Philox path and the injected-noise path run the same arithmetic."""
    ops, L = _ops()
    d, N, S, dt, gamma, seed = 8, 1000, 10, 0.01, 0.5, 99
    z0, _, _, mus = _setup_traj(d, 4, N, S, seed=8)
    z0 = z0.float().to(cuda)
    mus = mus.float().to(cuda)
    noise = ops.philox_normals(N, S + 1, d, seed)
    tau0 = ops.philox_uniforms(N, seed) * dt
    a = ops.kl_integrate(z0, S, dt, gamma, L.DRIFT_GMM, mus, n_gaussian=4, seed=seed)
    b = ops.kl_integrate(z0, S, dt, gamma, L.DRIFT_GMM, mus, n_gaussian=4, noise=noise, tau0=tau0)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    # sharding invariance: the second half computed with particle_offset reproduces the same particles
    c = ops.kl_integrate(z0[N // 2:].contiguous(), S, dt, gamma, L.DRIFT_GMM, mus, n_gaussian=4, seed=seed,
                         particle_offset=N // 2)
    assert torch.equal(c[0], a[0][N // 2:])


def test_ou_ensemble_moments_match_discrete_recursion(cuda):
    """KAT-2 / KAT-1 at config C2 scale-down: kinetic OU, d=4, S=100, dt=0.02, uniform schedule (tau0 = 0)."""
    ops, L = _ops()
    d, N, S, T = 4, 1 << 19, 100, 2.0
    dt = T / S
    cfg = o_mom.kinetic_ou_configuration(d)
    cfg["tilde_F"] = cfg["tilde_F"] / 4.0  # dt * lambda_max well inside the stability region
    Z, I = np.zeros((d, d)), np.eye(d)
    cfg["F"] = np.block([[Z, I], [-cfg["tilde_F"], -cfg["gamma_friction"] * I]])
    z0 = ops.gaussian_sample(N, 2 * d, None, None, seed=5)
    zl, _, _ = ops.kl_integrate(z0, S, dt, cfg["gamma_friction"], L.DRIFT_LINEAR,
                                torch.as_tensor(cfg["tilde_F"], dtype=torch.float32, device=cuda),
                                seed=17, schedule=L.SCHEDULE_UNIFORM, want_traj=False)
    s1, s2 = ops.ensemble_moments(zl)
    mean = (s1 / N).cpu().double().numpy()
    cov = (s2 / N).cpu().double().numpy() - np.outer(mean, mean)
    m_d, P_d = o_mom.discrete_mean_cov(S, dt, cfg, tau0=0.0)  # S-1 steps of dt plus (0, dt): S steps of dt
    m_c, P_c = o_mom.lyapunov_mean_cov(T, cfg)
    rel_d = np.linalg.norm(cov - P_d) / np.linalg.norm(P_d)
    rel_c = np.linalg.norm(cov - P_c) / np.linalg.norm(P_c)
    gap = np.linalg.norm(P_d - P_c) / np.linalg.norm(P_c)
    assert rel_d < 1e-2, rel_d                  # Monte-Carlo error at N = 2^19 (~0.4 %)
    assert np.abs(mean - m_d).max() < 1e-2
    assert abs(rel_c - gap) < 1e-2              # same O(dt) bias to the continuous solution as the scheme itself
    # moments kernel vs torch
    zc = zl.double()
    assert relmax(s2, (zc.T @ zc)) < 1e-4


@pytest.mark.parametrize("layout", ["particle", "time", "soa"])
def test_emit_drift_carries_grad_U_of_every_sample(cuda, layout):
    """emit_drift: each sample is [x, v, grad U(x)] — the drift the integrator evaluates at the following step."""
    ops, L = _ops()
    d, K, N, S, dt, gamma = 8, 5, 300, 12, 0.02, 0.5
    z0, noise, tau0, mus = _setup_traj(d, K, N, S, seed=21)
    tau0 = tau0 * dt
    lay = {"particle": L.TRAJ_PARTICLE_MAJOR, "time": L.TRAJ_TIME_MAJOR, "soa": L.TRAJ_TIME_SOA}[layout]
    _, traj64, _ = o_int.underdamped_langevin_dynamics_scan(z0, S, dt, noise, tau0, o_pot.GMMPotential(mus, 1.0).gradient, gamma)
    zl, tr, _ = ops.kl_integrate(z0.float().to(cuda), S, dt, gamma, L.DRIFT_GMM, mus.float().to(cuda), n_gaussian=K,
                                 noise=noise.float().to(cuda), tau0=tau0.float().to(cuda), traj_layout=lay,
                                 emit_drift=True, emit_every=3, emit_offset=1)
    if layout == "time":
        tr = tr.permute(1, 0, 2)
    elif layout == "soa":
        tr = tr.permute(2, 1, 0)
    assert tr.shape == (N, 4, 3 * d)
    ref = traj64[:, 1::3]
    assert relmax(tr[..., : 2 * d], ref) < 1e-5
    gref = o_pot.vg_gmm_V(ref[..., :d].reshape(-1, d), mus, 1.0).reshape(N, 4, d)
    assert relmax(tr[..., 2 * d:], gref) < 1e-5
    # UNIFORM schedule: the last sample has no following step, its drift is evaluated separately
    zl, tr, _ = ops.kl_integrate(z0.float().to(cuda), S, dt, gamma, L.DRIFT_GMM, mus.float().to(cuda), n_gaussian=K,
                                 noise=noise[:, :S].float().contiguous().to(cuda), schedule=L.SCHEDULE_UNIFORM,
                                 traj_layout=L.TRAJ_PARTICLE_MAJOR, emit_drift=True)
    x_last = tr[:, -1, :d].double().cpu()
    assert relmax(tr[:, -1, 2 * d:], o_pot.vg_gmm_V(x_last, mus, 1.0)) < 1e-5
    assert relmax(tr[:, -1, : 2 * d], zl) < 1e-7


@pytest.mark.parametrize("drift,d,K", [("gmm", 8, 16), ("gmm", 4, 3), ("gmm", 16, 20), ("linear", 4, 0), ("linear", 16, 0), ("gmm", 32, 64)])
def test_fast_production_kernel_matches_generic_kernel(cuda, drift, d, K):
    """The packed-fp32x2 production kernel (Philox noise, [3d][S][N] trajectory with grad U) against the generic
    kernel on the same seeds: identical noise stream, drift arithmetic equal up to rounding (MUFU ex2, fused damping
    factor).  Single steps agree to rtol 1e-5; over a short trajectory the difference stays at round-off level."""
    import os
    from pde_inverse_problem_b200 import ops, _lib as L
    n, S, T, gamma = 4096 + 37, 12, 0.24, 0.5
    g = torch.Generator().manual_seed(5 + d)
    z0 = (torch.randn(n, 2 * d, generator=g) * torch.cat([torch.full((d,), 2.0), torch.full((d,), 0.5)])).to(cuda)
    if drift == "gmm":
        params = (torch.rand(K, d, generator=g) * 8 - 4).to(cuda)
        kind = L.DRIFT_GMM
    else:
        f = torch.randn(d, d + 1, generator=g)
        params = ((f @ f.T) / d).to(cuda)
        kind = L.DRIFT_LINEAR

    def run():
        zl, tr, _ = ops.kl_integrate(z0, S, T / S, gamma, kind, params, n_gaussian=K, seed=77, particle_offset=1000,
                                     traj_layout=L.TRAJ_TIME_SOA, emit_drift=True)
        torch.cuda.synchronize()
        return zl.clone(), tr.clone()

    zl_fast, tr_fast = run()
    os.environ["PDEIP_NO_FAST_INTEGRATOR"] = "1"
    try:
        zl_ref, tr_ref = run()
    finally:
        del os.environ["PDEIP_NO_FAST_INTEGRATOR"]
    assert tr_fast.shape == tr_ref.shape == (3 * d, S, n)
    # first emitted sample = one step from identical states: rtol 1e-5 (max-norm per tensor)
    assert relmax(tr_fast[:, 0], tr_ref[:, 0]) < 1e-5
    assert relmax(tr_fast, tr_ref) < 5e-5
    assert relmax(zl_fast, zl_ref) < 5e-5


def unblock(tr):
    """[S][n/128][C][128] (TRAJ_BLOCK128) -> [C][S][n] (TRAJ_TIME_SOA)"""
    S, nb, C, _ = tr.shape
    return tr.permute(2, 0, 1, 3).reshape(C, S, nb * 128)


@pytest.mark.parametrize("d,K,n,S", [(32, 64, 4096 + 37, 12), (32, 64, 300, 200), (32, 40, 1000, 12), (32, 7, 129, 12),
                                     (16, 16, 1000, 12), (16, 64, 2048, 12), (16, 33, 515, 50), (32, 64, 4096, 12),
                                     (16, 48, 640, 30), (8, 16, 4096, 12), (8, 16, 1000, 200), (8, 3, 130, 12),
                                     (8, 40, 640, 30), (8, 64, 256, 12)])
def test_tensor_core_gmm_integrator_matches_fp32_kernel(cuda, d, K, n, S):
    """pdeip_kl_integrate_path(PDEIP_PATH_TENSOR): particle x centre contraction and softmax-weighted centre sum on
    tcgen05 with bf16 hi + lo split operands, against the fp32 production kernel on the same Philox stream.
    Tolerance class "bf16 GEMM paths" (BASELINE.json: rtol 1e-2, per-tensor max-norm); the split operands keep the
    measured difference two orders below that, asserted here at 2e-3 on whole trajectories and 2e-4 on one step."""
    from pde_inverse_problem_b200 import ops, _lib as L
    T, gamma = 0.01 * S, 0.5
    g = torch.Generator().manual_seed(11 + d + K)
    z0 = (torch.randn(n, 2 * d, generator=g) * torch.cat([torch.full((d,), 2.0), torch.full((d,), 0.3)])).to(cuda)
    mus = (torch.rand(K, d, generator=g) * 8 - 4).to(cuda)

    def run(path):
        zl, tr, _ = ops.kl_integrate(z0, S, T / S, gamma, L.DRIFT_GMM, mus, n_gaussian=K, seed=99, particle_offset=12345,
                                     traj_layout=L.TRAJ_TIME_SOA, emit_drift=True, path=path)
        torch.cuda.synchronize()
        return zl.clone(), tr.clone()

    zl_t, tr_t = run(L.PATH_TENSOR)
    assert ops.tensor_path_status() == 0
    zl_f, tr_f = run(L.PATH_FP32)
    if n % 128 == 0:  # the blocked trajectory layout holds the same numbers
        for path, zl_s, tr_s in ((L.PATH_TENSOR, zl_t, tr_t), (L.PATH_FP32, zl_f, tr_f)):
            zl_b, tr_b, _ = ops.kl_integrate(z0, S, T / S, gamma, L.DRIFT_GMM, mus, n_gaussian=K, seed=99,
                                             particle_offset=12345, traj_layout=L.TRAJ_BLOCK128, emit_drift=True, path=path)
            assert tr_b.shape == (S, n // 128, 3 * d, 128)
            assert torch.equal(unblock(tr_b), tr_s) and torch.equal(zl_b, zl_s)
    assert tr_t.shape == tr_f.shape == (3 * d, S, n)
    assert torch.isfinite(tr_t).all() and torch.isfinite(zl_t).all()
    # one step from identical states, then the emitted grad U of sample 0 (evaluated at step 1)
    assert relmax(tr_t[:2 * d, 0], tr_f[:2 * d, 0]) < 2e-4
    assert relmax(tr_t[2 * d:, 0], tr_f[2 * d:, 0]) < 2e-4
    assert relmax(tr_t, tr_f) < 2e-3
    assert relmax(zl_t, zl_f) < 2e-3
    # the emitted drift is grad U of the emitted position (closed form, float64)
    x = tr_t[:d, S // 2].t().double().cpu()
    gref = o_pot.vg_gmm_V(x, mus.double().cpu(), 1.0)
    assert relmax(tr_t[2 * d:, S // 2].t(), gref) < 2e-4


@pytest.mark.parametrize("drift,d,K", [("gmm", 8, 16), ("linear", 4, 0), ("linear", 16, 0), ("gmm", 32, 64)])
def test_block128_trajectory_layout(cuda, drift, d, K):
    """PDEIP_TRAJ_BLOCK128 ([S][N/128][3d][128]) carries exactly the numbers of PDEIP_TRAJ_TIME_SOA, from the
    production kernel and from the generic kernel; a ragged N is refused."""
    import os
    from pde_inverse_problem_b200 import ops, _lib as L
    n, S, T, gamma = 1024, 9, 0.18, 0.5
    g = torch.Generator().manual_seed(3 + d)
    z0 = torch.randn(n, 2 * d, generator=g).to(cuda)
    if drift == "gmm":
        params, kind = (torch.rand(K, d, generator=g) * 8 - 4).to(cuda), L.DRIFT_GMM
    else:
        f = torch.randn(d, d + 1, generator=g)
        params, kind = ((f @ f.T) / d).to(cuda), L.DRIFT_LINEAR

    def run(layout, nn=n):
        zl, tr, _ = ops.kl_integrate(z0[:nn], S, T / S, gamma, kind, params, n_gaussian=K, seed=5, traj_layout=layout,
                                     emit_drift=True)
        torch.cuda.synchronize()
        return zl.clone(), tr.clone()

    for generic in (False, True):
        if generic:
            os.environ["PDEIP_NO_FAST_INTEGRATOR"] = "1"
        try:
            zl_s, tr_s = run(L.TRAJ_TIME_SOA)
            zl_b, tr_b = run(L.TRAJ_BLOCK128)
        finally:
            os.environ.pop("PDEIP_NO_FAST_INTEGRATOR", None)
        assert torch.equal(unblock(tr_b), tr_s) and torch.equal(zl_b, zl_s)
    with pytest.raises(Exception):
        run(L.TRAJ_BLOCK128, nn=n - 1)


@pytest.mark.parametrize("d,K,n,layout_name", [(32, 64, 300, "soa"), (32, 64, 384, "block"), (8, 16, 1000, "soa"),
                                               (8, 16, 1024, "block"), (16, 20, 129, "soa")])
def test_tensor_core_integrator_writes_stay_in_bounds(cuda, d, K, n, layout_name):
    """Guard bands around the trajectory and final-state buffers (compute-sanitizer is not available on the pool):
    the tcgen05 integrator, ragged last tile included, writes every element of its outputs and nothing else."""
    from pde_inverse_problem_b200 import ops, _lib as L
    S, pad = 7, 4096
    layout = L.TRAJ_BLOCK128 if layout_name == "block" else L.TRAJ_TIME_SOA
    g = torch.Generator().manual_seed(1)
    z0 = torch.randn(n, 2 * d, generator=g).to(cuda)
    mus = (torch.rand(K, d, generator=g) * 8 - 4).to(cuda)
    need = 3 * d * S * n
    sentinel = float("nan")
    tbuf = torch.full((need + 2 * pad,), sentinel, device=cuda)
    zbuf = torch.full((2 * d * n + 2 * pad,), sentinel, device=cuda)
    zl, tr, _ = ops.kl_integrate(z0, S, 0.01, 0.5, L.DRIFT_GMM, mus, n_gaussian=K, seed=3, traj_layout=layout,
                                 emit_drift=True, path=L.PATH_TENSOR, traj_out=tbuf[pad:pad + need],
                                 z_last_out=zbuf[pad:pad + 2 * d * n].view(n, 2 * d))
    torch.cuda.synchronize()
    assert ops.tensor_path_status() == 0
    assert torch.isnan(tbuf[:pad]).all() and torch.isnan(tbuf[pad + need:]).all()
    assert torch.isnan(zbuf[:pad]).all() and torch.isnan(zbuf[pad + 2 * d * n:]).all()
    assert torch.isfinite(tr).all() and torch.isfinite(zl).all()


@pytest.mark.parametrize("d,K,n,S", [(8, 16, 256, 40), (32, 64, 256, 40), (16, 20, 384, 25)])
def test_tensor_core_integrator_matches_float64_oracle(cuda, d, K, n, S):
    """The production integrator (GMM drift on tcgen05, in-register Philox noise, BLOCK128 trajectory with grad U)
    against the float64 restatement of utils/sampling_utils.py:6-52 + core/potential.py:32-61, fed with the same Philox
    draws (pdeip_philox_normals / _uniforms).  Tolerance class "bf16 GEMM paths" (1e-2); asserted at 2e-3."""
    from pde_inverse_problem_b200 import ops, _lib as L
    T, gamma, seed, off = 0.01 * S, 0.5, 77, 4242
    dt = T / S
    g = torch.Generator().manual_seed(d + K)
    z0 = torch.randn(n, 2 * d, generator=g, dtype=torch.float64) * torch.cat([torch.full((d,), 2.0), torch.full((d,), 0.3)]).double()
    mus = torch.rand(K, d, generator=g, dtype=torch.float64) * 8 - 4
    noise = ops.philox_normals(n, S + 1, d, seed=seed, particle_offset=off, device=cuda).double().cpu()
    tau0 = ops.philox_uniforms(n, seed=seed, particle_offset=off, device=cuda).float().mul(dt).double().cpu()
    last, traj, _ = o_int.underdamped_langevin_dynamics_scan(z0.float().double(), S, dt, noise, tau0,
                                                             o_pot.GMMPotential(mus.float().double(), 1.0).gradient, gamma)
    zl, tr, _ = ops.kl_integrate(z0.float().to(cuda), S, dt, gamma, L.DRIFT_GMM, mus.float().to(cuda), n_gaussian=K,
                                 seed=seed, particle_offset=off, traj_layout=L.TRAJ_BLOCK128, emit_drift=True,
                                 path=L.PATH_TENSOR)
    torch.cuda.synchronize()
    assert ops.tensor_path_status() == 0
    got = tr.permute(1, 3, 0, 2).reshape(n, S, 3 * d)
    assert relmax(got[..., : 2 * d], traj) < 2e-3
    assert relmax(zl, last) < 2e-3
    gref = o_pot.vg_gmm_V(traj[..., :d].reshape(-1, d), mus.float().double(), 1.0).reshape(n, S, d)
    assert relmax(got[..., 2 * d:], gref) < 2e-3


@pytest.mark.parametrize("d,K", [(8, 16), (32, 64)])
def test_tensor_core_integrator_full_length_vs_float64_oracle_and_fp32_twin(cuda, d, K):
    """The C3 / C5 trajectory length (S = 200, dt = 0.01) on the tcgen05 integrator against the float64 oracle on the same
    Philox draws (round-1 verdict: checked only to S <= 40).  GMM trajectories are locally unstable between modes, so a
    few particles amplify any rounding difference (the float32 twin of the SAME code is 3e-5 off in the
    max-norm after 200 steps).  Criterion: per-step error statistics against float64, for the tensor
    kernel, the fp32 kernel and the CPU float32 twin; asserted: the fp32 kernel within 2x of the twin's band, the tensor
    kernel (operands split hi + lo, ~2^-17 per contraction) with median |err| < 1e-4, 99.9 % of the entries within 1e-2
    of the trajectory scale, and the ensemble mean / second moment of the terminal state within 1e-3."""
    from pde_inverse_problem_b200 import ops, _lib as L
    n, S, gamma, seed = 512, 200, 0.5, 123
    dt = 0.01
    g = torch.Generator().manual_seed(d * 3 + K)
    z0 = (torch.randn(n, 2 * d, generator=g) * torch.cat([torch.full((d,), 2.0), torch.full((d,), 0.316)])).double()
    mus = (torch.rand(K, d, generator=g) * 8 - 4).double()
    noise = ops.philox_normals(n, S + 1, d, seed=seed, device=cuda).double().cpu()
    tau0 = ops.philox_uniforms(n, seed=seed, device=cuda).float().mul(dt).double().cpu()
    grad64 = o_pot.GMMPotential(mus, 1.0).gradient
    last, traj, _ = o_int.underdamped_langevin_dynamics_scan(z0, S, dt, noise, tau0, grad64, gamma)
    # CPU float32 twin of the same restatement
    l32, t32, _ = o_int.underdamped_langevin_dynamics_scan(z0.float(), S, dt, noise.float(), tau0.float(),
                                                           o_pot.GMMPotential(mus.float(), 1.0).gradient, gamma)
    scale = traj.abs().max().item()

    def stats(tr):
        e = (tr.double().cpu() - traj).abs() / scale
        return e.max().item(), e.median().item(), (e < 1e-2).double().mean().item()

    out = {}
    for name, path in (("fp32", L.PATH_FP32), ("tensor", L.PATH_TENSOR)):
        zl, tr, _ = ops.kl_integrate(z0.float().to(cuda), S, dt, gamma, L.DRIFT_GMM, mus.float().to(cuda), n_gaussian=K,
                                     seed=seed, traj_layout=L.TRAJ_BLOCK128, emit_drift=True, path=path)
        torch.cuda.synchronize()
        got = tr.permute(1, 3, 0, 2).reshape(n, S, 3 * d)[..., : 2 * d]
        out[name] = stats(got) + (zl.double().cpu(),)
    assert ops.tensor_path_status() == 0
    twin = stats(t32)
    print(f"S=200 d={d} K={K} (max, median, frac<1e-2) vs float64: twin {twin}, fp32 kernel {out['fp32'][:3]}, "
          f"tensor kernel {out['tensor'][:3]}")
    assert out["fp32"][0] <= 2.0 * max(twin[0], 1e-5) or out["fp32"][2] > 0.999, (twin, out["fp32"][:3])
    assert out["fp32"][1] <= 2.0 * max(twin[1], 1e-7)
    assert out["tensor"][1] < 1e-4 and out["tensor"][2] > 0.999, out["tensor"][:3]
    zl = out["tensor"][3]
    assert (zl.mean(0) - last.mean(0)).abs().max() < 1e-3 * max(1.0, last.abs().max().item())
    m2, m2r = (zl ** 2).mean(0), (last ** 2).mean(0)
    assert ((m2 - m2r).abs() / m2r.abs().max()).max() < 1e-3
