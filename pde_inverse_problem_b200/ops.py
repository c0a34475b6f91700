"""Thin tensor-level wrappers over the C ABI (include/pdeip.h).

torch supplies device memory and the current CUDA stream; all arithmetic happens inside
libpdeip.so.  Every wrapper insists on CUDA float32 tensors and raises otherwise — there is no
CPU or eager fallback.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib as L
from ._lib import PdeipError  # noqa: F401  (re-export)

# count of libpdeip kernel launches issued through this module (bench.py reports it)
launch_counter = {"n": 0}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32(t: Optional[torch.Tensor], name: str, allow_none: bool = False) -> Optional[torch.Tensor]:
    if t is None:
        if allow_none:
            return None
        raise PdeipError(f"{name} is required")
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise PdeipError(f"{name} must be a CUDA tensor (got {type(t).__name__} on "
                         f"{getattr(t, 'device', 'n/a')}); the pdeip path has no CPU fallback")
    if t.dtype != torch.float32:
        raise PdeipError(f"{name} must be float32 (got {t.dtype})")
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------------
# K1 integrator
# ------------------------------------------------------------------------------------------------
def kl_integrate(z0: torch.Tensor, n_steps: int, dt: float, gamma: float, drift_kind: int,
                 drift_params: Optional[torch.Tensor] = None, n_gaussian: int = 0, sigma: float = 1.0,
                 noise: Optional[torch.Tensor] = None, tau0: Optional[torch.Tensor] = None, seed: int = 0,
                 particle_offset: int = 0, step_offset: int = 0, schedule: int = L.SCHEDULE_REFERENCE,
                 state_layout: int = L.LAYOUT_AOS, traj_layout: int = L.TRAJ_PARTICLE_MAJOR,
                 want_traj: bool = True, want_tau: bool = False, emit_every: int = 1, emit_offset: int = 0,
                 traj_out: Optional[torch.Tensor] = None, z_last_out: Optional[torch.Tensor] = None,
                 emit_drift: bool = False, path: int = L.PATH_FP32,
                 ) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
    """pdeip_kl_integrate.  z0: [N,2d] (AOS) or [2d,N] (SOA).  Returns (z_last, traj|None, tau|None).
    emit_drift: every trajectory sample is [x, v, grad U(x)] (3d components).
    path: PATH_TENSOR runs the GMM drift contraction on tcgen05 where that kernel exists (see pdeip.h)."""
    lib = L.load()
    z0 = _f32(z0, "z0")
    if state_layout == L.LAYOUT_AOS:
        n, two_d = z0.shape
    else:
        two_d, n = z0.shape
    d = two_d // 2
    drift_params = _f32(drift_params, "drift_params", allow_none=True)
    noise = _f32(noise, "noise", allow_none=True)
    tau0 = _f32(tau0, "tau0", allow_none=True)
    n_draws = n_steps + 1 if schedule == L.SCHEDULE_REFERENCE else n_steps
    if noise is not None and tuple(noise.shape) != (n, n_draws, d):
        raise PdeipError(f"noise must have shape {(n, n_draws, d)}, got {tuple(noise.shape)}")
    if tau0 is not None and tuple(tau0.shape) != (n,):
        raise PdeipError(f"tau0 must have shape {(n,)}, got {tuple(tau0.shape)}")
    z_last = z_last_out if z_last_out is not None else torch.empty_like(z0)
    s_emit = (n_steps - emit_offset + emit_every - 1) // emit_every
    traj = None
    if want_traj:
        width = two_d + d if emit_drift else two_d
        shape = {L.TRAJ_PARTICLE_MAJOR: (n, s_emit, width), L.TRAJ_TIME_MAJOR: (s_emit, n, width),
                 L.TRAJ_TIME_SOA: (width, s_emit, n), L.TRAJ_BLOCK128: (s_emit, n // 128, width, 128)}[traj_layout]
        if traj_layout == L.TRAJ_BLOCK128 and n % 128:
            raise PdeipError(f"TRAJ_BLOCK128 needs n % 128 == 0, got n={n}")
        if traj_out is not None:
            if traj_out.numel() < n * s_emit * width:
                raise PdeipError("traj_out too small")
            traj = _f32(traj_out, "traj_out").view(-1)[: n * s_emit * width].view(shape)
        else:
            traj = torch.empty(shape, device=z0.device, dtype=torch.float32)
    tau = torch.empty((n, n_steps), device=z0.device, dtype=torch.float32) if want_tau else None
    st = lib.pdeip_kl_integrate_path(_ptr(z0), _ptr(z_last), _ptr(traj), _ptr(tau), n, d, n_steps, dt, gamma,
                                     drift_kind, _ptr(drift_params), n_gaussian, sigma, _ptr(noise), _ptr(tau0),
                                     seed & 0xFFFFFFFFFFFFFFFF, particle_offset, step_offset, schedule,
                                     state_layout, traj_layout, emit_every, emit_offset, 1 if emit_drift else 0,
                                     path, _stream())
    L.check(st, "pdeip_kl_integrate_path")
    launch_counter["n"] += 1
    return z_last, traj, tau


def philox_normals(n: int, n_draws: int, d: int, seed: int, particle_offset: int = 0, step_offset: int = 0,
                   device="cuda") -> torch.Tensor:
    out = torch.empty((n, n_draws, d), device=device, dtype=torch.float32)
    L.check(L.load().pdeip_philox_normals(_ptr(out), n, n_draws, d, seed & 0xFFFFFFFFFFFFFFFF, particle_offset,
                                          step_offset, _stream()), "pdeip_philox_normals")
    launch_counter["n"] += 1
    return out


def philox_uniforms(n: int, seed: int, particle_offset: int = 0, device="cuda") -> torch.Tensor:
    out = torch.empty((n,), device=device, dtype=torch.float32)
    L.check(L.load().pdeip_philox_uniforms(_ptr(out), n, seed & 0xFFFFFFFFFFFFFFFF, particle_offset, _stream()),
            "pdeip_philox_uniforms")
    launch_counter["n"] += 1
    return out


def philox_raw(ctr: torch.Tensor, key: torch.Tensor) -> torch.Tensor:
    """ctr [n,4] int32 (bit pattern of uint32), key [2] int32 -> [n,4] int32."""
    if not (ctr.is_cuda and key.is_cuda and ctr.dtype == torch.int32 and key.dtype == torch.int32):
        raise PdeipError("philox_raw expects CUDA int32 tensors")
    ctr = ctr.contiguous()
    out = torch.empty_like(ctr)
    L.check(L.load().pdeip_philox_raw(_ptr(ctr), _ptr(key.contiguous()), _ptr(out), ctr.shape[0], _stream()),
            "pdeip_philox_raw")
    launch_counter["n"] += 1
    return out


def gaussian_sample(n: int, dim: int, mu: Optional[torch.Tensor], cov_half: Optional[torch.Tensor], seed: int,
                    particle_offset: int = 0, layout: int = L.LAYOUT_AOS, device="cuda") -> torch.Tensor:
    mu = _f32(mu, "mu", allow_none=True)
    cov_half = _f32(cov_half, "cov_half", allow_none=True)
    shape = (n, dim) if layout == L.LAYOUT_AOS else (dim, n)
    out = torch.empty(shape, device=device, dtype=torch.float32)
    L.check(L.load().pdeip_gaussian_sample(_ptr(out), n, dim, _ptr(mu), _ptr(cov_half),
                                           seed & 0xFFFFFFFFFFFFFFFF, particle_offset, layout, _stream()),
            "pdeip_gaussian_sample")
    launch_counter["n"] += 1
    return out


def gaussian_sample_grouped(n_groups: int, per_group: int, dim: int, mus: torch.Tensor, cov_halves: torch.Tensor,
                            seed: int, particle_offset: int = 0) -> torch.Tensor:
    """[n_groups, per_group, dim] samples, one (mean, cov_half) per group."""
    mus = _f32(mus, "mus")
    cov_halves = _f32(cov_halves, "cov_halves")
    if tuple(mus.shape) != (n_groups, dim) or tuple(cov_halves.shape) != (n_groups, dim, dim):
        raise PdeipError("mus must be [G,dim] and cov_halves [G,dim,dim]")
    out = torch.empty((n_groups, per_group, dim), device=mus.device, dtype=torch.float32)
    L.check(L.load().pdeip_gaussian_sample_grouped(_ptr(out), n_groups, per_group, dim, _ptr(mus), _ptr(cov_halves),
                                                   seed & 0xFFFFFFFFFFFFFFFF, particle_offset, _stream()),
            "pdeip_gaussian_sample_grouped")
    launch_counter["n"] += 1
    return out


def ou_exact_sample(n: int, U: torch.Tensor, s: torch.Tensor, B0: torch.Tensor, B: torch.Tensor, mt0: torch.Tensor,
                    t_min: float, t_max: float, seed: int, particle_offset: int = 0, want_t: bool = False):
    U, s, B0, B, mt0 = (_f32(t, nm) for t, nm in ((U, "U"), (s, "s"), (B0, "B0"), (B, "B"), (mt0, "mt0")))
    d = s.numel()
    out = torch.empty((n, d), device=U.device, dtype=torch.float32)
    out_t = torch.empty((n,), device=U.device, dtype=torch.float32) if want_t else None
    L.check(L.load().pdeip_ou_exact_sample(_ptr(out), _ptr(out_t), n, d, _ptr(U), _ptr(s), _ptr(B0), _ptr(B),
                                           _ptr(mt0), float(t_min), float(t_max), seed & 0xFFFFFFFFFFFFFFFF,
                                           particle_offset, _stream()), "pdeip_ou_exact_sample")
    launch_counter["n"] += 1
    return (out, out_t) if want_t else out


# ------------------------------------------------------------------------------------------------
# K2 potentials
# ------------------------------------------------------------------------------------------------
def gmm_value_grad(x: torch.Tensor, mus: torch.Tensor, sigma: float = 1.0, want_value: bool = False,
                   want_grad: bool = True) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    x = _f32(x, "x")
    mus = _f32(mus, "mus")
    n, d = x.shape
    if mus.shape[1] != d:
        raise PdeipError(f"mus must be [K,{d}], got {tuple(mus.shape)}")
    val = torch.empty((n,), device=x.device, dtype=torch.float32) if want_value else None
    grd = torch.empty_like(x) if want_grad else None
    L.check(L.load().pdeip_gmm_value_grad(_ptr(x), _ptr(mus), mus.shape[0], float(sigma), _ptr(val), _ptr(grd),
                                          n, d, _stream()), "pdeip_gmm_value_grad")
    launch_counter["n"] += 1
    return val, grd


def linear_grad(x: torch.Tensor, A: torch.Tensor) -> torch.Tensor:
    x = _f32(x, "x")
    A = _f32(A, "A")
    n, d = x.shape
    out = torch.empty_like(x)
    L.check(L.load().pdeip_linear_grad(_ptr(x), _ptr(A), _ptr(out), n, d, _stream()), "pdeip_linear_grad")
    launch_counter["n"] += 1
    return out


# ------------------------------------------------------------------------------------------------
# model evaluation and residual
# ------------------------------------------------------------------------------------------------
class ModelSpec:
    """(model_kind, d, hidden, layers, n_gaussian) + the flat parameter count."""

    def __init__(self, model_kind: int, d: int, hidden: int = 0, layers: int = 0, n_gaussian: int = 0):
        self.kind, self.d, self.hidden, self.layers, self.n_gaussian = model_kind, d, hidden, layers, n_gaussian
        self.num_params = int(L.load().pdeip_model_num_params(model_kind, d, hidden, layers, n_gaussian))
        if self.num_params <= 0:
            raise PdeipError(f"unknown model kind {model_kind}")

    def args(self):
        return (self.d, self.hidden, self.layers, self.n_gaussian)


def model_eval(spec: ModelSpec, params: torch.Tensor, x: torch.Tensor, v: Optional[torch.Tensor] = None,
               want: Sequence[str] = ("value", "grad")) -> Dict[str, torch.Tensor]:
    params = _f32(params, "params")
    x = _f32(x, "x")
    v = _f32(v, "v", allow_none=True)
    n, d = x.shape
    if d != spec.d or params.numel() != spec.num_params:
        raise PdeipError(f"shape mismatch: x is [{n},{d}], params has {params.numel()} (expected d={spec.d}, "
                         f"{spec.num_params} params)")
    out = {}
    if "value" in want:
        out["value"] = torch.empty((n,), device=x.device, dtype=torch.float32)
    if "grad" in want:
        out["grad"] = torch.empty((n, d), device=x.device, dtype=torch.float32)
    if "vHv" in want:
        out["vHv"] = torch.empty((n,), device=x.device, dtype=torch.float32)
    if "laplacian" in want:
        out["laplacian"] = torch.empty((n,), device=x.device, dtype=torch.float32)
    L.check(L.load().pdeip_model_eval(spec.kind, _ptr(params), spec.d, spec.hidden, spec.layers, spec.n_gaussian,
                                      _ptr(x), _ptr(v), _ptr(out.get("value")), _ptr(out.get("grad")),
                                      _ptr(out.get("vHv")), _ptr(out.get("laplacian")), n, _stream()),
            "pdeip_model_eval")
    launch_counter["n"] += 1
    return out


class TrueGrad:
    """Specification of grad V_true for the 0T sets (LINEAR: A [d,d]; GMM: mus [K,d], sigma)."""

    def __init__(self, kind: int = L.DRIFT_NONE, params: Optional[torch.Tensor] = None, sigma: float = 1.0):
        self.kind = kind
        self.params = _f32(params, "true_params", allow_none=(kind in (L.DRIFT_NONE, L.DRIFT_IN_POINTS)))
        self.n_gaussian = int(self.params.shape[0]) if kind == L.DRIFT_GMM else 0
        self.sigma = float(sigma)


class ResidualAccumulator:
    """begin / accumulate / finalize protocol of the residual kernels (include/pdeip.h)."""

    def __init__(self, spec: ModelSpec, device="cuda"):
        self.spec = spec
        lib = L.load()
        self.ws_bytes = int(lib.pdeip_residual_workspace_bytes(spec.kind, *spec.args()))
        self.ws = torch.empty((self.ws_bytes // 4,), device=device, dtype=torch.float32)
        self.sums = torch.empty((L.NUM_SUMS,), device=device, dtype=torch.float32)
        self.grad = torch.empty((spec.num_params,), device=device, dtype=torch.float32)

    def begin(self):
        L.check(L.load().pdeip_residual_begin(_ptr(self.ws), self.ws_bytes, self.spec.kind, *self.spec.args(),
                                              _stream()), "pdeip_residual_begin")
        return self

    def accumulate(self, set_kind: int, params: torch.Tensor, points: torch.Tensor, weight: float, coef: float = 0.0,
                   layout: int = L.LAYOUT_AOS, true_grad: Optional[TrueGrad] = None, path: int = L.PATH_FP32,
                   n_points: Optional[int] = None):
        params = _f32(params, "params")
        points = _f32(points, "points")
        if params.numel() != self.spec.num_params:
            raise PdeipError(f"params has {params.numel()} entries, expected {self.spec.num_params}")
        kinetic = set_kind in (L.SET_KFP_0T, L.SET_KFP_BOUNDARY)
        dim = 2 * self.spec.d if kinetic else self.spec.d
        if true_grad is not None and true_grad.kind == L.DRIFT_IN_POINTS:
            dim += self.spec.d  # grad V_true stored after the point's own components
        if n_points is None:
            if layout == L.LAYOUT_AOS:
                if points.ndim != 2 or points.shape[1] != dim:
                    raise PdeipError(f"points must be [n,{dim}], got {tuple(points.shape)}")
                n_points = points.shape[0]
            elif layout == L.LAYOUT_BLOCK128:
                if points.ndim != 3 or tuple(points.shape[1:]) != (dim, 128):
                    raise PdeipError(f"points must be [n/128,{dim},128], got {tuple(points.shape)}")
                n_points = points.shape[0] * 128
            else:
                if points.ndim != 2 or points.shape[0] != dim:
                    raise PdeipError(f"points must be [{dim},n], got {tuple(points.shape)}")
                n_points = points.shape[1]
        tg = true_grad or TrueGrad()
        L.check(L.load().pdeip_residual_accumulate(
            _ptr(self.ws), self.ws_bytes, set_kind, self.spec.kind, _ptr(params), *self.spec.args(),
            _ptr(points), n_points, layout, float(weight), float(coef), tg.kind, _ptr(tg.params), tg.n_gaussian,
            tg.sigma, path, _stream()), "pdeip_residual_accumulate")
        launch_counter["n"] += 1
        return self

    def finalize(self) -> Tuple[torch.Tensor, torch.Tensor]:
        L.check(L.load().pdeip_residual_finalize(_ptr(self.ws), self.ws_bytes, self.spec.kind, *self.spec.args(),
                                                 _ptr(self.sums), _ptr(self.grad), _stream()),
                "pdeip_residual_finalize")
        launch_counter["n"] += 2
        return self.sums, self.grad


# ------------------------------------------------------------------------------------------------
# K6 / K7 / gather
# ------------------------------------------------------------------------------------------------
def adam_l2_step(params: torch.Tensor, grad: torch.Tensor, m: torch.Tensor, v: torch.Tensor, count: int, lr: float,
                 b1: float = 0.9, b2: float = 0.999, eps: float = 1e-4, weight_decay: float = 1e-3,
                 grad_scale: float = 1.0, ema: Optional[torch.Tensor] = None, use_ema: bool = False,
                 ema_decay: float = 0.999, norms: Optional[torch.Tensor] = None) -> torch.Tensor:
    """In-place optimizer step on the flat parameter buffer; returns norms [grad_norm, params_norm]."""
    for name, t in (("params", params), ("grad", grad), ("m", m), ("v", v)):
        if _f32(t, name) is not t:
            raise PdeipError(f"{name} must be contiguous (updated in place)")
    if norms is None:
        norms = torch.empty((2,), device=params.device, dtype=torch.float32)
    L.check(L.load().pdeip_adam_l2_step(_ptr(params), _ptr(grad), _ptr(m), _ptr(v), _ptr(ema), params.numel(),
                                        float(lr), b1, b2, eps, weight_decay, int(count), float(grad_scale),
                                        1 if use_ema else 0, ema_decay, _ptr(norms), _stream()),
            "pdeip_adam_l2_step")
    launch_counter["n"] += 1
    return norms


def ensemble_moments(z: torch.Tensor, layout: int = L.LAYOUT_AOS) -> Tuple[torch.Tensor, torch.Tensor]:
    """Raw sums (sum z [dim], sum z z^T [dim,dim])."""
    z = _f32(z, "z")
    n, dim = z.shape if layout == L.LAYOUT_AOS else (z.shape[1], z.shape[0])
    lib = L.load()
    ws_bytes = int(lib.pdeip_moments_workspace_bytes(dim))
    ws = torch.empty((ws_bytes // 4,), device=z.device, dtype=torch.float32)
    out = torch.empty((dim + dim * dim,), device=z.device, dtype=torch.float32)
    L.check(lib.pdeip_ensemble_moments(_ptr(z), n, dim, layout, _ptr(out), _ptr(ws), ws_bytes, _stream()),
            "pdeip_ensemble_moments")
    launch_counter["n"] += 2
    return out[:dim], out[dim:].view(dim, dim)


def gather_0T(dataset: torch.Tensor, sample_index: torch.Tensor, interval: int, shift: int) -> torch.Tensor:
    """methods/consistency.py:102-118 as one device gather.  dataset [n_traj,n_time,dim]."""
    dataset = _f32(dataset, "dataset")
    if not (sample_index.is_cuda and sample_index.dtype == torch.int64):
        raise PdeipError("sample_index must be a CUDA int64 tensor")
    n_traj, n_time, dim = dataset.shape
    n_time_sel = n_time // interval
    n_sel = sample_index.numel()
    out = torch.empty((n_sel * n_time_sel, dim), device=dataset.device, dtype=torch.float32)
    L.check(L.load().pdeip_gather_0T(_ptr(dataset), n_traj, n_time, dim, _ptr(sample_index.contiguous()), n_sel,
                                     interval, shift, n_time_sel, _ptr(out), _stream()), "pdeip_gather_0T")
    launch_counter["n"] += 1
    return out


# ------------------------------------------------------------------------------------------------
# KMV pairwise residual (kinetic_mckean_vlasov.py:11-120)
# ------------------------------------------------------------------------------------------------
def kmv_mean_grad(spec: ModelSpec, params: torch.Tensor, xv: torch.Tensor, true_A: Optional[torch.Tensor] = None):
    """G[j,t] = mean_i grad Phi(x[j,t] - x[i,t]) and, if true_A is given, the same for Phi_true = D'AD/2.
    xv: [n, nt, 2d]."""
    params = _f32(params, "params")
    xv = _f32(xv, "xv")
    true_A = _f32(true_A, "true_A", allow_none=True)
    n, nt, two_d = xv.shape
    d = two_d // 2
    lib = L.load()
    ws_bytes = int(lib.pdeip_kmv_workspace_bytes(n, nt, d))
    ws = torch.empty((ws_bytes // 4,), device=xv.device, dtype=torch.float32)
    G = torch.empty((n, nt, d), device=xv.device, dtype=torch.float32)
    Gt = torch.empty_like(G) if true_A is not None else None
    L.check(lib.pdeip_kmv_mean_grad(spec.kind, _ptr(params), d, spec.hidden, spec.layers, _ptr(xv), n, nt, _ptr(G),
                                    _ptr(Gt), _ptr(true_A), _ptr(ws), ws_bytes, _stream()), "pdeip_kmv_mean_grad")
    launch_counter["n"] += 2
    return G, Gt


def kmv_value_and_grad(model, params, flat: torch.Tensor, xv: torch.Tensor, c: torch.Tensor, true_A: torch.Tensor,
                       acc: "ResidualAccumulator", result_dict):
    """Phase 1 (mean gradient per sample), then the pair set with the extra direction G_j and kappa = 2 c_j."""
    xv = _f32(xv, "xv")
    c = _f32(c, "c")
    n, nt, _ = xv.shape
    spec = model.spec
    G, Gt = kmv_mean_grad(spec, flat, xv, true_A)
    acc.begin()
    L.check(L.load().pdeip_residual_accumulate_kmv(_ptr(acc.ws), acc.ws_bytes, spec.kind, _ptr(flat), spec.d,
                                                   spec.hidden, spec.layers, _ptr(xv), n, nt, _ptr(G), _ptr(Gt),
                                                   _ptr(c), 1.0 / (float(n) * n * nt), _stream()),
            "pdeip_residual_accumulate_kmv")
    launch_counter["n"] += 2
    sums, grad = acc.finalize()
    return result_dict(model, params, sums, grad)


def tensor_path_status() -> int:
    """0 if every tcgen05 phase of the tensor-path launches so far completed (synchronises the stream)."""
    import ctypes as C
    out = C.c_int(-1)
    L.check(L.load().pdeip_tensor_path_status(_stream(), C.addressof(out)), "pdeip_tensor_path_status")
    return int(out.value)
