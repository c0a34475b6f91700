"""Mirror of core/model.py: hypothesis models whose forward / derivative evaluations run in libpdeip.

Parameter trees follow Flax ({"params": {"layers_i": {"kernel": [in,out], "bias": [out]}}}); every leaf
is a *view* into one flat float32 CUDA buffer so the kernels take a single pointer
(tree["_flat"] is that buffer; it is not a parameter leaf).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

from .. import _lib as L
from .. import ops

OUT_DIM = 40  # core/model.py:43


def _tree_from_flat(flat: torch.Tensor, layout: List[Tuple[str, str, Tuple[int, ...]]]) -> Dict:
    tree: Dict = {}
    off = 0
    for name, leaf, shape in layout:
        n = int(math.prod(shape))
        view = flat[off:off + n].view(shape)
        if leaf:
            tree.setdefault(name, {})[leaf] = view
        else:
            tree[name] = view
        off += n
    assert off == flat.numel()
    return {"params": tree, "_flat": flat}


def flat_of(params: Dict, layout) -> torch.Tensor:
    """The flat buffer behind a parameter tree (built on the fly if the tree is not one of ours)."""
    flat = params.get("_flat") if isinstance(params, dict) else None
    if flat is not None:
        return flat
    tree = params["params"]
    parts = []
    for name, leaf, shape in layout:
        t = tree[name][leaf] if leaf else tree[name]
        parts.append(t.reshape(-1))
    return torch.cat(parts).contiguous()


class _Model:
    spec: ops.ModelSpec

    def layout(self):
        raise NotImplementedError

    def tree(self, flat: torch.Tensor) -> Dict:
        return _tree_from_flat(flat, self.layout())

    def flat(self, params: Dict) -> torch.Tensor:
        return flat_of(params, self.layout())

    def apply(self, params: Dict, x: torch.Tensor) -> torch.Tensor:
        """forward_fn(params, x): x [d] -> [1] (as the reference, core/model.py:62) or x [N,d] -> [N]."""
        flat = self.flat(params)
        if x.ndim == 1:
            return ops.model_eval(self.spec, flat, x[None].contiguous(), want=("value",))["value"]
        return ops.model_eval(self.spec, flat, x, want=("value",))["value"]

    def gradient(self, params: Dict, x: torch.Tensor) -> torch.Tensor:
        """jax.vmap(jax.grad(V)) of the reference residual modules."""
        return ops.model_eval(self.spec, self.flat(params), x, want=("grad",))["grad"]


KERNEL_HIDDEN = 32  # width the CUDA MLP kernels are built for
MAX_LAYERS = 8      # deepest stack the fp32 kernels are instantiated for (tensor path: layers == 2)


class V_hypothesis(_Model):
    """core/model.py:32-62: Dense d -> [hidden]*layers -> 40, tanh, V = sum(out^2).

    Supported envelope: one common hidden width <= 32, 1 <= layers <= 8, dim <= 32 (the reference default
    configurations/neural_network/MLP.yaml is hidden_dim = 20, layers = 8; every script uses 32 x 2).  Narrower
    widths run on the 32-wide kernels through zero padding: the flat buffer the kernels see is the padded network
    (kernels [in_pad, out_pad], biases [out_pad]) and the parameter-tree leaves are the reference-shaped sub-views
    of it.  The padding is exactly invariant: a padded unit has z = 0, a = tanh 0 = 0 and outgoing weights 0, so
    every gradient entry of the padding is 0 and L2-in-Adam leaves it at 0 (u = 0 / (sqrt 0 + eps))."""

    def __init__(self, output_dim: int, hidden_dims, dim: int):
        self.output_dim = output_dim
        self.hidden_dims = list(hidden_dims)
        hs = set(self.hidden_dims)
        if len(hs) != 1:
            raise NotImplementedError(
                f"all hidden layers must share one width (got {self.hidden_dims}); supported: one width <= "
                f"{KERNEL_HIDDEN}, 1 <= layers <= {MAX_LAYERS}")
        self.hidden = self.hidden_dims[0]
        self.layers = len(self.hidden_dims)
        self.dim = dim
        if not 1 <= self.hidden <= KERNEL_HIDDEN or not 1 <= self.layers <= MAX_LAYERS:
            raise NotImplementedError(
                f"the CUDA MLP kernels support hidden_dim <= {KERNEL_HIDDEN} (zero-padded to {KERNEL_HIDDEN}) and "
                f"1 <= layers <= {MAX_LAYERS}; got hidden_dim = {self.hidden}, layers = {self.layers}. "
                f"Set neural_network.hidden_dim / neural_network.layers inside that envelope.")
        self.spec = ops.ModelSpec(L.MODEL_MLP, dim, KERNEL_HIDDEN, self.layers)

    def _dims(self, hidden):
        return [self.dim] + [hidden] * self.layers + [OUT_DIM]

    def layout(self):
        """(name, leaf, shape) of the flat buffer the kernels read (padded widths)."""
        dims = self._dims(KERNEL_HIDDEN)
        out = []
        for i in range(len(dims) - 1):
            out.append((f"layers_{i}", "kernel", (dims[i], dims[i + 1])))
            out.append((f"layers_{i}", "bias", (dims[i + 1],)))
        return out

    def tree(self, flat: torch.Tensor) -> Dict:
        t = _tree_from_flat(flat, self.layout())
        if self.hidden != KERNEL_HIDDEN:  # reference-shaped sub-views of the padded leaves
            dims = self._dims(self.hidden)
            for i in range(len(dims) - 1):
                leaf = t["params"][f"layers_{i}"]
                leaf["kernel"] = leaf["kernel"][: dims[i], : dims[i + 1]]
                leaf["bias"] = leaf["bias"][: dims[i + 1]]
        return t

    def flat(self, params: Dict) -> torch.Tensor:
        flat = params.get("_flat") if isinstance(params, dict) else None
        if flat is not None or self.hidden == KERNEL_HIDDEN:
            return flat_of(params, self.layout())
        # a foreign reference-shaped tree: scatter it into a zeroed padded buffer
        dev = params["params"]["layers_0"]["kernel"].device
        out = self.tree(torch.zeros(self.spec.num_params, device=dev, dtype=torch.float32))
        for name, leaf in params["params"].items():
            out["params"][name]["kernel"].copy_(leaf["kernel"])
            out["params"][name]["bias"].copy_(leaf["bias"])
        return out["_flat"]

    def init(self, rng, x: torch.Tensor) -> Dict:
        """net.init(PRNGKey(11), x): kaiming-normal kernels (std = sqrt(2/fan_in)), zero biases
        (core/model.py:42).  One-off host-side initialisation with torch's generator."""
        g = torch.Generator().manual_seed(int(rng) & 0x7FFFFFFFFFFFFFFF)
        dims = self._dims(self.hidden)
        tree = self.tree(torch.zeros(self.spec.num_params, device=x.device, dtype=torch.float32))
        for i in range(len(dims) - 1):
            w = torch.randn((dims[i], dims[i + 1]), generator=g, dtype=torch.float64) * math.sqrt(2.0 / dims[i])
            tree["params"][f"layers_{i}"]["kernel"].copy_(w.float())
        return tree


class V_parametric_GMM(_Model):
    """example_problems/kinetic_fokker_planck_example_GMM.py:214-234: learnable mus [K,d], sigma = 1."""

    def __init__(self, dim: int, n_Gaussians: int):
        self.dim, self.n_Gaussians = dim, n_Gaussians
        self.spec = ops.ModelSpec(L.MODEL_GMM, dim, n_gaussian=n_Gaussians)

    def layout(self):
        return [("mus", "", (self.n_Gaussians, self.dim))]

    def init(self, rng, x: torch.Tensor) -> Dict:
        g = torch.Generator().manual_seed(int(rng) & 0x7FFFFFFFFFFFFFFF)
        flat = torch.randn(self.n_Gaussians * self.dim, generator=g).to(x.device).contiguous()  # GMM.py:223
        return self.tree(flat)


class V_parametric_quadratic(_Model):
    """example_problems/kinetic_fokker_planck_example_OU.py:209-220: sum(y * Dense(d)(y))."""

    def __init__(self, dim: int):
        self.dim = dim
        self.spec = ops.ModelSpec(L.MODEL_QUADRATIC, dim)

    def layout(self):
        return [("tilde_F", "kernel", (self.dim, self.dim)), ("tilde_F", "bias", (self.dim,))]

    def init(self, rng, x: torch.Tensor) -> Dict:
        # flax Dense default: lecun_normal kernel (std = sqrt(1/fan_in)), zero bias
        g = torch.Generator().manual_seed(int(rng) & 0x7FFFFFFFFFFFFFFF)
        w = torch.randn(self.dim, self.dim, generator=g) * math.sqrt(1.0 / self.dim)
        flat = torch.cat([w.reshape(-1), torch.zeros(self.dim)]).to(x.device).contiguous()
        return self.tree(flat)


def get_model(cfg, DEBUG=False, pde_instance=None):
    """core/model.py:109-131."""
    if cfg.estimation_mode == "parametric":
        print("----Using parametric model----")
        return pde_instance.create_parametric_model()
    elif cfg.estimation_mode == "non-parametric":
        print("----Using non-parametric model----")
        if cfg.neural_network.n_resblocks > 0:
            raise NotImplementedError
        if DEBUG:
            raise NotImplementedError("V_hypothesis_DEBUG is never enabled in the reference (core/model.py:64-106)")
        return V_hypothesis(output_dim=1, hidden_dims=[cfg.neural_network.hidden_dim] * cfg.neural_network.layers,
                            dim=cfg.pde_instance.domain_dim)
    else:
        raise NotImplementedError


def model_of(forward_fn) -> _Model:
    """The residual modules receive `forward_fn = net.apply` (main.py:62).  The CUDA path needs the
    architecture, not a Python callable: recover the model object from the bound method."""
    owner = getattr(forward_fn, "__self__", None)
    if owner is None and hasattr(forward_fn, "func"):  # functools.partial(net.apply, params)
        owner = getattr(forward_fn.func, "__self__", None)
    if not isinstance(owner, _Model):
        raise NotImplementedError(
            "forward_fn must be the .apply of a pde_inverse_problem_b200 model (V_hypothesis / V_parametric_*): "
            "the residual runs as fused CUDA kernels and cannot trace an arbitrary Python callable")
    return owner
