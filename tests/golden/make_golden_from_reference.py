"""Regenerates the golden vectors of tests/golden/ from the REAL reference (JAX) — the pin the oracle is missing.

Status: this image has no jax / jaxlib / flax / optax (and no network), so this script CANNOT run here and the
committed tests/golden/*.npz are oracle outputs ("parity unpinned").  The day the reference's
dependencies are importable, ONE command closes that gap:

    python tests/golden/make_golden_from_reference.py /path/to/PDE-inverse-problem

It reads the INPUTS of the committed fixtures (same arrays, same keys), evaluates them with the reference's own
functions in float64 (jax_enable_x64), writes tests/golden/ref_<name>.npz with the SAME output keys, and
tests/test_oracle_kat.py::test_oracle_matches_reference_golden_when_present then checks the oracle against them
(the test is skipped while the ref_*.npz files are absent).

What is evaluated, by reference file:line:
  integrator_gmm / integrator_ou   utils/sampling_utils.py:6-52 — the un-vmapped body (`__wrapped__` of the jax.vmap
                                   decorator at :25) is run per particle with `random` and `scan` of that module
                                   replaced by shims that return the INJECTED normals / uniforms of the fixture, so
                                   the reference's own update_step arithmetic (:14-20) and step schedule (:32-46) run
                                   on identical noise;  core/potential.py:32-61 for the GMM gradient and value
  model_eval                       core/model.py:32-62 (flax V_hypothesis), jax.grad, utils/common_utils.py:6-14,
                                   jax.jacfwd(jax.grad) for the Laplacian
  kfp_residual / fp_residual       methods/consistency_instances/kinetic_fokker_planck.py:11-69, fokker_planck.py:33-63
                                   with stand-in problem objects carrying V_true_fn, gamma_friction, total_evolving_time
  adam                             main.py:11-29 (optax.chain(add_decayed_weights, adam), cosine schedule) and the update
                                   of core/trainer.py:61-64
philox.npz holds this repo's own counter-based stream (the reference uses threefry): nothing to regenerate.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _stub_missing_modules():
    """api.py / example_problems import plotting and logging packages that the hot path never calls."""
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn", "wandb", "hydra", "omegaconf"):
        try:
            __import__(name)
        except Exception:
            mod = types.ModuleType(name)
            mod.__dict__.setdefault("DictConfig", dict)
            mod.__dict__.setdefault("OmegaConf", object)
            mod.__dict__.setdefault("main", lambda *a, **k: (lambda f: f))
            sys.modules[name] = mod


def main(ref_root: str) -> None:
    import jax
    jax.config.update("jax_enable_x64", True)
    import jax.numpy as jnp
    _stub_missing_modules()
    sys.path.insert(0, ref_root)
    from core import potential as r_pot
    from core.model import V_hypothesis
    from utils import sampling_utils as r_su
    from utils.common_utils import hessian_vector_product

    def load(name):
        return dict(np.load(os.path.join(HERE, name + ".npz")))

    def save(name, **arrays):
        np.savez(os.path.join(HERE, "ref_" + name + ".npz"), **{k: np.asarray(v) for k, v in arrays.items()})

    # ---- integrator with injected noise: shims for `random` and `scan` inside utils/sampling_utils.py ----------------
    body = r_su.underdamped_langevin_dynamics_scan.__wrapped__   # the per-particle function under jax.vmap (:25)

    class Key:  # (particle, number of normals drawn so far); split() hands the SAME counter to key and subkey
        def __init__(self, n, s):
            self.n, self.s = n, s

    def run_integrator(g, potential_grad, gamma):
        z0, noise, tau0 = g["z0"], g["noise"], g["tau0"]
        S, dt = int(g["S"]), float(g["dt"])
        lasts, trajs, taus = [], [], []
        for n in range(z0.shape[0]):
            draws = {"i": 0}

            class FakeRandom:
                @staticmethod
                def split(key, num=2):
                    return tuple(key for _ in range(num))

                @staticmethod
                def normal(key, shape):          # one draw per update_step, in call order (:14)
                    xi = jnp.asarray(noise[n, draws["i"]])
                    draws["i"] += 1
                    return xi

                @staticmethod
                def uniform(key, shape):         # tau_0 = uniform * dt (:32)
                    return jnp.asarray(tau0[n] / dt)

            def fake_scan(f, init, xs, length):  # lax.scan(f, init, None, length) as a Python loop (:37-42)
                carry, ys = init, []
                for _ in range(length):
                    carry, y = f(carry, None)
                    ys.append(y)
                return carry, jnp.stack(ys)

            r_su.random, r_su.scan = FakeRandom, fake_scan
            last, traj, tau = body(jnp.asarray(z0[n]), S, dt, Key(n, 0), potential_grad, gamma)
            assert draws["i"] == S + 1
            lasts.append(last), trajs.append(traj), taus.append(tau)
        return np.stack(lasts), np.stack(trajs), np.stack(taus)

    g = load("integrator_gmm")
    mus = jnp.asarray(g["mus"])
    pot = r_pot.GMMPotential(mus, 1.0) if hasattr(r_pot, "GMMPotential") else None
    gmm_grad = (lambda x: r_pot.g_gmm_V(x, mus, 1.0)) if pot is None else pot.gradient
    last, traj, tau = run_integrator(g, gmm_grad, float(g["gamma"]))
    x0 = jnp.asarray(g["z0"][:, : g["mus"].shape[1]])
    save("integrator_gmm", last=last, traj=traj, tau=tau,
         gmm_grad=jax.vmap(lambda x: r_pot.g_gmm_V(x, mus, 1.0))(x0),
         gmm_value=jax.vmap(lambda x: r_pot.gmm_V(x, mus, 1.0))(x0))
    g = load("integrator_ou")
    F = jnp.asarray(g["F"])
    last, traj, _ = run_integrator(g, lambda x: F @ x, float(g["gamma"]))
    save("integrator_ou", last=last, traj=traj)

    # ---- model and residuals --------------------------------------------------------------------------------------
    def tree_of(flat, d):
        dims, off, tree = [d, 32, 32, 40], 0, {}
        for i in range(3):
            k = dims[i] * dims[i + 1]
            tree[f"layers_{i}"] = {"kernel": jnp.asarray(flat[off:off + k]).reshape(dims[i], dims[i + 1]),
                                   "bias": jnp.asarray(flat[off + k:off + k + dims[i + 1]])}
            off += k + dims[i + 1]
        return {"params": tree}

    def flat_of(tree):
        return np.concatenate([np.concatenate([np.asarray(tree["params"][f"layers_{i}"]["kernel"]).reshape(-1),
                                               np.asarray(tree["params"][f"layers_{i}"]["bias"]).reshape(-1)])
                               for i in range(3)])

    g = load("model_eval")
    d = g["x"].shape[1]
    net = V_hypothesis(output_dim=1, hidden_dims=[32, 32], dim=d)
    params = tree_of(g["params"], d)
    V = lambda x: net.apply(params, x)[0]
    x, v = jnp.asarray(g["x"]), jnp.asarray(g["v"])
    save("model_eval", value=jax.vmap(V)(x), grad=jax.vmap(jax.grad(V))(x),
         vHv=jax.vmap(lambda a, b: jnp.dot(b, hessian_vector_product(V, a, b)))(x, v),
         lap=jax.vmap(lambda a: jnp.trace(jax.jacfwd(jax.grad(V))(a)))(x))

    class Problem:  # what the residual modules read from pde_instance
        def __init__(self, V_true, gamma, T, dim):
            self.V_true_fn, self.total_evolving_time, self.dim = V_true, T, dim
            self.initial_configuration = {"gamma_friction": gamma}

    from methods.consistency_instances import fokker_planck as r_fp, kinetic_fokker_planck as r_kfp
    g = load("kfp_residual")
    tF = jnp.asarray(g["tilde_F"])
    pde = Problem(lambda x: jnp.dot(x, tF @ x) / 2, float(g["gamma"]), float(g["T"]), d)
    data = {k: jnp.asarray(g[f"data_{k}"]) for k in ("initial", "terminal", "0T")}
    out = r_kfp.value_and_grad_fn(net.apply, tree_of(g["params"], d), data, jax.random.PRNGKey(0), pde)
    save("kfp_residual", loss=out["loss"], grad=flat_of(out["grad"]), grad_norm=out["grad_norm"],
         loss_gt=out["loss ground truth"])
    g = load("fp_residual")
    Fm = jnp.asarray(g["F"])
    pde = Problem(lambda x: jnp.dot(x, Fm @ x) / 2, 0.0, float(g["T"]), d)
    data = {k: jnp.asarray(g[f"data_{k}"]) for k in ("initial", "terminal", "0T")}
    out = r_fp.value_and_grad_fn(net.apply, tree_of(g["params"], d), data, jax.random.PRNGKey(0), pde)
    save("fp_residual", loss=out["loss"], grad=flat_of(out["grad"]), grad_norm=out["grad_norm"],
         loss_gt=out["loss ground truth"])

    # ---- optimizer: main.py:11-29 + core/trainer.py:63-64 ------------------------------------------------------------
    import optax
    g = load("adam")
    lr = optax.cosine_decay_schedule(1e-2, 20000, 0.001)                                    # main.py:16
    opt = optax.chain(optax.add_decayed_weights(1e-3), optax.adam(learning_rate=lr, b1=0.9, eps=1e-4))  # main.py:20-26
    p = jnp.asarray(g["params"][0])
    st = opt.init(p)
    ps = [np.asarray(p)]
    for gr in g["grads"]:
        updates, st = opt.update(jnp.asarray(gr), st, p)
        p = optax.apply_updates(p, updates)
        ps.append(np.asarray(p))
    save("adam", params=np.stack(ps))
    print("reference golden vectors written to", HERE, "(ref_*.npz)")


if __name__ == "__main__":
    if len(sys.argv) != 2:
        raise SystemExit(__doc__)
    main(sys.argv[1])
