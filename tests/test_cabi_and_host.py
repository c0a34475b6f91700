"""CPU: the C-ABI library loads and exports every symbol of include/pdeip.h (no compute calls without a GPU),
argument validation returns error codes, and the host-side logic (config, rng, sharding, packing, schedules,
parameter layout) behaves like the reference's."""
import ctypes
import math
import os

import pytest
import torch

from oracle import model as o_model, optim as o_optim


def test_library_exports_every_header_symbol():
    from pde_inverse_problem_b200 import _lib as L
    lib = L.load()
    declared = L.header_functions()
    assert len(declared) >= 20
    assert set(declared) == set(L.SIGNATURES), set(declared) ^ set(L.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pdeip_abi_version() == 1
    assert lib.pdeip_sm_count() >= 1


def test_argument_validation_without_gpu():
    """Invalid arguments are rejected before any CUDA call, with a message."""
    from pde_inverse_problem_b200 import _lib as L
    lib = L.load()
    st = lib.pdeip_kl_integrate(None, None, None, None, 10, 4, 5, 0.1, 1.0, L.DRIFT_NONE, None, 0, 1.0, None, None,
                                0, 0, 0, 0, 0, 0, 1, 0, 0, None)
    assert st == -1 and b"z0" in lib.pdeip_last_error()
    st = lib.pdeip_gmm_value_grad(1, 1, 3, 1.0, None, None, 10, 64, None)   # d > 32
    assert st == -2 and b"32" in lib.pdeip_last_error()
    assert lib.pdeip_model_num_params(L.MODEL_MLP, 8, 32, 2, 0) == 8 * 32 + 32 + 32 * 32 + 32 + 32 * 40 + 40
    assert lib.pdeip_model_num_params(L.MODEL_GMM, 8, 0, 0, 16) == 128
    assert lib.pdeip_model_num_params(L.MODEL_QUADRATIC, 4, 0, 0, 0) == 20
    assert lib.pdeip_model_num_params(77, 4, 0, 0, 0) < 0
    need = lib.pdeip_residual_workspace_bytes(L.MODEL_MLP, 8, 32, 2, 0)
    assert need >= 4 * lib.pdeip_sm_count() * (2664 + 8)
    st = lib.pdeip_residual_begin(1, need - 4, L.MODEL_MLP, 8, 32, 2, 0, None)
    assert st == -3 and b"workspace" in lib.pdeip_last_error()
    st = lib.pdeip_adam_l2_step(None, None, None, None, None, 10, 0.1, 0.9, 0.999, 1e-4, 1e-3, 1, 1.0, 0, 0.999,
                                None, None)
    assert st == -1


def test_ops_refuse_cpu_tensors_and_missing_library(monkeypatch):
    from pde_inverse_problem_b200 import _lib as L
    from pde_inverse_problem_b200 import ops
    with pytest.raises(ops.PdeipError, match="CUDA"):
        ops.linear_grad(torch.zeros(3, 4), torch.zeros(4, 4))
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libpdeip.so")
    with pytest.raises(L.PdeipError, match="no CPU fallback"):
        L.load()


def test_config_composition_matches_reference_defaults():
    from pde_inverse_problem_b200.config import make_config
    cfg = make_config("kinetic_fokker_planck", **{"pde_instance.potential": "GMM", "neural_network.hidden_dim": 32,
                                                  "neural_network.layers": 2, "estimation_mode": "non-parametric"})
    assert cfg.pde_instance.name == "Kinetic-Fokker-Planck" and cfg.pde_instance.n_steps == 100
    assert cfg.solver.train.batch_size_0T == 50000 and cfg.train.optimizer.weight_decay == 0.001
    assert cfg.train.number_of_iterations == 80000 and cfg.seed == 1
    assert cfg.neural_network.hidden_dim == 32 and cfg.estimation_mode == "non-parametric"
    assert make_config().pde_instance.name == "Fokker-Planck"  # configurations/config.yaml:2


def test_rng_split_is_deterministic_and_distinct():
    from pde_inverse_problem_b200.utils import rng as R
    a = R.split(R.PRNGKey(1), 4)
    assert a == R.split(R.PRNGKey(1), 4) and len(set(a)) == 4
    assert set(a).isdisjoint(R.split(R.PRNGKey(2), 4))
    assert all(0 <= R.randint(k, 0, 5) < 5 for k in a)


def test_shard_bounds_partition():
    from pde_inverse_problem_b200.parallel import Shard
    for n, w in ((1 << 24, 8), (1000, 3), (5, 8)):
        spans = [Shard(r, w).bounds(n) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_pack_unpack_roundtrip():
    from pde_inverse_problem_b200 import parallel
    ts = [torch.arange(8.), torch.arange(12.).reshape(3, 4), torch.tensor(3.5)]
    flat, meta = parallel.pack(ts)
    assert flat.numel() == 21
    back = parallel.unpack(flat, meta)
    assert all(torch.equal(a, b) for a, b in zip(ts, back))
    assert parallel.allreduce_sum_packed(ts)[1] is ts[1]  # world == 1: no-op


def test_cosine_schedule_matches_oracle():
    from pde_inverse_problem_b200.core.optimizer import cosine_decay_schedule
    a, b = cosine_decay_schedule(1e-2, 20000, 0.001), o_optim.cosine_decay_schedule(1e-2, 20000, 0.001)
    for c in (0, 1, 777, 19999, 20000, 80000):
        assert a(c) == pytest.approx(b(c), rel=1e-15)


def test_model_layout_matches_flax_tree_order():
    """Flat parameter layout of the CUDA path == oracle.flatten_params (layers_i kernel [in,out] then bias)."""
    pytest.importorskip("torch")
    from pde_inverse_problem_b200.core import model as M

    class FakeSpec:
        pass

    net = M.V_hypothesis.__new__(M.V_hypothesis)
    net.hidden_dims, net.dim, net.hidden, net.layers = [32, 32], 4, 32, 2
    layout = net.layout()
    names = [(n, l, s) for n, l, s in layout]
    assert names == [("layers_0", "kernel", (4, 32)), ("layers_0", "bias", (32,)),
                     ("layers_1", "kernel", (32, 32)), ("layers_1", "bias", (32,)),
                     ("layers_2", "kernel", (32, 40)), ("layers_2", "bias", (40,))]
    p = o_model.init_mlp_params(4, 32, 2)
    flat = o_model.flatten_params(p).float()
    tree = net.tree(flat.clone())
    assert torch.equal(tree["params"]["layers_1"]["kernel"], p["params"]["layers_1"]["kernel"].float())
    assert torch.equal(net.flat(tree), flat)
    assert torch.equal(net.flat({"params": tree["params"]}), flat)  # tree without the cached flat buffer


def test_registry_dispatch_and_errors():
    from pde_inverse_problem_b200 import registry
    from pde_inverse_problem_b200.config import make_config
    assert registry.get_pde_instance(make_config("fokker_planck")).__name__ == "FokkerPlanck"
    cfg = make_config("kinetic_fokker_planck", **{"pde_instance.potential": "GMM"})
    assert "GMM" in registry.get_pde_instance(cfg).__module__
    assert registry.get_pde_instance(make_config("kinetic_mckean_vlasov")).__name__ == "KineticMcKeanVlasov"
    assert registry.get_method(cfg).__name__ == "ConsistencyBased"
    cfg.solver.name = "PINN"
    with pytest.raises(NotImplementedError):
        registry.get_method(cfg)


def test_forward_fn_must_be_a_pdeip_model():
    from pde_inverse_problem_b200.core.model import model_of
    with pytest.raises(NotImplementedError, match="fused CUDA"):
        model_of(lambda params, x: x)


def test_lyapunov_product_code_matches_oracle_van_loan():
    """Two independent implementations of OU.py:73-93: the product's stationary-solution formula vs the oracle's
    Van Loan block exponential."""
    import numpy as np
    from oracle import moments as o_mom
    from pde_inverse_problem_b200.utils import lyapunov
    cfg = o_mom.kinetic_ou_configuration(4)
    for t in (0.05, 0.7, 2.0):
        m1, P1 = o_mom.lyapunov_mean_cov(t, cfg)
        m2, P2 = lyapunov.kinetic_ou_mean_cov(t, cfg)
        assert np.abs(P1 - P2).max() < 1e-9 and np.abs(m1 - m2).max() < 1e-12


def test_pipeline_auto_chunk_is_whole_waves():
    """Host logic of pipeline.auto_chunk: chunks are whole waves of the integrator grid (no GPU needed:
    pdeip_sm_count() reports 148 without a device) and respect the trajectory-buffer budget."""
    from pde_inverse_problem_b200 import _lib as L
    from pde_inverse_problem_b200.pipeline import auto_chunk, integrator_wave
    assert integrator_wave(8, L.DRIFT_GMM, 16, L.PATH_TENSOR, sm_count=148) == 148 * 8 * 128
    assert integrator_wave(32, L.DRIFT_GMM, 64, L.PATH_TENSOR, sm_count=148) == 148 * 3 * 128
    assert integrator_wave(32, L.DRIFT_GMM, 100, L.PATH_TENSOR, sm_count=148) == 148 * 2 * 128  # K > 64: fp32 kernel
    assert integrator_wave(16, L.DRIFT_LINEAR, 0, L.PATH_TENSOR, sm_count=148) == 148 * 4 * 128
    c3 = auto_chunk(8, 200, L.DRIFT_GMM, 16, L.PATH_TENSOR, sm_count=148)
    assert c3 == 2 * 148 * 8 * 128  # 303 104 particles, 5.8 GB: what bench.py uses for C3
    c5 = auto_chunk(32, 200, L.DRIFT_GMM, 64, L.PATH_TENSOR, sm_count=148)
    assert c5 == 148 * 3 * 128      # 56 832 particles, 4.4 GB: C5
    for d, S, kind, K in [(4, 100, L.DRIFT_LINEAR, 0), (16, 100, L.DRIFT_LINEAR, 0), (8, 2000, L.DRIFT_GMM, 16)]:
        c = auto_chunk(d, S, kind, K, L.PATH_TENSOR, sm_count=148)
        w = integrator_wave(d, kind, K, L.PATH_TENSOR, sm_count=148)
        assert c % w == 0 and c >= w and (c == w or 3 * d * S * 4 * c <= 6.0e9)
    assert auto_chunk(8, 200, L.DRIFT_GMM, 16, L.PATH_TENSOR, sm_count=L.load().pdeip_sm_count()) % 128 == 0


def test_bench_reference_arm_prints_exactly_one_json_line():
    """The driver contract of bench.py (CPU arm, runnable without a GPU): stdout carries exactly one JSON line with the
    reference-arm keys, whatever libraries print (file descriptor 1 is pointed at stderr for the run)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "particle-steps/s" and j["unit"] == "particle-steps/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["n_gpus"] == 1 and j["steps"] == 1
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"]
    # ranks other than 0 of a torchrun launch exit 0 without work and without output
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r1 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1"],
                        capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_entry_points_registered_as_torch_custom_ops():
    """The C ABI is exposed as torch.library.custom_op's (torch.ops.pdeip.*) with mutable-output schemas and NO CPU
    kernel: a CPU tensor is refused by the dispatcher instead of falling back."""
    import pytest
    import torch
    from pde_inverse_problem_b200 import torch_ops
    for name in torch_ops.REGISTERED:
        op = getattr(torch.ops.pdeip, name)
        schema = str(op.default._schema)
        assert schema.startswith(f"pdeip::{name}(") and schema.endswith("-> ()") and "!" in schema, schema
    with pytest.raises(Exception):
        torch.ops.pdeip.linear_grad(torch.zeros(2, 2), torch.zeros(2, 2), torch.zeros(2, 2), 2, 2)
    assert torch_ops.as_i64(2 ** 64 - 1) == -1 and torch_ops._u64(-1) == 2 ** 64 - 1


def test_v_hypothesis_envelope_and_padding_layout():
    """hidden_dim <= 32 is zero-padded into the 32-wide kernels; wider / deeper / mixed networks fail with a message that
    names the supported envelope (the reference default MLP.yaml is hidden_dim 20, layers 8)."""
    import pytest
    import torch
    from pde_inverse_problem_b200.core.model import V_hypothesis
    m = V_hypothesis(1, [20] * 8, 4)
    assert m.spec.hidden == 32 and m.spec.layers == 8
    assert m.spec.num_params == 4 * 32 + 32 + 7 * (32 * 32 + 32) + 32 * 40 + 40
    tree = m.tree(torch.arange(m.spec.num_params, dtype=torch.float32))
    assert tuple(tree["params"]["layers_0"]["kernel"].shape) == (4, 20)
    assert tuple(tree["params"]["layers_3"]["kernel"].shape) == (20, 20)
    assert tuple(tree["params"]["layers_8"]["kernel"].shape) == (20, 40)
    assert tree["params"]["layers_1"]["kernel"][1, 2].item() == 4 * 32 + 32 + 1 * 32 + 2  # a view of the padded buffer
    for bad in ([64, 64], [32] * 9, [16, 32]):
        with pytest.raises(NotImplementedError, match="hidden"):
            V_hypothesis(1, bad, 4)
