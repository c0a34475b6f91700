"""GPU parity of the tcgen05 tensor-core residual path (PDEIP_PATH_TENSOR): rtol 1e-2 in the per-tensor max-norm
metric (BASELINE.json: "1e-2 for bf16 GEMM paths") against the float64 oracle, and against the fp32 CUDA path at
sizes the oracle cannot finish in seconds."""
import pytest
import torch

from conftest import relmax
from oracle import model as o_model, problems as o_prob, residuals as o_res

pytestmark = pytest.mark.gpu
TOL = 1e-2


def _ops():
    from pde_inverse_problem_b200 import ops, _lib
    return ops, _lib


def _params(d, seed=11):
    p = o_model.init_mlp_params(d, 32, 2, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    for k in p["params"]:
        b = p["params"][k]["bias"]
        p["params"][k]["bias"] = 0.1 * torch.randn(b.shape, generator=g, dtype=torch.float64)
    return p


def _run(ops, L, cuda, spec, flat, pts, n, gamma, tg, path, layout=None):
    acc = ops.ResidualAccumulator(spec, device=cuda).begin()
    acc.accumulate(L.SET_KFP_0T, flat, pts, 1.0 / n, coef=gamma, true_grad=tg, path=path,
                   layout=L.LAYOUT_AOS if layout is None else layout)
    s, g = acc.finalize()
    s, g = s.cpu().double(), g.cpu().double()
    assert ops.tensor_path_status() == 0, "a tcgen05 phase timed out"
    return s, g


@pytest.mark.parametrize("d,n", [(4, 128), (8, 1000), (16, 300), (2, 77), (32, 500), (8, 40000), (24, 129)])
def test_tensor_path_matches_oracle(cuda, d, n):
    ops, L = _ops()
    pde = o_prob.KineticOUProblem(d, T=2.0)
    p = _params(d)
    g = torch.Generator().manual_seed(300 + d)
    z = torch.randn(n, 2 * d, generator=g, dtype=torch.float64)
    data = {"initial": z[:1], "terminal": z[:1], "0T": z}
    # oracle 0T-only terms: subtract the boundary part by evaluating it with zero weight is not possible in the
    # reference API, so compare the pieces the kernel reports: loss (0T set only) and gradient (0T set only)
    from oracle import taylor as o_tay
    W, b = o_tay.unpack(p)
    val, dW, db, terms, gvec = o_tay.point_set(W, b, z[:, :d], [(z[:, d:], -2.0, 2.0 * 1.0)], 0.0, 1.0, 1.0 / n)
    gt = z[:, :d] @ pde.initial_configuration["tilde_F"].T
    ref_loss = val + (gt ** 2).sum(-1).mean()
    ref_grad = torch.cat([torch.cat([w_.reshape(-1), b_.reshape(-1)]) for w_, b_ in zip(dW, db)])
    ref_gt = ((gt - gvec) ** 2).sum(-1).mean()
    spec = ops.ModelSpec(L.MODEL_MLP, d, 32, 2)
    flat = o_model.flatten_params(p).float().to(cuda)
    tg = ops.TrueGrad(L.DRIFT_LINEAR, pde.initial_configuration["tilde_F"].float().to(cuda))
    s, gr = _run(ops, L, cuda, spec, flat, z.float().to(cuda), n, 1.0, tg, L.PATH_TENSOR)
    assert relmax(s[L.SUM_LOSS], ref_loss) < TOL
    assert relmax(s[L.SUM_GT], ref_gt) < TOL
    assert relmax(gr, ref_grad) < TOL
    # per-leaf
    off = 0
    for w_, b_ in zip(dW, db):
        for leaf in (w_, b_):
            assert relmax(gr[off:off + leaf.numel()], leaf.reshape(-1)) < TOL, (d, n, off)
            off += leaf.numel()


def test_tensor_path_matches_fp32_path_many_tiles(cuda):
    """More tiles than CTAs (persistent loop, TMEM-persistent bias regions), SoA layout, GMM true gradient."""
    ops, L = _ops()
    d, K, n = 8, 16, 148 * 128 * 3 + 517
    p = _params(d)
    flat = o_model.flatten_params(p).float().to(cuda)
    g = torch.Generator().manual_seed(9)
    pts = (torch.randn(n, 2 * d, generator=g) * torch.cat([torch.full((d,), 2.0), torch.full((d,), 0.6)])).to(cuda)
    mus = (torch.rand(K, d, generator=g) * 8 - 4).to(cuda)
    tg = ops.TrueGrad(L.DRIFT_GMM, mus, 1.0)
    spec = ops.ModelSpec(L.MODEL_MLP, d, 32, 2)
    s32, g32 = _run(ops, L, cuda, spec, flat, pts, n, 0.5, tg, L.PATH_FP32)
    stc, gtc = _run(ops, L, cuda, spec, flat, pts.t().contiguous(), n, 0.5, tg, L.PATH_TENSOR, layout=L.LAYOUT_SOA)
    assert relmax(stc[L.SUM_LOSS], s32[L.SUM_LOSS]) < TOL
    assert relmax(stc[L.SUM_GT], s32[L.SUM_GT]) < TOL
    assert relmax(gtc, g32) < TOL
    # run twice: bit-identical (no atomics across CTAs, fixed reduction order)
    stc2, gtc2 = _run(ops, L, cuda, spec, flat, pts.t().contiguous(), n, 0.5, tg, L.PATH_TENSOR, layout=L.LAYOUT_SOA)
    assert torch.equal(gtc, gtc2)


@pytest.mark.parametrize("path_name", ["fp32", "tensor"])
def test_true_gradient_stored_in_points(cuda, path_name):
    """PDEIP_DRIFT_IN_POINTS ([x, v, grad V_true] per point, as emitted by the integrator) gives the same sums as
    the inline true gradient."""
    ops, L = _ops()
    d, K, n = 8, 16, 9000
    path = L.PATH_FP32 if path_name == "fp32" else L.PATH_TENSOR
    p = _params(d)
    flat = o_model.flatten_params(p).float().to(cuda)
    g = torch.Generator().manual_seed(4)
    pts = (torch.randn(n, 2 * d, generator=g) * 1.5).to(cuda)
    mus = (torch.rand(K, d, generator=g) * 8 - 4).to(cuda)
    _, gt = ops.gmm_value_grad(pts[:, :d].contiguous(), mus, 1.0)
    spec = ops.ModelSpec(L.MODEL_MLP, d, 32, 2)
    s_a, g_a = _run(ops, L, cuda, spec, flat, pts, n, 0.5, ops.TrueGrad(L.DRIFT_GMM, mus, 1.0), path)
    pts3 = torch.cat([pts, gt], 1).contiguous()
    s_b, g_b = _run(ops, L, cuda, spec, flat, pts3, n, 0.5, ops.TrueGrad(L.DRIFT_IN_POINTS), path)
    s_c, g_c = _run(ops, L, cuda, spec, flat, pts3.t().contiguous(), n, 0.5, ops.TrueGrad(L.DRIFT_IN_POINTS), path,
                    layout=L.LAYOUT_SOA)
    for s_x, g_x in ((s_b, g_b), (s_c, g_c)):
        assert relmax(s_x[L.SUM_LOSS], s_a[L.SUM_LOSS]) < 1e-5
        assert relmax(s_x[L.SUM_GT], s_a[L.SUM_GT]) < 1e-5
        assert relmax(s_x[L.SUM_GTRUE2], s_a[L.SUM_GTRUE2]) < 1e-5
        assert relmax(g_x, g_a) < 1e-5


@pytest.mark.parametrize("n", [0, 1, 127, 128, 129, 255, 257])
def test_tensor_path_edge_sizes(cuda, n):
    """Empty, single-point, and tile-boundary point counts (one or two 128-point slots partially filled): the tensor path
    agrees with the fp32 path; n = 0 is a no-op that leaves the accumulator at zero."""
    ops, L = _ops()
    d = 8
    p = _params(d)
    flat = o_model.flatten_params(p).float().to(cuda)
    g = torch.Generator().manual_seed(70 + n)
    pts = torch.randn(n, 2 * d, generator=g).to(cuda)
    spec = ops.ModelSpec(L.MODEL_MLP, d, 32, 2)
    tg = ops.TrueGrad(L.DRIFT_NONE)
    s32, g32 = _run(ops, L, cuda, spec, flat, pts, max(n, 1), 0.5, tg, L.PATH_FP32)
    stc, gtc = _run(ops, L, cuda, spec, flat, pts, max(n, 1), 0.5, tg, L.PATH_TENSOR)
    if n == 0:
        assert float(stc.abs().sum()) == 0.0 and float(gtc.abs().sum()) == 0.0
        return
    assert relmax(stc[L.SUM_LOSS], s32[L.SUM_LOSS]) < TOL
    assert relmax(gtc, g32) < TOL


@pytest.mark.parametrize("path_name,d", [("fp32", 8), ("tensor", 8), ("tensor", 32), ("fp32", 16)])
def test_block128_point_layout(cuda, path_name, d):
    """PDEIP_LAYOUT_BLOCK128 ([n/128][dim][128], what the integrator emits with PDEIP_TRAJ_BLOCK128) gives bit-identical
    sums and gradient to the SoA layout of the same points; a ragged n is refused."""
    ops, L = _ops()
    n = 128 * 75
    path = L.PATH_FP32 if path_name == "fp32" else L.PATH_TENSOR
    flat = o_model.flatten_params(_params(d)).float().to(cuda)
    g = torch.Generator().manual_seed(40 + d)
    pts3 = (torch.randn(n, 3 * d, generator=g) * 1.3).to(cuda)  # [x, v, grad V_true]
    spec = ops.ModelSpec(L.MODEL_MLP, d, 32, 2)
    tg = ops.TrueGrad(L.DRIFT_IN_POINTS)
    s_s, g_s = _run(ops, L, cuda, spec, flat, pts3.t().contiguous(), n, 0.5, tg, path, layout=L.LAYOUT_SOA)
    blk = pts3.view(n // 128, 128, 3 * d).permute(0, 2, 1).contiguous()
    s_b, g_b = _run(ops, L, cuda, spec, flat, blk, n, 0.5, tg, path, layout=L.LAYOUT_BLOCK128)
    assert torch.equal(s_b, s_s) and torch.equal(g_b, g_s)
    acc = ops.ResidualAccumulator(spec, device=cuda).begin()
    with pytest.raises(Exception):
        acc.accumulate(L.SET_KFP_0T, flat, blk, 1.0 / n, coef=0.5, true_grad=tg, path=path, layout=L.LAYOUT_BLOCK128,
                       n_points=n - 5)
