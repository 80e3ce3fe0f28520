"""tcgen05 3xTF32 dense-layer kernel (gymrl_b200/csrc/linear_tc.cu) vs a plain PyTorch fp32 reference and vs the
FFMA kernel.  The parity bar is the fp32 one: the 3-way split must not cost accuracy relative to an fp32 GEMM."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture()
def modes():
    from gymrl_b200 import _ffi
    lib = _ffi.load()
    prev = lib.gymrl_get_gemm_mode()
    yield lib
    lib.gymrl_set_gemm_mode(prev)


def _ref64(x, w, b=None):
    y = x.double() @ w.double().T
    return y + b.double() if b is not None else y


@pytest.mark.parametrize("M,N,K,act", [(16384, 512, 256, 1), (16384, 256, 256, 1), (4096, 256, 256, 2), (4096, 512, 256, 0),
                                       (128, 64, 32, 0), (1000, 192, 96, 1), (4096, 64, 64, 2)])
def test_tc_forward_accuracy(modes, M, N, K, act):
    from gymrl_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    ref = _ref64(x, w, b)
    ref = torch.tanh(ref) if act == 1 else (torch.relu(ref) if act == 2 else ref)
    errs = {}
    for mode in (0, 1):
        modes.gymrl_set_gemm_mode(mode)
        y = ops.linear_forward(x.cuda(), w.cuda(), b.cuda(), act)
        errs[mode] = (y.cpu().double() - ref).abs().max().item()
        torch.testing.assert_close(y.cpu(), ref.float(), rtol=2e-5, atol=2e-5)
    # the tensor-core result is as accurate as the FFMA fp32 result (within a small factor), i.e. fp32-grade
    assert errs[1] <= 4 * errs[0] + 1e-6, errs


def test_tc_forward_gather_and_views(modes):
    from gymrl_b200 import ops
    modes.gymrl_set_gemm_mode(1)
    g = torch.Generator().manual_seed(1)
    X = torch.randn(20000, 256, generator=g)
    idx = torch.randperm(20000, generator=g)[:4096].to(torch.int32)
    w, b = torch.randn(512, 256, generator=g) / 16, torch.randn(512, generator=g)
    out = torch.zeros(4096, 1024, device="cuda")
    ops.linear_forward(X.cuda(), w.cuda(), b.cuda(), 1, row_index=idx.cuda(), out=out[:, 512:])   # strided output view
    torch.testing.assert_close(out[:, 512:].cpu(), torch.tanh(_ref64(X[idx.long()], w, b)).float(), rtol=2e-5, atol=2e-5)
    assert (out[:, :512] == 0).all()


@pytest.mark.parametrize("M,N,K,act", [(16384, 512, 256, 1), (16384, 256, 256, 2), (4096, 256, 256, 1), (4096, 64, 64, 0)])
def test_tc_backward_accuracy(modes, M, N, K, act):
    """dX = (dY W) * act'(h) [A K-major, B MN-major]  and  dW = dY^T X [both MN-major, split-K]."""
    from gymrl_b200 import ops
    g = torch.Generator().manual_seed(M * 3 + N + K)
    h = torch.tanh(torch.randn(M, K, generator=g)) if act == 1 else torch.relu(torch.randn(M, K, generator=g))
    w = torch.randn(N, K, generator=g) / K ** 0.5
    dy = torch.randn(M, N, generator=g) / M
    dx_ref = dy.double() @ w.double()
    if act == 1:
        dx_ref = dx_ref * (1 - h.double() ** 2)
    elif act == 2:
        dx_ref = dx_ref * (h > 0)
    dw_ref, db_ref = dy.double().T @ h.double(), dy.double().sum(0)
    for mode in (0, 1):
        modes.gymrl_set_gemm_mode(mode)
        dx = ops.linear_backward_input(dy.cuda(), w.cuda(), h.cuda() if act else None, act)
        torch.testing.assert_close(dx.cpu(), dx_ref.float(), rtol=1e-4, atol=1e-6 / M ** 0.5 + 1e-8)
        dw = torch.zeros(N, K, device="cuda"); db = torch.zeros(N, device="cuda")
        ops.linear_backward_weight(dy.cuda(), h.cuda(), dw, db)
        torch.testing.assert_close(dw.cpu(), dw_ref.float(), rtol=1e-4, atol=2e-6)
        torch.testing.assert_close(db.cpu(), db_ref.float(), rtol=1e-4, atol=2e-6)


def test_tc_and_ffma_trainers_agree(modes, golden):
    """The PPO reference-parity fixture passes on both engines (the tolerance is the reference's, not loosened)."""
    from conftest import trainer_from_golden as _trainer_from_golden
    g = golden("ppo_update.npz")
    for mode in (0, 1):
        modes.gymrl_set_gemm_mode(mode)
        t = _trainer_from_golden(g, use_graph=False, n_mb=1)
        t.cfg.max_grad_norm = 1e9
        t.optimizer.param_groups[0]["lr"] = 0.0
        t.update(float(g["next_value"]))
        for k, p in t.model.named_parameters():
            ref = g["g_" + k]
            np.testing.assert_allclose(p.grad.cpu().numpy(), ref, rtol=2e-3, atol=2e-6 + 1e-4 * np.abs(ref).max(), err_msg=f"mode {mode} {k}")
