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


# ---------------------------------------------------------------- warp-specialised TMA / 2-CTA kernel (gemm3x_ws_kernel)
@pytest.fixture(autouse=True, scope="module")
def _ws_everywhere():
    """The dispatcher only sends multi-tile-per-pair shapes to the ws kernel; the tests exercise it at every shape it supports."""
    import os
    os.environ["GYMRL_TC_WS_MIN_TILES"] = "1"
    os.environ["GYMRL_TC_WS_BN128_MIN_TILES"] = "1"     # widths that are multiples of 128 but not of 256: the BN = 128 pair tiles
    yield
    os.environ.pop("GYMRL_TC_WS_MIN_TILES", None)
    os.environ.pop("GYMRL_TC_WS_BN128_MIN_TILES", None)


def _ws_launches():
    import ctypes
    from gymrl_b200 import _ffi
    f = _ffi.load().gymrl_debug_ws_launches
    f.restype = ctypes.c_ulonglong
    return int(f())


def _flat_with_matrix(w, pad_front=8):
    """A flat parameter buffer holding W at a 16-byte aligned offset, registered for pre-split weight images."""
    from gymrl_b200 import ops
    N, K = w.shape
    flat = torch.zeros(pad_front + N * K + 4, device="cuda")
    flat[pad_front:pad_front + N * K] = w.reshape(-1).cuda()
    images = ops.weight_images_register(flat, [(pad_front, N, K)])
    return flat, flat[pad_front:pad_front + N * K].view(N, K), images


@pytest.mark.parametrize("M,N,K,act", [(16384, 512, 256, 1), (16384, 256, 256, 1), (16384 + 128, 256, 256, 0), (20000, 512, 512, 2),
                                       (16384, 256, 64, 0), (131072, 128, 128, 0), (4096 + 128, 384, 64, 1), (1000, 128, 256, 2)])
def test_ws_forward_matches_register_split_kernel_bitwise(modes, M, N, K, act):
    """Same 12 MMAs per 32-k slab on the same hi / lo operand bits, so the TMA-fed persistent kernel must reproduce the
    one-tile-per-CTA kernel BIT FOR BIT — and both are fp32-grade against float64."""
    from gymrl_b200 import ops
    modes.gymrl_set_gemm_mode(1)
    g = torch.Generator().manual_seed(M + 7 * N + K)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    xd, bd = x.cuda(), b.cuda()
    y_old = ops.linear_forward(xd, w.cuda(), bd, act)                    # unregistered weights: register-split kernel
    flat, wv, images = _flat_with_matrix(w)
    n0 = _ws_launches()
    try:
        y_new = ops.linear_forward(xd, wv, bd, act)                      # registered: TMA-fed warp-specialised kernel
        torch.cuda.synchronize()
    finally:
        ops.weight_images_unregister(flat)
    assert _ws_launches() == n0 + 1, "the warp-specialised kernel did not take this GEMM"
    ref = _ref64(x, w, b)
    ref = torch.tanh(ref) if act == 1 else (torch.relu(ref) if act == 2 else ref)
    torch.testing.assert_close(y_new.cpu(), ref.float(), rtol=2e-5, atol=2e-5)
    assert torch.equal(y_new, y_old)


@pytest.mark.parametrize("M,N,K,act", [(16384, 512, 256, 1), (16384, 256, 256, 2), (16384, 256, 512, 1), (12800, 256, 256, 0),
                                       (131072, 128, 128, 0), (8192, 256, 384, 1)])
def test_ws_backward_input_matches_register_split_kernel_bitwise(modes, M, N, K, act):
    """dX = (dY W) * act'(h): the ws kernel reads W^T from the transposed weight images as a K-major operand; the old kernel
    stages W MN-major.  The per-slab products are the same numbers, the results must agree to fp32 rounding of the
    accumulation (bitwise is not guaranteed across operand majors) and both are fp32-grade against float64."""
    from gymrl_b200 import ops
    modes.gymrl_set_gemm_mode(1)
    g = torch.Generator().manual_seed(M * 3 + N + K)
    h = torch.tanh(torch.randn(M, K, generator=g)) if act == 1 else torch.relu(torch.randn(M, K, generator=g))
    w = torch.randn(N, K, generator=g) / K ** 0.5
    dy = torch.randn(M, N, generator=g) / M
    dx_ref = dy.double() @ w.double()
    if act == 1:
        dx_ref = dx_ref * (1 - h.double() ** 2)
    elif act == 2:
        dx_ref = dx_ref * (h > 0)
    dyd, hd = dy.cuda(), h.cuda()
    dx_old = ops.linear_backward_input(dyd, w.cuda(), hd if act else None, act)
    flat, wv, images = _flat_with_matrix(w)
    n0 = _ws_launches()
    try:
        dx_new = ops.linear_backward_input(dyd, wv, hd if act else None, act)
        torch.cuda.synchronize()
    finally:
        ops.weight_images_unregister(flat)
    assert _ws_launches() == n0 + 1, "the warp-specialised kernel did not take this GEMM"
    torch.testing.assert_close(dx_new.cpu(), dx_ref.float(), rtol=1e-4, atol=1e-6 / M ** 0.5 + 1e-8)
    torch.testing.assert_close(dx_new, dx_old, rtol=1e-5, atol=1e-9)


def test_ws_images_follow_the_optimizer(modes):
    """gymrl_clip_adam_step / gymrl_adam_step / gymrl_polyak re-split the images of a registered buffer: a forward after the
    step uses the NEW weights."""
    from gymrl_b200 import ops
    from gymrl_b200.nn import FlatParams, FusedAdam
    modes.gymrl_set_gemm_mode(1)
    torch.manual_seed(0)
    lin = torch.nn.Linear(256, 256).cuda()
    fp = FlatParams(lin, device=torch.device("cuda"))
    fp.enable_weight_images([("weight", "weight", 256, 256)])
    opt = FusedAdam(fp, lr=0.05)
    x = torch.randn(16384, 256, device="cuda")
    y0 = ops.linear_forward(x, fp.p("weight"), fp.p("bias"), 0).clone()
    fp.grad.normal_()
    opt.step(max_norm=0.5)
    y1 = ops.linear_forward(x, fp.p("weight"), fp.p("bias"), 0)
    ref = (x.double() @ fp.p("weight").double().T + fp.p("bias").double()).float()
    torch.testing.assert_close(y1, ref, rtol=2e-5, atol=2e-5)
    assert (y1 - y0).abs().max() > 1e-2
    # a host-side write needs an explicit refresh
    fp.flat.mul_(0.5)
    fp.refresh_weight_images()
    y2 = ops.linear_forward(x, fp.p("weight"), fp.p("bias"), 0)
    torch.testing.assert_close(y2, 0.5 * ref, rtol=2e-5, atol=2e-5)


def test_ws_single_cta_mode_subprocess():
    """GYMRL_TC_WS=1 (single-CTA tiles, two 96 KB stages) computes the same bits as the CTA-pair mode."""
    import subprocess
    import sys
    code = (
        "import torch, sys; sys.path.insert(0, '.')\n"
        "from gymrl_b200 import ops\n"
        "g = torch.Generator().manual_seed(3)\n"
        "x, w, b = torch.randn(16384, 256, generator=g).cuda(), (torch.randn(512, 256, generator=g) / 16).cuda(), torch.randn(512, generator=g).cuda()\n"
        "y_old = ops.linear_forward(x, w, b, 1)\n"
        "flat = torch.zeros(512 * 256 + 8, device='cuda'); flat[4:4 + 512 * 256] = w.reshape(-1)\n"
        "img = ops.weight_images_register(flat, [(4, 512, 256)])\n"
        "y = ops.linear_forward(x, flat[4:4 + 512 * 256].view(512, 256), b, 1)\n"
        "torch.cuda.synchronize(); print('EQUAL', bool(torch.equal(y, y_old)))\n")
    from pathlib import Path
    import os
    env = dict(os.environ, GYMRL_TC_WS="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=str(Path(__file__).resolve().parent.parent), env=env)
    assert "EQUAL True" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
