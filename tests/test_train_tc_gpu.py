"""GPU tests of the tcgen05 training convolutions (csrc/train_gemm_tc.cu) against torch float64 autograd.
Tolerances (relative L2, stated per operand type as north_star asks): tf32 operands 1e-3, bf16 operands 1e-2;
accumulation is fp32 in TMEM either way.  The fp32 CUDA-core path of the same entry points is covered by
test_train_ops_gpu.py."""
import pytest

pytestmark = pytest.mark.gpu

TOL = {"tf32": 1e-3, "bf16": 1e-2}

SHAPES = [
    # N, H, W, Ci, k, Co, s, p
    (3, 13, 11, 5, 3, 7, 1, 1),        # ragged everything, scalar gather path
    (2, 38, 38, 23, 8, 40, 2, 3),      # generator first conv class (Ci = 23)
    (2, 34, 34, 32, 7, 64, 3, 1),      # critic pyramid conv, stride 3 (9 residue classes backward)
    (2, 12, 12, 16, 4, 24, 2, 1),
    (1, 9, 9, 70, 6, 66, 11, 4),       # stride > kernel: empty residue classes
    (2, 16, 16, 130, 1, 33, 1, 0),     # 1x1
    (2, 24, 24, 128, 3, 512, 1, 1),    # ConvLSTM gates: vector path, N tile 256 x 2
    (3, 24, 24, 128, 3, 64, 1, 1),     # N tile 64
    (2, 20, 20, 192, 3, 136, 1, 1),    # N = 136 -> two 128-wide tiles, second ragged
    (1, 40, 40, 2, 3, 8, 1, 1),        # critic hr ConvLSTM: N tile 16, K = 18
    (2, 30, 30, 21, 3, 64, 1, 1),      # critic mix ConvLSTM
    (1, 96, 96, 16, 3, 2, 1, 1),       # generator output conv
    (2, 48, 48, 36, 3, 20, 1, 1),      # Ci multiple of 4 but not 8 (mixed vector / scalar chunks with bf16)
]


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


@pytest.fixture(scope="module")
def ops():
    import torch
    assert torch.cuda.is_available()
    from wind_downscaling_gan_b200.train import ops
    yield ops
    ops.set_precision("fp32")


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
@pytest.mark.parametrize("N,H,W,Ci,k,Co,s,p", SHAPES)
def test_conv_fwd_bwd_tc(ops, prec, N, H, W, Ci, k, Co, s, p):
    import torch
    import torch.nn.functional as F
    tol = TOL[prec]
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((N, H, W, Ci), device="cuda", generator=g)
    w = torch.randn((k, k, Ci, Co), device="cuda", generator=g) * 0.1
    b = torch.randn((Co,), device="cuda", generator=g)
    Ho, Wo = ops.conv_out(H, k, s, p, p), ops.conv_out(W, k, s, p, p)
    xr = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.double().permute(3, 2, 0, 1).requires_grad_(True)
    yr = F.conv2d(xr, wr, b.double(), stride=s, padding=p)
    dy = torch.randn((N, Ho, Wo, Co), device="cuda", generator=g)
    yr.backward(dy.double().permute(0, 3, 1, 2))
    ops.set_precision(prec)
    try:
        y = ops.empty(N, Ho, Wo, Co)
        ops.conv2d_fwd(ops.full(x), w, b, ops.full(y), N, H, W, s, p, Ho, Wo)
        dx = ops.empty(N, H, W, Ci)
        ops.conv2d_bwd_data(ops.full(dy), w, ops.full(dx), N, H, W, s, p, Ho, Wo)
        dw = ops.empty(k, k, Ci, Co)
        ops.conv2d_bwd_weight(ops.full(x), ops.full(dy), dw, N, H, W, s, p, Ho, Wo)
        # channel views + accumulate
        wide_x = torch.randn((N, H, W, Ci + 3), device="cuda", generator=g)
        wide_x[..., 1:1 + Ci] = x
        wide_y = torch.ones((N, Ho, Wo, Co + 5), device="cuda")
        ops.conv2d_fwd(ops.View(wide_x, Ci, Ci + 3, 1), w, b, ops.View(wide_y, Co, Co + 5, 2), N, H, W, s, p, Ho, Wo,
                       accumulate=True)
        dx2 = torch.full((N, H, W, Ci + 4), 2.0, device="cuda")
        ops.conv2d_bwd_data(ops.View(wide_y, Co, Co + 5, 2), w, ops.View(dx2, Ci, Ci + 4, 4), N, H, W, s, p, Ho, Wo,
                            accumulate=True)
        torch.cuda.synchronize()
    finally:
        ops.set_precision("fp32")
    e = {"y": rel(y, yr.permute(0, 2, 3, 1)), "dx": rel(dx, xr.grad.permute(0, 2, 3, 1)),
         "dw": rel(dw, wr.grad.permute(2, 3, 1, 0)),
         "y_view": rel(wide_y[..., 2:2 + Co] - 1.0, yr.permute(0, 2, 3, 1))}
    print(prec, (N, H, W, Ci, k, Co, s, p), {k_: f"{v:.2e}" for k_, v in e.items()})
    assert all(v < tol for v in e.values()), e
    assert float((wide_y[..., :2] - 1.0).abs().max()) == 0 and float((wide_y[..., 2 + Co:] - 1.0).abs().max()) == 0
    assert float((dx2[..., :4] - 2.0).abs().max()) == 0
    # the accumulated view result: dx2 = 2 + bwd_data(wide_y view), linear in its input
    yv = wide_y[..., 2:2 + Co].double().permute(0, 3, 1, 2)
    ref2 = torch.autograd.grad(F.conv2d(xr, wr.detach(), None, stride=s, padding=p), xr, yv)[0].permute(0, 2, 3, 1)
    assert rel(dx2[..., 4:] - 2.0, ref2) < tol


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
def test_conv_transpose_via_bwd_data_tc(ops, prec):
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device="cuda").manual_seed(2)
    ops.set_precision(prec)
    try:
        x = torch.randn((2, 6, 6, 192), device="cuda", generator=g)
        w = torch.randn((2, 2, 32, 192), device="cuda", generator=g) * 0.1
        ref = F.conv_transpose2d(x.double().permute(0, 3, 1, 2), w.double().permute(3, 2, 0, 1), stride=2).permute(0, 2, 3, 1)
        y = ops.empty(2, 12, 12, 32)
        ops.conv2d_bwd_data(ops.full(x), w, ops.full(y), 2, 12, 12, 2, 0, 6, 6)
        assert rel(y, ref) < TOL[prec]
        x = torch.randn((2, 10, 10, 160), device="cuda", generator=g)
        w = torch.randn((5, 5, 16, 160), device="cuda", generator=g) * 0.1
        ref = F.conv_transpose2d(x.double().permute(0, 3, 1, 2), w.double().permute(3, 2, 0, 1), padding=2).permute(0, 2, 3, 1)
        y = ops.empty(2, 10, 10, 16)
        ops.conv2d_bwd_data(ops.full(x), w, ops.full(y), 2, 10, 10, 1, 2, 10, 10)
        assert rel(y, ref) < TOL[prec]
    finally:
        ops.set_precision("fp32")
