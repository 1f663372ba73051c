"""GPU tests of the fp32 training kernels (csrc/train_ops.cu) against torch float64 autograd as the checker.
Tolerances: fp32 accumulation, relative 1e-4 unless stated."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


@pytest.fixture(scope="module")
def ops():
    import torch
    assert torch.cuda.is_available()
    from wind_downscaling_gan_b200.train import ops
    return ops


@pytest.mark.parametrize("N,H,W,Ci,k,Co,s,p", [(3, 13, 11, 5, 3, 7, 1, 1), (2, 38, 38, 23, 8, 40, 2, 3), (2, 34, 34, 32, 7, 64, 3, 1),
                                               (2, 12, 12, 16, 4, 24, 2, 1), (1, 9, 9, 70, 6, 66, 11, 4), (2, 16, 16, 130, 1, 33, 1, 0),
                                               # narrow 3x3 / stride-1 / same layers: direct stencil kernels (train_direct.cuh)
                                               (1, 20, 17, 2, 3, 8, 1, 1), (2, 11, 13, 2, 3, 16, 1, 1), (1, 37, 12, 16, 3, 2, 1, 1),
                                               (3, 33, 35, 2, 3, 8, 1, 1), (1, 16, 16, 5, 3, 64, 1, 1)])
def test_conv_fwd_bwd(ops, N, H, W, Ci, k, Co, s, p):
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((N, H, W, Ci), device="cuda", generator=g)
    w = torch.randn((k, k, Ci, Co), device="cuda", generator=g) * 0.1
    b = torch.randn((Co,), device="cuda", generator=g)
    Ho, Wo = ops.conv_out(H, k, s, p, p), ops.conv_out(W, k, s, p, p)
    xr = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.double().permute(3, 2, 0, 1).requires_grad_(True)
    yr = F.conv2d(xr, wr, b.double(), stride=s, padding=p)
    assert yr.shape[2:] == (Ho, Wo)
    y = ops.empty(N, Ho, Wo, Co)
    ops.conv2d_fwd(ops.full(x), w, b, ops.full(y), N, H, W, s, p, Ho, Wo)
    assert rel(y, yr.permute(0, 2, 3, 1)) < 1e-5
    dy = torch.randn((N, Ho, Wo, Co), device="cuda", generator=g)
    yr.backward(dy.double().permute(0, 3, 1, 2))
    dx = ops.empty(N, H, W, Ci)
    ops.conv2d_bwd_data(ops.full(dy), w, ops.full(dx), N, H, W, s, p, Ho, Wo)
    assert rel(dx, xr.grad.permute(0, 2, 3, 1)) < 1e-5
    dw = ops.empty(k, k, Ci, Co)
    ops.conv2d_bwd_weight(ops.full(x), ops.full(dy), dw, N, H, W, s, p, Ho, Wo)
    assert rel(dw, wr.grad.permute(2, 3, 1, 0)) < 1e-4
    # accumulate + channel views: write y into channels [2, 2+Co) of a wider buffer, read x from a wider buffer
    wide_x = torch.randn((N, H, W, Ci + 3), device="cuda", generator=g)
    wide_x[..., 1:1 + Ci] = x
    wide_y = torch.zeros((N, Ho, Wo, Co + 5), device="cuda")
    ops.conv2d_fwd(ops.View(wide_x, Ci, Ci + 3, 1), w, b, ops.View(wide_y, Co, Co + 5, 2), N, H, W, s, p, Ho, Wo)
    assert rel(wide_y[..., 2:2 + Co], yr.permute(0, 2, 3, 1)) < 1e-5 and float(wide_y[..., :2].abs().max()) == 0
    db = ops.empty(Co)
    ops.colsum(ops.full(dy), db)
    assert rel(db, dy.double().sum((0, 1, 2))) < 1e-5


def test_conv_transpose_via_bwd_data(ops):
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device="cuda").manual_seed(2)
    # Conv2DTranspose(32, 2x2, s2): kernel (kh, kw, out, in) is the HWIO kernel of the conv it transposes
    x = torch.randn((2, 6, 6, 48), device="cuda", generator=g)
    w = torch.randn((2, 2, 32, 48), device="cuda", generator=g) * 0.1
    ref = F.conv_transpose2d(x.double().permute(0, 3, 1, 2), w.double().permute(3, 2, 0, 1), stride=2).permute(0, 2, 3, 1)
    y = ops.empty(2, 12, 12, 32)
    ops.conv2d_bwd_data(ops.full(x), w, ops.full(y), 2, 12, 12, 2, 0, 6, 6)
    assert rel(y, ref) < 1e-5
    # Conv2DTranspose(16, 5x5, same, s1)
    x = torch.randn((2, 10, 10, 40), device="cuda", generator=g)
    w = torch.randn((5, 5, 16, 40), device="cuda", generator=g) * 0.1
    ref = F.conv_transpose2d(x.double().permute(0, 3, 1, 2), w.double().permute(3, 2, 0, 1), padding=2).permute(0, 2, 3, 1)
    y = ops.empty(2, 10, 10, 16)
    ops.conv2d_bwd_data(ops.full(x), w, ops.full(y), 2, 10, 10, 1, 2, 10, 10)
    assert rel(y, ref) < 1e-5


@pytest.mark.parametrize("prec,tol", [("fp32", 1e-5), ("tf32", 1e-3), ("bf16", 1e-2)])
def test_conv_transpose_s1_gemm_col2im(ops, prec, tol):
    """Stride-1 Conv2DTranspose forward as a 1x1 GEMM into tap columns + overlap-add (wdg_col2im)."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device="cuda").manual_seed(7)
    for (N, H, W, cin, cout, k, p) in [(2, 10, 13, 160, 16, 5, 2), (1, 7, 9, 64, 6, 3, 1)]:
        x = torch.randn((N, H, W, cin), device="cuda", generator=g)
        w = torch.randn((k, k, cout, cin), device="cuda", generator=g) * 0.1
        ref = F.conv_transpose2d(x.double().permute(0, 3, 1, 2), w.double().permute(3, 2, 0, 1), padding=p).permute(0, 2, 3, 1)
        wide = torch.full((N, H, W, cout + 3), 5.0, device="cuda")
        ops.set_precision(prec)
        try:
            y = ops.empty(N, H, W, cout)
            ops.conv_transpose_s1_fwd(ops.full(x), w, ops.full(y), N, H, W, p)
            ops.conv_transpose_s1_fwd(ops.full(x), w, ops.View(wide, cout, cout + 3, 2), N, H, W, p)
        finally:
            ops.set_precision("fp32")
        assert rel(y, ref) < tol
        assert rel(wide[..., 2:2 + cout], ref) < tol and float((wide[..., :2] - 5).abs().max()) == 0


def test_batchnorm_train(ops):
    import torch
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((4, 9, 9, 24), device="cuda", generator=g) * 2 + 1
    gamma = torch.rand(24, device="cuda", generator=g) + 0.5
    beta = torch.randn(24, device="cuda", generator=g)
    mm, mv = torch.zeros(24, device="cuda"), torch.ones(24, device="cuda")
    y, sm, si = ops.empty(4, 9, 9, 24), ops.empty(24), ops.empty(24)
    ops.bn_train_fwd(x, y, gamma, beta, mm, mv, sm, si)
    xr = x.double().requires_grad_(True)
    mean, var = xr.mean((0, 1, 2)), xr.var((0, 1, 2), unbiased=False)
    yr = (xr - mean) / torch.sqrt(var + 1e-3) * gamma.double() + beta.double()
    assert rel(y, yr) < 1e-5
    n = 4 * 81
    assert rel(mm, mean * 0.01) < 1e-4 and rel(mv, 0.99 + 0.01 * var * n / (n - 1)) < 1e-5
    dy = torch.randn((4, 9, 9, 24), device="cuda", generator=g)
    gr = gamma.double().requires_grad_(True)
    yr2 = (xr - xr.mean((0, 1, 2))) / torch.sqrt(xr.var((0, 1, 2), unbiased=False) + 1e-3) * gr + beta.double()
    yr2.backward(dy.double())
    dx, dg, db = ops.empty(4, 9, 9, 24), ops.empty(24), ops.empty(24)
    ops.bn_train_bwd(dy, x, gamma, sm, si, dx, dg, db)
    assert rel(dx, xr.grad) < 1e-4 and rel(dg, gr.grad) < 1e-4 and rel(db, dy.double().sum((0, 1, 2))) < 1e-5


def test_layernorm(ops):
    import torch
    g = torch.Generator(device="cuda").manual_seed(4)
    for Cc in (16, 64, 256):
        x = torch.randn((3, 5, 5, Cc), device="cuda", generator=g)
        gamma = torch.rand(Cc, device="cuda", generator=g) + 0.5
        beta = torch.randn(Cc, device="cuda", generator=g)
        y, sm, si = ops.empty(3, 5, 5, Cc), ops.empty(75), ops.empty(75)
        ops.ln_fwd(x, ops.full(y), gamma, beta, sm, si)
        xr, gr = x.double().requires_grad_(True), gamma.double().requires_grad_(True)
        yr = torch.nn.functional.layer_norm(xr, (Cc,), gr, beta.double(), eps=1e-3)
        assert rel(y, yr) < 1e-5
        dy = torch.randn((3, 5, 5, Cc), device="cuda", generator=g)
        yr.backward(dy.double())
        dx, dg, db = ops.empty(3, 5, 5, Cc), ops.empty(Cc), ops.empty(Cc)
        ops.ln_bwd(ops.full(dy), x, gamma, sm, si, dx, dg, db)
        assert rel(dx, xr.grad) < 1e-4 and rel(dg, gr.grad) < 1e-4 and rel(db, dy.double().sum((0, 1, 2))) < 1e-5


def test_lstm_gates(ops):
    import torch
    g = torch.Generator(device="cuda").manual_seed(5)
    rows, Fc = 50, 12
    z = torch.randn((rows, 4 * Fc), device="cuda", generator=g) * 3
    cp = torch.randn((rows, Fc), device="cuda", generator=g)
    zr, cr = z.double().requires_grad_(True), cp.double().requires_grad_(True)
    zi, zf, zc, zo = zr.split(Fc, 1)
    hs = lambda v: torch.clamp(0.2 * v + 0.5, 0, 1)
    cn = hs(zf) * cr + hs(zi) * torch.tanh(zc)
    hn = hs(zo) * torch.tanh(cn)
    zz, c_out, h_out = z.clone(), ops.empty(rows, Fc), ops.empty(rows, Fc)
    ops.lstm_gates_fwd(zz, cp, c_out, h_out)
    assert rel(c_out, cn) < 1e-5 and rel(h_out, hn) < 1e-5
    dh = torch.randn((rows, Fc), device="cuda", generator=g)
    dc_next = torch.randn((rows, Fc), device="cuda", generator=g)
    (hn * dh.double()).sum().backward(retain_graph=True)
    (cn * dc_next.double()).sum().backward()
    dc = dc_next.clone()
    # the recurrent part of dL/dh may be passed separately and is added inside the kernel
    dh_a = torch.randn((rows, Fc), device="cuda", generator=g)
    zz2, dc2 = zz.clone(), dc_next.clone()
    ops.lstm_gates_bwd(zz2, cp, c_out, dh - dh_a, dc2, dh_a)
    ops.lstm_gates_bwd(zz, cp, c_out, dh, dc)
    assert rel(zz, zr.grad) < 1e-4 and rel(dc, cr.grad) < 1e-4
    assert rel(zz2, zr.grad) < 1e-4 and rel(dc2, cr.grad) < 1e-4


@pytest.mark.parametrize("Fc", [1, 2, 4])
def test_lstm_small_filters(ops, Fc):
    """Fused recurrent-conv + gate kernel for cells with 1/2/4 filters and its backward-data stencil against the
    generic path (conv2d_fwd accumulate + lstm_gates_fwd, conv2d_bwd_data)."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(6)
    N, H, W = 3, 13, 17
    xz = torch.randn((N, H, W, 4 * Fc), device="cuda", generator=g)
    hp = torch.randn((N, H, W, Fc), device="cuda", generator=g)
    cp = torch.randn((N, H, W, Fc), device="cuda", generator=g)
    R = torch.randn((3, 3, Fc, 4 * Fc), device="cuda", generator=g) * 0.5
    ref_z = xz.clone()
    ops.conv2d_fwd(ops.full(hp), R, None, ops.full(ref_z), N, H, W, 1, 1, H, W, accumulate=True)
    ref_c, ref_h = ops.empty(N, H, W, Fc), ops.empty(N, H, W, Fc)
    ops.lstm_gates_fwd(ref_z, cp, ref_c, ref_h)
    z, c, h = xz.clone(), ops.empty(N, H, W, Fc), ops.empty(N, H, W, Fc)
    ops.lstm_small_fwd(z, hp, R, cp, c, h)
    assert rel(z, ref_z) < 1e-5 and rel(c, ref_c) < 1e-5 and rel(h, ref_h) < 1e-5
    # first timestep: no previous state
    z0, c0, h0 = xz.clone(), ops.empty(N, H, W, Fc), ops.empty(N, H, W, Fc)
    ops.lstm_small_fwd(z0, None, R, None, c0, h0)
    r0, rc0, rh0 = xz.clone(), ops.empty(N, H, W, Fc), ops.empty(N, H, W, Fc)
    ops.lstm_gates_fwd(r0, None, rc0, rh0)
    assert rel(z0, r0) < 1e-6 and rel(h0, rh0) < 1e-6
    dz = torch.randn((N, H, W, 4 * Fc), device="cuda", generator=g)
    ref_dh, dh = ops.empty(N, H, W, Fc), ops.empty(N, H, W, Fc)
    ops.conv2d_bwd_data(ops.full(dz), R, ops.full(ref_dh), N, H, W, 1, 1, H, W)
    ops.lstm_small_bwd_data(dz, R, dh)
    assert rel(dh, ref_dh) < 1e-5


def test_upsample_and_adjoint(ops):
    import torch
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randn((2, 7, 5, 6), device="cuda", generator=g)
    y = ops.empty(2, 14, 10, 6)
    ops.upsample2x_fwd(x, y)
    xr = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    yr = torch.nn.functional.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=False)
    assert rel(y, yr.permute(0, 2, 3, 1)) < 1e-6
    dy = torch.randn((2, 14, 10, 6), device="cuda", generator=g)
    yr.backward(dy.double().permute(0, 3, 1, 2))
    dx = ops.empty(2, 7, 5, 6)
    ops.upsample2x_bwd(dy, dx)
    assert rel(dx, xr.grad.permute(0, 2, 3, 1)) < 1e-5


def test_dense_mean_reduce_gp_lerp(ops):
    import torch
    g = torch.Generator(device="cuda").manual_seed(7)
    B, Tn, D = 3, 4, 50
    flat = torch.randn((B, Tn, D), device="cuda", generator=g)
    w = torch.randn((D, 1), device="cuda", generator=g)
    b = torch.randn((1,), device="cuda", generator=g)
    score = ops.empty(B, 1)
    ops.dense_mean_fwd(flat, w, b, score, B, Tn, D)
    fr, wr = flat.double().requires_grad_(True), w.double().requires_grad_(True)
    sr = (fr @ wr + b.double()).mean(1)
    assert rel(score, sr) < 1e-5
    ds = torch.randn((B, 1), device="cuda", generator=g)
    sr.backward(ds.double())
    dflat, dw, db = ops.empty(B, Tn, D), ops.empty(D, 1), ops.empty(1)
    ops.dense_mean_bwd(ds, flat, w, dflat, dw, db, B, Tn, D)
    assert rel(dflat, fr.grad) < 1e-5 and rel(dw, wr.grad) < 1e-5 and rel(db, ds.double().sum().reshape(1)) < 1e-5
    a = torch.randn(100000, device="cuda", generator=g)
    assert rel(ops.reduce(a, 2, scale=0.5), (a.double() ** 2).sum().reshape(1) * 0.5) < 1e-6
    gimg = torch.randn((B, Tn, 6, 6, 2), device="cuda", generator=g)
    out = ops.empty(B, 2)
    ops.gp_norm(gimg, out)
    assert rel(out, torch.sqrt((gimg.double() ** 2).sum((1, 2, 3)))) < 1e-6
    eps = torch.rand(B, device="cuda", generator=g)
    real, fake = torch.randn_like(gimg), torch.randn_like(gimg)
    comb = torch.empty_like(gimg)
    ops.lerp_batch(comb, real, fake, eps)
    e = eps.view(B, 1, 1, 1, 1)
    assert rel(comb, e * real + (1 - e) * fake) < 1e-6


def test_adam_and_spectral_norm(ops):
    import torch
    from oracle import layers as L
    g = torch.Generator(device="cuda").manual_seed(8)
    w = torch.randn(1000, device="cuda", generator=g)
    m, v = torch.zeros(1000, device="cuda"), torch.zeros(1000, device="cuda")
    w0, wd = w.clone().double(), None
    md, vd = torch.zeros(1000, dtype=torch.float64), torch.zeros(1000, dtype=torch.float64)
    wd = w0.cpu()
    for t in range(1, 4):
        gr = torch.randn(1000, device="cuda", generator=g)
        lr_t = 4e-4 * np.sqrt(1 - 0.9 ** t) / (1 - 0.5 ** t)
        ops.adam(w, m, v, gr, lr_t, 0.5, 0.9, 0.1)
        gd = gr.double().cpu()
        md += (gd - md) * 0.5
        vd += (gd * gd - vd) * (1 - 0.9)
        wd = wd - lr_t * md / (torch.sqrt(vd) + 0.1)
    assert rel(w, wd) < 1e-6
    for shp in ((3, 3, 16, 16), (8, 8, 23, 128), (2, 2, 32, 192)):
        wk = torch.randn(shp, device="cuda", generator=g) * 0.1
        u = torch.randn((1, shp[-1]), device="cuda", generator=g) * 0.02
        w_ref, u_ref = L.spectral_norm_step(wk.cpu().numpy(), u.cpu().numpy())
        ops.sn_update(wk, u)
        assert rel(wk, torch.from_numpy(w_ref)) < 1e-4 and rel(u, torch.from_numpy(u_ref)) < 1e-4
