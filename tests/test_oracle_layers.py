"""CPU tests pinning the oracle (oracle/) with analytic known answers, brute-force
definitions on tiny shapes, and agreement between its two independent restatements."""
import numpy as np
import pytest

from oracle import layers as L
from oracle.generator import generator_forward, generator_weight_shapes, synthetic_generator_weights


def rl2(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b))


def test_conv2d_matches_bruteforce():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 9, 10, 3))
    w = rng.standard_normal((4, 4, 3, 5))
    b = rng.standard_normal(5)
    out = L.conv2d(x, w, b, stride=2)
    assert out.shape == (2, 3, 4, 5)   # floor((n-k)/s)+1
    ref = np.zeros_like(out)
    for n in range(2):
        for oy in range(3):
            for ox in range(4):
                patch = x[n, 2 * oy:2 * oy + 4, 2 * ox:2 * ox + 4, :]
                ref[n, oy, ox] = np.tensordot(patch, w, axes=([0, 1, 2], [0, 1, 2])) + b
    assert np.allclose(out, ref, atol=1e-12)


def test_conv_transpose_s2_is_adjoint_of_strided_conv():
    # <convT(x), y> == <x, conv_s2(y)> with the kernel read as (kh, kw, out, in)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((1, 5, 6, 4))
    w = rng.standard_normal((2, 2, 3, 4))
    y = rng.standard_normal((1, 10, 12, 3))
    lhs = np.sum(L.conv2d_transpose_s2k2(x, w) * y)
    rhs = np.sum(x * L.conv2d(y, w, stride=2))  # conv kernel HWIO = (2,2,3 in,4 out)
    assert abs(lhs - rhs) < 1e-9 * max(1, abs(lhs))


def test_conv_transpose_same_is_adjoint_of_same_conv():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((1, 7, 8, 4))
    w = rng.standard_normal((5, 5, 3, 4))   # (kh, kw, out, in)
    y = rng.standard_normal((1, 7, 8, 3))
    lhs = np.sum(L.conv2d_transpose_same_s1(x, w) * y)
    rhs = np.sum(x * L.conv2d(y, w, padding="same"))
    assert abs(lhs - rhs) < 1e-9 * max(1, abs(lhs))


def test_bilinear_constant_and_taps():
    c = np.full((1, 4, 5, 2), 3.25)
    assert np.allclose(L.upsample_bilinear_x2(c), 3.25)
    x = np.arange(4.0).reshape(1, 4, 1, 1)
    up = L.upsample_bilinear_x2(x)[0, :, 0, 0]
    assert np.allclose(up, [0, 0.25, 0.75, 1.25, 1.75, 2.25, 2.75, 3.0])  # half-pixel centres, edge clamp


def test_bilinear_matches_torch_interpolate():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 6, 7, 3))
    ref = torch.nn.functional.interpolate(torch.from_numpy(x).permute(0, 3, 1, 2), scale_factor=2, mode="bilinear",
                                          align_corners=False).permute(0, 2, 3, 1).numpy()
    assert np.allclose(L.upsample_bilinear_x2(x), ref, atol=1e-12)


def test_batchnorm_identity_stats():
    x = np.random.default_rng(4).standard_normal((3, 4, 4, 6))
    y = L.batchnorm_infer(x, np.ones(6), np.zeros(6), np.zeros(6), np.ones(6))
    assert np.allclose(y, x / np.sqrt(1.001))


def test_layernorm_axis_and_eps():
    x = np.random.default_rng(5).standard_normal((2, 3, 3, 16))
    y = L.layernorm(x, np.ones(16), np.zeros(16))
    assert np.allclose(y.mean(-1), 0, atol=1e-12)
    assert np.allclose((y ** 2).mean(-1), x.var(-1) / (x.var(-1) + 1e-3))


def test_convlstm_zero_recurrent_kernel_is_gated_pointwise():
    rng = np.random.default_rng(6)
    B, T, H, W, C, F = 1, 3, 5, 5, 2, 4
    x = rng.standard_normal((B, T, H, W, C))
    k = rng.standard_normal((3, 3, C, 4 * F)) * 0.3
    b = rng.standard_normal(4 * F) * 0.1
    out = L.conv_lstm2d(x, k, np.zeros((3, 3, F, 4 * F)), b)
    c = np.zeros((B, H, W, F))
    for t in range(T):
        z = L.conv2d(x[:, t], k, b, padding="same")
        i, f, g, o = (z[..., j * F:(j + 1) * F] for j in range(4))
        c = L.hard_sigmoid(f) * c + L.hard_sigmoid(i) * np.tanh(g)
        assert np.allclose(out[:, t], L.hard_sigmoid(o) * np.tanh(c), atol=1e-12)


def test_hard_sigmoid_and_leaky():
    assert np.allclose(L.hard_sigmoid(np.array([-3.0, -2.5, 0.0, 1.0, 2.5, 9.0])), [0, 0, 0.5, 0.7, 1, 1])
    assert np.allclose(L.leaky_relu(np.array([-1.0, 2.0])), [-np.float64(np.float32(0.2)), 2.0])


def test_spectral_norm_step_normalises_top_singular_value():
    rng = np.random.default_rng(7)
    w = rng.standard_normal((3, 3, 4, 6))
    u = rng.standard_normal((1, 6))
    for _ in range(60):
        w_n, u = L.spectral_norm_step(w, u)
    s = np.linalg.svd(w.reshape(-1, 6), compute_uv=False)[0]
    assert np.allclose(np.linalg.svd(w_n.reshape(-1, 6), compute_uv=False)[0], 1.0, atol=1e-6)
    assert np.allclose(w_n * s, w, atol=1e-5)


def test_generator_two_restatements_agree():
    torch = pytest.importorskip("torch")
    from oracle.torch_port import TorchGenerator
    w = synthetic_generator_weights(5)
    rng = np.random.default_rng(8)
    image = rng.standard_normal((1, 3, 32, 32, 3))
    noise = 0.1 * rng.standard_normal((1, 3, 32, 32, 20))
    a = generator_forward(w, image, noise)
    b = TorchGenerator(w, torch.float64).forward(image, noise).numpy()
    assert a.shape == (1, 3, 32, 32, 2)
    assert rl2(b, a) < 1e-7   # differs only by the fp32-stored LeakyReLU slope
    c = TorchGenerator(w, torch.float32, emulate_bf16=True).forward(image, noise).numpy()
    assert rl2(c, a) < 1e-2   # the bf16-operand budget the GPU tests use


def test_generator_golden_fixture_reproduces():
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "generator_golden.npz"))
    B, T, S, ws, xs = (int(v) for v in z["b1_t2_s32_meta"])
    rng = np.random.default_rng(xs)
    image = rng.standard_normal((B, T, S, S, 3)).astype(np.float32)
    noise = (0.1 * rng.standard_normal((B, T, S, S, 20))).astype(np.float32)
    y = generator_forward(synthetic_generator_weights(ws), image, noise)
    assert np.allclose(y, z["b1_t2_s32"], rtol=1e-5, atol=1e-6)


def test_weight_shapes_match_reference_checkpoint_index():
    """Shapes/names of make_generator (models.py:28-71) against the reference's own generator.index."""
    import json
    import os
    man = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ckpt_manifest.json")))["generator"]
    ours = generator_weight_shapes()
    assert set(ours) == set(man)
    for k, shp in ours.items():
        assert list(shp) == man[k]["shape"], k
        assert man[k]["dtype"] == 1
    assert sum(int(np.prod(s)) for s in ours.values()) - 128 - 128 - 64 - 192 == 1_795_154  # params excl. sn_u


def test_train_golden_fixture_is_what_the_oracle_computes():
    """tests/golden/train_step_s96_t24.npz (one WGAN step at 96 px x 24 timesteps) is reproducible from its script:
    guards both the committed numbers and the oracle against drift.  ~20 s of CPU."""
    import os
    import torch
    from oracle import torch_train as tt
    from tests.golden.make_train_golden import case, projections
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "train_step_s96_t24.npz"))
    B, T, S = (int(v) for v in z["meta"][:3])
    torch.set_num_threads(os.cpu_count() or 1)
    lr, hr, draws, gw, dw = case(B, T, S)
    st = tt.State(gw, dw)
    m = tt.train_step(st, lr, hr, draws)
    for k, v in m.items():
        assert abs(v - float(z["metric/" + k])) <= 1e-9 * max(1.0, abs(v)), k
    for prefix, old, new in (("g", gw, st.g), ("d", dw, st.d)):
        for k in (("layer_with_weights-0/layer/w", "layer_with_weights-4/cell/recurrent_kernel") if prefix == "g" else
                  ("layer_with_weights-2/layer/w", "layer_with_weights-1/cell/recurrent_kernel")):
            d = new[k].numpy() - np.asarray(old[k], np.float64)
            assert np.allclose(projections(f"{prefix}/{k}", d), z[f"{prefix}/{k}/proj"], rtol=1e-7, atol=1e-12), (prefix, k)


def test_operand_rounding_cascade():
    """Why a tensor-core run cannot be compared decision for decision with an operand-rounding emulation: with the
    operands of every convolution rounded to tf32 in BOTH runs, accumulating in float32 instead of float64 moves the
    first layer by ~5e-7, but every later layer re-rounds its input, a tiny difference flips the rounding of a
    fraction (difference / tf32 ulp) of the elements by a whole ulp, and after three layers the two runs carry
    independent rounding noise: outputs ~3e-4 apart -- as far as either is from the exact result within a factor ~2."""
    import torch
    from oracle import torch_train as tt
    from oracle.generator import synthetic_generator_weights
    B, T, S = 2, 2, 32
    rng = np.random.default_rng(2)
    lr = rng.standard_normal((B, T, S, S, 3)).astype(np.float32)
    noise = (0.1 * rng.standard_normal((B, T, S, S, 20))).astype(np.float32)
    gw = synthetic_generator_weights(3)

    def run(dt, operand):
        tt.DT, tt.OPERAND = dt, operand
        try:
            out, _ = tt.generator({k: tt.T(v).clone() for k, v in gw.items()}, tt.T(lr), tt.T(noise), training=False)
        finally:
            tt.DT, tt.OPERAND = torch.float64, None
        return out.double().numpy()

    def rel(a, b):
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))

    exact, emul64, emul32, plain32 = run(torch.float64, None), run(torch.float64, "tf32"), run(torch.float32, "tf32"), run(torch.float32, None)
    assert rel(plain32, exact) < 1e-5                     # fp32 accumulation alone: harmless
    assert 1e-4 < rel(emul64, exact) < 2e-3               # operand rounding: the tf32 error level
    assert 5e-5 < rel(emul32, emul64) < 2e-3              # same rounding rule, different accumulation: already decorrelated
    assert rel(emul32, emul64) > 20 * rel(plain32, exact)
