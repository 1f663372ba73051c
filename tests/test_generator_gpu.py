"""GPU parity tests of the sm_100a generator forward (through the C ABI) against the float64 oracle.

Tolerances are BASELINE.json's north_star bounds, stated per precision:
  bf16 (bf16 activations/weights, kind::f16 MMAs, fp32 accumulation in TMEM, fp32 cell state / BatchNorm math): rel-L2 <= 1e-2
  tf32 (fp32 activations rounded to tf32 by their producer, kind::tf32 MMAs, output conv in fp32):              rel-L2 <= 1e-3
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-2
TOLS = {"bf16": 1e-2, "tf32": 1e-3}
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "generator_golden.npz")


def rl2(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30))


def inputs(B, T, S, seed):
    rng = np.random.default_rng(seed)
    image = rng.standard_normal((B, T, S, S, 3)).astype(np.float32)
    noise = (0.1 * rng.standard_normal((B, T, S, S, 20))).astype(np.float32)
    return image, noise


@pytest.fixture(scope="module")
def mk():
    import torch
    assert torch.cuda.is_available()
    from wind_downscaling_gan_b200 import _lib
    _lib.lib()  # the CUDA extension must be the thing that runs: fail loudly if it is missing
    from wind_downscaling_gan_b200.gan.models import make_generator
    return make_generator


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
@pytest.mark.parametrize("B,T,S", [(2, 3, 96), (1, 1, 96), (3, 2, 64), (5, 4, 32)])
def test_forward_matches_oracle_per_layer(mk, B, T, S, precision):
    from oracle.generator import generator_forward, synthetic_generator_weights
    tol = TOLS[precision]
    w = synthetic_generator_weights(3)
    image, noise = inputs(B, T, S, 4)
    ref, inter = generator_forward(w, image, noise, return_intermediates=True)
    gen = mk(S, 3, 20, 2, T).set_precision(precision)
    gen.set_weights(w)
    out = gen.predict([image, noise])
    assert out.shape == (B, T, S, S, 2) and out.dtype == np.float32
    for k, name in enumerate(["res_2", "res_4", "lstm", "g5", "g7", "g9"]):
        assert rl2(gen.debug_intermediate(k), inter[name]) < tol, name
    assert rl2(out, ref) < tol


def test_precision_switch_on_one_handle(mk):
    """set_precision re-packs the weights and re-plans: bf16 -> tf32 -> bf16 on one handle reproduces each mode bit for bit."""
    from oracle.generator import generator_forward, synthetic_generator_weights
    w = synthetic_generator_weights(5)
    image, noise = inputs(2, 2, 96, 11)
    ref = generator_forward(w, image, noise)
    gen = mk(96, 3, 20, 2, 2)
    gen.set_weights(w)
    a = gen.predict([image, noise])
    b = gen.set_precision("tf32").predict([image, noise])
    c = gen.set_precision("bf16").predict([image, noise])
    assert np.array_equal(a, c) and not np.array_equal(a, b)
    assert rl2(b, ref) < 1e-3 < rl2(a, ref) < 1e-2
    with pytest.raises(ValueError):
        gen.set_precision("fp8")


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_golden_fixture(mk, precision):
    """Committed oracle outputs (tests/golden/make_generator_golden.py): no oracle import needed."""
    from oracle.generator import synthetic_generator_weights  # weights only
    z = np.load(GOLDEN)
    for name in ("b1_t2_s32", "b2_t3_s64"):
        B, T, S, ws, xs = (int(v) for v in z[name + "_meta"])
        image, noise = inputs(B, T, S, xs)
        gen = mk(S, 3, 20, 2, T).set_precision(precision)
        gen.set_weights(synthetic_generator_weights(ws))
        assert rl2(gen.predict([image, noise]), z[name].astype(np.float64)) < TOLS[precision]


def test_default_initialised_network_matches_oracle(mk):
    """Random-init generator as `get_network` builds it before load_weights (api.py:68-70)."""
    from oracle.generator import generator_forward
    gen = mk(96, 3, 20, 2, 2)
    w = gen.get_weights()
    assert np.all(w["layer_with_weights-1/gamma"] == 1) and np.all(w["layer_with_weights-4/cell/bias"][128:256] == 1)
    image, noise = inputs(2, 2, 96, 9)
    assert rl2(gen.predict([image, noise]), generator_forward(w, image, noise)) < TOL


def test_device_and_host_entry_points_agree_bitwise(mk):
    import torch
    from oracle.generator import synthetic_generator_weights
    gen = mk(96, 3, 20, 2, 3)
    gen.set_weights(synthetic_generator_weights(1))
    image, noise = inputs(2, 3, 96, 5)
    a = gen.predict([image, noise])
    b = gen.forward_device(torch.from_numpy(image).cuda(), torch.from_numpy(noise).cuda()).cpu().numpy()
    c = gen([torch.from_numpy(image), torch.from_numpy(noise)], training=False).numpy()
    assert np.array_equal(a, b) and np.array_equal(a, c)


def test_pipelined_host_path_is_bitwise_equal(mk):
    """B >= 32 goes through the chunked H2D | forward | D2H pipeline (chunks of 16 + a tail plan)."""
    import torch
    from oracle.generator import synthetic_generator_weights
    B, T, S = 37, 2, 32
    gen = mk(S, 3, 20, 2, T)
    gen.set_weights(synthetic_generator_weights(4))
    image, noise = inputs(B, T, S, 8)
    host = gen.predict_host(torch.from_numpy(image).pin_memory(), torch.from_numpy(noise).pin_memory()).numpy()
    dev = gen.forward_device(torch.from_numpy(image).cuda(), torch.from_numpy(noise).cuda()).cpu().numpy()
    assert np.array_equal(host, dev)
    again = gen.predict_host(image, noise).numpy()       # pageable buffers, second call
    assert np.array_equal(again, dev)


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
@pytest.mark.parametrize("B", [37, 64])
def test_in_kernel_noise_equals_materialised_noise(mk, precision, B):
    """Noise drawn inside the packing kernel (wdg_generator_forward_gen_noise) == the FlexibleNoiseGenerator tensor fed
    through forward(), bit for bit, on the device entry point and on the pipelined host entry point, whose pieces
    (B = 37: 8 | 21 | 8, B = 64: 8 | 16 | 32 | 8 sequences) each start at their own counter offset."""
    import torch
    from oracle.generator import synthetic_generator_weights
    from wind_downscaling_gan_b200.data.data_generator import FlexibleNoiseGenerator
    T, S = 2, 32
    gen = mk(S, 3, 20, 2, T).set_precision(precision)
    gen.set_weights(synthetic_generator_weights(4))
    image, _ = inputs(B, T, S, 8)
    img_d = torch.from_numpy(image).cuda()
    a_gen, b_gen, c_gen = (FlexibleNoiseGenerator((B, T, S, S, 20), std=0.1, random_seed=11) for _ in range(3))
    a_gen(3), b_gen(3), c_gen(3)                      # start from a non-zero stream offset
    noise = a_gen(B)
    ref = gen.forward_device(img_d, noise).clone()
    fused = gen.forward_device_gen_noise(img_d, b_gen)
    assert torch.equal(fused, ref)
    assert b_gen._offset == a_gen._offset             # the generator advanced as if the tensor had been drawn
    host = gen.predict_host_gen_noise(image, c_gen)
    assert torch.equal(host.cuda(), ref)
    assert float(noise.std()) == pytest.approx(0.1, rel=2e-2)


def test_full_size_properties(mk):
    """BASELINE configs[1] size (64 x 8 x 96 x 96): size-independent properties instead of the oracle.
    (1) sequences are independent: a sequence computed inside the batch of 64 equals the same sequence
        computed in a batch of 2, bit for bit;  (2) causality: changing inputs at t >= 5 leaves outputs
        at t < 5 bit-identical;  (3) no state leaks between calls."""
    import torch
    from oracle.generator import synthetic_generator_weights
    B, T, S = 64, 8, 96
    g = torch.Generator(device="cuda").manual_seed(0)
    image = torch.randn((B, T, S, S, 3), device="cuda", generator=g)
    noise = 0.1 * torch.randn((B, T, S, S, 20), device="cuda", generator=g)
    gen = mk(S, 3, 20, 2, T)
    gen.set_weights(synthetic_generator_weights(0))
    full = gen.forward_device(image, noise).clone()
    assert torch.isfinite(full).all()
    again = gen.forward_device(image, noise)
    assert torch.equal(full, again)
    sub = gen.forward_device(image[10:12].contiguous(), noise[10:12].contiguous())
    assert torch.equal(sub, full[10:12])
    image2, noise2 = image.clone(), noise.clone()
    image2[:, 5:] += 1.0
    noise2[:, 5:] *= -1.0
    changed = gen.forward_device(image2, noise2)
    assert torch.equal(changed[:, :5], full[:, :5])
    assert not torch.equal(changed[:, 5:], full[:, 5:])


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_full_size_batch_slices_match_oracle(mk, precision):
    """BASELINE configs[1] at full size (64 sequences x 8 timesteps x 96 x 96, the bench workload): sequences 0, 31
    and 63 of the batch against the float64 oracle run on those sequences (sequences are independent)."""
    import torch
    from oracle.generator import generator_forward, synthetic_generator_weights
    B, T, S = 64, 8, 96
    image, noise = inputs(B, T, S, 21)
    w = synthetic_generator_weights(0)
    gen = mk(S, 3, 20, 2, T).set_precision(precision)
    gen.set_weights(w)
    out = gen.forward_device(torch.from_numpy(image).cuda(), torch.from_numpy(noise).cuda()).cpu().numpy()
    pick = [0, 31, 63]
    ref = generator_forward(w, image[pick], noise[pick])
    for i, b in enumerate(pick):
        assert rl2(out[b], ref[i]) < TOLS[precision], (b, rl2(out[b], ref[i]))
    # and through the pipelined host entry point (chunks of 16 sequences)
    host = gen.predict_host(image, noise).numpy()
    assert np.array_equal(host, out)


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_24_timesteps_match_oracle(mk, precision):
    """The reference's sequence length (api.py:22): error does not build up over 24 recurrent steps."""
    from oracle.generator import generator_forward, synthetic_generator_weights
    B, T, S = 2, 24, 96
    image, noise = inputs(B, T, S, 22)
    w = synthetic_generator_weights(6)
    gen = mk(S, 3, 20, 2, T).set_precision(precision)
    gen.set_weights(w)
    out = gen.predict([image, noise])
    ref = generator_forward(w, image, noise)
    assert rl2(out, ref) < TOLS[precision]
    assert rl2(out[:, -1], ref[:, -1]) < TOLS[precision]      # the last timestep alone


def test_weights_roundtrip_and_reload(mk, tmp_path):
    from oracle.generator import synthetic_generator_weights
    w = synthetic_generator_weights(2)
    gen = mk(96, 3, 20, 2, 2)
    gen.set_weights(w)
    got = gen.get_weights()
    assert set(got) == set(w) and all(np.array_equal(got[k], w[k]) for k in w)
    image, noise = inputs(1, 2, 96, 6)
    a = gen.predict([image, noise])
    gen.save_weights(tmp_path / "ckpt" / "generator")
    gen2 = mk(96, 3, 20, 2, 2)
    gen2.load_weights(tmp_path / "ckpt" / "generator")
    assert np.array_equal(gen2.predict([image, noise]), a)
    # TF-checkpoint-V2 bundle with the reference's variable names
    from wind_downscaling_gan_b200.tf_checkpoint import write_bundle
    (tmp_path / "tf").mkdir()
    write_bundle(tmp_path / "tf" / "generator", w)
    gen3 = mk(96, 3, 20, 2, 2)
    gen3.load_weights(tmp_path / "tf" / "generator")
    assert np.array_equal(gen3.predict([image, noise]), a)


def test_error_behaviour(mk):
    from wind_downscaling_gan_b200._lib import WdgError
    with pytest.raises(AssertionError):
        mk(98, 3, 20, 2, 4)            # models.py:19
    with pytest.raises(WdgError):
        mk(96, 3, 4, 2, 4)             # 8*(3+4) < 128: not the configuration the kernels are built for
    gen = mk(96, 3, 20, 2, 2)
    image, noise = inputs(1, 2, 96, 1)
    with pytest.raises(ValueError):
        gen.predict([image[..., :2], noise])
    with pytest.raises(ValueError):
        gen.predict([image, noise[:, :1]])
    with pytest.raises(KeyError):
        gen.set_weights({"nope": np.zeros(3)})
    with pytest.raises(ValueError):
        gen.set_weights({"layer_with_weights-1/gamma": np.zeros(3)})
    # training-mode call (batch statistics + spectral-norm power iteration) runs on the fp32 training kernels and,
    # like the Keras wrapper, mutates the stored kernel / sn_u / BatchNorm moving statistics
    before = gen.get_weights()
    out = gen([image, noise], training=True)
    assert tuple(out.shape) == (1, 2, 96, 96, 2)
    after = gen.get_weights()
    assert not np.array_equal(before["layer_with_weights-0/layer/sn_u"], after["layer_with_weights-0/layer/sn_u"])
    assert not np.array_equal(before["layer_with_weights-1/moving_mean"], after["layer_with_weights-1/moving_mean"])


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
@pytest.mark.parametrize("B,T,S", [(3, 8, 96), (64, 8, 96), (1, 24, 32)])
def test_persistent_convlstm_equals_per_step_launches(mk, precision, B, T, S, monkeypatch):
    """All T ConvLSTM steps run in ONE cooperative launch (step counters in global memory order the recurrent half of step
    t after every tile of step t-1; the input half overlaps).  Same tiles, same K order, same arithmetic as one launch
    per step (WDG_NO_LSTM_PERSIST=1): outputs bit-identical, repeatedly (a stale h_{t-1} read would show up as noise)."""
    import torch
    from oracle.generator import synthetic_generator_weights
    w = synthetic_generator_weights(2)
    image, noise = inputs(B, T, S, 21)
    image, noise = torch.from_numpy(image).cuda(), torch.from_numpy(noise).cuda()
    gen = mk(S, 3, 20, 2, T)
    gen.set_weights(w)
    gen.set_precision(precision)
    outs = [gen.forward_device(image, noise).clone() for _ in range(3)]
    assert gen.launches_per_forward() == 10
    monkeypatch.setenv("WDG_NO_LSTM_PERSIST", "1")
    ref_gen = mk(S, 3, 20, 2, T)
    ref_gen.set_weights(w)
    ref_gen.set_precision(precision)
    ref = ref_gen.forward_device(image, noise)
    assert ref_gen.launches_per_forward() == 9 + T
    for o in outs:
        assert torch.equal(o, ref)
