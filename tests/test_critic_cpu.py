"""CPU tests of the `wdg_critic` handle's host logic (csrc/wdg_critic.cu builds the layer plan without a device): the
variable table against the reference's own discriminator.index (tests/golden/ckpt_manifest.json) and against the
oracle's walk of models.py:93-140 for both graph revisions, the flat layout, and checkpoint-driven topology selection."""
import ctypes as C
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_variable_table_equals_the_shipped_checkpoint():
    from wind_downscaling_gan_b200.train.nets import CriticHandle
    man = json.load(open(os.path.join(HERE, "golden", "ckpt_manifest.json")))["discriminator"]
    h = CriticHandle.get(96, 3, 2, 16, ckpt_topology=True)
    assert {n: list(s) for n, s in h.shapes().items()} == {n: e["shape"] for n, e in man.items()}
    sc = [n for n, s in h.shapes().items() if s[:2] == (6, 6)]
    assert sc == ["layer_with_weights-11/layer/w"] and h.shapes()[sc[0]] == (6, 6, 128, 256)
    cur = CriticHandle.get(96, 3, 2, 16, ckpt_topology=False)
    assert set(h.shapes()) - set(cur.shapes()) and len(cur.shapes()) == len(h.shapes()) - 5         # conv w/b/sn_u + LN gamma/beta


@pytest.mark.parametrize("size", [32, 64, 96, 112, 128])
@pytest.mark.parametrize("topo", [False, True])
def test_plan_follows_the_graph_building_loops(size, topo):
    from oracle.critic import critic_weight_shapes
    from wind_downscaling_gan_b200.train.nets import CriticHandle
    ref, _ = critic_weight_shapes(size, 3, 2, 16, topo)
    h = CriticHandle.get(size, 3, 2, 16, topo)
    assert h.shapes() == {k: tuple(v) for k, v in ref.items()}
    assert list(h.shapes()) == list(ref)           # creation order of the weighted layers = checkpoint numbering


def test_flat_layout():
    from wind_downscaling_gan_b200.train.nets import CriticHandle
    h = CriticHandle.get(96, 3, 2, 16, True)
    spans = sorted((off, off + int(np.prod(s)), tr, n) for n, (s, off, tr) in h.table.items())
    assert spans[0][0] == 0 and all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))      # disjoint
    assert all(off % 64 == 0 for off, _, _, _ in spans)                                    # 256-byte aligned
    assert max(e for _, e, tr, _ in spans if tr) <= h.n_train <= min(o for o, _, tr, _ in spans if not tr)
    assert all(n.endswith("sn_u") for _, _, tr, n in spans if not tr) and spans[-1][1] <= h.n_total
    # parameter count of the current-code graph: 2,124,905 trainable
    cur = CriticHandle.get(96, 3, 2, 16, False)
    assert sum(int(np.prod(s)) for s, _, tr in cur.table.values() if tr) == 2_124_905


def test_create_errors_mirror_the_reference():
    from wind_downscaling_gan_b200 import _lib
    from wind_downscaling_gan_b200.gan.models import make_discriminator
    L = _lib.lib()
    h = C.c_void_p()
    assert L.wdg_critic_create(C.byref(h), 96, 48, 3, 2, 24, 16, 0) != 0
    assert b"same size" in L.wdg_last_error()                       # models.py:89-91
    assert L.wdg_critic_create(C.byref(h), 16, 16, 3, 2, 24, 16, 0) != 0   # 16 -> 4 -> a 7x7 window no longer fits
    assert L.wdg_critic_create(C.byref(h), 17, 17, 3, 2, 24, 16, 0) == 0   # 17 -> 5 -> 1
    with pytest.raises(NotImplementedError):
        make_discriminator(96, 48, 3, 2, 24)


def test_loading_the_shortcut_checkpoint_rebuilds_the_critic(tmp_path, capsys):
    """get_network builds the current-code critic (api.py:71-73) and then loads weights-55.ckpt, whose discriminator has the
    shortcut branch: the load adopts the checkpoint's graph instead of leaving layers at their random initialisation."""
    from oracle.critic import synthetic_critic_weights
    from wind_downscaling_gan_b200.gan.models import make_discriminator
    from wind_downscaling_gan_b200.tf_checkpoint import write_bundle
    w = synthetic_critic_weights(3, size=96, ckpt_topology=True)
    extra = {"optimizer/iter": np.zeros((), np.float32),
             "layer_with_weights-0/cell/kernel/.OPTIMIZER_SLOT/optimizer/m": np.zeros((3, 3, 2, 8), np.float32)}
    write_bundle(tmp_path / "discriminator", {**w, **extra})
    d = make_discriminator(96, 96, 3, 2, 24)
    assert not d.ckpt_topology and "layer_with_weights-14/layer/kernel" not in d.weight_names()
    d.load_weights(tmp_path / "discriminator")
    assert d.ckpt_topology and "shortcut" in capsys.readouterr().out
    got = d.get_weights()
    assert set(got) == set(w) and all(np.array_equal(got[k], w[k]) for k in w)
    # and back: a current-code checkpoint into a shortcut-topology critic
    w0 = synthetic_critic_weights(4, size=96)
    write_bundle(tmp_path / "d0", w0)
    d.load_weights(tmp_path / "d0")
    assert not d.ckpt_topology and all(np.array_equal(d.get_weights()[k], w0[k]) for k in w0)
    # a checkpoint of neither graph still fails loudly
    bad = dict(w0)
    bad.pop("layer_with_weights-12/layer/bias")
    write_bundle(tmp_path / "bad", bad)
    with pytest.raises(ValueError, match="topology"):
        d.load_weights(tmp_path / "bad")


def test_save_weights_writes_checkpoint_bundles(tmp_path):
    from wind_downscaling_gan_b200.gan.models import make_discriminator
    from wind_downscaling_gan_b200.tf_checkpoint import read_index
    d = make_discriminator(32, 32, 3, 2, 4)
    d.save_weights(tmp_path / "ck" / "discriminator")
    idx = read_index(str(tmp_path / "ck" / "discriminator") + ".index")
    assert "layer_with_weights-0/cell/kernel/.ATTRIBUTES/VARIABLE_VALUE" in idx
    d2 = make_discriminator(32, 32, 3, 2, 4)
    d2.load_weights(tmp_path / "ck" / "discriminator")
    a, b = d.get_weights(), d2.get_weights()
    assert all(np.array_equal(a[k], b[k]) for k in a)


def test_buffer_sizes_are_host_arithmetic():
    """context / scratch sizes come from the same allocation walk the kernels' launcher uses; no device needed, and the
    process-wide training precision is not touched by asking."""
    from wind_downscaling_gan_b200 import _lib
    from wind_downscaling_gan_b200.train.nets import CriticHandle
    L = _lib.lib()
    h = CriticHandle.get(96, 3, 2, 16, True)
    before = L.wdg_train_get_precision()
    ctx_t, ctx_i, scr = C.c_size_t(), C.c_size_t(), C.c_size_t()
    assert L.wdg_critic_context_bytes(h.h, 8, 24, 1, C.byref(ctx_t)) == 0
    assert L.wdg_critic_context_bytes(h.h, 8, 24, 0, C.byref(ctx_i)) == 0
    assert L.wdg_critic_scratch_bytes(h.h, 8, 24, C.byref(scr)) == 0
    assert L.wdg_train_get_precision() == before
    px = 8 * 24 * 96 * 96
    # gates of the 16-filter cell alone are px * 64 floats; the training context additionally snapshots the variables
    assert ctx_i.value > px * 64 * 4 and ctx_t.value - ctx_i.value >= h.n_train * 4
    assert scr.value > px * 16 * 4 and scr.value % 256 == 0 or scr.value > 0
    small = C.c_size_t()
    assert L.wdg_critic_context_bytes(h.h, 1, 2, 1, C.byref(small)) == 0 and small.value < ctx_t.value
    assert L.wdg_critic_context_bytes(h.h, 0, 2, 1, C.byref(small)) != 0
