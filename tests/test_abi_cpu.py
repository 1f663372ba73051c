"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/wdg.h
declares; without a GPU the compute entry points fail loudly instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    txt = open(os.path.join(ROOT, "include", "wdg.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(wdg_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from wind_downscaling_gan_b200 import _lib
    return _lib.lib()


def test_library_exports_every_declared_symbol(lib):
    names = declared_functions()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_bindings_cover_header(lib):
    from wind_downscaling_gan_b200 import _lib
    for n in declared_functions():
        fn = getattr(lib, n)
        assert fn.argtypes is not None or n in ("wdg_last_error",), f"{n} has no ctypes prototype in _lib.py"


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.wdg_generator_create(ctypes.byref(h), 96, 3, 20, 2, 24, 128)
    assert rc != 0 and b"no CUDA device" in lib.wdg_last_error()
    from wind_downscaling_gan_b200.gan.models import make_generator
    from wind_downscaling_gan_b200._lib import WdgError
    with pytest.raises(WdgError):
        make_generator(96, 3, 20, 2, 24)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "wind_downscaling_gan_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(d, f)


def test_tf_checkpoint_bundle_roundtrip(tmp_path):
    import numpy as np
    from wind_downscaling_gan_b200.tf_checkpoint import read_bundle, read_index, write_bundle
    t = {"layer_with_weights-0/layer/w": np.random.rand(2, 2, 3, 4).astype(np.float32),
         "layer_with_weights-1/gamma": np.arange(5, dtype=np.float32)}
    write_bundle(tmp_path / "generator", t)
    r = read_bundle(tmp_path / "generator")
    assert set(r) == set(t) and all(np.array_equal(r[k], t[k]) for k in t)
    idx = read_index(str(tmp_path / "generator") + ".index")
    assert idx["layer_with_weights-1/gamma/.ATTRIBUTES/VARIABLE_VALUE"]["shape"] == [5]


def test_crc32c_reproduces_the_reference_checkpoint_trailers():
    """CRC-32C + TensorFlow's mask against every block trailer of the reference's own weights-55.ckpt/*.index
    (tests/golden/ckpt_crc.json, make_ckpt_crc.py) and the standard check value."""
    import json
    import os
    from wind_downscaling_gan_b200.tf_checkpoint import crc32c, masked_crc32c
    assert crc32c(b"123456789") == 0xE3069283
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ckpt_crc.json")))
    n = 0
    for model in ("generator", "discriminator"):
        for blk in gold[model]["blocks"]:
            assert masked_crc32c(bytes.fromhex(blk["bytes_and_type"])) == blk["stored_masked_crc32c"], (model, blk["kind"])
            n += 1
    assert n == 6


def test_written_bundle_carries_tensorflow_header_and_checksums(tmp_path):
    """What TensorFlow's BundleReader checks: header proto identical to the one in the reference's index files, a valid
    masked crc32c behind every block and in every entry; a flipped data byte is detected on read."""
    import json
    import os
    import struct
    import numpy as np
    from wind_downscaling_gan_b200.tf_checkpoint import _block, _varint, masked_crc32c, read_bundle, read_index, write_bundle
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ckpt_crc.json")))
    rng = np.random.default_rng(0)
    t = {"layer_with_weights-0/layer/w": rng.standard_normal((8, 8, 23, 128)).astype(np.float32),
         "layer_with_weights-4/cell/bias": rng.standard_normal(512).astype(np.float32)}
    prefix = str(tmp_path / "generator")
    write_bundle(prefix, t)
    b = open(prefix + ".index", "rb").read()
    foot, q = b[-48:], 0
    spans = []
    for _ in range(2):
        off, q = _varint(foot, q)
        size, q = _varint(foot, q)
        spans.append((off, size))
    for _, h in _block(b, *spans[1]):
        off, r = _varint(h, 0)
        size, r = _varint(h, r)
        spans.append((off, size))
        header = [v for k, v in _block(b, off, size) if k == b""][0]
        assert header.hex() == gold["generator"]["header_proto"] == gold["discriminator"]["header_proto"]
    assert len(spans) == 3
    for off, size in spans:
        assert b[off + size] == 0 and struct.unpack("<I", b[off + size + 1:off + size + 5])[0] == masked_crc32c(b[off:off + size + 1])
    raw = open(prefix + ".data-00000-of-00001", "rb").read()
    for key, e in read_index(prefix + ".index").items():
        assert e["crc32c"] == masked_crc32c(raw[e["offset"]:e["offset"] + e["size"]]), key
    back = read_bundle(prefix)
    assert all(np.array_equal(back[k], t[k]) for k in t)
    bad = bytearray(raw)
    bad[1000] ^= 0x10
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="crc32c"):
        read_bundle(prefix)
    bad_index = bytearray(b)
    bad_index[40] ^= 1
    open(prefix + ".index", "wb").write(bytes(bad_index))
    with pytest.raises(ValueError, match="crc32c"):
        read_index(prefix + ".index")


def test_philox_known_answers(lib):
    """Philox4x32-10 block function against the Random123 known-answer vectors (Salmon et al., SC'11)."""
    import ctypes as C
    def ph(ctr, key):
        c, k, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
        lib.wdg_philox4x32_10(c, k, o)
        return [int(v) for v in o]
    assert ph([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert ph([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert ph([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_host_pipeline_schedules(lib, monkeypatch):
    """The pieces predict_host* cuts a batch into (csrc/wdg_generator.cu: make_schedules): cover the batch exactly, the
    device-noise schedule starts and ends small, the host-noise one ends with the short piece."""
    def sched(B, host):
        buf = (ctypes.c_int * 32)()
        n = lib.wdg_generator_pipeline_schedule(B, host, buf, 32)
        return [buf[i] for i in range(n)]
    monkeypatch.delenv("WDG_CHUNK_B", raising=False)
    assert sched(64, 0) == [8, 16, 32, 8] and sched(64, 1) == [16, 16, 16, 12, 4]
    assert sched(37, 0) == [8, 21, 8] and sched(37, 1) == [16, 16, 5]
    assert sched(32, 0) == [8, 16, 8] and sched(31, 0) == [] and sched(16, 1) == []
    for B in range(32, 300, 7):
        for host in (0, 1):
            p = sched(B, host)
            assert sum(p) == B and min(p) >= 1, (B, host, p)
        d = sched(B, 0)
        assert d[0] == 8 and d[-1] == 8 and max(d) <= 39
    monkeypatch.setenv("WDG_CHUNK_B", "8")
    assert sched(24, 0) == [8, 8, 8] and sched(20, 1) == [8, 8, 4]
    assert lib.wdg_generator_pipeline_schedule(0, 0, None, 0) == -1
