"""Reader for TensorFlow checkpoint-V2 bundles (`<prefix>.index` + `<prefix>.data-00000-of-00001`),
the format `GAN.save_weights` / `load_weights` use in the reference (ganbase.py:132-140).

The index is a LevelDB-style SSTable with uncompressed, prefix-compressed blocks whose values are
`BundleEntryProto` messages (dtype, shape, shard, offset, size); tensor bytes are little-endian,
row-major in the data shard.  No TensorFlow needed.  (SURVEY.md Appendix B.)

Checksums are TensorFlow's: every SSTable block is followed by a 1-byte compression type and
`mask(crc32c(block + type))`, every BundleEntryProto carries `mask(crc32c(tensor bytes))` (field 6, fixed32), with
mask(c) = rotr(c, 15) + 0xa282ead8.  The CRC routine (csrc/wdg_crc32c.cu) reproduces all six block trailers of the
reference's own `weights-55.ckpt/*.index` files (tests/test_abi_cpu.py).  The reader verifies both kinds; the writer
emits both, plus the BundleHeaderProto TensorFlow writes (num_shards 1, little endian, version producer 1), so
`tf.train.load_checkpoint(prefix)` / BundleReader accept the files.  NOT written: the `_CHECKPOINTABLE_OBJECT_GRAPH`
entry Keras' object-based `load_weights` matches layers with -- a TensorFlow user restores by key
(`reader.get_tensor("layer_with_weights-0/layer/w/.ATTRIBUTES/VARIABLE_VALUE")`).
"""
import os
import struct

import numpy as np

_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}


def crc32c(data):
    from . import _lib
    b = bytes(data)
    return int(_lib.lib().wdg_crc32c(0, b, len(b)))


def masked_crc32c(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _check_block(b, off, size, path):
    stored = struct.unpack("<I", b[off + size + 1:off + size + 5])[0]
    if b[off + size] != 0:
        raise ValueError(f"{path}: compressed SSTable block (type {b[off + size]}) is not supported")
    if stored != masked_crc32c(b[off:off + size + 1]):
        raise ValueError(f"{path}: SSTable block at {off} fails its crc32c check (corrupt index)")


def _varint(b, p):
    r = s = 0
    while True:
        c = b[p]
        p += 1
        r |= (c & 0x7F) << s
        s += 7
        if c < 0x80:
            return r, p


def _block(b, off, size):
    blk = b[off:off + size]
    nrestart = struct.unpack("<I", blk[-4:])[0]
    end = len(blk) - 4 - 4 * nrestart
    p, key, out = 0, b"", []
    while p < end:
        shared, p = _varint(blk, p)
        non_shared, p = _varint(blk, p)
        vlen, p = _varint(blk, p)
        key = key[:shared] + blk[p:p + non_shared]
        p += non_shared
        out.append((key, blk[p:p + vlen]))
        p += vlen
    return out


def _proto(v):
    """Minimal protobuf wire decoder -> {field: [values]}."""
    p, d = 0, {}
    while p < len(v):
        tag, p = _varint(v, p)
        f, w = tag >> 3, tag & 7
        if w == 0:
            x, p = _varint(v, p)
        elif w == 2:
            n, p = _varint(v, p)
            x = v[p:p + n]
            p += n
        elif w == 5:
            x = v[p:p + 4]
            p += 4
        elif w == 1:
            x = v[p:p + 8]
            p += 8
        else:
            raise ValueError(f"unsupported wire type {w}")
        d.setdefault(f, []).append(x)
    return d


def read_index(path):
    """-> {key: {dtype, shape, shard, offset, size}} for every tensor entry of a .index file."""
    b = open(path, "rb").read()
    if b[-8:] != struct.pack("<Q", 0xDB4775248B80FB57):
        raise ValueError(f"{path}: not an SSTable (bad magic)")
    foot = b[-48:]
    p = 0
    _, p = _varint(foot, p)
    _, p = _varint(foot, p)
    ioff, p = _varint(foot, p)
    isize, p = _varint(foot, p)
    entries = {}
    _check_block(b, ioff, isize, path)
    for _, handle in _block(b, ioff, isize):
        off, q = _varint(handle, 0)
        size, q = _varint(handle, q)
        _check_block(b, off, size, path)
        for key, val in _block(b, off, size):
            if key == b"":
                continue  # BundleHeaderProto
            d = _proto(val)
            shape = []
            for dim in _proto(d.get(2, [b""])[0]).get(2, []):
                shape.append(_proto(dim).get(1, [0])[0])
            entries[key.decode()] = {"dtype": int(d.get(1, [0])[0]), "shape": shape, "shard": int(d.get(3, [0])[0]),
                                     "offset": int(d.get(4, [0])[0]), "size": int(d.get(5, [0])[0]),
                                     "crc32c": struct.unpack("<I", d[6][0])[0] if 6 in d else None}
    return entries


def read_bundle(prefix):
    """-> {variable name (without '/.ATTRIBUTES/VARIABLE_VALUE'): ndarray} of a checkpoint prefix."""
    prefix = str(prefix)
    entries = read_index(prefix + ".index")
    data_path = prefix + ".data-00000-of-00001"
    if not os.path.exists(data_path):
        raise FileNotFoundError(f"{data_path}: the tensor data shard is missing (only the .index was shipped)")
    out = {}
    with open(data_path, "rb") as f:
        for key, e in entries.items():
            if e["dtype"] not in _DTYPES or not key.endswith("/.ATTRIBUTES/VARIABLE_VALUE"):
                continue
            f.seek(e["offset"])
            raw = f.read(e["size"])
            if e["crc32c"] is not None and e["crc32c"] != masked_crc32c(raw):
                raise ValueError(f"{data_path}: tensor {key} fails its crc32c check")
            a = np.frombuffer(raw, dtype=np.dtype(_DTYPES[e["dtype"]]).newbyteorder("<")).reshape(e["shape"])
            out[key[:-len("/.ATTRIBUTES/VARIABLE_VALUE")]] = a.astype(_DTYPES[e["dtype"]])
    return out


def write_bundle(prefix, tensors):
    """Writes {name: float32 ndarray} as a single-shard checkpoint-V2 bundle: TensorFlow's header, per-tensor and
    per-block masked crc32c (one uncompressed data block)."""
    prefix = str(prefix)

    def venc(x):
        o = bytearray()
        while True:
            c = x & 0x7F
            x >>= 7
            if x:
                o.append(c | 0x80)
            else:
                o.append(c)
                return bytes(o)

    def field(f, w, payload):
        return venc((f << 3) | w) + payload

    data = bytearray()
    items = []
    for name in sorted(tensors):
        a = np.ascontiguousarray(tensors[name], dtype="<f4")
        shape = b"".join(field(2, 2, venc(len(field(1, 0, venc(d)))) + field(1, 0, venc(d))) for d in a.shape)
        entry = field(1, 0, venc(1)) + field(2, 2, venc(len(shape)) + shape) + field(4, 0, venc(len(data))) + \
            field(5, 0, venc(a.nbytes)) + field(6, 5, struct.pack("<I", masked_crc32c(a.tobytes())))
        items.append(((name + "/.ATTRIBUTES/VARIABLE_VALUE").encode(), entry))
        data += a.tobytes()
    # BundleHeaderProto as TensorFlow writes it (bytes 08 01 1a 02 08 01 in the reference's index files):
    # num_shards = 1, endianness LITTLE (default, omitted), version { producer: 1 }
    items.insert(0, (b"", field(1, 0, venc(1)) + field(3, 2, venc(2) + field(1, 0, venc(1)))))

    def make_block(kvs):
        blk = bytearray()
        for k, v in kvs:
            blk += venc(0) + venc(len(k)) + venc(len(v)) + k + v
        restarts = struct.pack("<I", 0) + struct.pack("<I", 1)
        return bytes(blk) + restarts

    def trailer(blk):                      # compression type 0 + masked crc32c over block and type byte
        return b"\x00" + struct.pack("<I", masked_crc32c(blk + b"\x00"))

    dblock = make_block(items)
    out = bytearray(dblock) + trailer(dblock)
    handle = venc(0) + venc(len(dblock))
    iblock = make_block([(items[-1][0] + b"\xff", handle)])
    ioff = len(out)
    out += iblock + trailer(iblock)
    mblock = make_block([])
    moff = len(out)
    out += mblock + trailer(mblock)
    foot = venc(moff) + venc(len(mblock)) + venc(ioff) + venc(len(iblock))
    foot += b"\x00" * (40 - len(foot)) + struct.pack("<Q", 0xDB4775248B80FB57)
    out += foot
    with open(prefix + ".index", "wb") as f:
        f.write(out)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(data)


def select_model_variables(tensors, shapes, model_name):
    """Model variables of a checkpoint bundle, checked against the model's own table (`shapes`: name -> shape).

    The bundle also holds optimizer slots and object-graph bookkeeping, which are skipped; but a model variable that
    the checkpoint lacks, a `layer_with_weights-*` variable the model does not have, or a shape mismatch means the
    checkpoint was written by a DIFFERENT topology (the reference's shipped critic checkpoint has the shortcut branch,
    SURVEY F6) and a silent partial load would leave layers at their random initialisation: raise instead."""
    def is_model_var(k):
        return k.startswith("layer_with_weights-") and "/.OPTIMIZER_SLOT" not in k and not k.endswith("/.ATTRIBUTES/OBJECT_CONFIG_JSON")
    missing = sorted(k for k in shapes if k not in tensors)
    extra = sorted(k for k in tensors if is_model_var(k) and k not in shapes)
    wrong = sorted(k for k in shapes if k in tensors and tuple(tensors[k].shape) != tuple(shapes[k]))
    if missing or extra or wrong:
        def head(v):
            return ", ".join(v[:4]) + (f", ... ({len(v)} in all)" if len(v) > 4 else "")
        raise ValueError(f"{model_name} checkpoint does not match this model's topology: "
                         + "; ".join(t for t in (f"missing from the checkpoint: {head(missing)}" if missing else "",
                                                 f"in the checkpoint but not in the model: {head(extra)}" if extra else "",
                                                 f"shape mismatch: {head(wrong)}" if wrong else "") if t))
    return {k: tensors[k] for k in shapes}
