"""Reader for TensorFlow checkpoint-V2 bundles (`<prefix>.index` + `<prefix>.data-00000-of-00001`),
the format `GAN.save_weights` / `load_weights` use in the reference (ganbase.py:132-140).

The index is a LevelDB-style SSTable with uncompressed, prefix-compressed blocks whose values are
`BundleEntryProto` messages (dtype, shape, shard, offset, size); tensor bytes are little-endian,
row-major in the data shard.  Pure Python, no TensorFlow.  (SURVEY.md Appendix B.)
"""
import os
import struct

import numpy as np

_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}


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
    for _, handle in _block(b, ioff, isize):
        off, q = _varint(handle, 0)
        size, q = _varint(handle, q)
        for key, val in _block(b, off, size):
            if key == b"":
                continue  # BundleHeaderProto
            d = _proto(val)
            shape = []
            for dim in _proto(d.get(2, [b""])[0]).get(2, []):
                shape.append(_proto(dim).get(1, [0])[0])
            entries[key.decode()] = {"dtype": int(d.get(1, [0])[0]), "shape": shape, "shard": int(d.get(3, [0])[0]),
                                     "offset": int(d.get(4, [0])[0]), "size": int(d.get(5, [0])[0])}
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
            a = np.frombuffer(raw, dtype=np.dtype(_DTYPES[e["dtype"]]).newbyteorder("<")).reshape(e["shape"])
            out[key[:-len("/.ATTRIBUTES/VARIABLE_VALUE")]] = a.astype(_DTYPES[e["dtype"]])
    return out


def write_bundle(prefix, tensors):
    """Writes {name: float32 ndarray} as a single-shard checkpoint-V2 bundle readable by read_bundle
    (one uncompressed data block; enough for round-trip tests and for exporting weights)."""
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
            field(5, 0, venc(a.nbytes))
        items.append(((name + "/.ATTRIBUTES/VARIABLE_VALUE").encode(), entry))
        data += a.tobytes()
    items.insert(0, (b"", field(1, 0, venc(1))))  # header: num_shards = 1

    def make_block(kvs):
        blk = bytearray()
        for k, v in kvs:
            blk += venc(0) + venc(len(k)) + venc(len(v)) + k + v
        restarts = struct.pack("<I", 0) + struct.pack("<I", 1)
        return bytes(blk) + restarts

    dblock = make_block(items)
    out = bytearray(dblock) + b"\x00" + b"\x00\x00\x00\x00"          # type 0 (no compression) + crc (unchecked)
    handle = venc(0) + venc(len(dblock))
    iblock = make_block([(items[-1][0] + b"\xff", handle)])
    ioff = len(out)
    out += iblock + b"\x00" + b"\x00\x00\x00\x00"
    mblock = make_block([])
    moff = len(out)
    out += mblock + b"\x00" + b"\x00\x00\x00\x00"
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
