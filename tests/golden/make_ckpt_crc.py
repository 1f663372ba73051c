"""Generates tests/golden/ckpt_crc.json: every SSTable block (bytes + compression-type byte, hex) of the reference's own
checkpoint index files together with the masked crc32c TensorFlow stored behind it, and the 6-byte BundleHeaderProto.
They are the known-answer vectors for csrc/wdg_crc32c.cu + tf_checkpoint.masked_crc32c and for the header the writer emits.

Run in the build container (the reference is not present on the GPU box):  python tests/golden/make_ckpt_crc.py
"""
import json
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from wind_downscaling_gan_b200.tf_checkpoint import _block, _varint  # noqa: E402

REF = "/root/reference/src/downscaling/weights-55.ckpt"


def main():
    out = {}
    for model in ("generator", "discriminator"):
        b = open(os.path.join(REF, model + ".index"), "rb").read()
        foot, q = b[-48:], 0
        moff, q = _varint(foot, q)
        msz, q = _varint(foot, q)
        ioff, q = _varint(foot, q)
        isz, q = _varint(foot, q)
        spans = [("metaindex", moff, msz), ("index", ioff, isz)]
        header = None
        for _, h in _block(b, ioff, isz):
            off, r = _varint(h, 0)
            size, r = _varint(h, r)
            spans.append(("data", off, size))
            for key, val in _block(b, off, size):
                if key == b"":
                    header = val.hex()
        out[model] = {"header_proto": header,
                      "blocks": [{"kind": k, "offset": o, "bytes_and_type": b[o:o + s + 1].hex(),
                                  "stored_masked_crc32c": struct.unpack("<I", b[o + s + 1:o + s + 5])[0]} for k, o, s in spans]}
    with open(os.path.join(HERE, "ckpt_crc.json"), "w") as f:
        json.dump(out, f, indent=1)
    print({k: len(v["blocks"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
