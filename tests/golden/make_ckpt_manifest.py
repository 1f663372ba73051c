"""Generates tests/golden/ckpt_manifest.json from the reference's own checkpoint index files
(/root/reference/src/downscaling/weights-55.ckpt/{generator,discriminator}.index).

Run in the build container (the reference is not present on the GPU box):
    python tests/golden/make_ckpt_manifest.py
The .index files are TF-checkpoint-V2 SSTables; the data shards are
absent (.MISSING_LARGE_BLOBS), so only names / dtypes / shapes are recorded.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from wind_downscaling_gan_b200.tf_checkpoint import read_index  # noqa: E402

REF = "/root/reference/src/downscaling/weights-55.ckpt"


def main():
    out = {}
    for model in ("generator", "discriminator"):
        entries = read_index(os.path.join(REF, model + ".index"))
        table = {}
        for key, e in entries.items():
            if "OPTIMIZER_SLOT" in key or key.startswith(("optimizer", "_CHECKPOINTABLE", "save_counter")):
                continue
            name = key.replace("/.ATTRIBUTES/VARIABLE_VALUE", "")
            table[name] = {"dtype": e["dtype"], "shape": e["shape"]}
        out[model] = table
    with open(os.path.join(HERE, "ckpt_manifest.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
