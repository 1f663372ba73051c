"""Generates tests/golden/generator_golden.npz: outputs of the float64 oracle (oracle/generator.py)
for small seeded cases.  ORACLE-GENERATED (parity unpinned: TensorFlow cannot run here); the file
guards against silent drift of the oracle itself and gives the GPU tests a fixture that does not
need the oracle at run time.

    python tests/golden/make_generator_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.generator import generator_forward, synthetic_generator_weights  # noqa: E402

CASES = {"b1_t2_s32": (1, 2, 32, 11, 12), "b2_t3_s64": (2, 3, 64, 21, 22)}  # B, T, S, weight seed, input seed


def inputs(B, T, S, seed):
    rng = np.random.default_rng(seed)
    image = rng.standard_normal((B, T, S, S, 3)).astype(np.float32)
    noise = (0.1 * rng.standard_normal((B, T, S, S, 20))).astype(np.float32)
    return image, noise


def main():
    out = {}
    for name, (B, T, S, ws, xs) in CASES.items():
        image, noise = inputs(B, T, S, xs)
        y = generator_forward(synthetic_generator_weights(ws), image, noise)
        out[name] = y.astype(np.float32)
        out[name + "_meta"] = np.array([B, T, S, ws, xs])
    np.savez_compressed(os.path.join(HERE, "generator_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
