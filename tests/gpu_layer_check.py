"""Per-layer parity report of the CUDA generator against the float64 oracle (run on a GPU box).

    python tests/gpu_layer_check.py [B] [T] [S] [bf16|tf32]

Prints the relative L2 error of every intermediate activation and of the output; exits
non-zero if the output misses the bf16 tolerance (1e-2, BASELINE.json north_star)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.generator import generator_forward, synthetic_generator_weights  # noqa: E402


def rl2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 96
    precision = sys.argv[4] if len(sys.argv) > 4 else "bf16"
    tol = {"bf16": 1e-2, "tf32": 1e-3}[precision]
    import torch
    from wind_downscaling_gan_b200.gan.models import make_generator
    print("device:", torch.cuda.get_device_name(0), flush=True)
    w = synthetic_generator_weights(3)
    rng = np.random.default_rng(4)
    image = rng.standard_normal((B, T, S, S, 3)).astype(np.float32)
    noise = (0.1 * rng.standard_normal((B, T, S, S, 20))).astype(np.float32)
    t0 = time.time()
    ref, inter = generator_forward(w, image, noise, return_intermediates=True)
    print(f"oracle: {time.time() - t0:.1f}s", flush=True)
    gen = make_generator(S, 3, 20, 2, T).set_precision(precision)
    gen.set_weights(w)
    print("precision:", precision, flush=True)
    out = gen.predict([image, noise])
    torch.cuda.synchronize()
    names = {0: "res_2", 1: "res_4", 2: "lstm", 3: "g5", 4: "g7", 5: "g9"}
    for k, name in names.items():
        got = gen.debug_intermediate(k)
        print(f"{name:6s} rel-L2 = {rl2(got, inter[name]):.3e}  (nan: {int(np.isnan(got).sum())})", flush=True)
    e = rl2(out, ref)
    print(f"output rel-L2 = {e:.3e}", flush=True)
    # second call must reproduce the first bit for bit (no stale state)
    out2 = gen.predict([image, noise])
    print("repeatable:", bool(np.array_equal(out, out2)), flush=True)
    sys.exit(0 if (e < tol and np.array_equal(out, out2)) else 1)


if __name__ == "__main__":
    main()
