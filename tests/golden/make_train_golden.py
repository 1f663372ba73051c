"""Golden fixture of ONE full WGAN `train_step` (ganbase.py:21-94) at the reference's real training shape --
96 x 96 pixels, 24 timesteps (api.py:22-23) -- computed by the float64 torch-autograd oracle
(oracle/torch_train.py) on the CPU.  Minutes of CPU time, so it is run once and committed:

    python tests/golden/make_train_golden.py            # writes tests/golden/train_step_s96_t24.npz

The updated weights (3.9 M parameters) are too large to commit; the fixture holds, per weight tensor, the norm of
the updated tensor, the norm of its change over the step and NPROJ projections of that change on seeded Gaussian
vectors (a difference e between two tensors moves a projection by ~|e|), plus the metrics dict.  Inputs, weights and
random draws are regenerated from the seeds by `case()` -- the same function the GPU test calls.

fp32 noise floor.  The reference trains in float32, and these piecewise-linear networks (LeakyReLU, hard-sigmoid,
gradients that are cancelling sums behind every normalisation layer) amplify rounding: the SAME oracle code run in
torch float32 moves the generator's update by ~2e-2 and the critic's by ~2e-3 (relative L2 of the change) away from
the float64 result.  The fixture stores that per-tensor distance (`floor`) so the GPU test can ask for "as close to
float64 as an independent float32 implementation is" instead of a made-up bound.
"""
import os
import sys
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

B, T, S = 1, 24, 96
NPROJ = 8
OUT = os.path.join(HERE, "train_step_s96_t24.npz")


def case(B=B, T=T, S=S, seed=16):
    """(low_res, high_res, draws, gen_weights, critic_weights) of the fixture."""
    from oracle.critic import synthetic_critic_weights
    from oracle.generator import synthetic_generator_weights
    rng = np.random.default_rng(seed)
    lr = rng.standard_normal((B, T, S, S, 3)).astype(np.float32)
    hr = rng.standard_normal((B, T, S, S, 2)).astype(np.float32)
    draws = []
    for _ in range(3):
        draws += [0.1 * rng.standard_normal((B, T, S, S, 20)), rng.uniform(0, 1, (B,)),
                  0.1 * rng.standard_normal((B, T, S, S, 2)), 0.1 * rng.standard_normal((B, T, S, S, 2))]
    draws += [0.1 * rng.standard_normal((B, T, S, S, 20)), 0.1 * rng.standard_normal((B, T, S, S, 20))]
    draws = [np.asarray(d, np.float32) for d in draws]
    return lr, hr, draws, synthetic_generator_weights(seed + 1), synthetic_critic_weights(seed + 2, size=S)


def projections(name, delta):
    """NPROJ projections of a float64 array on N(0,1) vectors seeded by the tensor name."""
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    flat = np.asarray(delta, np.float64).reshape(-1)
    return np.array([float(flat @ rng.standard_normal(flat.size)) for _ in range(NPROJ)])


def summarise(prefix, old, new, new32, blob):
    for k in sorted(new):
        a, b, c = np.asarray(old[k], np.float64), np.asarray(new[k], np.float64), np.asarray(new32[k], np.float64)
        blob[f"{prefix}/{k}/norm"] = np.array(np.linalg.norm(b))
        blob[f"{prefix}/{k}/dnorm"] = np.array(np.linalg.norm(b - a))
        blob[f"{prefix}/{k}/proj"] = projections(f"{prefix}/{k}", b - a)
        blob[f"{prefix}/{k}/floor"] = np.array(np.linalg.norm(c - b) / max(np.linalg.norm(b - a), 1e-300))


def main():
    import torch
    from oracle import torch_train as tt
    torch.set_num_threads(os.cpu_count() or 1)
    lr, hr, draws, gw, dw = case()
    st = tt.State(gw, dw)
    t0 = time.time()
    m = tt.train_step(st, lr, hr, draws)
    print(f"oracle train_step B={B} T={T} S={S}: {time.time() - t0:.0f} s", m, flush=True)
    tt.DT = torch.float32                      # the same step in float32: the noise floor of an fp32 implementation
    st32 = tt.State(gw, dw)
    m32 = tt.train_step(st32, lr, hr, draws)
    tt.DT = torch.float64
    blob = {"meta": np.array([B, T, S, 16, NPROJ])}
    for k, v in m.items():
        blob["metric/" + k] = np.array(v)
        blob["metric32/" + k] = np.array(m32[k])
    summarise("g", gw, {k: v.numpy() for k, v in st.g.items()}, {k: v.numpy() for k, v in st32.g.items()}, blob)
    summarise("d", dw, {k: v.numpy() for k, v in st.d.items()}, {k: v.numpy() for k, v in st32.d.items()}, blob)
    fl = {p: max(float(blob[k]) for k in blob if k.startswith(p + "/") and k.endswith("/floor")) for p in "gd"}
    print("fp32 noise floor (worst relative change distance): generator %.2e, critic %.2e" % (fl["g"], fl["d"]))
    np.savez_compressed(OUT, **blob)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
