#!/usr/bin/env python
"""One generator forward at the bench shape (64 x 8 x 96 x 96) for ncu captures.
    python tools/one_forward.py [bf16|tf32] [warm-up forwards] [B] [T]
Each forward is 17 launches (T = 8): skip 17 * warm-ups with `ncu -s`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from oracle.generator import synthetic_generator_weights
    from wind_downscaling_gan_b200.gan.models import make_generator
    prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    warm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    T = int(sys.argv[4]) if len(sys.argv) > 4 else 8
    g = torch.Generator(device="cuda").manual_seed(0)
    image = torch.randn((B, T, 96, 96, 3), device="cuda", generator=g)
    noise = 0.1 * torch.randn((B, T, 96, 96, 20), device="cuda", generator=g)
    gen = make_generator(96, 3, 20, 2, T).set_precision(prec)
    gen.set_weights(synthetic_generator_weights(0))
    for _ in range(warm + 1):
        out = gen.forward_device(image, noise)
    torch.cuda.synchronize()
    print(float(out.abs().mean()))


if __name__ == "__main__":
    main()
