"""Times GAN.train_step on the api.py constants (B=8, T=24, 96x96) -- BASELINE config 4, single GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wind_downscaling_gan_b200.data.data_generator import FlexibleNoiseGenerator
from wind_downscaling_gan_b200.gan import train
from wind_downscaling_gan_b200.gan.ganbase import GAN
from wind_downscaling_gan_b200.gan.models import make_discriminator, make_generator

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 24
S = 96
gen, disc = make_generator(S, 3, 20, 2, T), make_discriminator(S, S, 3, 2, T)
gan = GAN(gen, disc, FlexibleNoiseGenerator((B, T, S, S, 20), std=0.1, random_seed=0))
gan.compile(generator_optimizer=train.generator_optimizer(), discriminator_optimizer=train.discriminator_optimizer(),
            discriminator_loss=train.discriminator_loss)
g = torch.Generator(device="cuda").manual_seed(0)
lr = torch.randn((B, T, S, S, 3), device="cuda", generator=g)
hr = torch.randn((B, T, S, S, 2), device="cuda", generator=g)
for i in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    m = gan.train_step((lr, hr))
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"step {i}: {dt:.3f} s  ({B / dt:.2f} samples/s)  d_loss={m['d_loss']:.4f} g_loss={m['g_loss']:.4f} gp={m['d_gradient_pen']:.4f}", flush=True)
print("peak mem GB", torch.cuda.max_memory_allocated() / 1e9)
