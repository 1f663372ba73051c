"""Host-side overhead of GAN.train_step: same launch count as the full-size step (T = 24) on tiny images."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wind_downscaling_gan_b200.data.data_generator import FlexibleNoiseGenerator
from wind_downscaling_gan_b200.gan import train
from wind_downscaling_gan_b200.gan.ganbase import GAN
from wind_downscaling_gan_b200.gan.models import make_discriminator, make_generator
B, T, S = 1, 24, 32
gen, disc = make_generator(S, 3, 20, 2, T), make_discriminator(S, S, 3, 2, T)
gan = GAN(gen, disc, FlexibleNoiseGenerator((B, T, S, S, 20), std=0.1, random_seed=0))
gan.compile(generator_optimizer=train.generator_optimizer(), discriminator_optimizer=train.discriminator_optimizer(),
            discriminator_loss=train.discriminator_loss)
lr = torch.randn((B, T, S, S, 3), device="cuda"); hr = torch.randn((B, T, S, S, 2), device="cuda")
for i in range(3):
    torch.cuda.synchronize(); t0 = time.time(); gan.train_step((lr, hr)); torch.cuda.synchronize()
    print(f"tiny step {i}: {time.time()-t0:.3f} s")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); gan.train_step((lr, hr)); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
