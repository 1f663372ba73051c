#!/usr/bin/env python
"""One eager WGAN train_step (bench.py --workload train shapes: batch 8 x 24 x 96x96) between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file OUT.csv \
        python tools/ncu_train_step.py [tf32|bf16|fp32]
(the launch list of the training step; aggregate with tools/agg_launches.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["WDG_TRAIN_GRAPH"] = "0"
import numpy as np
import torch

from oracle.critic import synthetic_critic_weights
from oracle.generator import synthetic_generator_weights
from wind_downscaling_gan_b200.data.data_generator import FlexibleNoiseGenerator
from wind_downscaling_gan_b200.gan import train
from wind_downscaling_gan_b200.gan.ganbase import GAN
from wind_downscaling_gan_b200.gan.models import make_discriminator, make_generator

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
B, T, S = 8, 24, 96
gen, disc = make_generator(S, 3, 20, 2, T), make_discriminator(S, S, 3, 2, T)
gen.set_weights(synthetic_generator_weights(0))
disc.set_weights(synthetic_critic_weights(1, size=S))
gan = GAN(gen, disc, FlexibleNoiseGenerator((B, T, S, S, 20), std=0.1, random_seed=100))
gan.compile(generator_optimizer=train.generator_optimizer(), discriminator_optimizer=train.discriminator_optimizer(),
            discriminator_loss=train.discriminator_loss, train_precision=prec)
rng = np.random.default_rng(200)
lr = torch.from_numpy(rng.standard_normal((B, T, S, S, 3), dtype=np.float32)).cuda()
hr = torch.from_numpy(rng.standard_normal((B, T, S, S, 2), dtype=np.float32)).cuda()
for _ in range(2):
    gan.train_step((lr, hr))
torch.cuda.synchronize()
torch.cuda.profiler.start()
m = gan.train_step((lr, hr))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print({k: v for k, v in m.items() if v is not None})
