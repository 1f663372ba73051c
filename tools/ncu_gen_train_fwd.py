import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wind_downscaling_gan_b200.train.nets import GenNet, CriticNet, to_device
from oracle.generator import synthetic_generator_weights
from oracle.critic import synthetic_critic_weights
from wind_downscaling_gan_b200.train import ops
ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "fp32")
B, T, S = 8, 24, 96
gw, dw = to_device(synthetic_generator_weights(0)), to_device(synthetic_critic_weights(1, size=S))
g = torch.Generator(device="cuda").manual_seed(0)
lr = torch.randn((B, T, S, S, 3), device="cuda", generator=g)
hr = torch.randn((B, T, S, S, 2), device="cuda", generator=g)
nz = 0.1 * torch.randn((B, T, S, S, 20), device="cuda", generator=g)
gen = GenNet(gw)
fake = gen.forward(lr, nz, True)
gen.backward(torch.ones_like(fake))
d = CriticNet(dw, S)
d.forward(lr, hr, True)
d.backward(torch.ones((B, 1), device="cuda"), True, True)
torch.cuda.synchronize()
