"""Critic forward + backward (weights and input) once at B=8, T=24, 96x96 (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wind_downscaling_gan_b200.train import ops
from wind_downscaling_gan_b200.train.nets import CriticNet, to_device
from oracle.critic import synthetic_critic_weights
ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32")
B, T, S = 8, 24, 96
dw = to_device(synthetic_critic_weights(1, size=S))
g = torch.Generator(device="cuda").manual_seed(0)
lr = torch.randn((B, T, S, S, 3), device="cuda", generator=g)
hr = torch.randn((B, T, S, S, 2), device="cuda", generator=g)
d = CriticNet(dw, S)
for _ in range(2):
    d.forward(lr, hr, True)
    d.backward(torch.ones((B, 1), device="cuda"), True, True)
torch.cuda.synchronize()
