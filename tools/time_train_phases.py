"""Phase timing of the fp32 training path at the api.py constants (B=8, T=24, 96x96)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wind_downscaling_gan_b200.train.nets import GenNet, CriticNet, to_device
from oracle.generator import synthetic_generator_weights
from oracle.critic import synthetic_critic_weights

from wind_downscaling_gan_b200.train import ops
ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "fp32")
print("precision", ops.get_precision())
B, T, S = 8, 24, 96
gw, dw = to_device(synthetic_generator_weights(0)), to_device(synthetic_critic_weights(1, size=S))
g = torch.Generator(device="cuda").manual_seed(0)
lr = torch.randn((B, T, S, S, 3), device="cuda", generator=g)
hr = torch.randn((B, T, S, S, 2), device="cuda", generator=g)
nz = 0.1 * torch.randn((B, T, S, S, 20), device="cuda", generator=g)
ones = torch.ones((B, 1), device="cuda")

def timed(name, fn, n=2):
    fn(); torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(n): r = fn()
    torch.cuda.synchronize()
    print(f"{name:28s} {(time.time()-t0)/n*1e3:9.1f} ms", flush=True)
    return r

gen = GenNet(gw)
fake = timed("G forward (training)", lambda: gen.forward(lr, nz, True))
timed("G forward (inference)", lambda: gen.forward(lr, nz, False))
gen.forward(lr, nz, True)
timed("G backward", lambda: gen.backward(torch.ones_like(fake)), n=1)
d = CriticNet(dw, S)
timed("D forward (training)", lambda: d.forward(lr, hr, True))
# a context is consumed by its backward (wdg_critic_backward): time forward + backward pairs, subtract the forward above
timed("D forward + backward input only", lambda: (d.forward(lr, hr, True), d.backward(ones, False, True)), n=1)
timed("D forward + backward weights only", lambda: (d.forward(lr, hr, True), d.backward(ones, True, False)), n=1)
