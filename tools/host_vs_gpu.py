"""Host launch time vs GPU time of the training-mode generator forward (B=8, T=24, 96x96)."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wind_downscaling_gan_b200.train import ops
from wind_downscaling_gan_b200.train.nets import GenNet, CriticNet, to_device
from oracle.generator import synthetic_generator_weights
from oracle.critic import synthetic_critic_weights
ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32")
B, T, S = 8, 24, 96
gw = to_device(synthetic_generator_weights(0))
g = torch.Generator(device="cuda").manual_seed(0)
lr = torch.randn((B, T, S, S, 3), device="cuda", generator=g)
nz = 0.1 * torch.randn((B, T, S, S, 20), device="cuda", generator=g)
gen = GenNet(gw)
for training in (True, False):
    for i in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        gen.forward(lr, nz, training)
        e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"training={training} iter {i}: host issue {1e3*(t1-t0):7.2f} ms, wall {1e3*(t2-t0):7.2f} ms, gpu events {e0.elapsed_time(e1):7.2f} ms", flush=True)
pr = cProfile.Profile(); pr.enable(); gen.forward(lr, nz, True); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
