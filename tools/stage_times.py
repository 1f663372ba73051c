#!/usr/bin/env python
"""Per-stage CUDA-event times of the generator forward at the bench shape: python tools/stage_times.py [bf16|tf32] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.generator import synthetic_generator_weights
from wind_downscaling_gan_b200.gan.models import make_generator
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
B, T = 64, 8
g = torch.Generator(device="cuda").manual_seed(0)
image = torch.randn((B, T, 96, 96, 3), device="cuda", generator=g)
noise = 0.1 * torch.randn((B, T, 96, 96, 20), device="cuda", generator=g)
gen = make_generator(96, 3, 20, 2, T).set_precision(prec)
gen.set_weights(synthetic_generator_weights(0))
out = torch.empty((B, T, 96, 96, 2), device="cuda")
for _ in range(5):
    gen.forward_device(image, noise, out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(steps):
    gen.forward_device(image, noise, out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
gen.set_profiling(True)
acc = {}
for _ in range(steps):
    gen.forward_device(image, noise, out)
    for k, v in gen.stage_ms().items():
        acc[k] = acc.get(k, 0) + v / steps
print(prec, os.environ.get("WDG_SHALLOW"), os.environ.get("WDG_HALO11"), "ms/step %.4f  fields/s %.0f " % (ms, B * T / ms * 1e3),
      " ".join(f"{k}={v:.3f}" for k, v in acc.items()), "chk %.6f" % float(out.abs().mean()))
