"""BASELINE configs[1] batch sweep: generator forward at B in {1..64}, T = 8, device-resident inputs."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wind_downscaling_gan_b200.gan.models import make_generator
from oracle.generator import synthetic_generator_weights
T, S = 8, 96
gen = make_generator(S, 3, 20, 2, T)
gen.set_weights(synthetic_generator_weights(0))
rows = []
for B in (1, 2, 4, 8, 16, 32, 64):
    g = torch.Generator(device="cuda").manual_seed(B)
    image = torch.randn((B, T, S, S, 3), device="cuda", generator=g)
    noise = 0.1 * torch.randn((B, T, S, S, 20), device="cuda", generator=g)
    out = torch.empty((B, T, S, S, 2), device="cuda")
    for _ in range(5): gen.forward_device(image, noise, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): gen.forward_device(image, noise, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    rows.append({"B": B, "ms_per_forward": ms, "fields_per_sec": B * T / ms * 1e3})
    print(rows[-1], flush=True)
json.dump(rows, open("gpurun_out/batch_sweep.json", "w"), indent=1)
