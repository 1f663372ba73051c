#!/usr/bin/env python
"""Runs the training graphs once warm, then ONE phase between cudaProfilerStart/Stop (ncu --profile-from-start off):
    python tools/phase_run.py {gfwd|gbwd|dfwd|dbwd_in|dbwd_w} [fp32|tf32|bf16]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wind_downscaling_gan_b200.train.nets import GenNet, CriticNet, to_device
from oracle.generator import synthetic_generator_weights
from oracle.critic import synthetic_critic_weights
from wind_downscaling_gan_b200.train import ops
phase = sys.argv[1]
ops.set_precision(sys.argv[2] if len(sys.argv) > 2 else "tf32")
B, T, S = 8, 24, 96
gw, dw = to_device(synthetic_generator_weights(0)), to_device(synthetic_critic_weights(1, size=S))
g = torch.Generator(device="cuda").manual_seed(0)
lr = torch.randn((B, T, S, S, 3), device="cuda", generator=g)
hr = torch.randn((B, T, S, S, 2), device="cuda", generator=g)
nz = 0.1 * torch.randn((B, T, S, S, 20), device="cuda", generator=g)
ones = torch.ones((B, 1), device="cuda")
gen, d = GenNet(gw), CriticNet(dw, S)
def run(ph, prof):
    if ph.startswith("g"):
        if ph == "gfwd" and prof: torch.cuda.profiler.start()
        fake = gen.forward(lr, nz, True)
        if ph == "gfwd" and prof: torch.cuda.synchronize(); torch.cuda.profiler.stop()
        if ph == "gbwd":
            if prof: torch.cuda.profiler.start()
            gen.backward(torch.ones_like(fake))
            if prof: torch.cuda.synchronize(); torch.cuda.profiler.stop()
    else:
        if ph == "dfwd" and prof: torch.cuda.profiler.start()
        d.forward(lr, hr, True)
        if ph == "dfwd" and prof: torch.cuda.synchronize(); torch.cuda.profiler.stop()
        if ph.startswith("dbwd"):
            if prof: torch.cuda.profiler.start()
            d.backward(ones, ph == "dbwd_w", ph == "dbwd_in")
            if prof: torch.cuda.synchronize(); torch.cuda.profiler.stop()
run(phase, False); torch.cuda.synchronize()
run(phase, True); torch.cuda.synchronize()
