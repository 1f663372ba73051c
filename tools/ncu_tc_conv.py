"""One forward / backward-data / backward-weight launch of a training conv at a given layer shape (for ncu).
Usage: python tools/ncu_tc_conv.py PREC N H W Ci k Co s p"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wind_downscaling_gan_b200.train import ops
prec = sys.argv[1]
N, H, W, Ci, k, Co, s, p = map(int, sys.argv[2:10])
Ho, Wo = ops.conv_out(H, k, s, p, p), ops.conv_out(W, k, s, p, p)
x = torch.randn((N, H, W, Ci), device="cuda"); w = torch.randn((k, k, Ci, Co), device="cuda") * 0.05
b = torch.randn((Co,), device="cuda"); dy = torch.randn((N, Ho, Wo, Co), device="cuda")
y, dx, dw = ops.empty(N, Ho, Wo, Co), ops.empty(N, H, W, Ci), ops.empty(k, k, Ci, Co)
ops.set_precision(prec)
for _ in range(2):
    ops.conv2d_fwd(ops.full(x), w, b, ops.full(y), N, H, W, s, p, Ho, Wo)
    ops.conv2d_bwd_data(ops.full(dy), w, ops.full(dx), N, H, W, s, p, Ho, Wo)
    ops.conv2d_bwd_weight(ops.full(x), ops.full(dy), dw, N, H, W, s, p, Ho, Wo)
torch.cuda.synchronize()
