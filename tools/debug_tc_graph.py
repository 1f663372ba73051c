"""Per-weight gradient deviations of the tcgen05 training graphs vs the same-rounding oracle (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import torch_train as tt
from oracle.critic import synthetic_critic_weights
from oracle.generator import synthetic_generator_weights
from wind_downscaling_gan_b200.train import ops
from wind_downscaling_gan_b200.train.nets import CriticNet, GenNet, to_device

def rel(a, b):
    a = np.asarray(a.detach().cpu().numpy() if hasattr(a, "detach") else a, np.float64)
    b = np.asarray(b.detach().cpu().numpy() if hasattr(b, "detach") else b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))

B, T, S = 2, 2, 32
rng = np.random.default_rng(2)
lr = rng.standard_normal((B, T, S, S, 3)).astype(np.float32)
hr = rng.standard_normal((B, T, S, S, 2)).astype(np.float32)
noise = (0.1 * rng.standard_normal((B, T, S, S, 20))).astype(np.float32)
gw = synthetic_generator_weights(3)
dout = rng.standard_normal((B, T, S, S, 2))
for prec in sys.argv[1:] or ["fp32", "tf32"]:
    tt.OPERAND = None if prec == "fp32" else prec
    ref_w = {k: tt.T(v).clone() for k, v in gw.items()}
    out_ref, reads = tt.generator(ref_w, tt.T(lr), tt.T(noise), training=True)
    names = tt.trainable(ref_w)
    grads_ref = dict(zip(names, torch.autograd.grad((out_ref * tt.T(dout)).sum(), [reads[n] for n in names])))
    tt.OPERAND = None
    ops.set_precision(prec)
    w = to_device(gw)
    net = GenNet(w)
    out = net.forward(torch.from_numpy(lr).cuda(), torch.from_numpy(noise).cuda(), training=True)
    grads = net.backward(torch.from_numpy(dout.astype(np.float32)).cuda())
    ops.set_precision("fp32")
    print(prec, "G out", f"{rel(out, out_ref):.2e}")
    for k in ("layer_with_weights-0/layer/w", "layer_with_weights-1/moving_mean", "layer_with_weights-10/moving_variance",
              "layer_with_weights-3/moving_variance", "layer_with_weights-6/moving_mean", "layer_with_weights-8/moving_mean"):
        print("   state", k, f"{rel(w[k], ref_w[k]):.2e}")
    for n in names:
        print(f"   {rel(grads[n], grads_ref[n]):.2e}  |ref|={float(grads_ref[n].norm()):.3e}  {n} {tuple(grads_ref[n].shape)}")
