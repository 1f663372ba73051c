#!/usr/bin/env python
"""Per-tensor gradient parity of the training graphs against the float64 autograd oracle at a chosen shape.
    python tools/diag_train_parity.py [B] [T] [S] [fp32|tf32|bf16]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rel(a, b):
    a = np.asarray(a.detach().cpu().numpy() if hasattr(a, "detach") else a, np.float64)
    b = np.asarray(b.detach().cpu().numpy() if hasattr(b, "detach") else b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def main():
    import torch
    from oracle import torch_train as tt
    from oracle.critic import synthetic_critic_weights
    from oracle.generator import synthetic_generator_weights
    from wind_downscaling_gan_b200.train import ops
    from wind_downscaling_gan_b200.train.nets import CriticNet, GenNet, to_device
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 96
    prec = sys.argv[4] if len(sys.argv) > 4 else "fp32"
    torch.set_num_threads(os.cpu_count() or 1)
    rng = np.random.default_rng(2)
    lr = rng.standard_normal((B, T, S, S, 3)).astype(np.float32)
    hr = rng.standard_normal((B, T, S, S, 2)).astype(np.float32)
    noise = (0.1 * rng.standard_normal((B, T, S, S, 20))).astype(np.float32)
    gw, dw_ = synthetic_generator_weights(3), synthetic_critic_weights(5, size=S)
    dout = rng.standard_normal((B, T, S, S, 2)).astype(np.float32)
    ds = rng.standard_normal((B, 1)).astype(np.float32)
    t0 = time.time()
    ref_w = {k: tt.T(v).clone() for k, v in gw.items()}
    out_ref, reads = tt.generator(ref_w, tt.T(lr), tt.T(noise), training=True)
    names = tt.trainable(ref_w)
    gg_ref = dict(zip(names, torch.autograd.grad((out_ref * tt.T(dout)).sum(), [reads[n] for n in names])))
    ref_d = {k: tt.T(v).clone() for k, v in dw_.items()}
    hr_t = tt.T(hr).requires_grad_(True)
    s_ref, dreads = tt.critic(ref_d, tt.T(lr), hr_t, training=True)
    dnames = tt.trainable(ref_d)
    dgr = torch.autograd.grad((s_ref * tt.T(ds)).sum(), [dreads[n] for n in dnames] + [hr_t])
    dg_ref = dict(zip(dnames, dgr[:-1]))
    print(f"oracle {time.time() - t0:.1f}s", flush=True)
    ops.set_precision(prec)
    ops.use_current_stream()
    net = GenNet(to_device(gw))
    out = net.forward(torch.from_numpy(lr).cuda(), torch.from_numpy(noise).cuda(), training=True)
    grads = net.backward(torch.from_numpy(dout).cuda())
    cnet = CriticNet(to_device(dw_), S)
    s = cnet.forward(torch.from_numpy(lr).cuda(), torch.from_numpy(hr).cuda(), training=True)
    g, dhr = cnet.backward(torch.from_numpy(ds).cuda(), need_weight_grads=True, need_input_grad=True)
    torch.cuda.synchronize()
    print(f"B={B} T={T} S={S} {prec}: G out {rel(out, out_ref):.2e}  D score {rel(s, s_ref):.2e}  D dhr {rel(dhr, dgr[-1]):.2e}")
    for n in names:
        print(f"  G {n:48s} {rel(grads[n], gg_ref[n]):.2e}")
    for n in dnames:
        print(f"  D {n:48s} {rel(g[n], dg_ref[n]):.2e}")


if __name__ == "__main__":
    main()
