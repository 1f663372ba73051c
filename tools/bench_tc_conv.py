"""Times the three convolution GEMMs (forward, backward-data, backward-weight) of the training path at the layer
shapes of one WGAN step (B=8, T=24 -> 192 images, ConvLSTM recurrent convs: 8 images) in fp32 (CUDA cores), tf32 and
bf16 (tcgen05).  Prints one JSON line per (layer, precision) with TFLOP/s of each GEMM.  Usage:
    python tools/bench_tc_conv.py [--precisions fp32,tf32,bf16] [--reps 5]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

LAYERS = [
    # name, N, H, W, Ci, k, Co, s, p
    ("G0 conv8x8s2 23->128", 192, 96, 96, 23, 8, 128, 2, 3),
    ("G1 conv4x4s2 128->128", 192, 48, 48, 128, 4, 128, 2, 1),
    ("G lstm x-conv 128->512", 192, 24, 24, 128, 3, 512, 1, 1),
    ("G lstm h-conv 128->512 (1 step)", 8, 24, 24, 128, 3, 512, 1, 1),
    ("G5 conv3x3 128->64", 192, 24, 24, 128, 3, 64, 1, 1),
    ("G8 convT5x5 160->16 (as conv 16->160)", 192, 96, 96, 16, 5, 160, 1, 2),
    ("G9 conv3x3 16->2", 192, 96, 96, 16, 3, 2, 1, 1),
    ("D lstm2 x-conv 2->8", 192, 96, 96, 2, 3, 8, 1, 1),
    ("D conv 2->16", 192, 96, 96, 2, 3, 16, 1, 1),
    ("D lstm16 x-conv 5->64", 192, 96, 96, 5, 3, 64, 1, 1),
    ("D lstm16 h-conv 16->64 (1 step)", 8, 96, 96, 16, 3, 64, 1, 1),
    ("D conv 16->16", 192, 96, 96, 16, 3, 16, 1, 1),
    ("D pyr 7x7s3 32->64", 192, 96, 96, 32, 7, 64, 3, 1),
    ("D pyr 7x7s3 64->128", 192, 31, 31, 64, 7, 128, 3, 1),
    ("D pyr 7x7s3 128->256", 192, 9, 9, 128, 7, 256, 3, 1),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precisions", default="fp32,tf32,bf16")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch
    from wind_downscaling_gan_b200.train import ops
    torch.cuda.set_device(0)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.reps

    for name, N, H, W, Ci, k, Co, s, p in LAYERS:
        Ho, Wo = ops.conv_out(H, k, s, p, p), ops.conv_out(W, k, s, p, p)
        x = torch.randn((N, H, W, Ci), device="cuda")
        w = torch.randn((k, k, Ci, Co), device="cuda") * 0.05
        b = torch.randn((Co,), device="cuda")
        dy = torch.randn((N, Ho, Wo, Co), device="cuda")
        y, dx, dw = ops.empty(N, Ho, Wo, Co), ops.empty(N, H, W, Ci), ops.empty(k, k, Ci, Co)
        flop = 2.0 * N * Ho * Wo * k * k * Ci * Co
        for prec in a.precisions.split(","):
            ops.set_precision(prec)
            ops.use_current_stream()
            t_f = timed(lambda: ops.conv2d_fwd(ops.full(x), w, b, ops.full(y), N, H, W, s, p, Ho, Wo))
            t_d = timed(lambda: ops.conv2d_bwd_data(ops.full(dy), w, ops.full(dx), N, H, W, s, p, Ho, Wo))
            t_w = timed(lambda: ops.conv2d_bwd_weight(ops.full(x), ops.full(dy), dw, N, H, W, s, p, Ho, Wo))
            print(json.dumps({"layer": name, "prec": prec, "gflop": round(flop / 1e9, 2),
                              "fwd_ms": round(t_f, 4), "bwd_data_ms": round(t_d, 4), "bwd_weight_ms": round(t_w, 4),
                              "fwd_tflops": round(flop / t_f / 1e9, 1), "bwd_data_tflops": round(flop / t_d / 1e9, 1),
                              "bwd_weight_tflops": round(flop / t_w / 1e9, 1)}), flush=True)
        ops.set_precision("fp32")


if __name__ == "__main__":
    main()
