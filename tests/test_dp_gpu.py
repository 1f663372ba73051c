"""Data-parallel training equivalence (needs >= 2 GPUs, skipped otherwise): two ranks with half the batch each,
NCCL gradient all-reduce and synchronised BatchNorm, must reproduce the single-process step on the full batch."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(S, T, Bg):
    from oracle.critic import synthetic_critic_weights
    from oracle.generator import synthetic_generator_weights
    rng = np.random.default_rng(11)
    lr = rng.standard_normal((Bg, T, S, S, 3)).astype(np.float32)
    hr = rng.standard_normal((Bg, T, S, S, 2)).astype(np.float32)
    draws = []
    for _ in range(3):
        draws += [0.1 * rng.standard_normal((Bg, T, S, S, 20)), rng.uniform(0, 1, (Bg,)),
                  0.1 * rng.standard_normal((Bg, T, S, S, 2)), 0.1 * rng.standard_normal((Bg, T, S, S, 2))]
    draws += [0.1 * rng.standard_normal((Bg, T, S, S, 20)), 0.1 * rng.standard_normal((Bg, T, S, S, 20))]
    return synthetic_generator_weights(12), synthetic_critic_weights(13, size=S), lr, hr, [np.asarray(d, np.float32) for d in draws]


def _make_gan(S, T, B, gw, dw):
    from wind_downscaling_gan_b200.data.data_generator import FlexibleNoiseGenerator
    from wind_downscaling_gan_b200.gan import train
    from wind_downscaling_gan_b200.gan.ganbase import GAN
    from wind_downscaling_gan_b200.gan.models import make_discriminator, make_generator
    gen, disc = make_generator(S, 3, 20, 2, T), make_discriminator(S, S, 3, 2, T)
    gen.set_weights(gw)
    disc.set_weights(dw)
    gan = GAN(gen, disc, FlexibleNoiseGenerator((B, T, S, S, 20), std=0.1))
    gan.compile(generator_optimizer=train.generator_optimizer(), discriminator_optimizer=train.discriminator_optimizer(),
                discriminator_loss=train.discriminator_loss)
    return gan


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from wind_downscaling_gan_b200.train.dist import Comm, shard_batch
    S, T, Bg = 32, 2, 4
    gw, dw, lr, hr, draws = _setup(S, T, Bg)
    a, b = shard_batch(Bg, rank, world)
    gan = _make_gan(S, T, b - a, gw, dw)
    gan.train_step((lr[a:b], hr[a:b]), draws=[d[a:b] for d in draws], comm=Comm())
    gan.sync_weights()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **{"g/" + k: v for k, v in gan.generator.get_weights().items()},
             **{"d/" + k: v for k, v in gan.discriminator.get_weights().items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    S, T, Bg = 32, 2, 4
    gw, dw, lr, hr, draws = _setup(S, T, Bg)
    gan = _make_gan(S, T, Bg, gw, dw)
    gan.train_step((lr, hr), draws=draws)
    gan.sync_weights()
    ref = {**{"g/" + k: v for k, v in gan.generator.get_weights().items()},
           **{"d/" + k: v for k, v in gan.discriminator.get_weights().items()}}
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for k, v in ref.items():
        assert np.array_equal(r0[k], r1[k]), f"replicas diverged: {k}"
        err = np.linalg.norm(r0[k].astype(np.float64) - v) / max(np.linalg.norm(v), 1e-30)
        assert err < 1e-4, (k, err)
