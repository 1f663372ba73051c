"""`GAN` with the reference's constructor / compile / call / weight-I/O surface (`gan/ganbase.py`).

The inference path (`.generator`, `.noise_generator`, `call`, `save_weights`, `load_weights`) is what
`api.predict` uses.  `train_step` / `test_step` need the critic and the backward kernels, which are
the next rows of SURVEY.md §8 (A14-A16) and raise NotImplementedError until they are built: there is
deliberately no PyTorch-autograd fallback.
"""
import os
from pathlib import Path


class GAN:
    def __init__(self, generator, discriminator, noise_generator, n_critic=3, reconstruction_loss=None, *args, **kwargs):
        self.generator = generator
        self.discriminator = discriminator
        self.noise_generator = noise_generator
        self.reconstruction_loss = reconstruction_loss
        self._n_critic = n_critic
        self.compiled_metrics = None
        self.metrics = []

    def compile(self, generator_optimizer, discriminator_optimizer, generator_loss=None, generator_metrics=None,
                discriminator_loss=None, **kwargs):
        self.metrics = list(kwargs.get("metrics") or [])
        self.generator.compile(generator_optimizer, generator_loss, metrics=generator_metrics)
        if self.discriminator is not None:
            self.discriminator.compile(discriminator_optimizer, discriminator_loss)

    def call(self, inputs, training=None, mask=None):
        """ganbase.py:126-130: draw noise for the batch and run the generator."""
        low_res = inputs[0] if isinstance(inputs, (tuple, list)) else inputs
        noise = self.noise_generator(int(low_res.shape[0]))
        return self.generator([low_res, noise], training=bool(training))

    __call__ = call

    def train_step(self, data):
        raise NotImplementedError("WGAN train_step (ganbase.py:21-94) needs the critic + backward kernels (SURVEY §8 A14-A16)")

    def test_step(self, data):
        raise NotImplementedError("test_step (ganbase.py:96-113) needs the critic forward (SURVEY §8 A14)")

    def save_weights(self, filepath, *args, **kwargs):
        self.generator.save_weights(os.path.join(filepath, 'generator'), *args, **kwargs)
        if self.discriminator is not None:
            self.discriminator.save_weights(os.path.join(filepath, 'discriminator'), *args, **kwargs)

    def load_weights(self, filepath, *args, **kwargs):
        self.generator.load_weights(Path(filepath) / 'generator', *args, **kwargs)
        if self.discriminator is not None:
            self.discriminator.load_weights(Path(filepath) / 'discriminator', *args, **kwargs)
