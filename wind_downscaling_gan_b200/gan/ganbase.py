"""`GAN` with the reference's constructor / compile / call / train_step / test_step / weight-I/O surface
(`gan/ganbase.py`).  Inference (`call`, `.generator.predict`) runs the bf16 tcgen05 path; `train_step` /
`test_step` run the training kernels (train/nets.py: fp32 tensors; convolution GEMMs in fp32 on CUDA cores by default,
tf32 / bf16 on tcgen05 with `compile(..., train_precision="tf32")` or `train.ops.set_precision`).  There is no
autograd / PyTorch math fallback.
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
        self._train = None   # lazily built device-side training state
        self._graphed = None
        self.use_cuda_graph = os.environ.get("WDG_TRAIN_GRAPH", "1") != "0"
        # extension, off by default: skip the gradient-penalty passes whose result nothing uses (train/step.py)
        self.skip_dead_gradient_penalty = bool(kwargs.get("skip_dead_gradient_penalty", False))
        # One set of variables, as in Keras: the trained weights live in the device-side TrainState; the model handles
        # pull them in lazily before anything reads them, and writing weights into a handle drops the TrainState.
        for m in (generator, discriminator):
            if m is not None and hasattr(m, "_before_read"):
                m._before_read, m._after_write = self.sync_weights, self._invalidate_state

    def compile(self, generator_optimizer, discriminator_optimizer, generator_loss=None, generator_metrics=None,
                discriminator_loss=None, **kwargs):
        self.metrics = list(kwargs.get("metrics") or [])
        if kwargs.get("train_precision"):          # extension: arithmetic of the training convolution GEMMs
            from ..train import ops as _train_ops
            _train_ops.set_precision(kwargs["train_precision"])
        self.generator.compile(generator_optimizer, generator_loss, metrics=generator_metrics)
        if self.discriminator is not None:
            self.discriminator.compile(discriminator_optimizer, discriminator_loss)

    def call(self, inputs, training=None, mask=None):
        """ganbase.py:126-130: draw noise for the batch and run the generator."""
        low_res = inputs[0] if isinstance(inputs, (tuple, list)) else inputs
        noise = self.noise_generator(int(low_res.shape[0]))
        return self.generator([low_res, noise], training=bool(training))

    __call__ = call

    # ------------------------------------------------------------------ training
    def _state(self):
        if self._train is None:
            from ..train.step import TrainState
            if self.reconstruction_loss is not None:
                raise NotImplementedError("reconstruction_loss (autoencoder features) is outside the built path")
            self._train = TrainState(self.generator, self.discriminator, self.generator.optimizer, self.discriminator.optimizer)
            self._graphed = None
        return self._train

    def _invalidate_state(self):
        """A handle's weights were written (set_weights / load_weights / training-mode call): the next train_step starts
        from them, with fresh Adam slots and the optimizers' step counters reset -- a reload restarts the optimizer."""
        if self._train is not None:
            self._train = None
            self._graphed = None
            for m in (self.generator, self.discriminator):
                if m is not None and getattr(m, "optimizer", None) is not None and hasattr(m.optimizer, "iterations"):
                    m.optimizer.iterations = 0

    def train_step(self, data, draws=None, comm=None):
        """ganbase.py:21-94.  data = (low_res, high_res[, sample_weight]); `draws` optionally replaces the random
        tensors (parity tests): per critic iteration [G noise, eps (B,), noise on real, noise on fake], then
        G noise for the generator update and for the metric recompute.  `comm` (train/dist.py Comm): data-parallel
        training, `data` being this rank's shard of the global batch."""
        from .. import _lib
        from ..train.step import GraphedStep, train_step
        st = self._state()
        before = _lib.calls
        if draws is None and self.use_cuda_graph:
            # steady state: the whole step is ONE captured CUDA graph (train/step.py: GraphedStep); the first calls run
            # eagerly and size every buffer
            if self._graphed is None or self._graphed.comm is not comm:
                self._graphed = GraphedStep(st, self.noise_generator, self._n_critic, comm, self.skip_dead_gradient_penalty)
            out = self._graphed(data[0], data[1])
            if self._graphed.graph is None or self._graphed.calls == GraphedStep.WARMUP + 1:
                self._last_step_calls = _lib.calls - before      # launches of one step, counted while it ran eagerly / was captured
        else:
            out = train_step(st, data[0], data[1], self.noise_generator, self._n_critic, draws, comm=comm,
                             skip_dead_gp=self.skip_dead_gradient_penalty)
            self._last_step_calls = _lib.calls - before
        return out

    def launches_per_step(self):
        """Library entry points the last train_step called (each launches at least one kernel): bench.py's gpu_launches."""
        return getattr(self, "_last_step_calls", None)

    def test_step(self, data, draws=None):
        """ganbase.py:96-113."""
        from ..train.step import test_step
        return test_step(self._state(), data[0], data[1], self.noise_generator, draws)

    def sync_weights(self):
        """Copies the trained fp32 device weights back into the model handles.  Called automatically (lazily) before a
        handle is read -- predict / call / get_weights / save_weights -- so trained weights are what inference runs."""
        if self._train is not None and self._train.dirty:
            self._train.push_weights()

    def save_weights(self, filepath, *args, **kwargs):
        self.sync_weights()
        self.generator.save_weights(os.path.join(filepath, 'generator'), *args, **kwargs)
        if self.discriminator is not None:
            self.discriminator.save_weights(os.path.join(filepath, 'discriminator'), *args, **kwargs)

    def load_weights(self, filepath, *args, **kwargs):
        self.generator.load_weights(Path(filepath) / 'generator', *args, **kwargs)
        if self.discriminator is not None:
            self.discriminator.load_weights(Path(filepath) / 'discriminator', *args, **kwargs)
        self._invalidate_state()
