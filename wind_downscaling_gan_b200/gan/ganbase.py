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
        # extension: arithmetic of THIS model's training convolution GEMMs ("fp32" | "tf32" | "bf16").  The library
        # switch is process-wide, so the model re-asserts its own choice at the start of every train_step / test_step.
        self.train_precision = kwargs.get("train_precision") or getattr(self, "train_precision", None)
        self._assert_precision()
        self.generator.compile(generator_optimizer, generator_loss, metrics=generator_metrics)
        if self.discriminator is not None:
            self.discriminator.compile(discriminator_optimizer, discriminator_loss)

    def _assert_precision(self):
        if getattr(self, "train_precision", None):
            from ..train import ops as _train_ops
            if _train_ops.get_precision() != self.train_precision:
                _train_ops.set_precision(self.train_precision)
                self._graphed = None          # a captured step replays the arithmetic it was captured with

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
        training, `data` being this rank's shard of the global batch.

        Returns the reference's dict (:75-93): the seven logged scalars, then `result()` of every compiled metric object
        (`metrics=` of compile under its own name -- d_real / d_fake running means in api.py:84 -- and every generator
        metric as `g_<name>`), all of them updated first with this step's inference-mode recompute (:71-72)."""
        from .. import _lib
        from ..train.step import GraphedStep, train_step
        if len(data) > 2 and data[2] is not None:
            raise NotImplementedError("sample_weight is not supported by the CUDA train_step (the reference never passes one: "
                                      "its generators yield (low_res, high_res) pairs)")
        self._assert_precision()
        st = self._state()
        before = _lib.calls
        if draws is None and self.use_cuda_graph:
            # steady state: the whole step is ONE captured CUDA graph (train/step.py: GraphedStep); the first calls run
            # eagerly and size every buffer
            if self._graphed is None or self._graphed.comm is not comm:
                self._graphed = GraphedStep(st, self.noise_generator, self._n_critic, comm, self.skip_dead_gradient_penalty)
            out, extras = self._graphed(data[0], data[1])
            if self._graphed.graph is None or self._graphed.calls == GraphedStep.WARMUP + 1:
                self._last_step_calls = _lib.calls - before      # launches of one step, counted while it ran eagerly / was captured
        else:
            out, extras = train_step(st, data[0], data[1], self.noise_generator, self._n_critic, draws, comm=comm,
                                     skip_dead_gp=self.skip_dead_gradient_penalty)
            self._last_step_calls = _lib.calls - before
        self._last_recompute = extras      # (generated images, D(real), D(generated)) of the inference-mode recompute
        self._update_metrics(out, data[1], extras)
        return out

    def _update_metrics(self, out, high_res, extras):
        """ganbase.py:71-72, :82-93.  With a data-parallel `comm` each rank's metric objects see its own shard."""
        gen_metrics = list(getattr(self.generator, "metrics", None) or [])
        if not gen_metrics and not self.metrics:
            return
        fake, s_real, s_fake = extras
        for metric in gen_metrics:
            metric.update_state(high_res, fake)
        if self.metrics:
            s_real, s_fake = s_real.detach().cpu().numpy(), s_fake.detach().cpu().numpy()
            for metric in self.metrics:
                metric.update_state(s_real, s_fake)
        self._collect_metrics(out, gen_metrics)

    def _collect_metrics(self, out, gen_metrics=()):
        for metric in self.metrics:
            result = metric.result()
            if isinstance(result, dict):
                out.update(result)
            else:
                out[metric.name] = result
        for metric in gen_metrics:
            result = metric.result()
            if isinstance(result, dict):
                out.update(result)
            else:
                out[f'g_{metric.name}'] = result

    def launches_per_step(self):
        """Library entry points the last train_step called (each launches at least one kernel): bench.py's gpu_launches."""
        return getattr(self, "_last_step_calls", None)

    def test_step(self, data, draws=None):
        """ganbase.py:96-113."""
        from ..train.step import test_step
        self._assert_precision()
        out = test_step(self._state(), data[0], data[1], self.noise_generator, draws)
        self._collect_metrics(out)          # :105-111: results of the compiled metrics (test_step does not update them)
        return out

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
