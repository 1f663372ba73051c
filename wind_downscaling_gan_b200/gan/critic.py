"""`make_discriminator` product object (reference `gan/models.py:76-142`): weight bookkeeping with the checkpoint's
variable names and a forward pass on the fp32 CUDA training kernels (train/nets.py)."""
import os

import numpy as np

from .models import _glorot_uniform, _orthogonal

LW = "layer_with_weights-%d/"


def critic_weight_shapes(size, lr_ch, hr_ch, F, ckpt_topology=False):
    """Variable name -> shape of the graph `make_discriminator` builds, read from the `wdg_critic` handle's own table
    (csrc/wdg_critic.cu walks models.py:93-140).  ckpt_topology: the graph the shipped discriminator checkpoint was
    written from (shortcut branch of tf_utils.py:15-32 around the last 7x7 stage)."""
    from ..train.nets import CriticHandle
    return CriticHandle.get(size, lr_ch, hr_ch, F, ckpt_topology).shapes()


class Critic:
    """Drop-in for the Keras critic `Model` of models.py:142."""

    name = "discriminator"

    def __init__(self, low_res_size, high_res_size, low_res_channels, high_res_channels, n_timesteps, batch_size=None,
                 feature_channels=16, seed=None, ckpt_topology=False):
        if low_res_size != high_res_size:
            raise NotImplementedError("The discriminator assumes that the low res and high res images have the same size."
                                      "Perhaps you should upsample your low res image first?")
        self.image_size, self.low_res_channels, self.high_res_channels = high_res_size, low_res_channels, high_res_channels
        self.n_timesteps, self.feature_channels = n_timesteps, feature_channels
        self.input_names = ["low_resolution_image", "high_resolution_image"]
        self.output_names = ["score"]
        self.optimizer = self.compiled_loss = self.compiled_metrics = None
        self.metrics = []
        self._dev = None
        self._before_read = None     # hooks of an owning GAN, see gan/models.py
        self._after_write = None
        self._seed = seed
        self._set_topology(ckpt_topology)

    def _handle(self):
        from ..train.nets import CriticHandle
        return CriticHandle.get(self.image_size, self.low_res_channels, self.high_res_channels, self.feature_channels,
                                self.ckpt_topology)

    def _set_topology(self, ckpt_topology):
        """(Re)builds the variable table and the Keras-default initial values for one of the two graph revisions."""
        self.ckpt_topology = bool(ckpt_topology)
        self._shapes = critic_weight_shapes(self.image_size, self.low_res_channels, self.high_res_channels,
                                            self.feature_channels, self.ckpt_topology)
        self._dev = None
        rng = np.random.default_rng(self._seed)
        w = {}
        for name, shp in self._shapes.items():
            leaf = name.rsplit("/", 1)[1]
            if leaf == "recurrent_kernel":
                a = _orthogonal(rng, shp)
            elif leaf in ("w", "kernel"):
                a = _glorot_uniform(rng, shp)
            elif leaf == "bias":
                a = np.zeros(shp, np.float32)
                if "cell" in name:
                    F = shp[0] // 4
                    a[F:2 * F] = 1.0
            elif leaf == "sn_u":
                a = np.clip(rng.standard_normal(shp) * 0.02, -0.04, 0.04).astype(np.float32)
            elif leaf == "gamma":
                a = np.ones(shp, np.float32)
            else:
                a = np.zeros(shp, np.float32)
            w[name] = a
        self._w = w

    def weight_names(self):
        return list(self._shapes)

    @property
    def trainable_weights(self):
        return [n for n in self._shapes if not n.endswith("sn_u")]

    def set_weights(self, weights, _from_state=False):
        if not _from_state:
            if self._before_read is not None:
                self._before_read()
            if self._after_write is not None:
                self._after_write()
        for name, arr in weights.items():
            if name not in self._shapes:
                raise KeyError(name)
            a = np.ascontiguousarray(np.asarray(arr, np.float32))
            if tuple(a.shape) != tuple(self._shapes[name]):
                raise ValueError(f"{name}: expected shape {self._shapes[name]}, got {tuple(a.shape)}")
            self._w[name] = a
        self._dev = None

    def get_weights(self):
        if self._before_read is not None:
            self._before_read()
        return {k: v.copy() for k, v in self._w.items()}

    def save_weights(self, filepath, *args, **kwargs):
        """ganbase.py:134 (`<dir>/discriminator`): TensorFlow checkpoint-V2 bundle, see Generator.save_weights."""
        from ..tf_checkpoint import write_bundle
        filepath = str(filepath)
        os.makedirs(os.path.dirname(filepath) or ".", exist_ok=True)
        write_bundle(filepath, self.get_weights())

    def load_weights(self, filepath, *args, **kwargs):
        filepath = str(filepath)
        if os.path.exists(filepath + ".index"):
            from ..tf_checkpoint import read_bundle, select_model_variables
            tensors = read_bundle(filepath)
            try:
                sel = select_model_variables(tensors, self._shapes, "discriminator")
            except ValueError:
                # the checkpoint was written by the other revision of make_discriminator (the reference ships a critic
                # checkpoint WITH the shortcut branch its current code never builds): adopt that graph
                other = critic_weight_shapes(self.image_size, self.low_res_channels, self.high_res_channels,
                                             self.feature_channels, not self.ckpt_topology)
                sel = select_model_variables(tensors, other, "discriminator")      # raises if neither graph matches
                print(f"  discriminator checkpoint has the {'shortcut' if not self.ckpt_topology else 'current-code'} "
                      "topology: rebuilding the critic with it")
                self._set_topology(not self.ckpt_topology)
                if self._after_write is not None:
                    self._after_write()
            self.set_weights(sel)
            return
        if os.path.exists(filepath + ".npz"):
            with np.load(filepath + ".npz") as z:
                self.set_weights({k: z[k] for k in z.files})
            return
        raise FileNotFoundError(filepath)

    def compile(self, optimizer=None, loss=None, metrics=None, **kwargs):
        self.optimizer, self.compiled_loss, self.compiled_metrics = optimizer, loss, metrics
        self.metrics = list(metrics or [])

    def __call__(self, inputs, training=False, mask=None):
        """`discriminator([low_res, high_res], training=False)` -> score (B, 1) CUDA tensor."""
        from ..train.nets import CriticNet
        from ..train.step import _dev
        from ..train import ops as _ops
        _ops.use_current_stream()
        if self._before_read is not None:
            self._before_read()
        if self._dev is None:
            self._dev = self._handle().pack(self._w)
        low_res, high_res = _dev(inputs[0]), _dev(inputs[1])
        if tuple(low_res.shape[:-1]) != tuple(high_res.shape[:-1]):
            raise ValueError("low_resolution_image and high_resolution_image must share (B, T, H, W)")
        score = CriticNet(self._dev, self.image_size).forward(low_res, high_res, training=bool(training))
        if training:   # the spectral-norm power iteration mutated w / sn_u in place, as the Keras wrapper does
            self._w = {k: v.cpu().numpy() for k, v in self._dev.items()}
            if self._after_write is not None:
                self._after_write()
        return score

    call = __call__

    def predict(self, inputs, **kwargs):
        return self(inputs, training=False).cpu().numpy()
