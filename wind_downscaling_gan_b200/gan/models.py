"""`make_generator` / `make_discriminator` with the reference's signatures
(`/root/reference/src/downscaling/gan/models.py:9-17,76-84`), backed by libwdg.so.

The returned objects mirror the slice of the Keras `Model` surface that
`api.py` and `ganbase.py` use for this path: `.predict([image, noise])`,
`__call__([a, b], training=False)`, `.get_weights()/.set_weights()`,
`.save_weights(path)/.load_weights(path)`, `.name`, `.trainable_weights`.
PyTorch is used only to allocate device memory and streams.
"""
import ctypes as C
import os

import numpy as np

from .. import _lib

LW = "layer_with_weights-%d/"


def _glorot_uniform(rng, shape):
    # Keras VarianceScaling(fan_avg, uniform): fans from the last two axes times the receptive field
    rf = int(np.prod(shape[:-2]))
    fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, shape).astype(np.float32)


def _orthogonal(rng, shape):
    rows, cols = int(np.prod(shape[:-1])), shape[-1]
    a = rng.standard_normal((max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    if rows < cols:
        q = q.T
    return q[:rows, :cols].reshape(shape).astype(np.float32)


class _NativeModel:
    """Shared weight bookkeeping for the handles exported by libwdg.so."""

    def _weight_table(self):
        L = _lib.lib()
        n = L.wdg_generator_num_weights(self._h)
        table = []
        for i in range(n):
            name = C.c_char_p()
            dims = (C.c_int64 * 4)()
            nd = C.c_int()
            _lib.check(L.wdg_generator_weight_info(self._h, i, C.byref(name), dims, C.byref(nd)))
            table.append((name.value.decode(), tuple(int(dims[k]) for k in range(nd.value))))
        return table


class Generator(_NativeModel):
    """Drop-in for the Keras generator `Model` of models.py:73 on the inference path."""

    name = "generator"

    PRECISIONS = {"bf16": 0, "tf32": 1}

    def __init__(self, image_size, in_channels, noise_channels, out_channels, n_timesteps, batch_size=None,
                 feature_channels=128, seed=None, precision=None):
        assert image_size % 4 == 0          # models.py:19
        assert feature_channels % 8 == 0    # models.py:20
        self.image_size, self.in_channels, self.noise_channels = image_size, in_channels, noise_channels
        self.out_channels, self.n_timesteps, self.batch_size = out_channels, n_timesteps, batch_size
        self.feature_channels = feature_channels
        self.input_names = ["input_image", "input_noise"]   # models.py:24-25
        self.output_names = ["predicted_image"]             # models.py:71
        self.optimizer = None
        self.compiled_loss = None
        self.compiled_metrics = None
        self.metrics = []
        h = C.c_void_p()
        _lib.check(_lib.lib().wdg_generator_create(C.byref(h), image_size, in_channels, noise_channels, out_channels,
                                                   n_timesteps, feature_channels))
        self._h = h
        self._shapes = dict(self._weight_table())
        self._dirty = True
        self._plan = None      # (B, T)
        self._ws = None
        self._io = None
        # hooks of an owning GAN: trained weights live in its device-side TrainState and are pulled in lazily before
        # anything reads this handle; writing weights here invalidates that state (one set of variables, as in Keras)
        self._before_read = None
        self._after_write = None
        self.precision = "bf16"
        self.set_precision(precision or os.environ.get("WDG_PRECISION", "bf16"))
        self._init_weights(np.random.default_rng(seed))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and _lib._lib is not None:
            _lib._lib.wdg_generator_destroy(h)
            self._h = None

    def set_precision(self, precision):
        """Operand precision of the GEMM stages: "bf16" (kind::f16 MMAs, rel-L2 <= 1e-2 of the fp32 reference) or
        "tf32" (kind::tf32 MMAs on fp32 activations, output convolution in fp32: rel-L2 <= 1e-3).  Default "bf16",
        or the WDG_PRECISION environment variable."""
        if precision not in self.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(self.PRECISIONS)}, got {precision!r}")
        _lib.check(_lib.lib().wdg_generator_set_precision(self._h, self.PRECISIONS[precision]))
        if precision != self.precision:
            self._dirty = True
            self._plan = None
        self.precision = precision
        return self

    # ------------------------------------------------------------ weights
    def _init_weights(self, rng):
        """Keras default initialisers: glorot_uniform kernels, orthogonal recurrent kernel, zero biases
        with unit forget bias, BN (1, 0, 0, 1), SN u ~ TruncatedNormal(0.02)."""
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
            elif leaf in ("gamma", "moving_variance"):
                a = np.ones(shp, np.float32)
            else:
                a = np.zeros(shp, np.float32)
            w[name] = a
        self.set_weights(w)

    def weight_names(self):
        return list(self._shapes)

    @property
    def trainable_weights(self):
        return [n for n in self._shapes if not n.endswith(("moving_mean", "moving_variance", "sn_u"))]

    def set_weights(self, weights, _from_state=False):
        """weights: dict name -> array, names/layouts of weights-55.ckpt/generator.index."""
        if not _from_state:
            if self._before_read is not None:
                self._before_read()          # a partial update must land on top of the trained weights
            if self._after_write is not None:
                self._after_write()
        L = _lib.lib()
        for name, arr in weights.items():
            if name not in self._shapes:
                raise KeyError(name)
            a = np.ascontiguousarray(np.asarray(arr, np.float32))
            if tuple(a.shape) != self._shapes[name]:
                raise ValueError(f"{name}: expected shape {self._shapes[name]}, got {tuple(a.shape)}")
            dims = (C.c_int64 * 4)(*a.shape)
            _lib.check(L.wdg_generator_set_weight(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), dims, a.ndim))
        self._dirty = True

    def get_weights(self):
        if self._before_read is not None:
            self._before_read()
        L = _lib.lib()
        out = {}
        for name, shp in self._shapes.items():
            a = np.empty(shp, np.float32)
            _lib.check(L.wdg_generator_get_weight(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), a.size))
            out[name] = a
        return out

    def save_weights(self, filepath, *args, **kwargs):
        """Counterpart of ganbase.py:133 (`<dir>/generator`): a TensorFlow checkpoint-V2 bundle `<filepath>.index` +
        `<filepath>.data-00000-of-00001` keyed `<variable>/.ATTRIBUTES/VARIABLE_VALUE` like the reference's own
        checkpoints (tf_checkpoint.write_bundle: TensorFlow's header and crc32c checksums)."""
        from ..tf_checkpoint import write_bundle
        filepath = str(filepath)
        os.makedirs(os.path.dirname(filepath) or ".", exist_ok=True)
        write_bundle(filepath, self.get_weights())

    def load_weights(self, filepath, *args, **kwargs):
        """Counterpart of ganbase.py:139: reads the TF-checkpoint-V2 prefix `<filepath>.index` + data shard (the
        reference's format and what save_weights writes); `<filepath>.npz` (round-1 exports) is still accepted."""
        filepath = str(filepath)
        if os.path.exists(filepath + ".index"):
            from ..tf_checkpoint import read_bundle, select_model_variables
            self.set_weights(select_model_variables(read_bundle(filepath), self._shapes, "generator"))
            return
        if os.path.exists(filepath + ".npz"):
            with np.load(filepath + ".npz") as z:
                self.set_weights({k: z[k] for k in z.files})
            return
        raise FileNotFoundError(filepath)

    # ------------------------------------------------------------ forward
    def _ensure_plan(self, B, T, stream=None):
        import torch
        L = _lib.lib()
        if self._before_read is not None:
            self._before_read()
        if self._dirty:
            _lib.check(L.wdg_generator_finalize(self._h))
            self._dirty = False
            self._plan = None
        if self._plan != (B, T):
            nbytes = C.c_size_t()
            _lib.check(L.wdg_generator_workspace_bytes(self._h, B, T, C.byref(nbytes)))
            self._ws = None
            self._ws = torch.empty(nbytes.value + 1024, dtype=torch.uint8, device="cuda")
            base = (self._ws.data_ptr() + 1023) // 1024 * 1024
            s = torch.cuda.current_stream().cuda_stream if stream is None else stream
            _lib.check(L.wdg_generator_bind(self._h, B, T, C.c_void_p(base), nbytes.value, C.c_void_p(s)))
            _lib.check(L.wdg_generator_io_bytes(self._h, B, T, C.byref(nbytes)))
            self._io = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
            self._plan = (B, T)

    def launches_per_forward(self):
        return _lib.lib().wdg_generator_launches_per_forward(self._h)

    def forward_device(self, image, noise, out=None):
        """image (B,T,S,S,Cin), noise (B,T,S,S,Cn): contiguous fp32 CUDA tensors -> (B,T,S,S,Cout) CUDA tensor.
        Asynchronous on the current torch stream."""
        import torch
        assert image.is_cuda and noise.is_cuda and image.dtype == torch.float32 and noise.dtype == torch.float32
        image, noise = image.contiguous(), noise.contiguous()
        B, T, S = image.shape[:3]
        if tuple(image.shape[2:]) != (S, S, self.in_channels) or S != self.image_size:
            raise ValueError(f"input_image: expected (B,T,{self.image_size},{self.image_size},{self.in_channels}), got {tuple(image.shape)}")
        if tuple(noise.shape) != (B, T, S, S, self.noise_channels):
            raise ValueError(f"input_noise: expected {(B, T, S, S, self.noise_channels)}, got {tuple(noise.shape)}")
        self._ensure_plan(B, T)
        if out is None:
            out = torch.empty((B, T, S, S, self.out_channels), dtype=torch.float32, device=image.device)
        s = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().wdg_generator_forward(self._h, C.c_void_p(image.data_ptr()), C.c_void_p(noise.data_ptr()),
                                                    C.c_void_p(out.data_ptr()), C.c_void_p(s)))
        return out

    def forward_device_gen_noise(self, image, noise_generator, out=None):
        """`generator([image, noise_generator(bs, channels=Cn)])` with the noise drawn inside the input-packing kernel:
        same values as materialising the tensor from `noise_generator` first (it is advanced identically), but the
        (B,T,S,S,Cn) fp32 tensor is never written to or read from HBM.  image: contiguous fp32 CUDA tensor."""
        import torch
        assert image.is_cuda and image.dtype == torch.float32
        image = image.contiguous()
        B, T, S = image.shape[:3]
        if tuple(image.shape[2:]) != (S, S, self.in_channels) or S != self.image_size:
            raise ValueError(f"input_image: expected (B,T,{self.image_size},{self.image_size},{self.in_channels}), got {tuple(image.shape)}")
        self._ensure_plan(B, T)
        if out is None:
            out = torch.empty((B, T, S, S, self.out_channels), dtype=torch.float32, device=image.device)
        n = B * T * S * S * self.noise_channels
        std, seed, offset = noise_generator.reserve(n)
        scratch = None
        if (self.in_channels, self.noise_channels) != (3, 20):
            scratch = torch.empty(n, dtype=torch.float32, device=image.device)
        s = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().wdg_generator_forward_gen_noise(
            self._h, C.c_void_p(image.data_ptr()), std, C.c_uint64(seed), C.c_uint64(offset), C.c_void_p(out.data_ptr()),
            C.c_void_p(scratch.data_ptr()) if scratch is not None else None, C.c_void_p(s)))
        return out

    def predict_host(self, image, noise, out=None):
        """Host numpy (or CPU torch, ideally pinned) in, host numpy out: H2D copy, forward, D2H copy, sync --
        the call `gen.predict([tensor, noise])` of api.py:137 makes."""
        import torch
        img = image if isinstance(image, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(image, np.float32))
        noi = noise if isinstance(noise, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(noise, np.float32))
        img, noi = img.contiguous(), noi.contiguous()
        assert img.dtype == torch.float32 and noi.dtype == torch.float32 and not img.is_cuda
        B, T, S = img.shape[:3]
        if tuple(img.shape[2:]) != (S, S, self.in_channels) or S != self.image_size:
            raise ValueError(f"input_image: bad shape {tuple(img.shape)}")
        if tuple(noi.shape) != (B, T, S, S, self.noise_channels):
            raise ValueError(f"input_noise: bad shape {tuple(noi.shape)}")
        self._ensure_plan(B, T)
        if out is None:
            out = torch.empty((B, T, S, S, self.out_channels), dtype=torch.float32)
        s = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().wdg_generator_predict_host(self._h, C.c_void_p(img.data_ptr()), C.c_void_p(noi.data_ptr()),
                                                         C.c_void_p(out.data_ptr()), C.c_void_p(self._io.data_ptr()),
                                                         C.c_void_p(s)))
        return out

    def predict_host_gen_noise(self, image, noise_generator, out=None):
        """Host image in, host result out, noise drawn ON THE DEVICE from `noise_generator`'s Philox stream (what
        `gen.predict([tensor, network.noise_generator(...)])` does in the reference, api.py:136-137)."""
        import torch
        img = image if isinstance(image, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(image, np.float32))
        img = img.contiguous()
        B, T, S = img.shape[:3]
        self._ensure_plan(B, T)
        if out is None:
            out = torch.empty((B, T, S, S, self.out_channels), dtype=torch.float32)
        std, seed, offset = noise_generator.reserve(B * T * S * S * self.noise_channels)
        s = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().wdg_generator_predict_host_gen_noise(
            self._h, C.c_void_p(img.data_ptr()), std, C.c_uint64(seed), C.c_uint64(offset), C.c_void_p(out.data_ptr()),
            C.c_void_p(self._io.data_ptr()), C.c_void_p(s)))
        return out

    def predict(self, inputs, batch_size=None, verbose=0, **kwargs):
        """Keras `Model.predict([image, noise])` (api.py:137): returns a numpy array."""
        image, noise = inputs
        import torch
        if isinstance(image, torch.Tensor) and image.is_cuda:
            return self.forward_device(image, noise).cpu().numpy()
        return self.predict_host(image, noise).numpy()

    def __call__(self, inputs, training=False, mask=None):
        """`generator([low_res, noise], training=False)` (ganbase.py:65).  Training-mode forward
        (batch statistics + spectral-norm power iteration) is not part of the inference path."""
        import torch
        image, noise = inputs
        if training:
            # training-mode call: spectral-norm power iteration (mutates w / sn_u), BatchNorm batch statistics and
            # moving-average update -- on the fp32 training kernels (train/nets.py)
            from ..train.nets import GenNet, to_device
            from ..train.step import _dev
            from ..train import ops as _ops
            _ops.use_current_stream()
            w = to_device(self.get_weights())
            out = GenNet(w).forward(_dev(image), _dev(noise), training=True)
            self.set_weights({k: v.cpu().numpy() for k, v in w.items()})
            return out
        if isinstance(image, torch.Tensor) and image.is_cuda:
            return self.forward_device(image, noise)
        return self.predict_host(image, noise)

    call = __call__

    def compile(self, optimizer=None, loss=None, metrics=None, **kwargs):
        self.optimizer, self.compiled_loss, self.compiled_metrics = optimizer, loss, metrics
        self.metrics = list(metrics or [])

    STAGES = ("pack_input", "conv8x8s2", "conv4x4s2", "convlstm", "conv3x3", "convT2x2s2", "border_fix",
              "upconvT5x5", "conv3x3_out")

    def set_profiling(self, enable=True):
        _lib.check(_lib.lib().wdg_generator_profile(self._h, int(enable)))

    def stage_ms(self):
        """Device time (ms) of each stage of the last forward, from CUDA events on the launch stream."""
        ms = (C.c_float * len(self.STAGES))()
        _lib.check(_lib.lib().wdg_generator_stage_ms(self._h, ms, len(self.STAGES)))
        return dict(zip(self.STAGES, (float(v) for v in ms)))

    def debug_intermediate(self, which):
        """Intermediate activation of the last forward as fp32 numpy (parity tests)."""
        B, T = self._plan
        N, S, F = B * T, self.image_size, self.feature_channels
        shape = {0: (N, S // 2, S // 2, 128), 1: (N, S // 4, S // 4, F), 2: (N, S // 4, S // 4, F),
                 3: (N, S // 4, S // 4, F // 2), 4: (N, S // 2, S // 2, F // 4), 5: (N, S, S, F // 8)}[which]
        a = np.empty(shape, np.float32)
        _lib.check(_lib.lib().wdg_generator_debug_read(self._h, which, a.ctypes.data_as(C.c_void_p), a.size))
        return a


def make_generator(image_size: int, in_channels: int, noise_channels: int, out_channels: int, n_timesteps: int,
                   batch_size: int = None, feature_channels=128):
    """Same signature as the reference's `make_generator` (models.py:9-17)."""
    return Generator(image_size, in_channels, noise_channels, out_channels, n_timesteps, batch_size, feature_channels)


def make_discriminator(low_res_size: int, high_res_size: int, low_res_channels: int, high_res_channels: int,
                       n_timesteps: int, batch_size: int = None, feature_channels: int = 16, ckpt_topology: bool = False):
    """Same signature as the reference's `make_discriminator` (models.py:76-84).  ckpt_topology (extension): build the
    graph revision the shipped weights-55.ckpt/discriminator was written from (shortcut branch); loading
    such a checkpoint switches to it automatically."""
    if low_res_size != high_res_size:
        raise NotImplementedError("The discriminator assumes that the low res and high res images have the same size."
                                  "Perhaps you should upsample your low res image first?")  # models.py:89-91
    from .critic import Critic
    return Critic(low_res_size, high_res_size, low_res_channels, high_res_channels, n_timesteps, batch_size,
                  feature_channels, ckpt_topology=ckpt_topology)
