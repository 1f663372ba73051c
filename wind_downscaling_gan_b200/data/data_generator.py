"""`FlexibleNoiseGenerator` with the reference's signature (`data/data_generator.py:319-335`), drawing on the GPU
with the library's Philox4x32-10 + Box-Muller kernel (csrc/wdg_noise.cu) so the noise never crosses PCIe.
The reference draws from TensorFlow's global Philox generator; the streams differ, so parity runs pass explicit
noise tensors instead.

The stream position (key, next counter block) lives in DEVICE memory and is advanced by a one-thread kernel after every
draw, so a captured CUDA graph of the training step draws fresh noise on every replay; `_offset` is the host mirror of
that counter (the training step bumps it after each replay)."""
import ctypes as C
import os


class FlexibleNoiseGenerator(object):
    def __init__(self, noise_shape, std=1, random_seed=None, device="cuda"):
        self.noise_shape = noise_shape
        self.std = std
        self.device = device
        self.random_seed = random_seed
        self._seed = int(random_seed) if random_seed is not None else int.from_bytes(os.urandom(8), "little")
        self._offset = 0          # Philox counter blocks consumed so far (host mirror of the device counter)
        self._state = None        # device uint64[2] = (key, next counter block)

    def _device_state(self):
        import numpy as np
        import torch
        if self._state is None:
            host = np.array([self._seed & (2 ** 64 - 1), self._offset], dtype=np.uint64).view(np.int64)
            self._state = torch.from_numpy(host).to(self.device)
        return self._state

    def __call__(self, bs=None, channels=None, std=None):
        import torch
        from .. import _lib
        bs = self.noise_shape[0] if bs is None else int(bs)
        t, x, y = self.noise_shape[1], self.noise_shape[2], self.noise_shape[3]
        channels = self.noise_shape[4] if channels is None else channels
        std = std or self.std
        out = torch.empty((bs, t, x, y, channels), dtype=torch.float32, device=self.device)
        n = out.numel()
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().wdg_noise_normal_state(C.c_void_p(out.data_ptr()), n, float(std),
                                                     C.c_void_p(self._device_state().data_ptr()), stream))
        self._offset += (n + 3) // 4
        return out

    def uniform(self, n):
        """n draws from U[0, 1) out of the same stream (the interpolation weights of ganbase.py:30)."""
        import torch
        from .. import _lib
        out = torch.empty(int(n), dtype=torch.float32, device=self.device)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().wdg_uniform_state(C.c_void_p(out.data_ptr()), int(n), C.c_void_p(self._device_state().data_ptr()), stream))
        self._offset += (int(n) + 3) // 4
        return out

    def reserve(self, n_elements):
        """Hands out the next `n_elements` of the stream WITHOUT materialising them: returns (std, seed, offset) for a
        kernel that draws the same values itself (the generator's input packing, wdg_generator_forward_gen_noise) and
        advances the generator exactly as `__call__` would."""
        import torch
        from .. import _lib
        spec = (float(self.std), self._seed & (2 ** 64 - 1), self._offset)
        blocks = (int(n_elements) + 3) // 4
        if self._state is not None:
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().wdg_rng_advance(C.c_void_p(self._state.data_ptr()), C.c_uint64(blocks), stream))
        self._offset += blocks
        return spec
