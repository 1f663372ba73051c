"""`FlexibleNoiseGenerator` with the reference's signature (`data/data_generator.py:319-335`), drawing on the GPU
with the library's Philox4x32-10 + Box-Muller kernel (csrc/wdg_noise.cu) so the noise never crosses PCIe.
The reference draws from TensorFlow's global Philox generator; the streams differ, so parity runs pass explicit
noise tensors instead."""
import ctypes as C
import os


class FlexibleNoiseGenerator(object):
    def __init__(self, noise_shape, std=1, random_seed=None, device="cuda"):
        self.noise_shape = noise_shape
        self.std = std
        self.device = device
        self.random_seed = random_seed
        self._seed = int(random_seed) if random_seed is not None else int.from_bytes(os.urandom(8), "little")
        self._offset = 0          # Philox counter blocks consumed so far

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
        _lib.check(_lib.lib().wdg_noise_normal(C.c_void_p(out.data_ptr()), n, float(std), C.c_uint64(self._seed & (2 ** 64 - 1)),
                                               C.c_uint64(self._offset), stream))
        self._offset += (n + 3) // 4
        return out

    def reserve(self, n_elements):
        """Hands out the next `n_elements` of the stream WITHOUT materialising them: returns (std, seed, offset) for a
        kernel that draws the same values itself (the generator's input packing, wdg_generator_forward_gen_noise) and
        advances the generator exactly as `__call__` would."""
        spec = (float(self.std), self._seed & (2 ** 64 - 1), self._offset)
        self._offset += (int(n_elements) + 3) // 4
        return spec
