"""`FlexibleNoiseGenerator` with the reference's signature (`data/data_generator.py:319-335`),
drawing on the GPU so the noise never crosses PCIe.  The reference draws from TensorFlow's global
Philox generator; the streams differ, so parity runs pass explicit noise tensors instead."""


class FlexibleNoiseGenerator(object):
    def __init__(self, noise_shape, std=1, random_seed=None, device="cuda"):
        self.noise_shape = noise_shape
        self.std = std
        self.device = device
        self.random_seed = random_seed
        self._gen = None

    def _generator(self):
        import torch
        if self._gen is None:
            self._gen = torch.Generator(device=self.device)
            if self.random_seed is not None:
                self._gen.manual_seed(int(self.random_seed))
            else:
                self._gen.seed()
        return self._gen

    def __call__(self, bs=None, channels=None, std=None):
        import torch
        bs = self.noise_shape[0] if bs is None else int(bs)
        t, x, y = self.noise_shape[1], self.noise_shape[2], self.noise_shape[3]
        channels = self.noise_shape[4] if channels is None else channels
        std = std or self.std
        out = torch.empty((bs, t, x, y, channels), dtype=torch.float32, device=self.device)
        return out.normal_(mean=0.0, std=float(std), generator=self._generator())
