"""Losses and optimizer settings of the reference (`gan/train.py:11-12,34-35,57-58`)."""


def discriminator_loss(real_output, fake_output):
    """Wasserstein critic loss: -(mean(real) - mean(fake))  (train.py:11-12)."""
    return -(real_output.mean() - fake_output.mean())


class AdamSettings:
    """Keras Adam hyper-parameters; update rule var -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)."""

    def __init__(self, lr, beta_1, beta_2, epsilon):
        self.lr, self.beta_1, self.beta_2, self.epsilon = lr, beta_1, beta_2, epsilon
        self.iterations = 0


def generator_optimizer():
    return AdamSettings(lr=1e-4, beta_1=0.5, beta_2=0.9, epsilon=0.1)      # train.py:35


def discriminator_optimizer():
    return AdamSettings(lr=4e-4, beta_1=0.5, beta_2=0.9, epsilon=0.1)      # train.py:58
