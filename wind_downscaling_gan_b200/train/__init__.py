"""fp32 training path (critic forward/backward, training-mode generator, WGAN step) on the generic CUDA ops of
csrc/train_ops.cu.  PyTorch allocates the device buffers; every arithmetic op is a libwdg kernel."""
