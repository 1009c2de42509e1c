import torch


class SeedGenerator:
    def __init__(self, seed=None):
        self.generator = torch.Generator()
        if seed is not None:
            self.generator.manual_seed(int(seed))


def normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None):
    gen = seed.generator if isinstance(seed, SeedGenerator) else None
    shape = [int(s) for s in shape]
    return torch.randn(shape, generator=gen, dtype=torch.get_default_dtype()) * stddev + mean


def dropout(inputs, rate, noise_shape=None, seed=None):
    gen = seed.generator if isinstance(seed, SeedGenerator) else None
    keep = (torch.rand(inputs.shape, generator=gen) >= rate).to(inputs.dtype)
    return inputs * keep / (1.0 - rate)
