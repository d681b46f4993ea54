"""Random-number policy of the drop-in classes.

The reference draws from torch's global generators: normal_ for the
reparameterisation (cvae.py:81), Exp(1) inside Categorical.sample for sampled
pivots (pivotcvae.py:349-351) and Bernoulli for the mask (train_generative.py:39).
The kernels accept either caller-supplied noise tensors (parity mode, identical
streams to the reference) or a Philox4x32-10 counter (throughput mode).
"""
import torch


class NoiseSource:
    """mode 'philox' (default): in-kernel counter RNG, seeded from torch.initial_seed()
    on first use so torch.manual_seed() still controls reproducibility.
    External tensors queued with push() take precedence and are consumed in order."""

    def __init__(self, seed=None):
        self.seed = seed
        self.counter = 0
        self._queue = {"eps": [], "race": [], "mask": []}

    def push(self, kind, tensor):
        self._queue[kind].append(tensor)

    def pop(self, kind):
        q = self._queue[kind]
        return q.pop(0) if q else None

    def reseed(self, seed):
        self.seed, self.counter = int(seed), 0

    def next_stream(self, rows):
        """-> (seed, offset): a fresh block of `rows` Philox row counters."""
        if self.seed is None:
            self.seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        off = self.counter
        self.counter += int(rows)
        return self.seed, off
