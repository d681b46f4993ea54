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
        self.device_counter = None   # int64[1] CUDA tensor while a CUDA graph is captured / replayed
        self._queue = {"eps": [], "race": [], "mask": []}

    def __getstate__(self):
        st = self.__dict__.copy()
        st["device_counter"] = None
        return st

    # ---- CUDA-graph mode: offsets inside the captured step are relative, the base lives on the
    # device and is advanced by one tiny kernel at the end of every replay (pcv_counter_add)
    def begin_graph(self, device):
        import torch as _t
        if self.seed is None:
            self.seed = int(_t.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        self.device_counter = _t.full((1,), self.counter, dtype=_t.int64, device=device)
        self._graph_base = self.counter
        self.counter = 0

    def end_graph_step(self):
        """Call as the last op of the captured step: advances the device counter by the rows consumed."""
        from . import ops
        used = self.counter
        ops.counter_add(self.device_counter, used)
        self.counter = 0
        return used

    def flush_eager(self):
        """End of an EAGER top-level op (recommend / get_gen_loss) while graph mode is installed: advance the
        device counter by the rows this op drew, exactly as the captured step's last node does, so eager calls
        and graph replays never reuse a Philox row."""
        if self.device_counter is not None and self.counter and not torch.cuda.is_current_stream_capturing():
            self.end_graph_step()

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

    def stream_args(self, rows):
        """kwargs (seed, offset[, offset_dev]) for one op consuming `rows` Philox row counters."""
        seed, off = self.next_stream(rows)
        kw = dict(seed=seed, offset=off)
        if self.device_counter is not None:
            kw["offset_dev"] = self.device_counter
        return kw
