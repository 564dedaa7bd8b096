"""Helpers shared by the parity tests."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


class Replay:
    """Replays the sample indices the real reference drew (recorded in the golden file)."""

    def __init__(self, gold, step):
        p = f"s{step}_"
        self.high = gold[p + "call_high"].tolist()
        self.shape = gold[p + "call_shape"].tolist()
        self.idx = [torch.from_numpy(gold[p + f"call{i}"]) for i in range(len(self.high))]
        self.pos = 0

    def __call__(self, high, shape):
        i = self.pos
        assert i < len(self.high), "more sampler calls than the reference made"
        assert (int(high), int(shape)) == (self.high[i], self.shape[i]), \
            f"sampler call {i}: got (high={high}, shape={shape}), reference made ({self.high[i]}, {self.shape[i]})"
        self.pos += 1
        return self.idx[i]

    def done(self):
        return self.pos == len(self.high)

    def split(self):
        """Indices as the CUDA op takes them: per active slot, (anchor, negative) in call order."""
        return self.idx[0::2], self.idx[1::2]


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).flatten()
    b = torch.as_tensor(b, dtype=torch.float64).flatten()
    denom = max(float(b.norm()), 1e-30)
    return float((a - b).norm()) / denom
