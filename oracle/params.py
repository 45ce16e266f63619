"""TEST INFRASTRUCTURE — deterministic, reproducible weights for parity tests.

Golden fixtures are produced by the real reference in the build container and re-checked on
the GPU box where the reference does not exist, so weights cannot be shipped as a 140 MB
checkpoint.  Instead every tensor of a reference-keyed state dict is filled from a CPU
generator seeded by (seed, position in sorted key order); the same call on the reference's,
the oracle's and the product model's state dict gives bit-identical weights.
Scales are chosen so activations stay O(1) through the networks (fan-in scaled normals for
conv / linear weights, BN affine near 1, positive running variances).
"""
from typing import Dict

import torch


def synth_fill_(state: Dict[str, torch.Tensor], seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Fill `state` in place (no grad) and return it."""
    with torch.no_grad():
        for idx, key in enumerate(sorted(state.keys())):
            t = state[key]
            g = torch.Generator(device="cpu")
            g.manual_seed(seed * 100003 + idx)
            if key.endswith("num_batches_tracked"):
                t.zero_()
                continue
            shape = tuple(t.shape)
            if key.endswith("running_var"):
                v = 0.5 + torch.rand(shape, generator=g)
            elif key.endswith("running_mean"):
                v = 0.1 * torch.randn(shape, generator=g)
            elif "_routing_fn.fc" in key:
                v = 0.5 * torch.randn(shape, generator=g)
            elif key.endswith(".bias"):
                v = 0.05 * torch.randn(shape, generator=g)
            elif t.dim() == 1:  # BN / norm scale
                v = 1.0 + 0.1 * torch.randn(shape, generator=g)
            else:
                # conv (O,I,kh,kw), CondConv experts (E,O,I,kh,kw) or linear (O,I)
                if t.dim() == 5:
                    fan_in = shape[2] * shape[3] * shape[4]
                    gain = 1.0 / 1.5  # three sigmoid-weighted experts are summed
                elif t.dim() == 4:
                    fan_in = shape[1] * shape[2] * shape[3]
                    gain = 1.0
                else:
                    fan_in = shape[-1]
                    gain = 1.0
                v = gain * torch.randn(shape, generator=g) / (fan_in ** 0.5)
            t.copy_(v.to(t.dtype))
    return state
