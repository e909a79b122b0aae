"""ORACLE TOOLING (not product code): deterministic synthetic weights for parity tests and benchmarks.

There are no checkpoints in the reference tree, and a freshly constructed reference network returns eps == 0 because
every temporal / control / output projection is zero-initialised.  Parity therefore uses weights
generated from a committed manifest (tests/golden/manifest_*.json: state-dict key -> shape, plus whether the reference
initialises the tensor to zeros / ones), one independent seeded generator per key:
    ones-initialised (norm scales)            -> 1 + 0.1 * N(0,1)
    matrices / conv kernels (default init)    -> U(-1/sqrt(fan_in), 1/sqrt(fan_in))   (nn.Linear / nn.Conv default)
    zero-initialised tensors and all vectors  -> N(0, 0.02)
The same function feeds the reference (make_golden.py), the oracle and the CUDA build, on any machine.
"""
import json
import math
import os
import zlib

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_manifest(kind: str) -> dict:
    with open(os.path.join(GOLDEN_DIR, f"manifest_{kind}.json")) as f:
        return json.load(f)


def seeded_tensor(key: str, shape, init: str, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    shape = tuple(shape)
    if init == "ones":
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if len(shape) >= 2 and init != "zeros":
        bound = 1.0 / math.sqrt(math.prod(shape[1:]))
        return (torch.rand(shape, generator=g) * 2 - 1) * bound
    return 0.02 * torch.randn(shape, generator=g)


def seeded_state_dict(manifest: dict, seed: int = 0, prefix_filter: str = "") -> dict:
    """manifest: key -> [shape, init]; returns key -> fp32 tensor (only keys starting with prefix_filter)."""
    return {k: seeded_tensor(k, shp, init, seed) for k, (shp, init) in manifest.items() if k.startswith(prefix_filter)}
