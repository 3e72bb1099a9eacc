"""ORACLE helper (test infrastructure): deterministic synthetic weights for a Unet3D state_dict.

The reference zero-initialises ``to_out`` (dalle2_video.py:644), ties the pixel-shuffle conv rows x4
(:55-62) and leaves GroupNorm/LayerNorm affine at identity, so a freshly constructed model makes a
vacuous parity test.  ``synth_state_dict`` fills every tensor of a given
{key: shape} layout from one seeded generator, in sorted key order, with a scale that keeps
activations O(1).  It depends only on the layout, so fixtures need not store weights.
"""
from __future__ import annotations

import math

import torch


def synth_state_dict(shapes: dict, seed: int = 0, dtype=torch.float32) -> dict:
    g = torch.Generator().manual_seed(seed)
    out = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "g" or (leaf == "weight" and len(shape) == 1):  # norm gains
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            t = 0.05 * torch.randn(shape, generator=g)
        elif leaf == "weight":
            fan_in = max(1, math.prod(shape[1:]))
            t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        elif leaf == "null_kv":
            t = torch.randn(shape, generator=g)
        else:  # null_video_embed / null_video_hiddens / null_text_embed
            t = torch.randn(shape, generator=g)
        out[key] = t.to(dtype)
    return out
