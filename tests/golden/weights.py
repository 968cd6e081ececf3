"""Deterministic, name-keyed test weights shared by the golden generators (which load them
into the REFERENCE classes) and the tests (which load them into the oracle and the CUDA
modules).  Keeps golden files small: they store inputs and outputs, never state dicts.

Magnitudes follow the reference initialisers (xavier-uniform matrices,
`painn_pyg/painn.py:467-473,528-533`; uniform(-sqrt3, sqrt3) embeddings, `layers.py:213`)
with small non-zero biases so that every bias path is exercised.
"""
import zlib

import numpy as np


def golden_state_dict(template: dict, bias_std: float = 0.02, weight_scale: float = 0.5, style: str = "mlp") -> dict:
    """template: name -> array-like (only .shape is read). Returns name -> float64 ndarray.
    style="e3": e3nn-style parameters (flat o3.Linear / TensorProduct weights and
    FullyConnectedNet `layer{i}.weight`) are N(0,1) as e3nn initialises them."""
    out = {}
    for name, ref in template.items():
        shape = tuple(ref.shape)
        rng = np.random.default_rng(zlib.crc32(name.encode()))
        if name.endswith(("offset", "offsets", "widths", "atomref", "cutoff_fn.cutoff")):
            continue  # buffers keep their constructor values
        if style == "e3":
            leaf = name.rsplit(".", 1)[-1]
            if leaf in ("cutoff", "logc", "n", "v", "_alpha") or len(shape) == 0:
                continue
            if (leaf in ("weight", "weights") and len(shape) == 1) or (len(shape) == 2 and ".layer" in name):
                out[name] = rng.standard_normal(size=shape).astype(np.float64)
                continue
        if "emb" in name and len(shape) == 2:
            w = rng.uniform(-np.sqrt(3.0), np.sqrt(3.0), size=shape)
        elif len(shape) == 2:
            bound = np.sqrt(6.0 / (shape[0] + shape[1])) * weight_scale
            w = rng.uniform(-bound, bound, size=shape)
        else:
            w = bias_std * rng.standard_normal(size=shape)
        out[name] = w.astype(np.float64)
    return out
