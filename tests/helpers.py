"""Shared helpers for the test-suite: fixture molecules, golden weights, batches."""
import os

import numpy as np
import torch

from weights import golden_state_dict  # tests/golden/weights.py

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(mol_ids, dtype=torch.float64):
    fx = np.load(os.path.join(GOLDEN, "fixture_molecules.npz"))
    z, pos, b = [], [], []
    for k, m in enumerate(mol_ids):
        a, e = fx["ptr"][m], fx["ptr"][m + 1]
        z.append(fx["z"][a:e])
        pos.append(fx["pos"][a:e])
        b.append(np.full(e - a, k))
    return (
        torch.from_numpy(np.concatenate(z)).long(),
        torch.from_numpy(np.concatenate(pos)).to(dtype),
        torch.from_numpy(np.concatenate(b)).long(),
    )


def load_golden_weights(module: torch.nn.Module, dtype=torch.float64, **kw):
    sd = module.state_dict()
    for k, v in golden_state_dict(sd, **kw).items():
        sd[k] = torch.from_numpy(v).to(dtype)
    module.load_state_dict(sd, strict=True)
    return module


def random_rotation(seed=0, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    q, r = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
    q = q * torch.sign(torch.diagonal(r))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q.to(dtype)
