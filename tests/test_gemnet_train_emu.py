"""GemNet-OC training (direct forces: first-order back-propagation from dLoss/dE and dLoss/dF) checked on the CPU: csrc/gemnet_oc_train.inc through the
host-emulation build, driven by the product's own host code (differentiable flat export, `GemNetOCFn`), against the autograd of the pinned oracle
(oracle/gemnet_oc.py) for EVERY reference-named parameter.  Same caveats as tests/test_gemnet_emu.py."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, os.path.join(HERE, "emu"))

from test_gemnet_emu import _models, emu  # noqa: E402,F401  (fixture + model builders)


def test_gemnet_oc_every_parameter_gradient_matches_oracle_autograd(emu):
    g = np.load(os.path.join(HERE, "golden", "gemnet_oc_f32.npz"))
    z, pos, batch = torch.from_numpy(g["z"]).long(), torch.from_numpy(g["pos"]), torch.from_numpy(g["batch"]).long()
    net, ora = _models(True)  # scale factors != 1: the gradient has to pass through the folded basis matrices
    ora = ora.double().train()
    for p in ora.parameters():
        p.requires_grad_(p.dtype.is_floating_point and p.dim() > 0)
    gen = torch.Generator().manual_seed(11)
    c = torch.randn(2, generator=gen, dtype=torch.float64)
    v = torch.randn(z.shape[0], 3, generator=gen, dtype=torch.float64)
    E0, F0 = ora(z, pos.double(), batch)
    ((E0 * c).sum() + (F0 * v).sum()).backward()

    class D:
        pass

    d = D()
    d.z, d.pos, d.batch = z, pos, batch
    net.train()
    E, F = net._train_with(emu(), d)
    assert (E.detach().double() - E0.detach()).abs().max() < 2e-5 * E0.abs().max() and (F.detach().double() - F0.detach()).abs().max() < 2e-5 * F0.abs().max()
    ((E * c.float()).sum() + (F * v.float()).sum()).backward()
    refp = dict(ora.named_parameters())
    worst, n_checked = (0.0, ""), 0
    for name, p in net.named_parameters():
        g_ref = refp[name].grad
        if name.endswith("scale_factor"):
            assert p.grad is None  # fitted constants, requires_grad False in the reference too
            continue
        assert p.grad is not None and g_ref is not None, name
        scale = g_ref.abs().max().item()
        err = (p.grad.double() - g_ref).abs().max().item()
        if err / max(scale, 1e-30) > worst[0]:
            worst = (err / max(scale, 1e-30), name)
        assert err <= 2e-4 * scale + 1e-10, (name, err, scale)
        n_checked += 1
    print(f"{n_checked} parameter tensors; worst relative gradient error {worst[0]:.2e} ({worst[1]})")
    assert n_checked > 300
