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


def test_kept_forward_and_recompute_fallback_give_the_same_gradient(emu):
    """Autograd flow on a small batch: (1) forward keeps the tape, backward replays it; (2) a second training forward on the same runner
    invalidates the kept one, so the first graph's backward falls back to the one-call form -- both must produce the same gradients."""
    from nabladft_b200.synth import synth_batch

    b = synth_batch(7, 2, heavy_min=3, heavy_max=5)
    net, _ = _models(False)
    net.train()

    class D:
        z, pos, batch = torch.from_numpy(b["z"]).long(), torch.from_numpy(b["pos"]), torch.from_numpy(b["batch"]).long()

    r = emu()
    v = torch.randn(D.z.shape[0], 3, generator=torch.Generator().manual_seed(2))

    def loss(E, F):
        return E.sum() + (F * v).sum()

    E, F = net._train_with(r, D())
    loss(E, F).backward()                       # replayed tape
    g_kept = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    net.zero_grad()
    E1, F1 = net._train_with(r, D())
    E2, F2 = net._train_with(r, D())             # replaces the kept forward of (E1, F1)
    assert r.backward(10**9, torch.ones(2), None) is False  # unknown token: refused, nothing replayed
    loss(E1, F1).backward()                      # falls back to forward + backward in one call
    g_fallback = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    assert set(g_kept) == set(g_fallback) and len(g_kept) > 300
    for n in g_kept:
        scale = g_kept[n].abs().max().item()
        assert (g_kept[n] - g_fallback[n]).abs().max().item() <= 1e-5 * scale + 1e-12, n
    assert torch.equal(E1, E2) and torch.allclose(F1, F2)


def test_small_batch_parameter_gradients_match_oracle_autograd(emu):
    """The same comparison as the first test on a 2-molecule synthetic batch (fast): used by tests/test_emu_libgemm_dispatch.py to check the
    arguments of the library GEMM calls of the training engine against the oracle."""
    from nabladft_b200.synth import synth_batch

    b = synth_batch(7, 2, heavy_min=3, heavy_max=5)
    z, pos, batch = torch.from_numpy(b["z"]).long(), torch.from_numpy(b["pos"]), torch.from_numpy(b["batch"]).long()
    net, ora = _models(True)
    ora = ora.double().train()
    for p in ora.parameters():
        p.requires_grad_(p.dtype.is_floating_point and p.dim() > 0)
    gen = torch.Generator().manual_seed(4)
    c, v = torch.randn(2, generator=gen, dtype=torch.float64), torch.randn(z.shape[0], 3, generator=gen, dtype=torch.float64)
    E0, F0 = ora(z, pos.double(), batch)
    ((E0 * c).sum() + (F0 * v).sum()).backward()

    class D:
        pass

    d = D()
    d.z, d.pos, d.batch = z, pos, batch
    E, F = net.train()._train_with(emu(), d)
    ((E * c.float()).sum() + (F * v.float()).sum()).backward()
    refp = dict(ora.named_parameters())
    n_checked = 0
    for name, p in net.named_parameters():
        if name.endswith("scale_factor"):
            continue
        g_ref = refp[name].grad
        scale = g_ref.abs().max().item()
        assert (p.grad.double() - g_ref).abs().max().item() <= 2e-4 * scale + 1e-10, name
        n_checked += 1
    assert n_checked > 300
