"""Test-only: evaluates the engine's CANONICAL weight layout (include/nabla_b200.h,
`struct nb200_painn_weights`) with plain torch on CPU, mirroring engine.cu step by step.
Used to validate the host-side weight export (role permutations) without a GPU and to
localise GPU mismatches.  Never imported by the product."""
import math

import torch

from oracle.graph import radius_graph


def canonical_energy_forces(t, s, z, pos, batch):
    """t: canonical tensors (any dtype/device cpu), s: scalars dict (as passed to the engine)."""
    dt = pos.dtype
    t = {k: v.to(dt) for k, v in t.items()}
    L, F = s["n_layers"], s["n_feat"]
    pos = pos.clone().requires_grad_(True)
    ei = radius_graph(pos, s["cutoff"], batch, 10**9)
    j, i = ei
    r = pos[j] - pos[i]
    d = r.norm(dim=1)
    u = r / d[:, None]
    x = d * s["rbf_xscale"]
    phi = torch.exp(s["rbf_coeff"] * (x[:, None] - t["rbf_offsets"][None]) ** 2)
    if s["radial_mode"] == 0:
        s1 = 0.5 * (torch.cos(d * math.pi / s["cutoff"]) + 1) * (d < s["cutoff"])
        s2 = s1
    else:
        xs = d / s["cutoff"]
        s1 = torch.where(xs < 1, 1 - 21 * xs**5 + 35 * xs**6 - 15 * xs**7, torch.zeros_like(xs))
        s2 = torch.ones_like(s1)
    q = t["emb"][z - s["z_offset"]]
    mu = torch.zeros(z.numel(), 3, F, dtype=dt)
    for l in range(L):
        W = s1[:, None] * (phi @ t["w_rbf"][l]) + s2[:, None] * t["b_rbf"][l]
        xh = torch.nn.functional.silu(q @ t["A1"][l].T + t["c1"][l]) @ t["A2"][l].T + t["c2"][l]
        p = xh[j] * W
        a, b, c = p[:, :F], p[:, F:2 * F], p[:, 2 * F:]
        q = q + torch.zeros_like(q).index_add_(0, i, a)
        mu = mu + torch.zeros_like(mu).index_add_(0, i, b[:, None, :] * u[:, :, None] + c[:, None, :] * mu[j])
        VW = mu @ t["U"][l].T
        V, Wv = VW[..., :F], VW[..., F:]
        n = torch.sqrt((V**2).sum(1) + s["epsilon"])
        y = torch.nn.functional.silu(torch.cat([q, n], -1) @ t["B1"][l].T + t["d1"][l]) @ t["B2"][l].T + t["d2"][l]
        y0, y1, y2 = y[:, :F], y[:, F:2 * F], y[:, 2 * F:]
        q = q + y0 + y2 * (V * Wv).sum(1)
        mu = mu + y1[:, None, :] * Wv
    eps = torch.nn.functional.silu(q @ t["R1"].T + t["e1"]) @ t["R2"].T + t["e2"]
    n_mol = int(batch.max()) + 1
    e = torch.zeros(n_mol, dtype=dt).index_add_(0, batch, eps.squeeze(-1))
    f = -torch.autograd.grad(e.sum(), pos)[0]
    e = e + s["energy_shift_per_atom"] * torch.bincount(batch, minlength=n_mol).to(dt)
    return e.detach(), f
