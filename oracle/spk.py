"""Oracle: SchNet / PaiNN as configured by `config/model/{schnet,painn}.yaml`
-- TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

The arithmetic lives in **schnetpack==2.0.4** (`/root/reference/setup.py:32`), which is not
vendored under /root/reference and cannot be installed here.  This file restates its
published algorithm anchored on the reference's call sites:

  config/model/painn.yaml:5-28   NeuralNetworkPotential(PaiNN(128, 6, GaussianRBF(100, 5.0),
                                 CosineCutoff(5.0)), [PairwiseDistances], [Atomwise, Forces],
                                 [AddOffsets(energy, add_mean)])
  config/model/schnet.yaml:5-28  same with representation SchNet
  nablaDFT/ase_model/task.py:34-65  test/predict call `self(batch)` (post-processing on)

PARITY UNPINNED by the reference's own tests (shape-only, tests/model/test_torch_models.py:31-40).
Cross-check available offline: `SpkPaiNN` and `oracle.painn_oc.PaiNNOC` are the same layer up
to the weight-role permutation in `tests/test_oracle.py::test_spk_painn_matches_painn_oc_roles`.

schnetpack modules restated (names are state_dict-compatible with 2.0.4):
  nn/radial.py GaussianRBF          -> gaussian_rbf           offsets=linspace(0,rc,n), widths=|o1-o0|
  nn/cutoff.py CosineCutoff         -> cosine_cutoff          0.5(cos(pi d/rc)+1)(d<rc)
  nn/activations.py shifted_softplus-> ssp
  nn/base.py Dense                  -> nn.Linear (xavier weight, zero bias) + activation
  representation/painn.py           -> SpkPaiNN (PaiNNInteraction, PaiNNMixing)
  representation/schnet.py          -> SpkSchNet (SchNetInteraction)
  atomistic/atomwise.py Atomwise    -> outnet 128->64->1 silu, scatter_add over idx_m
  atomistic/response.py Forces      -> -dE/dR
  transform/atomistic.py AddOffsets -> E += mean * n_atoms (eval-time post-processing)
"""
import math

import torch
import torch.nn.functional as F
from torch import nn


def ssp(x):
    return F.softplus(x) - math.log(2.0)


def gaussian_rbf(d, offsets, widths):
    coeff = -0.5 / widths**2
    return torch.exp(coeff * (d[..., None] - offsets) ** 2)


def cosine_cutoff(d, cutoff):
    return 0.5 * (torch.cos(d * math.pi / cutoff) + 1.0) * (d < cutoff).to(d.dtype)


def _dense(n_in, n_out, bias=True):
    lin = nn.Linear(n_in, n_out, bias=bias)
    nn.init.xavier_uniform_(lin.weight)
    if bias:
        nn.init.zeros_(lin.bias)
    return lin


class _RBF(nn.Module):
    def __init__(self, n_rbf, cutoff):
        super().__init__()
        offsets = torch.linspace(0.0, cutoff, n_rbf)
        self.register_buffer("offsets", offsets)
        self.register_buffer("widths", torch.abs(offsets[1] - offsets[0]) * torch.ones_like(offsets))

    def forward(self, d):
        return gaussian_rbf(d, self.offsets, self.widths)


class _PaiNNInteraction(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.n = n
        self.interatomic_context_net = nn.ModuleList([_dense(n, n), _dense(n, 3 * n)])

    def forward(self, q, mu, Wij, dir_ij, idx_i, idx_j, n_atoms):
        x = self.interatomic_context_net[1](F.silu(self.interatomic_context_net[0](q)))
        xj, muj = x[idx_j], mu[idx_j]
        x = Wij * xj
        dq, dmuR, dmumu = torch.split(x, self.n, dim=-1)
        dq = torch.zeros_like(q).index_add_(0, idx_i, dq)
        dmu = dmuR * dir_ij[..., None] + dmumu * muj
        dmu = torch.zeros_like(mu).index_add_(0, idx_i, dmu)
        return q + dq, mu + dmu


class _PaiNNMixing(nn.Module):
    def __init__(self, n, epsilon=1e-8):
        super().__init__()
        self.n, self.epsilon = n, epsilon
        self.intraatomic_context_net = nn.ModuleList([_dense(2 * n, n), _dense(n, 3 * n)])
        self.mu_channel_mix = _dense(n, 2 * n, bias=False)

    def forward(self, q, mu):
        mu_V, mu_W = torch.split(self.mu_channel_mix(mu), self.n, dim=-1)
        mu_Vn = torch.sqrt(torch.sum(mu_V**2, dim=-2, keepdim=True) + self.epsilon)
        ctx = torch.cat([q, mu_Vn], dim=-1)
        x = self.intraatomic_context_net[1](F.silu(self.intraatomic_context_net[0](ctx)))
        dq_intra, dmu_intra, dqmu_intra = torch.split(x, self.n, dim=-1)
        dmu_intra = dmu_intra * mu_W
        dqmu_intra = dqmu_intra * torch.sum(mu_V * mu_W, dim=1, keepdim=True)
        return q + dq_intra + dqmu_intra, mu + dmu_intra


class SpkPaiNN(nn.Module):
    """schnetpack.representation.PaiNN(n_atom_basis, n_interactions, radial_basis, cutoff_fn)."""

    def __init__(self, n_atom_basis=128, n_interactions=6, n_rbf=100, cutoff=5.0, max_z=100, epsilon=1e-8):
        super().__init__()
        self.n_atom_basis, self.n_interactions, self.cutoff = n_atom_basis, n_interactions, cutoff
        self.embedding = nn.Embedding(max_z, n_atom_basis, padding_idx=0)
        self.radial_basis = _RBF(n_rbf, cutoff)
        self.filter_net = _dense(n_rbf, n_interactions * 3 * n_atom_basis)
        self.interactions = nn.ModuleList(_PaiNNInteraction(n_atom_basis) for _ in range(n_interactions))
        self.mixing = nn.ModuleList(_PaiNNMixing(n_atom_basis, epsilon) for _ in range(n_interactions))

    def forward(self, z, r_ij, idx_i, idx_j):
        n_atoms = z.shape[0]
        d_ij = torch.norm(r_ij, dim=1, keepdim=True)
        dir_ij = r_ij / d_ij
        phi_ij = self.radial_basis(d_ij)  # [E,1,n_rbf]
        fcut = cosine_cutoff(d_ij, self.cutoff)
        filters = self.filter_net(phi_ij) * fcut[..., None]
        filter_list = torch.split(filters, 3 * self.n_atom_basis, dim=-1)
        q = self.embedding(z)[:, None]
        mu = torch.zeros((n_atoms, 3, self.n_atom_basis), dtype=q.dtype)
        for i in range(self.n_interactions):
            q, mu = self.interactions[i](q, mu, filter_list[i], dir_ij, idx_i, idx_j, n_atoms)
            q, mu = self.mixing[i](q, mu)
        return q.squeeze(1), mu


class _SchNetInteraction(nn.Module):
    def __init__(self, n, n_rbf, n_filters):
        super().__init__()
        self.in2f = _dense(n, n_filters, bias=False)
        self.f2out = nn.ModuleList([_dense(n_filters, n), _dense(n, n)])
        self.filter_network = nn.ModuleList([_dense(n_rbf, n_filters), _dense(n_filters, n_filters)])

    def forward(self, x, f_ij, idx_i, idx_j, rcut_ij):
        x = self.in2f(x)
        Wij = self.filter_network[1](ssp(self.filter_network[0](f_ij)))
        Wij = Wij * rcut_ij[:, None]
        x_ij = x[idx_j] * Wij
        x = torch.zeros_like(x).index_add_(0, idx_i, x_ij)
        return self.f2out[1](ssp(self.f2out[0](x)))


class SpkSchNet(nn.Module):
    """schnetpack.representation.SchNet(n_atom_basis, n_interactions, radial_basis, cutoff_fn)."""

    def __init__(self, n_atom_basis=128, n_interactions=6, n_rbf=100, cutoff=5.0, n_filters=None, max_z=100):
        super().__init__()
        self.n_atom_basis, self.cutoff = n_atom_basis, cutoff
        n_filters = n_filters or n_atom_basis
        self.embedding = nn.Embedding(max_z, n_atom_basis, padding_idx=0)
        self.radial_basis = _RBF(n_rbf, cutoff)
        self.interactions = nn.ModuleList(_SchNetInteraction(n_atom_basis, n_rbf, n_filters) for _ in range(n_interactions))

    def forward(self, z, r_ij, idx_i, idx_j):
        d_ij = torch.norm(r_ij, dim=1)
        f_ij = self.radial_basis(d_ij)
        rcut_ij = cosine_cutoff(d_ij, self.cutoff)
        x = self.embedding(z)
        for inter in self.interactions:
            x = x + inter(x, f_ij, idx_i, idx_j, rcut_ij)
        return x, None


class _Atomwise(nn.Module):
    def __init__(self, n_in=128):
        super().__init__()
        self.outnet = nn.ModuleList([_dense(n_in, n_in // 2), _dense(n_in // 2, 1)])

    def forward(self, x, idx_m, n_mol):
        y = self.outnet[1](F.silu(self.outnet[0](x))).squeeze(-1)
        return torch.zeros(n_mol, dtype=y.dtype).index_add_(0, idx_m, y)


class _AddOffsets(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("mean", torch.zeros(1))


class NeuralNetworkPotential(nn.Module):
    """`schnetpack.model.NeuralNetworkPotential` as instantiated by config/model/*.yaml:
    input_modules=[PairwiseDistances], output_modules=[Atomwise('energy'), Forces()],
    postprocessors=[AddOffsets('energy', add_mean=True)]."""

    def __init__(self, representation: nn.Module):
        super().__init__()
        self.representation = representation
        self.output_modules = nn.ModuleList([_Atomwise(representation.n_atom_basis)])
        self.postprocessors = nn.ModuleList([_AddOffsets()])

    @torch.enable_grad()
    def forward(self, inputs: dict, postprocess: bool = True, create_graph: bool = False):
        z = inputs["_atomic_numbers"].long()
        R = inputs["_positions"].requires_grad_(True)
        idx_i, idx_j, idx_m = inputs["_idx_i"].long(), inputs["_idx_j"].long(), inputs["_idx_m"].long()
        offsets = inputs.get("_offsets")
        r_ij = R[idx_j] - R[idx_i]  # PairwiseDistances
        if offsets is not None:
            r_ij = r_ij + offsets
        x, _ = self.representation(z, r_ij, idx_i, idx_j)
        n_mol = int(idx_m.max().item()) + 1
        energy = self.output_modules[0](x, idx_m, n_mol)
        forces = -torch.autograd.grad(energy, R, grad_outputs=torch.ones_like(energy), create_graph=create_graph)[0]
        if postprocess:
            n_atoms = torch.bincount(idx_m, minlength=n_mol).to(energy.dtype)
            energy = energy + self.postprocessors[0].mean.to(energy.dtype) * n_atoms
        return {"energy": energy, "forces": forces}
