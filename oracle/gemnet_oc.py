"""CPU restatement of GemNet-OC, in progress.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Built and PINNED so far (against intermediates recorded from the reference's own classes, tests/golden/gemnet_oc_f32.npz):
    graphs and all index structures      oracle/gemnet_graph.py
    radial basis, atom / edge embedding, output block 0   this file  (gemnet_oc/layers/radial_basis.py:19-39,57-77,176-220; embedding_block.py:14-92;
                                         base_layers.py:15-75; gemnet_oc.py:1165-1167)
Not restated yet: circular / spherical bases, the interaction blocks (output blocks 1-4 reuse OutputBlock on their outputs).
Parameter names are the reference's (strict state-dict loading of the restated sub-modules).
"""
import math

import torch
from torch import nn

from .gemnet_graph import build_all_indices


class ScaledSiLU(nn.Module):
    def forward(self, x):  # base_layers.py:66-75: silu(x) / 0.6
        return torch.nn.functional.silu(x) * (1 / 0.6)


class Dense(nn.Module):
    def __init__(self, n_in, n_out, bias=False, activation=None):  # base_layers.py:15-63
        super().__init__()
        self.linear = nn.Linear(n_in, n_out, bias=bias)
        self._activation = ScaledSiLU() if activation in ("silu", "swish") else nn.Identity()

    def forward(self, x):
        return self._activation(self.linear(x))


class _Scale(nn.Module):  # scale_factor.py: a scalar parameter multiplied onto the value (fitted offline; 1 in the golden run)
    def __init__(self):
        super().__init__()
        self.scale_factor = nn.Parameter(torch.tensor(0.0), requires_grad=False)

    def forward(self, x):
        return x * self.scale_factor


class _Gaussian(nn.Module):
    def __init__(self, num):  # radial_basis.py:57-77 on the scaled distance: start 0, stop 1
        super().__init__()
        self.register_buffer("offset", torch.linspace(0.0, 1.0, num))
        self.coeff = -0.5 / (1.0 / (num - 1)) ** 2

    def forward(self, d):
        return torch.exp(self.coeff * (d[:, None] - self.offset[None, :]) ** 2)


class RadialBasis(nn.Module):
    def __init__(self, num_radial=128, cutoff=12.0, exponent=5, scale_basis=True):  # radial_basis.py:176-220
        super().__init__()
        self.inv_cutoff, p = 1.0 / cutoff, float(exponent)
        self.p, self.a, self.b, self.c = p, -(p + 1) * (p + 2) / 2, p * (p + 2), -p * (p + 1) / 2
        self.rbf = _Gaussian(num_radial)
        self.scale_basis = scale_basis
        if scale_basis:
            self.scale_rbf = _Scale()

    def forward(self, d):
        x = d * self.inv_cutoff
        env = torch.where(x < 1, 1 + self.a * x ** self.p + self.b * x ** (self.p + 1) + self.c * x ** (self.p + 2), torch.zeros_like(x))
        res = env[:, None] * self.rbf(x)
        return self.scale_rbf(res) if self.scale_basis else res


class AtomEmbedding(nn.Module):
    def __init__(self, emb_size=256, num_elements=83):
        super().__init__()
        self.embeddings = nn.Embedding(num_elements, emb_size)

    def forward(self, z):
        return self.embeddings(z - 1)


class EdgeEmbedding(nn.Module):
    def __init__(self, atom_features=256, edge_features=128, out_features=512, activation="silu"):
        super().__init__()
        self.dense = Dense(2 * atom_features + edge_features, out_features, activation=activation)

    def forward(self, h, m, edge_index):
        return self.dense(torch.cat([h[edge_index[0]], h[edge_index[1]], m], dim=-1))


class ResidualLayer(nn.Module):
    def __init__(self, units, n_layers=2, activation="silu"):  # base_layers.py:78-97: (x + mlp(x)) / sqrt(2)
        super().__init__()
        self.dense_mlp = nn.Sequential(*[Dense(units, units, activation=activation) for _ in range(n_layers)])

    def forward(self, x):
        return (x + self.dense_mlp(x)) * (1 / math.sqrt(2.0))


def _mlp(units_in, units, n_hidden, activation="silu"):  # atom_update_block.py get_mlp
    layers = [Dense(units_in, units, activation=activation)] if units_in != units else []
    return nn.ModuleList(layers + [ResidualLayer(units, 2, activation) for _ in range(n_hidden)])


class OutputBlock(nn.Module):
    """atom_update_block.py:93-172 (direct forces): per-atom energy features x_E and per-edge force features x_F."""

    def __init__(self, emb_size_atom=256, emb_size_edge=512, emb_size_rbf=16, n_hidden=3, n_hidden_afteratom=3):
        super().__init__()
        self.dense_rbf = Dense(emb_size_rbf, emb_size_edge)
        self.scale_sum = _Scale()
        self.layers = _mlp(emb_size_edge, emb_size_atom, n_hidden)
        self.seq_energy_pre = self.layers  # the reference registers the same list under both names
        self.seq_energy2 = _mlp(emb_size_atom, emb_size_atom, n_hidden_afteratom)
        self.scale_rbf_F = _Scale()
        self.seq_forces = _mlp(emb_size_edge, emb_size_edge, n_hidden)
        self.dense_rbf_F = Dense(emb_size_rbf, emb_size_edge)

    def forward(self, h, m, basis_rad, idx_atom):
        x = m * self.dense_rbf(basis_rad)
        x_E = self.scale_sum(torch.zeros(h.shape[0], x.shape[1], dtype=x.dtype).index_add_(0, idx_atom, x))
        for layer in self.seq_energy_pre:
            x_E = layer(x_E)
        x_E = (x_E + h) * (1 / math.sqrt(2.0))
        for layer in self.seq_energy2:
            x_E = layer(x_E)
        x_F = m
        for layer in self.seq_forces:
            x_F = layer(x_F)
        return x_E, self.scale_rbf_F(x_F * self.dense_rbf_F(basis_rad))


class GemNetOCStem(nn.Module):
    """Graphs -> radial basis -> h0, m0 (gemnet_oc.py:1121-1167).  The rest of the network follows in the next round."""

    def __init__(self, num_radial=128, cutoff=12.0, emb_size_atom=256, emb_size_edge=512, num_elements=83):
        super().__init__()
        self.radial_basis = RadialBasis(num_radial, cutoff)
        self.atom_emb = AtomEmbedding(emb_size_atom, num_elements)
        self.edge_emb = EdgeEmbedding(emb_size_atom, num_radial, emb_size_edge)
        self.mlp_rbf_out = Dense(num_radial, 16)  # shared down-projection of the radial basis for the output blocks (gemnet_oc.py:1112)
        self.out_blocks = nn.ModuleList([OutputBlock(emb_size_atom, emb_size_edge, 16, 3, 3)])  # block 0 only so far

    def forward(self, z, pos, batch):
        g = build_all_indices(pos, batch)
        rbf = self.radial_basis(g["main"]["distance"])
        h = self.atom_emb(z)
        m = self.edge_emb(h, rbf, g["main"]["edge_index"])
        x_E, x_F = self.out_blocks[0](h, m, self.mlp_rbf_out(rbf), g["main"]["edge_index"][1])
        return g, rbf, h, m, x_E, x_F
