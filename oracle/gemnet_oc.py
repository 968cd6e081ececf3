"""CPU restatement of GemNet-OC, in progress.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Built and PINNED so far (against intermediates recorded from the reference's own classes, tests/golden/gemnet_oc_f32.npz):
    graphs and all index structures      oracle/gemnet_graph.py
    radial basis, atom / edge embedding  this file  (gemnet_oc/layers/radial_basis.py:19-39,57-77,176-220; embedding_block.py:14-92;
                                         base_layers.py:15-75; gemnet_oc.py:1165-1167)
Not restated yet: circular / spherical bases, the interaction blocks, the output blocks.
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


class GemNetOCStem(nn.Module):
    """Graphs -> radial basis -> h0, m0 (gemnet_oc.py:1121-1167).  The rest of the network follows in the next round."""

    def __init__(self, num_radial=128, cutoff=12.0, emb_size_atom=256, emb_size_edge=512, num_elements=83):
        super().__init__()
        self.radial_basis = RadialBasis(num_radial, cutoff)
        self.atom_emb = AtomEmbedding(emb_size_atom, num_elements)
        self.edge_emb = EdgeEmbedding(emb_size_atom, num_radial, emb_size_edge)

    def forward(self, z, pos, batch):
        g = build_all_indices(pos, batch)
        rbf = self.radial_basis(g["main"]["distance"])
        h = self.atom_emb(z)
        m = self.edge_emb(h, rbf, g["main"]["edge_index"])
        return g, rbf, h, m
