"""Oracle: PaiNN-OC (`config/painn-oc.yaml`) -- TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

CPU restatement of `nablaDFT/painn_pyg/painn.py` + `layers.py`; parameter names and shapes
are the reference's, so a reference state_dict loads with `strict=True`.

Pinned: `tests/golden/painn_oc_*.npz` hold outputs of the reference's own classes run in the
build container (generator: `tests/golden/make_golden_painn_oc.py`).

Line map (reference -> here):
  painn.py:89-148   PaiNN.forward                      -> PaiNNOC.forward
  painn.py:306-349  generate_graph_values               -> _graph (radius graph, unit vectors)
  painn.py:168-304  symmetrize_edges                    -> omitted: with uncapped neighbour
        lists (max degree 42 < max_neighbors 100) it returns the same multiset of
        (j, i, dist, vec) edges in another order; sums are order-independent
        up to fp rounding, which the golden test bounds.
  painn.py:475-509  PaiNNMessage.forward/message/aggr   -> MessageOC
  painn.py:535-548  PaiNNUpdate.forward                 -> UpdateOC
  layers.py:14-33   PolynomialEnvelope                  -> poly_envelope
  layers.py:129-185 RadialBasis (+ PyG GaussianSmearing)-> RadialBasisOC
  layers.py:198-222 AtomEmbedding                       -> AtomEmbeddingOC
"""
import math

import torch
from torch import nn

from .graph import radius_graph


def poly_envelope(d_scaled: torch.Tensor, p: float = 5.0) -> torch.Tensor:
    a = -(p + 1) * (p + 2) / 2
    b = p * (p + 2)
    c = -p * (p + 1) / 2
    env = 1 + a * d_scaled**p + b * d_scaled ** (p + 1) + c * d_scaled ** (p + 2)
    return torch.where(d_scaled < 1, env, torch.zeros_like(d_scaled))


class _GaussianSmearing(nn.Module):
    """PyG `GaussianSmearing(start, stop, num_gaussians)` semantics."""

    def __init__(self, start=0.0, stop=1.0, num_gaussians=100):
        super().__init__()
        offset = torch.linspace(start, stop, num_gaussians)
        self.coeff = -0.5 / (offset[1] - offset[0]).item() ** 2
        self.register_buffer("offset", offset)

    def forward(self, dist):
        dist = dist.view(-1, 1) - self.offset.view(1, -1)
        return torch.exp(self.coeff * torch.pow(dist, 2))


class RadialBasisOC(nn.Module):
    def __init__(self, num_radial: int, cutoff: float, exponent: int = 5):
        super().__init__()
        self.inv_cutoff = 1 / cutoff
        self.exponent = float(exponent)
        self.rbf = _GaussianSmearing(0.0, 1.0, num_radial)

    def forward(self, d):
        d_scaled = d * self.inv_cutoff
        env = poly_envelope(d_scaled, self.exponent)
        return env[:, None] * self.rbf(d_scaled)


class AtomEmbeddingOC(nn.Module):
    def __init__(self, emb_size: int, num_elements: int):
        super().__init__()
        self.embeddings = nn.Embedding(num_elements, emb_size)
        nn.init.uniform_(self.embeddings.weight, a=-math.sqrt(3), b=math.sqrt(3))

    def forward(self, z):
        return self.embeddings(z - 1)


class MessageOC(nn.Module):
    def __init__(self, hidden: int, num_rbf: int):
        super().__init__()
        self.hidden_channels = hidden
        self.x_proj = nn.Sequential(nn.Linear(hidden, hidden), nn.SiLU(), nn.Linear(hidden, hidden * 3))
        self.rbf_proj = nn.Linear(num_rbf, hidden * 3)
        for lin in (self.x_proj[0], self.x_proj[2], self.rbf_proj):
            nn.init.xavier_uniform_(lin.weight)
            lin.bias.data.fill_(0)

    def forward(self, x, vec, edge_index, edge_rbf, edge_vector):
        xh = self.x_proj(x)
        rbfh = self.rbf_proj(edge_rbf)
        j, i = edge_index  # PyG flow source_to_target: *_j = x[edge_index[0]], aggregate at [1]
        s, xh2, xh3 = torch.split(xh[j] * rbfh, self.hidden_channels, dim=-1)
        v = vec[j] * xh2.unsqueeze(1) + xh3.unsqueeze(1) * edge_vector.unsqueeze(2)
        dx = torch.zeros_like(x).index_add_(0, i, s)
        dvec = torch.zeros_like(vec).index_add_(0, i, v)
        return dx, dvec


class UpdateOC(nn.Module):
    def __init__(self, hidden: int):
        super().__init__()
        self.hidden_channels = hidden
        self.vec_proj = nn.Linear(hidden, hidden * 2, bias=False)
        self.xvec_proj = nn.Sequential(nn.Linear(hidden * 2, hidden), nn.SiLU(), nn.Linear(hidden, hidden * 3))
        nn.init.xavier_uniform_(self.vec_proj.weight)
        for lin in (self.xvec_proj[0], self.xvec_proj[2]):
            nn.init.xavier_uniform_(lin.weight)
            lin.bias.data.fill_(0)

    def forward(self, x, vec):
        vec1, vec2 = torch.split(self.vec_proj(vec), self.hidden_channels, dim=-1)
        vec_dot = (vec1 * vec2).sum(dim=1)
        h = self.xvec_proj(torch.cat([x, torch.sqrt(torch.sum(vec2**2, dim=-2) + 1e-8)], dim=-1))
        h1, h2, h3 = torch.split(h, self.hidden_channels, dim=-1)
        return h1 + h2 * vec_dot, h3.unsqueeze(1) * vec1


class PaiNNOC(nn.Module):
    """`nablaDFT.painn_pyg.PaiNN` with the `config/model/painn-oc.yaml` flags
    (regress_forces=True, direct_forces=False, use_pbc=False, otf_graph=True)."""

    def __init__(self, hidden_channels=128, num_layers=6, num_rbf=100, cutoff=5.0, max_neighbors=100, num_elements=100):
        super().__init__()
        self.hidden_channels, self.num_layers, self.num_rbf = hidden_channels, num_layers, num_rbf
        self.cutoff, self.max_neighbors = cutoff, max_neighbors
        self.atom_emb = AtomEmbeddingOC(hidden_channels, num_elements)
        self.radial_basis = RadialBasisOC(num_rbf, cutoff)
        self.message_layers = nn.ModuleList(MessageOC(hidden_channels, num_rbf) for _ in range(num_layers))
        self.update_layers = nn.ModuleList(UpdateOC(hidden_channels) for _ in range(num_layers))
        self.out_energy = nn.Sequential(
            nn.Linear(hidden_channels, hidden_channels // 2), nn.SiLU(), nn.Linear(hidden_channels // 2, 1)
        )
        for lin in (self.out_energy[0], self.out_energy[2]):
            nn.init.xavier_uniform_(lin.weight)
            lin.bias.data.fill_(0)

    def _graph(self, pos, batch):
        edge_index = radius_graph(pos, self.cutoff, batch, self.max_neighbors)
        j, i = edge_index
        distance_vec = pos[j] - pos[i]
        edge_dist = (pos[i] - pos[j]).pow(2).sum(dim=-1).sqrt()
        mask_zero = torch.isclose(edge_dist, torch.zeros((), dtype=edge_dist.dtype), atol=1e-6).to(pos.dtype) * 1e-6
        edge_vector = distance_vec / (edge_dist + mask_zero)[:, None]
        return edge_index, edge_dist, edge_vector

    @torch.enable_grad()
    def forward(self, z, pos, batch, create_graph=False, return_intermediates=False):
        pos = pos.requires_grad_(True)
        z = z.long()
        edge_index, edge_dist, edge_vector = self._graph(pos, batch)
        edge_rbf = self.radial_basis(edge_dist)
        x = self.atom_emb(z)
        vec = torch.zeros(x.size(0), 3, x.size(1), dtype=x.dtype)
        inter = []
        for l in range(self.num_layers):
            dx, dvec = self.message_layers[l](x, vec, edge_index, edge_rbf, edge_vector)
            x, vec = x + dx, vec + dvec
            dx, dvec = self.update_layers[l](x, vec)
            x, vec = x + dx, vec + dvec
            if return_intermediates:
                inter.append((x.detach().clone(), vec.detach().clone()))
        per_atom = self.out_energy(x).squeeze(1)
        n_mol = int(batch.max().item()) + 1
        energy = torch.zeros(n_mol, dtype=x.dtype).index_add_(0, batch, per_atom)
        forces = -torch.autograd.grad(energy, pos, grad_outputs=torch.ones_like(energy), create_graph=create_graph)[0]
        if return_intermediates:
            return energy, forces, inter
        return energy, forces
