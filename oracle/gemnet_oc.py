"""CPU restatement of GemNet-OC.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The whole forward of config/model/gemnet-oc.yaml (non-periodic, direct coupled forces), PINNED against the energies, forces and per-block
intermediates recorded from the reference's own classes (tests/golden/gemnet_oc_f32.npz, tests/golden/make_golden_gemnet_oc.py):
    graphs and all index structures            oracle/gemnet_graph.py
    radial / circular / spherical bases        gemnet_oc/layers/radial_basis.py:19-39,57-77,176-220; spherical_basis.py:18-127; basis.py:84-106,273-295
    shared basis embeddings, bilinear layers   layers/efficient.py:15-253
    embedding, output, atom-update blocks      layers/embedding_block.py:14-92; atom_update_block.py:15-172; base_layers.py:15-97
    interaction block (triplet / quadruplet / atom-edge / edge-atom / atom-atom)   layers/interaction_block.py:19-739
    forward, angles, coupled force assembly    gemnet_oc.py:596-656,1001-1120,1121-1251
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


class CircularBasis(nn.Module):
    """Y_l0(z) = sqrt((2l+1)/(4 pi)) P_l(z), l = 0..L-1 (`get_sph_harm_basis(L, zero_m_only=True)`, basis.py:84-106,273-295)."""

    def __init__(self, num_spherical=7):
        super().__init__()
        self.L = num_spherical

    def forward(self, z):
        P = [torch.ones_like(z), z]
        for l in range(1, self.L - 1):
            P.append(((2 * l + 1) * z * P[l] - l * P[l - 1]) / (l + 1))
        return torch.stack([math.sqrt((2 * l + 1) / (4 * math.pi)) * P[l] for l in range(self.L)], dim=1)


class CircularBasisLayer(nn.Module):
    def __init__(self, num_spherical, radial_basis):
        super().__init__()
        self.radial_basis, self.scale_cbf, self.cos_basis = radial_basis, _Scale(), CircularBasis(num_spherical)

    def forward(self, d, cos_phi):
        return self.radial_basis(d), self.scale_cbf(self.cos_basis(cos_phi))


class SphericalBasisLayer(nn.Module):  # sbf "legendre_outer": outer product of the circular basis of cos(phi) and of cos(theta)
    def __init__(self, num_spherical, radial_basis):
        super().__init__()
        self.radial_basis, self.scale_sbf, self.cos_basis = radial_basis, _Scale(), CircularBasis(num_spherical)

    def forward(self, d, cos_phi, theta):
        sph = (self.cos_basis(cos_phi)[:, :, None] * self.cos_basis(torch.cos(theta))[:, None, :]).reshape(cos_phi.shape[0], -1)
        return self.radial_basis(d), self.scale_sbf(sph)


class BasisEmbedding(nn.Module):
    """efficient.py:15-140: radial (x spherical) basis -> interaction embedding, optionally scattered into padded per-edge / per-atom
    matrices so that the later aggregation is a matmul."""

    def __init__(self, num_radial, emb_size_interm, num_spherical=None):
        super().__init__()
        self.num_spherical = num_spherical
        shape = (emb_size_interm, num_radial) if num_spherical is None else (num_radial, num_spherical, emb_size_interm)
        self.weight = nn.Parameter(torch.zeros(shape))

    def forward(self, rad_basis, sph_basis=None, idx_rad_outer=None, idx_rad_inner=None, idx_sph_outer=None, idx_sph_inner=None, num_atoms=None):
        n_edges = rad_basis.shape[0]
        if self.num_spherical is not None:
            rad_W1 = (rad_basis @ self.weight.reshape(self.weight.shape[0], -1)).reshape(n_edges, -1, sph_basis.shape[-1])
        else:
            rad_W1 = rad_basis @ self.weight.T
        if idx_rad_inner is not None:
            kmax = int(idx_rad_inner.max()) + 1 if idx_rad_outer.shape[0] else 0
            pad = rad_W1.new_zeros([num_atoms, kmax] + list(rad_W1.shape[1:]))
            pad[idx_rad_outer, idx_rad_inner] = rad_W1
            rad_W1 = pad.transpose(1, 2).reshape(num_atoms, rad_W1.shape[1], -1)
        if idx_sph_inner is not None:
            kmax = int(idx_sph_inner.max()) + 1 if idx_sph_outer.shape[0] else 0
            sph2 = sph_basis.new_zeros(n_edges, kmax, sph_basis.shape[-1])
            sph2[idx_sph_outer, idx_sph_inner] = sph_basis
            sph2 = sph2.transpose(1, 2)
        if sph_basis is None:
            return rad_W1
        if idx_sph_inner is None:
            return (rad_W1[idx_sph_outer] @ sph_basis[:, :, None]).squeeze(-1)
        return rad_W1, sph2


class EfficientInteractionBilinear(nn.Module):
    def __init__(self, emb_size_in, emb_size_interm, emb_size_out):  # efficient.py:143-253
        super().__init__()
        self.emb_size_in = emb_size_in
        self.bilinear = Dense(emb_size_in * emb_size_interm, emb_size_out)

    def forward(self, basis, m, idx_agg_outer, idx_agg_inner, idx_agg2_outer=None, idx_agg2_inner=None, agg2_out_size=None):
        rad_W1, sph = basis
        n_edges = sph.shape[0]
        kmax = int(idx_agg_inner.max()) + 1
        m_pad = m.new_zeros(n_edges, kmax, self.emb_size_in)
        m_pad[idx_agg_outer, idx_agg_inner] = m
        sph_m = sph @ m_pad
        if idx_agg2_outer is not None:
            kmax2 = int(idx_agg2_inner.max()) + 1
            pad = sph_m.new_zeros(agg2_out_size, kmax2, sph_m.shape[1], sph_m.shape[2])
            pad[idx_agg2_outer, idx_agg2_inner] = sph_m
            out = rad_W1 @ pad.reshape(agg2_out_size, -1, sph_m.shape[-1])
        else:
            out = rad_W1 @ sph_m
        return self.bilinear(out.reshape(-1, out.shape[1:].numel()))


class AtomUpdateBlock(nn.Module):
    def __init__(self, emb_size_atom, emb_size_edge, emb_size_rbf, n_hidden):  # atom_update_block.py:15-91
        super().__init__()
        self.dense_rbf, self.scale_sum = Dense(emb_size_rbf, emb_size_edge), _Scale()
        self.layers = _mlp(emb_size_edge, emb_size_atom, n_hidden)

    def forward(self, h, m, basis_rad, idx_atom):
        x = m * self.dense_rbf(basis_rad)
        x = self.scale_sum(torch.zeros(h.shape[0], x.shape[1], dtype=x.dtype).index_add_(0, idx_atom, x))
        for layer in self.layers:
            x = layer(x)
        return x


_ISQ2 = 1 / math.sqrt(2.0)


class TripletInteraction(nn.Module):
    def __init__(self, emb_in, emb_out, emb_trip_in, emb_trip_out, emb_rbf, emb_cbf, symmetric_mp=True, swap_output=True):
        super().__init__()
        self.symmetric_mp, self.swap_output = symmetric_mp, swap_output
        self.dense_ba = Dense(emb_in, emb_in, activation="silu")
        self.mlp_rbf, self.scale_rbf = Dense(emb_rbf, emb_in), _Scale()
        self.mlp_cbf, self.scale_cbf_sum = EfficientInteractionBilinear(emb_trip_in, emb_cbf, emb_trip_out), _Scale()
        self.down_projection = Dense(emb_in, emb_trip_in, activation="silu")
        self.up_projection_ca = Dense(emb_trip_out, emb_out, activation="silu")
        if symmetric_mp:
            self.up_projection_ac = Dense(emb_trip_out, emb_out, activation="silu")

    def forward(self, m, bases, idx, id_swap, expand_idx=None, idx_agg2=None, idx_agg2_inner=None, agg2_out_size=None):
        x = self.dense_ba(m)
        if expand_idx is not None:
            x = x[expand_idx]
        x = self.down_projection(self.scale_rbf(x * self.mlp_rbf(bases["rad"])))[idx["in"]]
        x = self.scale_cbf_sum(self.mlp_cbf(bases["cir"], x, idx["out"], idx["out_agg"], idx_agg2, idx_agg2_inner, agg2_out_size))
        if self.symmetric_mp:
            return (self.up_projection_ca(x) + self.up_projection_ac(x)[id_swap]) * _ISQ2
        if self.swap_output:
            x = x[id_swap]
        return self.up_projection_ca(x)


class QuadrupletInteraction(nn.Module):
    def __init__(self, emb_edge, emb_quad_in, emb_quad_out, emb_rbf, emb_cbf, emb_sbf):
        super().__init__()
        self.dense_db = Dense(emb_edge, emb_edge, activation="silu")
        self.mlp_rbf, self.scale_rbf = Dense(emb_rbf, emb_edge), _Scale()
        self.mlp_cbf, self.scale_cbf = Dense(emb_cbf, emb_quad_in), _Scale()
        self.mlp_sbf, self.scale_sbf_sum = EfficientInteractionBilinear(emb_quad_in, emb_sbf, emb_quad_out), _Scale()
        self.down_projection = Dense(emb_edge, emb_quad_in, activation="silu")
        self.up_projection_ca = Dense(emb_quad_out, emb_edge, activation="silu")
        self.up_projection_ac = Dense(emb_quad_out, emb_edge, activation="silu")

    def forward(self, m, bases, idx, id_swap):
        x = self.dense_db(m)
        x = self.down_projection(self.scale_rbf(x * self.mlp_rbf(bases["rad"])))[idx["triplet_in"]["in"]]
        x = self.scale_cbf(x * self.mlp_cbf(bases["cir"]))[idx["trip_in_to_quad"]]
        x = self.scale_sbf_sum(self.mlp_sbf(bases["sph"], x, idx["out"], idx["out_agg"]))
        return (self.up_projection_ca(x) + self.up_projection_ac(x)[id_swap]) * _ISQ2


class PairInteraction(nn.Module):
    def __init__(self, emb_atom, emb_pair_in, emb_pair_out, emb_rbf):
        super().__init__()
        self.bilinear, self.scale_rbf_sum = Dense(emb_rbf * emb_pair_in, emb_pair_out), _Scale()
        self.down_projection = Dense(emb_atom, emb_pair_in, activation="silu")
        self.up_projection = Dense(emb_pair_out, emb_atom, activation="silu")

    def forward(self, h, rad_basis, edge_index, target_neighbor_idx):
        n = h.shape[0]
        x_ba = self.down_projection(h)[edge_index[0]]
        pad = x_ba.new_zeros(n, int(target_neighbor_idx.max()) + 1, x_ba.shape[-1])
        pad[edge_index[1], target_neighbor_idx] = x_ba
        return self.up_projection(self.scale_rbf_sum(self.bilinear((rad_basis @ pad).reshape(n, -1))))


class InteractionBlock(nn.Module):
    def __init__(self, ea=256, ee=512, trip_in=64, trip_out=64, quad_in=32, quad_out=32, a2a_in=64, a2a_out=64, rbf=16, cbf=16, sbf=32,
                 num_before_skip=2, num_after_skip=2, num_concat=1, num_atom=3):
        super().__init__()
        self.dense_ca = Dense(ee, ee, activation="silu")
        self.trip_interaction = TripletInteraction(ee, ee, trip_in, trip_out, rbf, cbf)
        self.quad_interaction = QuadrupletInteraction(ee, quad_in, quad_out, rbf, cbf, sbf)
        self.atom_edge_interaction = TripletInteraction(ea, ee, trip_in, trip_out, rbf, cbf)
        self.edge_atom_interaction = TripletInteraction(ee, ea, trip_in, trip_out, rbf, cbf, symmetric_mp=False, swap_output=False)
        self.atom_interaction = PairInteraction(ea, a2a_in, a2a_out, rbf)
        self.layers_before_skip = nn.ModuleList([ResidualLayer(ee) for _ in range(num_before_skip)])
        self.layers_after_skip = nn.ModuleList([ResidualLayer(ee) for _ in range(num_after_skip)])
        self.atom_emb_layers = nn.ModuleList([])
        self.atom_update = AtomUpdateBlock(ea, ee, rbf, num_atom)
        self.concat_layer = EdgeEmbedding(ea, ee, ee)
        self.residual_m = nn.ModuleList([ResidualLayer(ee) for _ in range(num_concat)])

    def forward(self, h, m, B, g):
        n = h.shape[0]
        main_ei, a2ee2a, a2a, id_swap = g["main"]["edge_index"], g["a2ee2a"], g["a2a"], g["id_swap"]
        x = self.dense_ca(m) + self.trip_interaction(m, B["e2e"], g["trip_e2e"], id_swap)
        x = x + self.quad_interaction(m, B["qint"], g["quad"], id_swap)
        x = x + self.atom_edge_interaction(h, B["a2e"], g["trip_a2e"], id_swap, expand_idx=a2ee2a["edge_index"][0])
        x = x * (1 / math.sqrt(4.0))
        h_e2a = self.edge_atom_interaction(m, B["e2a"], g["trip_e2a"], id_swap, idx_agg2=a2ee2a["edge_index"][1],
                                           idx_agg2_inner=a2ee2a["target_neighbor_idx"], agg2_out_size=n)
        h_a2a = self.atom_interaction(h, B["a2a_rad"], a2a["edge_index"], a2a["target_neighbor_idx"])
        h = (h + h_e2a + h_a2a) * (1 / math.sqrt(3.0))
        for layer in self.layers_before_skip:
            x = layer(x)
        m = (m + x) * _ISQ2
        for layer in self.layers_after_skip:
            m = layer(m)
        h = (h + self.atom_update(h, m, B["atom_update"], main_ei[1])) * _ISQ2
        m2 = self.concat_layer(h, m, main_ei)
        for layer in self.residual_m:
            m2 = layer(m2)
        return h, (m + m2) * _ISQ2


def _clamped_dot(x, y):
    return (x * y).sum(-1).clamp(min=-1, max=1)


class GemNetOCOracle(nn.Module):
    """config/model/gemnet-oc.yaml.  forward(z, pos, batch) -> (energy [B], forces [N, 3]) plus the per-block intermediates in `.trace`."""

    def __init__(self, num_spherical=7, num_radial=128, num_blocks=4, ea=256, ee=512, rbf=16, cbf=16, sbf=32, cutoff=12.0, num_elements=83):
        super().__init__()
        rb = lambda: RadialBasis(num_radial, cutoff)
        self.radial_basis = rb()
        shared_sph = rb()                                            # `radial_basis_spherical`: one instance under three parents
        self.cbf_basis_qint = CircularBasisLayer(num_spherical, rb())
        self.sbf_basis_qint = SphericalBasisLayer(num_spherical, shared_sph)
        self.radial_basis_aeaint = rb()
        self.cbf_basis_aeint = CircularBasisLayer(num_spherical, shared_sph)
        self.cbf_basis_eaint = CircularBasisLayer(num_spherical, rb())
        self.radial_basis_aint = rb()
        self.cbf_basis_tint = CircularBasisLayer(num_spherical, shared_sph)
        self.mlp_rbf_qint = Dense(num_radial, rbf)
        self.mlp_cbf_qint = BasisEmbedding(num_radial, cbf, num_spherical)
        self.mlp_sbf_qint = BasisEmbedding(num_radial, sbf, num_spherical ** 2)
        self.mlp_rbf_aeint = Dense(num_radial, rbf)
        self.mlp_cbf_aeint = BasisEmbedding(num_radial, cbf, num_spherical)
        self.mlp_rbf_eaint = Dense(num_radial, rbf)
        self.mlp_cbf_eaint = BasisEmbedding(num_radial, cbf, num_spherical)
        self.mlp_rbf_aint = BasisEmbedding(num_radial, rbf)
        self.mlp_rbf_tint = Dense(num_radial, rbf)
        self.mlp_cbf_tint = BasisEmbedding(num_radial, cbf, num_spherical)
        self.mlp_rbf_h = Dense(num_radial, rbf)
        self.mlp_rbf_out = Dense(num_radial, rbf)
        self.atom_emb = AtomEmbedding(ea, num_elements)
        self.edge_emb = EdgeEmbedding(ea, num_radial, ee)
        self.int_blocks = nn.ModuleList([InteractionBlock(ea, ee, rbf=rbf, cbf=cbf, sbf=sbf) for _ in range(num_blocks)])
        self.out_blocks = nn.ModuleList([OutputBlock(ea, ee, rbf, 3, 3) for _ in range(num_blocks + 1)])
        self.out_mlp_E = nn.Sequential(Dense(ea * (num_blocks + 1), ea, activation="silu"), ResidualLayer(ea), ResidualLayer(ea))
        self.out_energy = Dense(ea, 1)
        self.out_mlp_F = nn.Sequential(Dense(ee * (num_blocks + 1), ee, activation="silu"), ResidualLayer(ee), ResidualLayer(ee))
        self.out_forces = Dense(ee, 1)

    def bases(self, g, n_atoms):  # gemnet_oc.py:1001-1120
        main, a2a, a2ee2a, qint, q = g["main"], g["a2a"], g["a2ee2a"], g["qint"], g["quad"]
        V, Vq = main["vector"], qint["vector"]
        rad_main = self.radial_basis(main["distance"])
        rad_cir_e2e, cir_e2e = self.cbf_basis_tint(main["distance"], _clamped_dot(V[g["trip_e2e"]["out"]], V[g["trip_e2e"]["in"]]))
        # quadruplet angles (gemnet_oc.py:596-656)
        V_ba, V_db = Vq[q["triplet_in"]["out"]], V[q["triplet_in"]["in"]]
        cos_abd = _clamped_dot(V_ba, V_db)
        V_db_cross = torch.cross(V_db, V_ba, dim=-1)[q["trip_in_to_quad"]]
        V_ca, V_ba2 = V[q["triplet_out"]["out"]], Vq[q["triplet_out"]["in"]]
        cos_cab_q = _clamped_dot(V_ca, V_ba2)
        V_ca_cross = torch.cross(V_ca, V_ba2, dim=-1)[q["trip_out_to_quad"]]
        x = (V_ca_cross * V_db_cross).sum(-1)
        y = torch.cross(V_ca_cross, V_db_cross, dim=-1).norm(dim=-1).clamp(min=1e-9)
        angle_cabd = torch.atan2(y, x)
        rad_cir_q, cir_q = self.cbf_basis_qint(qint["distance"], cos_abd)
        rad_sph_q, sph_q = self.sbf_basis_qint(main["distance"], cos_cab_q[q["trip_out_to_quad"]], angle_cabd)
        rad_a2ee2a = self.radial_basis_aeaint(a2ee2a["distance"])
        rad_cir_a2e, cir_a2e = self.cbf_basis_aeint(main["distance"], _clamped_dot(V[g["trip_a2e"]["out"]], a2ee2a["vector"][g["trip_a2e"]["in"]]))
        rad_cir_e2a, cir_e2a = self.cbf_basis_eaint(a2ee2a["distance"], _clamped_dot(a2ee2a["vector"][g["trip_e2a"]["out"]], V[g["trip_e2a"]["in"]]))
        rad_a2a = self.radial_basis_aint(a2a["distance"])
        B = {"qint": {"rad": self.mlp_rbf_qint(rad_main),
                      "cir": self.mlp_cbf_qint(rad_basis=rad_cir_q, sph_basis=cir_q, idx_sph_outer=q["triplet_in"]["out"]),
                      "sph": self.mlp_sbf_qint(rad_basis=rad_sph_q, sph_basis=sph_q, idx_sph_outer=q["out"], idx_sph_inner=q["out_agg"])},
             "a2e": {"rad": self.mlp_rbf_aeint(rad_a2ee2a),
                     "cir": self.mlp_cbf_aeint(rad_basis=rad_cir_a2e, sph_basis=cir_a2e, idx_sph_outer=g["trip_a2e"]["out"],
                                               idx_sph_inner=g["trip_a2e"]["out_agg"])},
             "e2a": {"rad": self.mlp_rbf_eaint(rad_main),
                     "cir": self.mlp_cbf_eaint(rad_basis=rad_cir_e2a, sph_basis=cir_e2a, idx_rad_outer=a2ee2a["edge_index"][1],
                                               idx_rad_inner=a2ee2a["target_neighbor_idx"], idx_sph_outer=g["trip_e2a"]["out"],
                                               idx_sph_inner=g["trip_e2a"]["out_agg"], num_atoms=n_atoms)},
             "a2a_rad": self.mlp_rbf_aint(rad_basis=rad_a2a, idx_rad_outer=a2a["edge_index"][1], idx_rad_inner=a2a["target_neighbor_idx"],
                                          num_atoms=n_atoms),
             "e2e": {"rad": self.mlp_rbf_tint(rad_main),
                     "cir": self.mlp_cbf_tint(rad_basis=rad_cir_e2e, sph_basis=cir_e2e, idx_sph_outer=g["trip_e2e"]["out"],
                                              idx_sph_inner=g["trip_e2e"]["out_agg"])},
             "atom_update": self.mlp_rbf_h(rad_main), "output": self.mlp_rbf_out(rad_main)}
        return rad_main, B

    def forward(self, z, pos, batch):
        g = build_all_indices(pos, batch)
        n = z.shape[0]
        rad_main, B = self.bases(g, n)
        ei = g["main"]["edge_index"]
        h = self.atom_emb(z)
        m = self.edge_emb(h, rad_main, ei)
        self.trace = {"atom_emb/h": h, "edge_emb/m": m}
        x_E, x_F = self.out_blocks[0](h, m, B["output"], ei[1])
        xs_E, xs_F = [x_E], [x_F]
        self.trace["out0/x_E"], self.trace["out0/x_F"] = x_E, x_F
        for i, blk in enumerate(self.int_blocks):
            h, m = blk(h, m, B, g)
            x_E, x_F = self.out_blocks[i + 1](h, m, B["output"], ei[1])
            xs_E.append(x_E); xs_F.append(x_F)
            self.trace.update({f"int{i}/h": h, f"int{i}/m": m, f"out{i + 1}/x_E": x_E, f"out{i + 1}/x_F": x_F})
        E_atom = self.out_energy(self.out_mlp_E(torch.cat(xs_E, dim=-1)))
        F_st = self.out_forces(self.out_mlp_F(torch.cat(xs_F, dim=-1)))
        n_mol = int(batch.max()) + 1
        E = torch.zeros(n_mol, 1, dtype=E_atom.dtype).index_add_(0, batch, E_atom).squeeze(1)
        # coupled forces (gemnet_oc.py:1217-1242): edge k of a molecule's directed half and its flipped copy share the mean
        rev = g["id_swap"]
        F_st = 0.5 * (F_st + F_st[rev])
        F = torch.zeros(n, 3, dtype=F_st.dtype).index_add_(0, ei[1], F_st * g["main"]["vector"])
        return E, F
