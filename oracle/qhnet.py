"""Oracle: QHNet Hamiltonian prediction (`config/model/qhnet.yaml`) -- TEST INFRASTRUCTURE ONLY.

CPU restatement of `nablaDFT/qhnet/qhnet.py` + `layers.py` on top of `oracle.e3` (the e3nn
primitives).  Parameter names / shapes follow the reference so its state_dict loads.
Pinned: `tests/golden/qhnet_f64.npz` holds outputs of the reference's OWN classes executed in the
build container with `oracle.e3` standing in for the absent e3nn wheel
(tests/golden/make_golden_qhnet.py) -- this pins the in-repo model code; e3nn itself stays
[3P-memory] (oracle/e3.py header).

Line map (reference -> here):
  layers.py:44-83    get_feasible_irrep                 -> feasible_irrep
  layers.py:86-120   cutoff_function, ExpBernstein RBF   -> ExpBernstein
  layers.py:123-147  NormGate                            -> NormGate
  layers.py:150-274  ConvLayer                           -> ConvLayer
  layers.py:277-294  InnerProduct                        -> inner_product
  layers.py:297-343  ConvNetLayer                        -> (residual inside QHNetOracle.forward)
  layers.py:346-492  PairNetLayer                        -> PairNetLayer
  layers.py:495-582  SelfNetLayer                        -> SelfNetLayer
  layers.py:585-681  Expansion                           -> Expansion
  qhnet.py:186-252   QHNet.forward                       -> QHNetOracle.forward
  qhnet.py:254-291   build_graph                         -> QHNetOracle.build_graph
  qhnet.py:293-321   build_final_matrix                  -> assemble (vectorised; same block placement)
  qhnet.py:323-342   _get_mask                           -> orbital_masks
"""
import math
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .e3 import ElementwiseTensorProduct, FullyConnectedNet, Irreps, Linear, Norm, TensorProduct, spherical_harmonics, wigner_3j
from .graph import radius_graph


def ssp(x):
    return F.softplus(x) - math.log(2.0)


def feasible_irrep(in1: Irreps, in2: Irreps, cutoff_out: Irreps, mode="uvu"):
    mid, ins = [], []
    for i, (_, ir1) in enumerate(in1):
        for j, (_, ir2) in enumerate(in2):
            for ir3 in ir1 * ir2:
                if ir3 in cutoff_out:
                    key = (cutoff_out.count(ir3), ir3)
                    if key not in mid:
                        mid.append(key)
                    ins.append((i, j, mid.index(key), mode, True))
    mid = Irreps(mid)
    nel = {"uvu": lambda a: in2[a[1]].mul, "uuu": lambda a: 1}[mode]
    total = sum(nel(a) for a in ins)  # the reference sums over ALL instructions (layers.py:73)
    alphas = [math.sqrt(mid[a[2]].ir.dim / total if total > 0 else mid[a[2]].ir.dim) for a in ins]
    mid, p, _ = mid.sort()
    return mid, [(a[0], a[1], p[a[2]], a[3], a[4], al) for a, al in zip(ins, alphas)]


class ExpBernstein(nn.Module):
    def __init__(self, k, cutoff, ini_alpha=0.5):
        super().__init__()
        logfact = np.zeros(k)
        for i in range(2, k):
            logfact[i] = logfact[i - 1] + np.log(i)
        v = np.arange(k)
        n = (k - 1) - v
        self.register_buffer("cutoff", torch.tensor(cutoff, dtype=torch.float32))
        self.register_buffer("logc", torch.tensor(logfact[-1] - logfact[v] - logfact[n], dtype=torch.float32))
        self.register_buffer("n", torch.tensor(n, dtype=torch.float32))
        self.register_buffer("v", torch.tensor(v, dtype=torch.float32))
        a = torch.tensor(ini_alpha)
        self._alpha = nn.Parameter((a + torch.log(-torch.expm1(-a))).float())  # softplus^-1

    def forward(self, r):  # r [E,1]
        alpha = F.softplus(self._alpha)
        x = -alpha * r
        x = self.logc + self.n * x + self.v * torch.log(-torch.expm1(x))
        rc = self.cutoff
        r_ = torch.where(r < rc, r, torch.zeros_like(r))
        fcut = torch.where(r < rc, torch.exp(-(r_**2) / ((rc - r_) * (rc + r_))), torch.zeros_like(r))
        return fcut * torch.exp(x)


def inner_product(irreps: Irreps, a, b):
    """InnerProduct: per channel sum_m a b / (2l+1) -> one scalar per channel of every block."""
    B = a.shape[0]
    return torch.cat([(a[:, s].reshape(B, m.mul, m.ir.dim) * b[:, s].reshape(B, m.mul, m.ir.dim)).sum(-1) / m.ir.dim
                      for s, m in zip(irreps.slices(), irreps)], dim=-1)


class NormGate(nn.Module):
    def __init__(self, irreps: Irreps):
        super().__init__()
        self.irrep = irreps
        self.norm = Norm(irreps)
        n_all = sum(m.mul for m in irreps)
        n_wo0 = sum(m.mul for m in irreps if m.ir.l != 0)
        self.mul = ElementwiseTensorProduct(irreps[1:], Irreps(f"{n_wo0}x0e"))
        self.fc = nn.Sequential(nn.Linear(n_all, n_all), nn.SiLU(), nn.Linear(n_all, n_all))

    def forward(self, x):
        s0 = self.irrep.slices()[0]
        gates = self.fc(torch.cat([x[:, s0], self.norm(x)[:, s0.stop:]], dim=-1))
        return torch.cat([gates[:, s0], self.mul(x[:, s0.stop:], gates[:, s0.stop:])], dim=-1)


class ConvLayer(nn.Module):
    def __init__(self, irrep_in, irrep_hidden, irrep_out, sh_irrep, edge_attr_dim, use_norm_gate=True, invariant_neurons=32):
        super().__init__()
        self.irrep_in_node, self.irrep_out, self.use_norm_gate = irrep_in, irrep_out, use_norm_gate
        self.irrep_tp_out_node, ins = feasible_irrep(irrep_in, sh_irrep, irrep_hidden, "uvu")
        self.tp_node = TensorProduct(irrep_in, sh_irrep, self.irrep_tp_out_node, ins, shared_weights=False, internal_weights=False)
        self.fc_node = FullyConnectedNet([edge_attr_dim, invariant_neurons, self.tp_node.weight_numel], ssp)
        n_mul = sum(m.mul for m in irrep_in)
        self.layer_l0 = FullyConnectedNet([n_mul + irrep_in[0][0], invariant_neurons, self.tp_node.weight_numel], ssp)
        self.linear_out = Linear(self.irrep_tp_out_node, irrep_out)
        lin_out, _ = feasible_irrep(irrep_in, Irreps("0e"), irrep_in)
        if use_norm_gate:
            self.norm_gate = NormGate(irrep_in)
            self.linear_node = Linear(irrep_in, lin_out)
            self.linear_node_pre = Linear(irrep_in, lin_out)

    def forward(self, x, edge_dst, edge_src, edge_attr, edge_sh):
        s0 = self.irrep_in_node.slices()[0]
        if self.use_norm_gate:
            pre = self.linear_node_pre(x)
            ip = inner_product(self.irrep_in_node, pre[edge_dst], pre[edge_src])[:, s0.stop:]
            inv = torch.cat([pre[edge_dst][:, s0], pre[edge_dst][:, s0], ip], dim=-1)  # dst scalars twice (layers.py:240-247)
            x = self.linear_node(self.norm_gate(x))
        else:
            ip = inner_product(self.irrep_in_node, x[edge_dst], x[edge_src])[:, s0.stop:]
            inv = torch.cat([x[edge_dst][:, s0], x[edge_dst][:, s0], ip], dim=-1)
        self_x = x
        msg = self.tp_node(x[edge_src], edge_sh, self.fc_node(edge_attr) * self.layer_l0(inv))
        out = torch.zeros(x.shape[0], msg.shape[1], dtype=x.dtype).index_add_(0, edge_dst, msg)
        if self.irrep_in_node == self.irrep_out:
            out = out + self_x
        return self.linear_out(out)


class ConvNetLayer(nn.Module):
    def __init__(self, irrep_in, irrep_hidden, irrep_out, sh_irrep, edge_attr_dim, use_norm_gate=True):
        super().__init__()
        self.resnet = irrep_in == irrep_out
        self.conv = ConvLayer(irrep_in, irrep_hidden, irrep_out, sh_irrep, edge_attr_dim, use_norm_gate)

    def forward(self, x, *graph):
        y = self.conv(x, *graph)
        return x + y if self.resnet else y


class PairNetLayer(nn.Module):
    def __init__(self, irrep_in, irrep_bottle, irrep_out, edge_attr_dim, invariant_neurons):
        super().__init__()
        self.irrep_in_node = irrep_in
        tp_in, _ = feasible_irrep(irrep_in, Irreps("0e"), irrep_bottle)
        self.irrep_tp_out_node_pair, ins = feasible_irrep(tp_in, tp_in, irrep_bottle, "uuu")
        self.linear_node_pair_n = Linear(irrep_in, irrep_in)
        self.linear_node_pair_inner = Linear(irrep_in, irrep_in)
        self.tp_node_pair = TensorProduct(tp_in, tp_in, self.irrep_tp_out_node_pair, ins, shared_weights=False, internal_weights=False)
        self.fc_node_pair = FullyConnectedNet([edge_attr_dim, invariant_neurons, self.tp_node_pair.weight_numel], ssp)
        self.resnet = irrep_in == irrep_out
        self.linear_node_pair = Linear(self.irrep_tp_out_node_pair, irrep_out)
        self.norm_gate = NormGate(self.irrep_tp_out_node_pair)
        self.norm_gate_pre = NormGate(self.irrep_tp_out_node_pair)
        n_mul = sum(m.mul for m in irrep_in)
        h = irrep_in[0][0]
        self.fc = nn.Sequential(nn.Linear(h + n_mul, h), nn.SiLU(), nn.Linear(h, self.tp_node_pair.weight_numel))

    def forward(self, node_attr, dst, src, full_edge_attr, pair_attr=None):
        s0 = self.irrep_in_node.slices()[0]
        a0 = self.linear_node_pair_inner(node_attr)
        ip = inner_product(self.irrep_in_node, a0[dst], a0[src])[:, s0.stop:]
        inv = torch.cat([a0[dst][:, s0], a0[src][:, s0], ip], dim=-1)
        x = self.linear_node_pair_n(self.norm_gate_pre(node_attr))
        pair = self.tp_node_pair(x[src], x[dst], self.fc_node_pair(full_edge_attr) * self.fc(inv))
        pair = self.linear_node_pair(self.norm_gate(pair))
        if self.resnet and pair_attr is not None:
            pair = pair + pair_attr
        return pair


class SelfNetLayer(nn.Module):
    def __init__(self, irrep_in, irrep_bottle, irrep_out):
        super().__init__()
        tp_in, _ = feasible_irrep(irrep_in, Irreps("0e"), irrep_bottle)
        tp_out, ins = feasible_irrep(tp_in, tp_in, irrep_bottle, "uuu")
        self.linear_node_1 = Linear(irrep_in, irrep_in)
        self.linear_node_2 = Linear(irrep_in, irrep_in)
        self.tp = TensorProduct(tp_in, tp_in, tp_out, ins, shared_weights=True, internal_weights=True)
        self.norm_gate = NormGate(irrep_out)
        self.norm_gate_1 = NormGate(irrep_in)
        self.norm_gate_2 = NormGate(irrep_in)
        self.linear_node_3 = Linear(tp_out, irrep_out)

    def forward(self, x, old_fii):
        xl = self.linear_node_1(self.norm_gate_1(x))
        xr = self.linear_node_2(self.norm_gate_2(x))
        y = self.tp(xl, xr) + x
        y = self.linear_node_3(self.norm_gate(y))
        return old_fii + y if old_fii is not None else y


class Expansion(nn.Module):
    """bottle irreps -> (n_s x 0e + n_p x 1e + n_d x 2e)^2 block; path weights / biases are inputs."""

    def __init__(self, irrep_in: Irreps, irrep_out_1: Irreps, irrep_out_2: Irreps):
        super().__init__()
        self.irrep_in, self.irrep_out_1, self.irrep_out_2 = irrep_in, irrep_out_1, irrep_out_2
        self.instructions = [(i, j, k, [mi.mul, mj.mul, mk.mul]) for i, mi in enumerate(irrep_in) for j, mj in enumerate(irrep_out_1)
                             for k, mk in enumerate(irrep_out_2) if mi.ir in mj.ir * mk.ir]
        self.num_path_weight = sum(int(np.prod(s)) for *_, s in self.instructions)
        self.num_bias = sum(int(np.prod(s[1:])) for i, _, _, s in self.instructions if i == 0)
        self.weights = nn.Parameter(torch.rand(self.num_path_weight + self.num_bias))  # unused by forward (layers.py:595)

    def forward(self, x_in, weights, bias_weights):
        B = x_in.shape[0]
        xs = [x_in[:, s].reshape(B, m.mul, m.ir.dim) for s, m in zip(self.irrep_in.slices(), self.irrep_in)]
        tiles, woff, boff = {}, 0, 0
        for i, j, k, shape in self.instructions:
            nw = int(np.prod(shape))
            Wt = weights[:, woff:woff + nw].reshape([B] + shape)
            woff += nw
            r = torch.einsum("bwuv,bwk->buvk", Wt, xs[i])
            if i == 0:
                nb = int(np.prod(shape[1:]))
                r = r + bias_weights[:, boff:boff + nb].reshape([B] + shape[1:]).unsqueeze(-1)
                boff += nb
            C = wigner_3j(j, k, i, dtype=x_in.dtype)  # irreps indices == l here (layers.py:617)
            r = torch.einsum("ijk,buvk->buivj", C, r) / self.irrep_in[i].mul
            r = r.reshape(B, self.irrep_out_1[j].dim, self.irrep_out_2[k].dim)
            tiles[(j, k)] = tiles[(j, k)] + r if (j, k) in tiles else r
        rows = []
        for j, mj in enumerate(self.irrep_out_1):
            rows.append(torch.cat([tiles.get((j, k), x_in.new_zeros(B, mj.dim, mk.dim)) for k, mk in enumerate(self.irrep_out_2)], dim=-1))
        return torch.cat(rows, dim=-2)


def orbital_masks(orbitals: Dict[int, List[int]]):
    max_z = max(orbitals.keys())
    _, counts = np.unique(orbitals[max_z], return_counts=True)
    s_max, p_max, d_max = (int(c) for c in counts)
    ranges = [list(range(s_max)), [s_max + i for i in range(3 * p_max)], [s_max + 3 * p_max + i for i in range(5 * d_max)]]
    masks = {}
    for z, ls in orbitals.items():
        _, cnt = np.unique(ls, return_counts=True)
        m = []
        for t, c in enumerate(cnt):
            m += ranges[t][: int(c) * (1, 3, 5)[t]]
        masks[z] = torch.tensor(m)
    return masks, s_max, p_max, d_max


class QHNetOracle(nn.Module):
    def __init__(self, sh_lmax=4, hidden_size=128, bottle_hidden_size=32, num_gnn_layers=5, max_radius=12, num_nodes=83,
                 radius_embed_dim=32, orbitals: Dict[int, List[int]] = None):
        super().__init__()
        hs, hbs = hidden_size, bottle_hidden_size
        self.sh_irrep = Irreps.spherical_harmonics(sh_lmax)
        self.hs, self.hbs, self.max_radius, self.num_gnn_layers = hs, hbs, max_radius, num_gnn_layers
        self.node_embedding = nn.Embedding(num_nodes, hs)
        self.hidden_irrep = Irreps(f"{hs}x0e+{hs}x1o+{hs}x2e+{hs}x3o+{hs}x4e")
        self.hidden_bottle_irrep = Irreps(f"{hbs}x0e+{hbs}x1o+{hbs}x2e+{hbs}x3o+{hbs}x4e")
        self.hidden_irrep_base = Irreps(f"{hs}x0e+{hs}x1e+{hs}x2e+{hs}x3e+{hs}x4e")
        self.distance_expansion = ExpBernstein(radius_embed_dim, max_radius)
        self.orbital_mask, max_s, max_p, max_d = orbital_masks(orbitals)
        self.start_layer = 2
        self.e3_gnn_layer, self.e3_gnn_node_layer, self.e3_gnn_node_pair_layer = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for i in range(num_gnn_layers):
            irr_in = Irreps(f"{hs}x0e") if i == 0 else self.hidden_irrep
            self.e3_gnn_layer.append(ConvNetLayer(irr_in, self.hidden_irrep, self.hidden_irrep, self.sh_irrep, radius_embed_dim, use_norm_gate=i != 0))
            if i > self.start_layer:
                b = self.hidden_irrep_base
                self.e3_gnn_node_layer.append(SelfNetLayer(b, b, b))
                self.e3_gnn_node_pair_layer.append(PairNetLayer(b, b, b, radius_embed_dim, invariant_neurons=hs))
        bottle_e = Irreps(f"{hbs}x0e+{hbs}x1e+{hbs}x2e+{hbs}x3e+{hbs}x4e")
        out_irr = Irreps(f"{max_s}x0e+{max_p}x1e+{max_d}x2e")
        mk = lambda n_in, n_out: nn.Sequential(nn.Linear(n_in, hs), nn.SiLU(), nn.Linear(hs, n_out))
        self.expand_ii = nn.ModuleDict({"hamiltonian": Expansion(bottle_e, out_irr, out_irr)})
        self.expand_ij = nn.ModuleDict({"hamiltonian": Expansion(bottle_e, out_irr, out_irr)})
        e = self.expand_ii["hamiltonian"]
        self.fc_ii = nn.ModuleDict({"hamiltonian": mk(hs, e.num_path_weight)})
        self.fc_ii_bias = nn.ModuleDict({"hamiltonian": mk(hs, e.num_bias)})
        self.fc_ij = nn.ModuleDict({"hamiltonian": mk(2 * hs, e.num_path_weight)})
        self.fc_ij_bias = nn.ModuleDict({"hamiltonian": mk(2 * hs, e.num_bias)})
        self.output_ii = Linear(self.hidden_irrep, self.hidden_bottle_irrep)
        self.output_ij = Linear(self.hidden_irrep, self.hidden_bottle_irrep)

    def build_graph(self, pos, batch, max_radius, max_num_neighbors):
        ei = radius_graph(pos, max_radius, batch, max_num_neighbors)
        dst, src = ei[0], ei[1]
        vec = pos[dst] - pos[src]
        rbf = self.distance_expansion(vec.norm(dim=-1, keepdim=True)).to(pos.dtype)
        sh = spherical_harmonics(self.sh_irrep, vec[:, [1, 2, 0]]).to(pos.dtype)
        return dst, src, rbf, sh

    def blocks(self, z, pos, batch):
        """-> diagonal blocks [N,32,32], off-diagonal blocks [P,32,32] and the full-graph (dst, src)."""
        n_total = z.shape[0]
        dst, src, rbf, sh = self.build_graph(pos, batch, self.max_radius, n_total)
        emb = self.node_embedding(z)
        fdst, fsrc, frbf, _ = self.build_graph(pos, batch, 10000.0, n_total)
        x, fii, fij = emb, None, None
        for li, layer in enumerate(self.e3_gnn_layer):
            x = layer(x, dst, src, rbf, sh)
            if li > self.start_layer:
                k = li - self.start_layer - 1
                fii = self.e3_gnn_node_layer[k](x, fii)
                fij = self.e3_gnn_node_pair_layer[k](x, fdst, fsrc, frbf, fij)
        fii, fij = self.output_ii(fii), self.output_ij(fij)
        diag = self.expand_ii["hamiltonian"](fii, self.fc_ii["hamiltonian"](emb), self.fc_ii_bias["hamiltonian"](emb))
        pe = torch.cat([emb[fdst], emb[fsrc]], dim=-1)
        offd = self.expand_ij["hamiltonian"](fij, self.fc_ij["hamiltonian"](pe), self.fc_ij_bias["hamiltonian"](pe))
        return diag, offd, fdst, fsrc

    def assemble(self, z, batch, diag, offd, fdst, fsrc):
        """build_final_matrix + H + H^T: block (row atom a, col atom b) = offd[edge with dst=a, src=b] masked."""
        masks = [self.orbital_mask[int(t)] for t in z.tolist()]
        norb = torch.tensor([len(m) for m in masks])
        off = torch.zeros(z.shape[0] + 1, dtype=torch.long)
        off[1:] = torch.cumsum(norb, 0)
        H = torch.zeros(int(off[-1]), int(off[-1]), dtype=diag.dtype)
        for a in range(z.shape[0]):
            H[off[a]:off[a + 1], off[a]:off[a + 1]] = diag[a][masks[a]][:, masks[a]]
        for e in range(fdst.shape[0]):
            a, b = int(fdst[e]), int(fsrc[e])
            H[off[a]:off[a + 1], off[b]:off[b + 1]] = offd[e][masks[a]][:, masks[b]]
        return H + H.T

    def forward(self, z, pos, batch):
        diag, offd, fdst, fsrc = self.blocks(z, pos, batch)
        return self.assemble(z, batch, diag, offd, fdst, fsrc)
