"""CPU restatement of GemNet-OC's graph and index construction.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/nablaDFT/gemnet_oc/gemnet_oc.py and utils.py / interaction_indices.py for the non-periodic case
(`use_pbc: false`, config/model/gemnet-oc.yaml):
    generate_graph_dict        gemnet_oc.py:820-868   radius graph at cutoff_aint, edge vector pointing target -> source ... negated
    subselect_edges            gemnet_oc.py:777-818   stricter cutoff, then the `max_neighbors` NEAREST neighbours per target atom
                               (utils.get_max_neighbors_mask with enforce_max_strictly, utils.py:408-500)
    get_graphs_and_indices     gemnet_oc.py:897-1000  a2a (12 A, 1000) -> main (30), a2ee2a (20), qint (8)
    symmetrize_edges           gemnet_oc.py:694-775   keep source < target, append the flipped copies, per molecule: [directed | flipped]
    id_swap                                           index of the opposite edge
    get_triplets               interaction_indices.py:14-62   all (b->a, c->a) with distinct edges, grouped by the output edge c->a
Every index array is compared EXACTLY with what the reference's own classes produce (tests/golden/gemnet_oc_f32.npz written by
tests/golden/make_golden_gemnet_oc.py): PINNED.  The network itself (bases, interaction blocks) is restated in oracle/gemnet_oc.py.
"""
import numpy as np
import torch

from .graph import radius_graph


def _max_neighbors_mask(target: torch.Tensor, dist: torch.Tensor, n_atoms: int, max_nb: int) -> torch.Tensor:
    """Keep, for every target atom, its `max_nb` nearest incoming edges (strict).  `target` is sorted (radius_graph order).
    Ties: the reference sorts a padded [atoms, max_degree] distance matrix with torch.sort (utils.py:448-466); distances of distinct pairs
    of a real geometry do not tie, and the symmetric pair i<->j lives in different rows."""
    keep = torch.zeros_like(target, dtype=torch.bool)
    counts = torch.bincount(target, minlength=n_atoms)
    start = torch.cumsum(counts, 0) - counts
    for a in range(n_atoms):
        c = int(counts[a])
        if c == 0:
            continue
        seg = slice(int(start[a]), int(start[a]) + c)
        if c <= max_nb:
            keep[seg] = True
        else:
            order = torch.argsort(dist[seg], stable=True)[:max_nb]
            keep[int(start[a]) + order] = True
    return keep


def build_graphs(pos: torch.Tensor, batch: torch.Tensor, cutoff=12.0, max_neighbors=30, cutoff_aint=12.0, max_neighbors_aint=1000,
                 cutoff_aeaint=12.0, max_neighbors_aeaint=20, cutoff_qint=12.0, max_neighbors_qint=8):
    n = pos.shape[0]
    ei = radius_graph(pos, cutoff_aint, batch, max_neighbors_aint)  # row 0 = source, row 1 = target, sorted by target
    j, i = ei
    vec = pos[j] - pos[i]
    dist = vec.norm(dim=1)
    a2a = {"edge_index": ei, "distance": dist, "vector": -vec / dist[:, None]}

    def subselect(g, cut, max_nb, cut_orig, max_orig):
        m = torch.ones_like(g["distance"], dtype=torch.bool)
        if not np.isclose(cut, cut_orig):
            m &= g["distance"] <= cut
        sub = {k: (v[:, m] if k == "edge_index" else v[m]) for k, v in g.items()}
        if max_nb != max_orig:
            k2 = _max_neighbors_mask(sub["edge_index"][1], sub["distance"], n, max_nb)
            sub = {k: (v[:, k2] if k == "edge_index" else v[k2]) for k, v in sub.items()}
        return sub

    main = subselect(a2a, cutoff, max_neighbors, cutoff_aint, max_neighbors_aint)
    a2ee2a = subselect(a2a, cutoff_aeaint, max_neighbors_aeaint, cutoff_aint, max_neighbors_aint)
    qint = subselect(a2a, cutoff_qint, max_neighbors_qint, cutoff_aint, max_neighbors_aint)
    main, id_swap = symmetrize(main, batch)
    return {"main": main, "a2a": a2a, "a2ee2a": a2ee2a, "qint": qint, "id_swap": id_swap, "trip_e2e": triplets(main["edge_index"], n)}


def symmetrize(g, batch):
    """gemnet_oc.py:694-775 without periodic images: edges with source < target survive; per molecule the result is
    [its directed edges (in the original order) | the same edges flipped]."""
    s, t = g["edge_index"]
    mask = s < t
    sd, td = s[mask], t[mask]
    mol_of_edge = batch[td]                      # both ends are in the same molecule
    n_mol = int(batch.max()) + 1
    per_mol = torch.bincount(mol_of_edge, minlength=n_mol)
    n_dir = int(mask.sum())
    # edges are sorted by target, hence by molecule: molecule m owns a contiguous block of the directed list
    starts = torch.cumsum(per_mol, 0) - per_mol
    order = []
    for m in range(n_mol):
        blk = torch.arange(int(starts[m]), int(starts[m]) + int(per_mol[m]))
        order.append(blk); order.append(blk + n_dir)
    order = torch.cat(order) if order else torch.zeros(0, dtype=torch.long)
    cat_idx = torch.cat([torch.stack([sd, td]), torch.stack([td, sd])], dim=1)
    new = {"edge_index": cat_idx[:, order], "distance": torch.cat([g["distance"][mask]] * 2)[order],
           "vector": torch.cat([g["vector"][mask], -g["vector"][mask]])[order]}
    # id_swap: position of the edge (t, s) for every edge (s, t)
    n = int(batch.shape[0])
    s2, t2 = new["edge_index"]
    eid = s2 + t2 * n
    eid_rev = t2 + s2 * n
    pos_of = torch.empty(n * n, dtype=torch.long).fill_(-1)
    pos_of[eid] = torch.arange(eid.numel())
    return new, pos_of[eid_rev]


def triplets(edge_index, n_atoms):
    """interaction_indices.py:14-62: for every output edge c->a (in edge order) all input edges b->a with b sorted ascending (the
    SparseTensor row of a is sorted by column = source), the edge itself excluded."""
    s, t = edge_index
    e = s.numel()
    order = torch.argsort(t * n_atoms + s, stable=True)        # adjacency rows (target) with columns (source) ascending
    t_sorted = t[order]
    counts = torch.bincount(t_sorted, minlength=n_atoms)
    ptr = torch.cumsum(counts, 0) - counts
    n_per = counts[t]                                           # inputs of each output edge = in-degree of its target
    out = torch.repeat_interleave(torch.arange(e), n_per)
    inner = torch.arange(int(n_per.sum())) - torch.repeat_interleave(torch.cumsum(n_per, 0) - n_per, n_per)
    inp = order[torch.repeat_interleave(ptr[t], n_per) + inner]
    keep = inp != out
    inp, out = inp[keep], out[keep]
    agg = torch.arange(out.numel()) - torch.repeat_interleave(torch.cumsum(torch.bincount(out, minlength=e), 0) - torch.bincount(out, minlength=e),
                                                               torch.bincount(out, minlength=e))
    return {"in": inp, "out": out, "out_agg": agg}


def _inner_idx(sorted_idx, dim_size):
    """utils.get_inner_idx: 0,1,2,... inside every run of equal (sorted) indices."""
    counts = torch.bincount(sorted_idx, minlength=dim_size)
    return torch.arange(sorted_idx.numel()) - torch.repeat_interleave(torch.cumsum(counts, 0) - counts, counts)


def mixed_triplets(edge_index_in, edge_index_out, n_atoms, to_outedge=False, with_agg=False):
    """interaction_indices.get_mixed_triplets (non-periodic): for every OUTPUT edge k, in order, the INPUT edges whose target is
    the output edge's target (source if to_outedge), sorted by their source; pairs with identical end atoms removed."""
    in_s, in_t = edge_index_in
    out_s, out_t = edge_index_out
    order = torch.argsort(in_t * n_atoms + in_s, stable=True)
    counts = torch.bincount(in_t, minlength=n_atoms)
    ptr = torch.cumsum(counts, 0) - counts
    pivot = out_s if to_outedge else out_t
    n_per = counts[pivot]
    idx_out = torch.repeat_interleave(torch.arange(out_s.numel()), n_per)
    inner = torch.arange(int(n_per.sum())) - torch.repeat_interleave(torch.cumsum(n_per, 0) - n_per, n_per)
    idx_in = order[torch.repeat_interleave(ptr[pivot], n_per) + inner]
    atom_in = in_s[idx_in]
    atom_out = out_t[idx_out] if to_outedge else out_s[idx_out]
    keep = atom_in != atom_out
    res = {"in": idx_in[keep], "out": idx_out[keep]}
    if with_agg:
        res["out_agg"] = _inner_idx(res["out"], out_s.numel())
    return res


def quadruplets(main_edge_index, qint_edge_index, n_atoms):
    """interaction_indices.get_quadruplets (non-periodic): d->b->a<-c with b->a from the quadruplet-interaction graph."""
    idx_s = main_edge_index[0]
    n_qint = qint_edge_index.shape[1]
    trip_in = mixed_triplets(main_edge_index, qint_edge_index, n_atoms, to_outedge=True)     # d->b for every b->a
    trip_out = mixed_triplets(qint_edge_index, main_edge_index, n_atoms, to_outedge=False)   # b->a for every c->a
    n_in_per_inter = torch.bincount(trip_in["out"], minlength=n_qint)
    n_out = n_in_per_inter[trip_out["in"]]
    out = torch.repeat_interleave(trip_out["out"], n_out)
    trip_out_to_quad = torch.repeat_interleave(torch.arange(trip_out["out"].numel()), n_out)
    # input triplets of every intermediate edge, in stored order (trip_in["out"] is sorted)
    start = torch.cumsum(n_in_per_inter, 0) - n_in_per_inter
    inner = torch.arange(int(n_out.sum())) - torch.repeat_interleave(torch.cumsum(n_out, 0) - n_out, n_out)
    trip_in_to_quad = torch.repeat_interleave(start[trip_out["in"]], n_out) + inner
    idx_in = trip_in["in"][trip_in_to_quad]
    keep = idx_s[out] != idx_s[idx_in]   # c != d
    res = {"triplet_in": trip_in, "triplet_out": trip_out, "out": out[keep], "trip_out_to_quad": trip_out_to_quad[keep],
           "trip_in_to_quad": trip_in_to_quad[keep]}
    res["out_agg"] = _inner_idx(res["out"], main_edge_index.shape[1])
    return res


def build_all_indices(pos, batch, **kw):
    """Everything GemNetOC.get_graphs_and_indices returns (gemnet_oc.py:897-1000), as plain index tensors."""
    g = build_graphs(pos, batch, **kw)
    n = int(pos.shape[0])
    g["trip_a2e"] = mixed_triplets(g["a2ee2a"]["edge_index"], g["main"]["edge_index"], n, with_agg=True)
    g["trip_e2a"] = mixed_triplets(g["main"]["edge_index"], g["a2ee2a"]["edge_index"], n, with_agg=True)
    g["quad"] = quadruplets(g["main"]["edge_index"], g["qint"]["edge_index"], n)
    g["a2a"]["target_neighbor_idx"] = _inner_idx(g["a2a"]["edge_index"][1], n)
    g["a2ee2a"]["target_neighbor_idx"] = _inner_idx(g["a2ee2a"]["edge_index"][1], n)
    return g
