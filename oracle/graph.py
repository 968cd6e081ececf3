"""Oracle graph construction (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Restates the third-party neighbour searches the reference calls:

* `radius_graph`  -- `torch_cluster.radius_graph(x, r, batch, loop=False,
  max_num_neighbors=K, flow='source_to_target')` as called at
  `nablaDFT/painn_pyg/painn.py:411-416` and `nablaDFT/qhnet/qhnet.py:258`:
  `edge_index[0]` = source j, `edge_index[1]` = target i, same-molecule pairs with squared
  distance strictly below r^2, no self loops, at most K sources per target (the first K in
  ascending source index -- GPU semantics), grouped by ascending target.
* `ase_neighbor_list` -- `ase.neighborlist.neighbor_list('ijS', atoms, cutoff)` as wrapped by
  `schnetpack.transform.ASENeighborList` (`config/datamodule/nablaDFT_ase.yaml:13-14`) for
  non-periodic molecules: both directions, d < cutoff, sorted by centre atom i.
"""
from typing import Tuple

import torch


def batch_to_ptr(batch: torch.Tensor) -> torch.Tensor:
    n_mol = int(batch.max().item()) + 1 if batch.numel() else 0
    counts = torch.bincount(batch, minlength=n_mol)
    ptr = torch.zeros(n_mol + 1, dtype=torch.long)
    ptr[1:] = torch.cumsum(counts, 0)
    return ptr


def radius_graph(pos: torch.Tensor, r: float, batch: torch.Tensor, max_num_neighbors: int = 32) -> torch.Tensor:
    ptr = batch_to_ptr(batch)
    src_all, tgt_all = [], []
    for m in range(ptr.numel() - 1):
        a, b = int(ptr[m]), int(ptr[m + 1])
        p = pos[a:b].detach()
        diff = p[:, None, :] - p[None, :, :]
        d2 = (diff * diff).sum(-1)
        mask = d2 < (r * r)
        mask.fill_diagonal_(False)
        # rows = target i, cols = source j (ascending); keep the first K per target
        rank = torch.cumsum(mask.to(torch.long), dim=1)
        mask &= rank <= max_num_neighbors
        tgt, src = mask.nonzero(as_tuple=True)
        src_all.append(src + a)
        tgt_all.append(tgt + a)
    if not src_all:
        return torch.zeros(2, 0, dtype=torch.long)
    return torch.stack([torch.cat(src_all), torch.cat(tgt_all)])


def ase_neighbor_list(pos: torch.Tensor, ptr: torch.Tensor, cutoff: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (idx_i, idx_j): centre i, neighbour j, sorted by i then j."""
    batch = torch.repeat_interleave(torch.arange(ptr.numel() - 1), ptr[1:] - ptr[:-1])
    ei = radius_graph(pos, cutoff, batch, max_num_neighbors=10**9)
    return ei[1], ei[0]
