"""Loss / metric steps of the hot path restated on the engine's
outputs.  These are scalar reductions over tensors the kernels already produced; they stay in
PyTorch (device-side reductions, no host round trip).

* `L2Loss`                -- nablaDFT/gemnet_oc/loss.py:5-22 (mean per-atom |dF|_2), PaiNN-OC forces loss.
* `HamiltonianLoss`       -- nablaDFT/qhnet/loss.py:5-16: RMSE + MAE over the batch block-diagonal,
                             rescaled by numel/mask.sum.  Off-block entries of prediction and target are
                             structurally zero, so the same number is obtained from the PACKED
                             per-molecule matrices (`QHNet.last_blocks`) without materialising the
                             [sum Norb]^2 dense matrix nor the CPU block_diag of the targets
                             (qhnet.py:368-373): sqrt(sum |dH_m|^2 / sum Norb_m^2) + sum |dH_m| / sum Norb_m^2.
* `masked_mae`            -- nablaDFT/qhnet/masked_mae.py:12-20 times norm_coef = numel/mask.sum of qhnet.py:490-495.
"""
from typing import List, Sequence

import torch
from torch import nn


class L2Loss(nn.Module):
    def __init__(self, reduction: str = "mean"):
        super().__init__()
        self.reduction = reduction

    def forward(self, pred, target):
        dist = torch.linalg.vector_norm(pred - target, dim=-1)
        return dist.mean() if self.reduction == "mean" else dist.sum() if self.reduction == "sum" else dist


class HamiltonianLoss(nn.Module):
    def forward(self, pred, target, mask=None):
        """Reference signature (dense block-diagonal pred/target + mask) or packed lists of per-molecule matrices."""
        if isinstance(pred, (list, tuple)):
            return self.packed(pred, target)
        diff = pred - target
        scale = pred.numel() / mask.sum()
        return torch.sqrt(torch.mean(diff**2) * scale) + torch.mean(torch.abs(diff)) * scale

    @staticmethod
    def packed(pred: Sequence[torch.Tensor], target: Sequence[torch.Tensor]):
        n = sum(p.numel() for p in pred)
        sq = sum(((p - t.to(p)) ** 2).sum() for p, t in zip(pred, target))
        ab = sum((p - t.to(p)).abs().sum() for p, t in zip(pred, target))
        return torch.sqrt(sq / n) + ab / n


def masked_mae(pred: List[torch.Tensor], target: List[torch.Tensor]) -> torch.Tensor:
    """The reference's per-step Hamiltonian metric, from packed per-molecule matrices.

    nablaDFT/qhnet/qhnet.py:490-495 returns `MaskedMeanAbsoluteError(pred, target) * norm_coef` on the dense batch
    block-diagonal: the metric (masked_mae.py:12-20) is sum|dH| / count_nonzero(target) and
    norm_coef = numel(block_diag) / mask.sum() = (sum_m Norb_m)^2 / sum_m Norb_m^2 -- 1 for a single molecule, ~B for a
    batch of B similar molecules.  Off-block entries are structurally zero in prediction and target, so both factors
    follow from the packed matrices without materialising the [sum Norb]^2 dense matrix.
    """
    ab = sum((p - t.to(p)).abs().sum() for p, t in zip(pred, target))
    nnz = sum(torch.count_nonzero(t) for t in target)
    n_orb = [int(p.shape[-1]) for p in pred]
    norm_coef = float(sum(n_orb)) ** 2 / float(sum(n * n for n in n_orb))
    return ab / nnz * norm_coef
