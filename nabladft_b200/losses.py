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
* `masked_mae`            -- nablaDFT/qhnet/masked_mae.py:12-20 with the numel/mask rescale of qhnet.py:490-495.
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
    """MaskedMeanAbsoluteError over packed matrices: sum |dH| / count_nonzero(target), times numel/mask.sum
    of the block diagonal (qhnet.py:490-495) -- the block-diagonal numel cancels against the dense mean only
    when taken over the packed entries, which is what is returned here."""
    ab = sum((p - t.to(p)).abs().sum() for p, t in zip(pred, target))
    nnz = sum(torch.count_nonzero(t) for t in target)
    return ab / nnz
