#!/usr/bin/env python
"""QHNet Hamiltonians for a nablaDFT Hamiltonian database + the reference's losses against the stored matrices (qhnet.py:375-395)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nabladft_b200.data import PackedHamiltonianDataset  # noqa: E402
from nabladft_b200.losses import HamiltonianLoss, masked_mae  # noqa: E402
from nabladft_b200.qhnet import QHNet  # noqa: E402

ORBITALS = {1: [0, 0, 1], 6: [0, 0, 0, 1, 1, 2], 7: [0, 0, 0, 1, 1, 2], 8: [0, 0, 0, 1, 1, 2], 9: [0, 0, 0, 1, 1, 2],
            16: [0, 0, 0, 0, 1, 1, 1, 2], 17: [0, 0, 0, 0, 1, 1, 1, 2], 35: [0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2]}  # config/model/qhnet.yaml:14-22


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("db")
    ap.add_argument("--weights", help="state dict with the reference QHNet parameter names")
    ap.add_argument("--batch", type=int, default=8)
    a = ap.parse_args()
    net = QHNet(sh_lmax=4, hidden_size=128, bottle_hidden_size=32, num_gnn_layers=5, max_radius=12, num_nodes=83, radius_embed_dim=32, orbitals=ORBITALS)
    if a.weights:
        net.load_state_dict(torch.load(a.weights, map_location="cpu"), strict=True)
    net = net.eval().to("cuda:0")
    ds = PackedHamiltonianDataset.from_db(a.db)
    loss, mae, n = 0.0, 0.0, 0
    for start in range(0, len(ds), a.batch):
        data, targets = ds.batch(range(start, min(len(ds), start + a.batch)), device="cuda:0")
        pred = net(data, packed=True)
        loss += float(HamiltonianLoss.packed(pred, targets)); mae += float(masked_mae(pred, targets)); n += 1
    print(f"{len(ds)} molecules: HamiltonianLoss {loss / n:.6f}  masked MAE {mae / n:.6f} Ha")


if __name__ == "__main__":
    main()
