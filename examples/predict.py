#!/usr/bin/env python
"""Energies and forces for a packed dataset + MAE against the stored DFT values (reference: job_type test / predict)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nabladft_b200.data import DeviceBatcher, PackedEnergyDataset  # noqa: E402
from train_painn import build_model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cache")
    ap.add_argument("--weights", help="state dict with schnetpack parameter names (e.g. written by train_painn.py or a reference checkpoint)")
    ap.add_argument("--batch", type=int, default=256)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    model = build_model()
    if a.weights:
        model.load_state_dict(torch.load(a.weights, map_location="cpu"), strict=True)
    model = model.to(dev).eval()
    ds = PackedEnergyDataset.load(a.cache)
    e_pred = np.zeros(len(ds), dtype=np.float32)
    f_err, n_at = 0.0, 0
    for b in DeviceBatcher(ds, a.batch, device=dev):
        out = model(b.as_spk())
        e_pred[b.index.cpu().numpy()] = out["energy"].cpu().numpy()
        f_err += float((out["forces"] - b.forces).abs().sum()); n_at += b.z.shape[0]
    print(f"{len(ds)} molecules: MAE(E) {np.abs(e_pred - np.asarray(ds.energy)).mean():.6f} Ha, MAE(F) {f_err / (3 * n_at):.6f} Ha/A")
    np.save("energy_pred.npy", e_pred)


if __name__ == "__main__":
    main()
