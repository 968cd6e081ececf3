"""CUDA side of the data path: pinned staging + side-stream prefetch deliver exactly what a direct upload delivers, and the batches
drive both model mirrors."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN
from test_gpu_painn import _oc_model, _spk_model, dev

pytestmark = pytest.mark.gpu


def _dataset():
    from nabladft_b200.data import PackedEnergyDataset

    fx = np.load(os.path.join(GOLDEN, "fixture_molecules.npz"))
    return PackedEnergyDataset(fx["z"].astype(np.int32), fx["pos"].astype(np.float32), fx["forces"].astype(np.float32),
                               fx["energy"].astype(np.float32), fx["ptr"].astype(np.int64))


def test_cuda_batcher_prefetch_matches_direct_upload_and_feeds_the_models():
    from nabladft_b200.data import DeviceBatcher

    ds = _dataset()
    oc, spk = _oc_model(2).to(dev()), _spk_model(2).to(dev())
    it = DeviceBatcher(ds, batch_size=16, device=dev(), shuffle=True, seed=3)
    seen = []
    for b in it:
        assert b.z.is_cuda and b.pos.is_cuda
        idx = b.index.cpu().numpy()
        ref_z = np.concatenate([ds.molecule(int(m))["z"] for m in idx])
        ref_p = np.concatenate([ds.molecule(int(m))["pos"] for m in idx])
        assert np.array_equal(b.z.cpu().numpy(), ref_z) and np.array_equal(b.pos.cpu().numpy(), ref_p)
        e1, f1 = oc(b.as_pyg())
        out = spk(b.as_spk())
        assert e1.shape[0] == b.n_mol == out["energy"].shape[0] and f1.shape == b.forces.shape == out["forces"].shape
        assert torch.isfinite(e1).all() and torch.isfinite(out["forces"]).all()
        seen += idx.tolist()
    assert sorted(seen) == list(range(len(ds)))
