"""The reference's own model / dataset tests, restated for the mirror classes (tests/model/test_torch_models.py:31-40,56-62 and
tests/dataset/assertions.py:15-17 of the reference): same assertions, our classes, batches produced by nabladft_b200.data."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN


def _packed():
    from nabladft_b200.data import PackedEnergyDataset

    fx = np.load(os.path.join(GOLDEN, "fixture_molecules.npz"))
    return PackedEnergyDataset(fx["z"].astype(np.int32), fx["pos"].astype(np.float32), fx["forces"].astype(np.float32),
                               fx["energy"].astype(np.float32), fx["ptr"].astype(np.int64))


def test_assert_shapes_spk_on_our_batches():
    """tests/dataset/assertions.py:15-17 `assert_shapes_spk`."""
    from nabladft_b200.data import DeviceBatcher

    batch = next(iter(DeviceBatcher(_packed(), batch_size=4, device="cpu"))).as_spk()
    assert batch["energy"].shape == batch["_idx"].shape
    assert batch["forces"].shape == batch["_positions"].shape == torch.Size([batch["_atomic_numbers"].shape[0], 3])


@pytest.mark.gpu
@pytest.mark.parametrize("model_name", ["SchNet", "PaiNN"])
def test_spk_models(model_name):
    """tests/model/test_torch_models.py:31-40 (`AtomsLoader(dataset_spk, batch_size=4)` -> our DeviceBatcher)."""
    from nabladft_b200 import spk
    from nabladft_b200.data import DeviceBatcher
    from helpers import load_golden_weights

    rep = spk.SchNet if model_name == "SchNet" else spk.PaiNN
    model = spk.NeuralNetworkPotential(
        representation=rep(n_atom_basis=128, n_interactions=6, radial_basis=spk.GaussianRBF(n_rbf=100, cutoff=5.0), cutoff_fn=spk.CosineCutoff(cutoff=5.0)),
        input_modules=[spk.PairwiseDistances()], output_modules=[spk.Atomwise(n_in=128, output_key="energy"), spk.Forces()],
        postprocessors=[spk.AddOffsets(property="energy", add_mean=True)])
    model = load_golden_weights(model, torch.float32).eval().to("cuda:0")
    batch = next(iter(DeviceBatcher(_packed(), batch_size=4, device="cuda:0"))).as_spk()
    output = model(batch)
    energy, forces = output["energy"], output["forces"]
    assert energy.shape == batch["energy"].shape
    assert forces.shape == batch["forces"].shape


@pytest.mark.gpu
def test_pyg_model():
    """tests/model/test_torch_models.py:21-27 for the in-repo PaiNN (`Batch.from_data_list([dataset_pyg[0]])` -> one-molecule batch)."""
    from nabladft_b200.data import DeviceBatcher
    from test_gpu_painn import _oc_model

    model = _oc_model(6).to("cuda:0")
    batch = next(iter(DeviceBatcher(_packed(), batch_size=1, device="cuda:0"))).as_pyg()
    energy, forces = model(batch)
    assert energy.shape == batch.y.shape
    assert forces.shape == batch.forces.shape


@pytest.mark.gpu
def test_hamiltonian_model():
    """tests/model/test_torch_models.py:56-62: QHNet returns the block-diagonal matrix of the batch; for one molecule its shape is the
    Hamiltonian's (Norb x Norb, Norb from the def2-SVP table of config/model/qhnet.yaml:14-22)."""
    from nabladft_b200.qhnet import QHNet
    from helpers import load_golden_weights
    from test_gpu_qhnet import ORBITALS, _Data

    net = QHNet(sh_lmax=4, hidden_size=128, bottle_hidden_size=32, num_gnn_layers=5, max_radius=12, num_nodes=83, radius_embed_dim=32, orbitals=ORBITALS)
    load_golden_weights(net, torch.float32, style="e3")
    net = net.eval().to("cuda:0")
    m = _packed().molecule(2)
    z = torch.from_numpy(m["z"].astype(np.int64)).to("cuda:0")
    pos = torch.from_numpy(m["pos"] * 1.8897261).to("cuda:0")  # Hamiltonian DBs are in bohr
    output = net(_Data(z, pos, torch.zeros(len(m["z"]), dtype=torch.long, device="cuda:0")))
    norb = sum(sum(2 * l + 1 for l in ORBITALS[int(a)]) for a in m["z"])
    assert output.shape == (norb, norb)


@pytest.mark.gpu
def test_spk_optimization():
    """tests/optimization/test_optim_pipelines.py:21-30 without the ASE database: relax the first 8 fixture molecules in two batches with
    the PaiNN mirror; the reference asserts on shapes and that the model energy at the relaxed geometry is below the starting one."""
    from nabladft_b200.optimization import ASEBatchwiseLBFGS, PackedOptimizeTask, SimpleAtoms, SpkBatchwiseCalculator
    from nabladft_b200.data import PackedEnergyDataset
    from test_gpu_painn import _spk_model

    full = _packed()
    n = int(full.ptr[8])
    ds = PackedEnergyDataset(full.z[:n], full.pos[:n], full.forces[:n], full.energy[:8], full.ptr[:9])
    model = _spk_model(3).to("cuda:0")
    calc = SpkBatchwiseCalculator(model, device="cuda:0", energy_unit="Hartree", position_unit="Ang")
    e_start = calc.get_potential_energy([SimpleAtoms(ds.molecule(i)["pos"], ds.molecule(i)["z"]) for i in range(8)]).copy()
    out = PackedOptimizeTask(ds, ASEBatchwiseLBFGS(calc, logfile=None, maxstep=0.05, check_every=5), batch_size=5, fmax=1e-3, steps=15).run()
    assert out["model_forces"].shape == ds.forces.shape and out["positions"].shape == ds.pos.shape and out["model_energy"].shape == (8,)
    assert len(out["nsteps"]) == 2 and np.isfinite(out["model_energy"]).all()
    assert (out["model_energy"] < e_start).all()          # 15 small quasi-Newton steps lower every molecule's model energy
    assert np.abs(out["positions"] - ds.pos).max() > 1e-3  # and the geometry did move
