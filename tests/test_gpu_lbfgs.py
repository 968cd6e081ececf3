"""Device L-BFGS loop (csrc/lbfgs.cu through nabladft_b200.optimization) against the reference-pinned oracle."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_lbfgs import scenarios, start_geometry  # noqa: E402
from toy_potential import ToyPotential  # noqa: E402

pytestmark = pytest.mark.gpu


class _ToyEngine:
    """Stands in for PainnEngine: same run / launch / e_cap / raise_on_status surface, forces from the analytic potential."""

    def __init__(self, pot):
        self.pot, self.e_cap, self.n_calls = pot, 0, 0

    def run(self, z, pos32, mol_ptr, n_mol, with_forces=True):
        e, f, st = self.launch(z, pos32, mol_ptr, n_mol)
        return e, f, st.cpu()

    def launch(self, z, pos32, mol_ptr, n_mol, with_forces=True, e_cap=None):
        e, f = self.pot.torch(pos32)
        self.n_calls += 1
        return e.float(), f.contiguous(), torch.zeros(4, dtype=torch.int32, device=pos32.device)

    @staticmethod
    def raise_on_status(st):
        assert int(st[1]) == 0


def _toy_run(name, check_every, record=False):
    from nabladft_b200.optimization import ASEBatchwiseLBFGS, BatchwiseCalculator, SimpleAtoms

    sc = scenarios()[name]
    fix = np.load(os.path.join(HERE, "golden", "fixture_molecules.npz"))
    si = list(scenarios()).index(name)
    zs, ps = start_geometry(fix, sc["mols"], sc["jitter"], seed=100 + si)
    pot = ToyPotential(zs, [fix["pos"][int(fix["ptr"][m]):int(fix["ptr"][m + 1])] for m in sc["mols"]])

    class ToyCalc(BatchwiseCalculator):
        def engine(self_inner):
            return eng

    eng = _ToyEngine(pot)
    calc = ToyCalc(torch.nn.Identity(), device="cuda:0", energy_unit="Hartree", position_unit="Ang")
    opt = ASEBatchwiseLBFGS(calc, logfile=None, maxstep=sc["maxstep"], memory=sc["memory"], damping=sc["damping"], alpha=sc["alpha"],
                            fixed_atoms_mask=sc["fixed"], check_every=check_every)
    opt.record_positions = record
    conv = opt.run([SimpleAtoms(p, z) for p, z in zip(ps, zs)], fmax=sc["fmax"], steps=sc["steps"])
    return opt, conv, calc


def _setup(name):
    sc = scenarios()[name]
    fix = np.load(os.path.join(HERE, "golden", "fixture_molecules.npz"))
    si = list(scenarios()).index(name)
    zs, ps = start_geometry(fix, sc["mols"], sc["jitter"], seed=100 + si)
    pot = ToyPotential(zs, [fix["pos"][int(fix["ptr"][m]):int(fix["ptr"][m + 1])] for m in sc["mols"]])
    return sc, zs, ps, pot


@pytest.mark.parametrize("name", list(scenarios()))
def test_lbfgs_step_kernel_teacher_forced(name):
    """Every step of the reference trajectory, one kernel call each, through the C ABI: positions and float32 forces of step k
    are uploaded exactly as the reference saw them (the history the kernel builds from them is then bit-identical to the
    reference's), so the comparison with step k+1 isolates the arithmetic of ONE step -- no trajectory amplification."""
    from nabladft_b200 import _lib

    lib = _lib.load()
    gold = np.load(os.path.join(HERE, "golden", "lbfgs_ref.npz"))
    ref = gold[f"{name}/traj"]
    sc, zs, ps, pot = _setup(name)
    sizes = np.array([len(z) for z in zs])
    n_mol, n_atoms = len(sizes), int(sizes.sum())
    dev = "cuda:0"
    mol_ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)).to(dev)
    state = torch.empty(int(lib.nb200_lbfgs_state_bytes(n_mol, n_atoms, sc["memory"])), dtype=torch.uint8, device=dev)
    fixed = None
    if sc["fixed"] is not None:
        fixed = torch.zeros(n_atoms, dtype=torch.uint8, device=dev)
        fixed[torch.tensor(sc["fixed"], device=dev)] = 1
    unconv = torch.zeros(1, dtype=torch.int32, device=dev)
    n_norm = torch.zeros(1, dtype=torch.int32, device=dev)
    pos32 = torch.empty(n_atoms, 3, dtype=torch.float32, device=dev)
    worst = 0.0
    for k in range(len(ref) - 1):
        pos = torch.from_numpy(ref[k].copy()).to(dev)
        f = torch.from_numpy(pot.numpy(ref[k])[1].copy()).to(dev)
        rc = lib.nb200_lbfgs_step(_lib.ptr(state), state.numel(), _lib.ptr(mol_ptr), n_mol, n_atoms, int(sizes.max()), sc["memory"], k,
                                  float(sc["fmax"]), float(sc["maxstep"]), float(sc["damping"]), 1.0 / sc["alpha"], _lib.ptr(fixed),
                                  _lib.ptr(pos), _lib.ptr(f), _lib.ptr(pos32), _lib.ptr(unconv), _lib.ptr(n_norm), _lib.current_stream())
        _lib.check(rc, "nb200_lbfgs_step")
        got = pos.cpu().numpy()
        worst = max(worst, np.abs(got - ref[k + 1]).max())
        assert np.array_equal(pos32.cpu().numpy(), got.astype(np.float32))
        assert int(unconv.item()) > 0  # the reference took this step, so it was not converged
    # one step: float32 direction (|dr| <= 0.2 A, 24-bit) added to float64 positions; tree- vs sequential float64 sums
    assert worst < 2e-8, worst
    assert int(n_norm.item()) == int(gold[f"{name}/n_normalizations"])


# Free-running loop.  The reference algorithm itself amplifies a 1-ulp perturbation of the float32 forces to the tolerances
# below (measured with oracle/lbfgs.py: forces * (1 + 6e-8 * randn) -> 2.0e-3 A on "basic" after 40 steps, <= 1.5e-6 A on the
# other three), so these are the meaningful bounds for an implementation that reduces in a different order.
_FREE_TOL = {"basic": 2e-2, "short_memory": 2e-5, "converging": 2e-5, "fixed_atoms": 2e-5}


@pytest.mark.parametrize("name", list(scenarios()))
def test_device_lbfgs_loop_follows_reference_trajectory(name):
    gold = np.load(os.path.join(HERE, "golden", "lbfgs_ref.npz"))
    ref = gold[f"{name}/traj"]
    opt, conv, calc = _toy_run(name, check_every=1, record=True)
    traj = np.stack(opt.positions_history)
    assert opt.nsteps == int(gold[f"{name}/nsteps"]) and conv == bool(gold[f"{name}/converged"])
    n = min(len(traj), len(ref))
    assert n >= opt.nsteps
    err = np.abs(traj[:n] - ref[:n]).reshape(n, -1).max(axis=1)
    assert err[:4].max() < 1e-7 and err.max() < _FREE_TOL[name], err
    final = np.concatenate([a.get_positions() for a in opt.atoms])
    assert np.abs(final - ref[-1]).max() < _FREE_TOL[name]
    assert abs(opt.n_normalizations - int(gold[f"{name}/n_normalizations"])) <= (1 if name == "basic" else 0)
    assert np.abs(calc.results["energy"] - gold[f"{name}/final_energy"]).max() < 50 * _FREE_TOL[name] ** 2 + 1e-6


@pytest.mark.parametrize("name", ["converging", "fixed_atoms"])
def test_check_every_does_not_change_the_result(name):
    a, conv_a, _ = _toy_run(name, check_every=1)
    b, conv_b, _ = _toy_run(name, check_every=7)
    assert conv_a == conv_b and a.nsteps == b.nsteps and a.n_normalizations == b.n_normalizations
    pa = np.concatenate([x.get_positions() for x in a.atoms]); pb = np.concatenate([x.get_positions() for x in b.atoms])
    assert np.array_equal(pa, pb)  # steps taken after global convergence move nothing, bit for bit


def test_painn_relaxation_matches_oracle_loop():
    """The whole caller: nabladft_b200 PaiNN-OC forces driving the device loop vs the oracle loop driven by the oracle model."""
    from helpers import load_fixture, load_golden_weights
    from nabladft_b200.optimization import ASEBatchwiseLBFGS, PyGBatchwiseCalculator, SimpleAtoms
    from nabladft_b200.painn_oc import PaiNN
    from oracle.lbfgs import BatchLBFGS
    from oracle.painn_oc import PaiNNOC

    mols = [0, 5]
    zcat, pcat, batch = load_fixture(mols, dtype=torch.float64)
    sizes = torch.bincount(batch).tolist()
    off = np.concatenate([[0], np.cumsum(sizes)])
    zs = [zcat[off[i]:off[i + 1]].numpy() for i in range(len(mols))]
    ps = [pcat[off[i]:off[i + 1]].numpy() for i in range(len(mols))]
    kw = dict(hidden_channels=128, num_layers=3, num_rbf=100, cutoff=5.0, max_neighbors=100, num_elements=100)
    net = load_golden_weights(PaiNN(direct_forces=False, use_pbc=False, **kw), torch.float32)
    ref = PaiNNOC(**kw).float()
    ref.load_state_dict(net.state_dict(), strict=True)
    net = net.cuda().eval()

    def oracle_forces(pos):
        e, f = ref(zcat, torch.from_numpy(np.asarray(pos, dtype=np.float32)), batch)
        return e.detach().numpy(), f.detach().numpy()

    steps = 6
    orc = BatchLBFGS(oracle_forces, sizes, memory=100, maxstep=0.2)
    pos_o, conv_o, traj_o = orc.run(np.concatenate(ps), fmax=1e-4, steps=steps)
    calc = PyGBatchwiseCalculator(net, device="cuda:0", energy_unit="Hartree", position_unit="Ang")
    opt = ASEBatchwiseLBFGS(calc, logfile=None, check_every=3)
    conv = opt.run([SimpleAtoms(p, z) for p, z in zip(ps, zs)], fmax=1e-4, steps=steps)
    pos_d = np.concatenate([a.get_positions() for a in opt.atoms])
    assert opt.nsteps == orc.nsteps and conv == conv_o
    # forces agree to ~1e-6 Ha/A per call (test_gpu_painn); six quasi-Newton steps of <= 0.2 A amplify that mildly
    assert np.abs(pos_d - pos_o).max() < 2e-4, np.abs(pos_d - pos_o).max()
    assert np.abs(calc.results["forces"] - orc.final_forces).max() < 2e-3
