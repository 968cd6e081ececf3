"""Golden trajectories for the batched L-BFGS geometry optimiser, produced by the REFERENCE'S OWN CLASS
(`/root/reference/nablaDFT/optimization/optimizers.py`, `calculator.py`, unmodified, imported where they lie).

ASE / schnetpack / torch_geometric are not installable here, so the few names the two files import are provided as shims:
    ase.Atoms                       positions/numbers/pbc/cell container with ==, copy (only what optimizers.py touches)
    ase.optimize.optimize.Dynamics  __init__ stores logfile/trajectory, nsteps = 0, max_steps (ASE 3.22 semantics)
    ase.io.write, ase.parallel.{barrier, world}   unused on this path (trajectory=None, restart=None)
    schnetpack.units.convert_units  identity for equal units (config/calculator/*.yaml: Hartree, Ang)
    schnetpack.interfaces.ase_interface.{AtomsConverter, AtomsConverterError}, torch_geometric.data.{Batch, Data}  names only
The "model" is tests/golden/toy_potential.py (float64 energy, float32 forces), wrapped in a subclass of the reference's
BatchwiseCalculator, so every line of ASEBatchwiseLBFGS.{run, step, update, determine_step, converged} that executes is the reference's.

    python tests/golden/make_golden_lbfgs.py      # writes tests/golden/lbfgs_ref.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
REF = "/root/reference/nablaDFT/optimization"
from toy_potential import ToyPotential  # noqa: E402


class Atoms:
    def __init__(self, positions=None, numbers=None):
        self.positions = np.array(positions, dtype=np.float64)
        self.numbers = np.array(numbers, dtype=np.int64)
        self.pbc = np.zeros(3, dtype=bool)
        self.cell = np.zeros((3, 3))

    def get_positions(self): return self.positions.copy()
    def get_atomic_numbers(self): return self.numbers.copy()
    def copy(self): return Atoms(self.positions, self.numbers)
    def __len__(self): return len(self.numbers)
    def __eq__(self, o): return np.array_equal(self.numbers, o.numbers) and np.array_equal(self.positions, o.positions)
    def __ne__(self, o): return not self.__eq__(o)


class Dynamics:
    def __init__(self, atoms, logfile, trajectory, append_trajectory=False, master=None):
        self.atoms, self.logfile, self.trajectory = atoms, (sys.stdout if logfile == "-" else None), trajectory
        self.nsteps, self.max_steps = 0, 100000000


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_shims():
    _mod("ase", Atoms=Atoms)
    _mod("ase.io", write=lambda *a, **k: None)
    _mod("ase.optimize"); _mod("ase.optimize.optimize", Dynamics=Dynamics)
    _mod("ase.parallel", barrier=lambda: None, world=types.SimpleNamespace(rank=0))
    _mod("schnetpack"); _mod("schnetpack.interfaces")
    _mod("schnetpack.interfaces.ase_interface", AtomsConverter=object, AtomsConverterError=RuntimeError)

    def convert_units(a, b):
        norm = lambda u: {"ang": "angstrom"}.get(u.lower(), u.lower())
        assert norm(a) == norm(b), (a, b)
        return 1.0

    _mod("schnetpack.units", convert_units=convert_units)
    _mod("torch_geometric"); _mod("torch_geometric.data", Batch=object, Data=object)
    pkg = _mod("nablaDFT"); pkg.__path__ = []
    sub = _mod("nablaDFT.optimization"); sub.__path__ = [REF]
    for name in ("opt_utils", "line_search", "calculator", "optimizers"):
        spec = importlib.util.spec_from_file_location(f"nablaDFT.optimization.{name}", os.path.join(REF, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
    return sys.modules["nablaDFT.optimization.optimizers"], sys.modules["nablaDFT.optimization.calculator"]


def scenarios():
    """name -> dict(mols=[fixture indices], memory, maxstep, damping, alpha, fmax, steps, fixed, jitter)"""
    return {
        "basic": dict(mols=[0, 3, 7, 11], memory=100, maxstep=0.2, damping=1.0, alpha=1.0, fmax=1e-6, steps=40, fixed=None, jitter=0.15),
        "short_memory": dict(mols=[1, 2, 5], memory=4, maxstep=0.05, damping=0.8, alpha=2.0, fmax=1e-4, steps=30, fixed=None, jitter=0.08),
        "converging": dict(mols=[4, 6, 8, 9, 10], memory=20, maxstep=0.2, damping=1.0, alpha=1.0, fmax=2e-3, steps=200, fixed=None, jitter=0.12),
        "fixed_atoms": dict(mols=[12, 13], memory=10, maxstep=0.2, damping=1.0, alpha=1.0, fmax=1e-3, steps=15, fixed=[0, 5, 40, 47], jitter=0.05),
    }


def start_geometry(fix, mols, jitter, seed):
    rng = np.random.default_rng(seed)
    zs, ps = [], []
    for m in mols:
        a, b = int(fix["ptr"][m]), int(fix["ptr"][m + 1])
        zs.append(fix["z"][a:b].astype(np.int64))
        ps.append(fix["pos"][a:b].astype(np.float64) + jitter * rng.standard_normal((b - a, 3)))
    return zs, ps


def main():
    opt_mod, calc_mod = install_shims()
    fix = np.load(os.path.join(HERE, "fixture_molecules.npz"))

    class ToyCalculator(calc_mod.BatchwiseCalculator):
        def __init__(self, pot):
            super().__init__(torch.nn.Identity(), device="cpu", energy_unit="Hartree", position_unit="Ang")
            self.pot, self.n_calls = pot, 0

        def calculate(self, atoms):
            e, f = self.pot.numpy(np.concatenate([a.get_positions() for a in atoms]))
            self.n_calls += 1
            self.results = {"energy": e * self.property_units["energy"], "forces": f * self.property_units["forces"]}
            self.atoms = [a.copy() for a in atoms]

    out = {}
    for si, (name, sc) in enumerate(scenarios().items()):
        zs, ps = start_geometry(fix, sc["mols"], sc["jitter"], seed=100 + si)
        pot = ToyPotential(zs, [fix["pos"][int(fix["ptr"][m]):int(fix["ptr"][m + 1])] for m in sc["mols"]])
        calc = ToyCalculator(pot)
        opt = opt_mod.ASEBatchwiseLBFGS(calc, logfile=None, maxstep=sc["maxstep"], memory=sc["memory"], damping=sc["damping"],
                                        alpha=sc["alpha"], fixed_atoms_mask=sc["fixed"])
        atoms = [Atoms(p, z) for p, z in zip(ps, zs)]
        traj = [np.concatenate(ps)]
        # drive the reference's own run(); positions are recorded after every step() through a wrapper
        orig_step = opt.step

        def step_and_record(f=None, _o=orig_step):
            _o(f)
            traj.append(np.concatenate([a.get_positions() for a in opt.atoms]))

        opt.step = step_and_record
        conv = opt.run(atoms, fmax=sc["fmax"], steps=sc["steps"])
        e, f = pot.numpy(traj[-1])
        out[f"{name}/traj"] = np.stack(traj)
        out[f"{name}/converged"] = np.array(bool(conv))
        out[f"{name}/nsteps"] = np.array(opt.nsteps)
        out[f"{name}/n_normalizations"] = np.array(opt.n_normalizations)
        out[f"{name}/final_energy"] = e
        out[f"{name}/final_forces"] = calc.results["forces"]
        print(name, "steps", opt.nsteps, "converged", conv, "normalizations", opt.n_normalizations, "fmax", float(np.sqrt((calc.results['forces'] ** 2).sum(1).max())),
              "E0->E", pot.numpy(traj[0])[0].sum(), e.sum())
    np.savez_compressed(os.path.join(HERE, "lbfgs_ref.npz"), **out)


if __name__ == "__main__":
    main()
