"""What does the REFERENCE do with `use_line_search=True`?  (missing #7: optimizers.py:511, line_search.py)

Runs the reference's own ASEBatchwiseLBFGS (imported where it lies, third-party names shimmed exactly as for the L-BFGS goldens,
tests/golden/make_golden_lbfgs.py) on the four golden scenarios with the line search switched on.  Result in this container
(numpy 2.3; recorded in DESIGN.md section 3.6):

    basic          25 steps, no crash, 54 model calls
    short_memory   25 steps, no crash, 58 model calls
    converging     TypeError: 'bool' object is not subscriptable   (line_search.py:81, after line_search.py:287 replaced the per-molecule
    fixed_atoms    TypeError: 'bool' object is not subscriptable    `no_update` list by the scalar True)
    + RuntimeWarning "invalid value encountered in divide" at line_search.py:73: the direction of an already converged molecule is 0 and
      is divided by |p| = 0, so its positions become NaN for the rest of the run.

The reference's docstring says "use_line_search: Not implemented yet" (optimizers.py:360-361) and every shipped config sets it False
(config/optimizer/*.yaml:3).  There is no well-defined reference behaviour to be identical to, so `nabladft_b200.optimization.ASEBatchwiseLBFGS`
keeps raising NotImplementedError for it.

    python tools/probe_reference_line_search.py        # needs /root/reference (this container only)
"""
import os
import sys
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, ROOT)
import make_golden_lbfgs as mg  # noqa: E402


def main():
    opt_mod, calc_mod = mg.install_shims()
    from toy_potential import ToyPotential

    fix = np.load(os.path.join(ROOT, "tests", "golden", "fixture_molecules.npz"))

    class ToyCalculator(calc_mod.BatchwiseCalculator):
        def __init__(self, pot):
            super().__init__(torch.nn.Identity(), device="cpu", energy_unit="Hartree", position_unit="Ang")
            self.pot, self.n_calls = pot, 0

        def calculate(self, atoms):
            e, f = self.pot.numpy(np.concatenate([a.get_positions() for a in atoms]))
            self.n_calls += 1
            self.results = {"energy": e * self.property_units["energy"], "forces": f * self.property_units["forces"]}
            self.atoms = [a.copy() for a in atoms]

    for si, (name, sc) in enumerate(mg.scenarios().items()):
        zs, ps = mg.start_geometry(fix, sc["mols"], sc["jitter"], seed=100 + si)
        pot = ToyPotential(zs, [fix["pos"][int(fix["ptr"][m]):int(fix["ptr"][m + 1])] for m in sc["mols"]])
        calc = ToyCalculator(pot)
        opt = opt_mod.ASEBatchwiseLBFGS(calc, logfile=None, maxstep=sc["maxstep"], memory=sc["memory"], damping=sc["damping"], alpha=sc["alpha"],
                                        fixed_atoms_mask=sc["fixed"], use_line_search=True)
        atoms = [mg.Atoms(p, z) for p, z in zip(ps, zs)]
        try:
            conv = opt.run(atoms, fmax=sc["fmax"], steps=min(sc["steps"], 25))
            pos = np.concatenate([a.get_positions() for a in opt.atoms])
            print(f"{name}: {opt.nsteps} steps, converged {conv}, {calc.n_calls} model calls, NaN positions {bool(np.isnan(pos).any())}")
        except Exception as ex:  # noqa: BLE001
            tb = traceback.extract_tb(ex.__traceback__)[-1]
            print(f"{name}: {type(ex).__name__}: {ex}  ({os.path.basename(tb.filename)}:{tb.lineno})")


if __name__ == "__main__":
    main()
