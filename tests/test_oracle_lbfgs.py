"""The L-BFGS oracle against trajectories produced by the reference's own optimiser class (tests/golden/make_golden_lbfgs.py)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_lbfgs import scenarios, start_geometry  # noqa: E402  (no reference import happens at module import)
from toy_potential import ToyPotential  # noqa: E402

from oracle.lbfgs import BatchLBFGS


def _setup(name):
    sc = scenarios()[name]
    fix = np.load(os.path.join(HERE, "golden", "fixture_molecules.npz"))
    si = list(scenarios()).index(name)
    zs, ps = start_geometry(fix, sc["mols"], sc["jitter"], seed=100 + si)
    pot = ToyPotential(zs, [fix["pos"][int(fix["ptr"][m]):int(fix["ptr"][m + 1])] for m in sc["mols"]])
    return sc, zs, ps, pot


@pytest.mark.parametrize("name", list(scenarios()))
def test_oracle_lbfgs_matches_reference_trajectory(name):
    gold = np.load(os.path.join(HERE, "golden", "lbfgs_ref.npz"))
    sc, zs, ps, pot = _setup(name)
    opt = BatchLBFGS(pot.numpy, [len(z) for z in zs], memory=sc["memory"], maxstep=sc["maxstep"], damping=sc["damping"], alpha=sc["alpha"],
                     fixed_atoms_mask=sc["fixed"])
    pos, conv, traj = opt.run(np.concatenate(ps), fmax=sc["fmax"], steps=sc["steps"])
    ref = gold[f"{name}/traj"]
    assert traj.shape == ref.shape
    assert opt.nsteps == int(gold[f"{name}/nsteps"]) and conv == bool(gold[f"{name}/converged"])
    assert opt.n_normalizations == int(gold[f"{name}/n_normalizations"])
    # same arithmetic, same dtypes: bit-level agreement is expected; allow 1e-12 A for BLAS dot ordering
    assert np.abs(traj - ref).max() < 1e-12
    assert np.abs(opt.final_forces - gold[f"{name}/final_forces"]).max() == 0.0


def test_toy_potential_forces_are_gradients():
    sc, zs, ps, pot = _setup("basic")
    pos = np.concatenate(ps).astype(np.float32).astype(np.float64)
    e0, f = pot.numpy(pos)
    rng = np.random.default_rng(0)
    d = rng.standard_normal(pos.shape)
    h = 1e-3  # float32 positions inside the potential: central difference at a float32-representable step
    ep = pot.numpy((pos + h * d).astype(np.float32))[0].sum()
    em = pot.numpy((pos - h * d).astype(np.float32))[0].sum()
    dd = ((pos + h * d).astype(np.float32).astype(np.float64) - (pos - h * d).astype(np.float32).astype(np.float64))
    assert abs((ep - em) + (f.astype(np.float64) * dd).sum()) < 5e-5
