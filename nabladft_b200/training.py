"""Training through the CUDA engine: autograd bridge for energy losses (BASELINE configs[2]).

The reference trains by `loss.backward()` through the eager graph (painn_pyg/painn.py:642-653; schnetpack AtomisticTask via
ase_model/task.py).  Here the model's `forward` in training mode returns `energy` attached to ONE autograd node
(`PainnEnergyFn`): its backward hands dLoss/dE_m to `nb200_painn_energy_forces_grads`, which returns the gradient w.r.t. the
canonical weight tensors; autograd then carries it through the (differentiable) export permutations back to the module's
reference-named parameters, so `torch.optim.*`, Lightning's optimiser loop and DDP's gradient all-reduce work unchanged.

Built: gradients of any loss of the ENERGIES.  Not built: the force-loss term (second-order: d/dtheta of -dE/dR, `create_graph=True`
in painn.py:142) -- using `forces` in the loss raises NotImplementedError in backward instead of silently dropping the term;
`forces.detach()` gives the values.  DESIGN.md section 3.7 has the plan (forward-over-reverse tangent pass).
"""
from typing import Dict, List

import torch

from .engine import PainnEngine


class PainnEnergyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine: PainnEngine, scalars: Dict, z, pos, mol_ptr, n_mol: int, names: List[str], *canon):
        tensors = {n: t.detach().contiguous() for n, t in zip(names, canon)}
        engine._wkey = None  # weights change every optimiser step: always re-bind
        engine.set_weights(object(), tensors, scalars)
        energy, forces, _ = engine.run(z, pos, mol_ptr, n_mol)
        ctx.engine, ctx.names, ctx.n_mol = engine, names, n_mol
        ctx.tensors, ctx.scalars = tensors, scalars
        ctx.save_for_backward(z, pos, mol_ptr)
        ctx.set_materialize_grads(False)
        return energy, forces

    @staticmethod
    def backward(ctx, g_energy, g_forces):
        if g_forces is not None:
            raise NotImplementedError(
                "gradients through the forces (force-loss term, create_graph=True in the reference) are not built in the CUDA path; "
                "use forces.detach() or set the force loss coefficient to 0")
        z, pos, mol_ptr = ctx.saved_tensors
        n_fixed = 7
        if g_energy is None:
            return (None,) * (n_fixed + len(ctx.names))
        eng = ctx.engine
        eng._wkey = None
        eng.set_weights(object(), ctx.tensors, ctx.scalars)  # another forward may have re-bound the engine since
        _, _, grads = eng.run_train(z, pos, mol_ptr, ctx.n_mol, g_energy.to(torch.float32).contiguous())
        return (None,) * n_fixed + tuple(grads.get(n) for n in ctx.names)


def energy_forces_training(engine: PainnEngine, tensors: Dict[str, torch.Tensor], scalars: Dict, z, pos, mol_ptr, n_mol: int):
    names = list(tensors)
    return PainnEnergyFn.apply(engine, scalars, z, pos, mol_ptr, n_mol, names, *[tensors[n] for n in names])
