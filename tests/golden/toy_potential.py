"""A deterministic analytic "model" used to exercise the batched L-BFGS loop without a neural network.

Per molecule: a spring network over the atom pairs closer than 2.0 A in the starting geometry, with rest lengths perturbed by a
fixed pseudo-random +-4 %, plus a soft r^-6 repulsion between all pairs.  Energies in float64, forces returned as float32 (what a
model hands the reference calculator, nablaDFT/optimization/calculator.py:125-129).  `ToyPotential.numpy` is used by the golden
script / oracle tests; `ToyPotential.torch` evaluates the same formula on a device tensor for the GPU tests.
"""
import numpy as np


class ToyPotential:
    def __init__(self, numbers_list, positions_list, k=0.6, c6=0.002):
        self.k, self.c6 = float(k), float(c6)
        self.sizes = [len(z) for z in numbers_list]
        self.ptr = np.concatenate([[0], np.cumsum(self.sizes)]).astype(np.int64)
        bi, bj, rest, ri, rj = [], [], [], [], []
        for m, pos in enumerate(positions_list):
            pos = np.asarray(pos, dtype=np.float64)
            n = len(pos)
            iu, ju = np.triu_indices(n, 1)
            d = np.linalg.norm(pos[iu] - pos[ju], axis=1)
            bond = d < 2.0
            h = np.sin(12.9898 * (iu + 1) + 78.233 * (ju + 1) + 3.7 * m)
            bi.append(iu[bond] + self.ptr[m]); bj.append(ju[bond] + self.ptr[m]); rest.append(d[bond] * (1.0 + 0.04 * h[bond]))
            ri.append(iu + self.ptr[m]); rj.append(ju + self.ptr[m])
        self.bi, self.bj, self.rest = np.concatenate(bi), np.concatenate(bj), np.concatenate(rest)
        self.ri, self.rj = np.concatenate(ri), np.concatenate(rj)
        self.pair_mol = np.searchsorted(self.ptr, self.ri, side="right") - 1
        self.bond_mol = np.searchsorted(self.ptr, self.bi, side="right") - 1
        self._torch = None

    def numpy(self, pos):
        """pos [N,3] float64 -> (energy [B] float64, forces [N,3] float32)"""
        pos = np.asarray(pos, dtype=np.float32).astype(np.float64)  # the model sees float32 positions (opt_utils.py:21)
        n_mol = len(self.sizes)
        v = pos[self.bi] - pos[self.bj]
        d = np.linalg.norm(v, axis=1)
        e = np.zeros(n_mol)
        np.add.at(e, self.bond_mol, 0.5 * self.k * (d - self.rest) ** 2)
        g = (self.k * (d - self.rest) / d)[:, None] * v
        grad = np.zeros_like(pos)
        np.add.at(grad, self.bi, g); np.add.at(grad, self.bj, -g)
        w = pos[self.ri] - pos[self.rj]
        r2 = (w * w).sum(1)
        np.add.at(e, self.pair_mol, self.c6 / r2 ** 3)
        gr = (-6.0 * self.c6 / r2 ** 4)[:, None] * w
        np.add.at(grad, self.ri, gr); np.add.at(grad, self.rj, -gr)
        return e, (-grad).astype(np.float32)

    def torch(self, pos32):
        """pos32 [N,3] float32 device tensor -> (energy [B] float64, forces [N,3] float32), same formula"""
        import torch

        dev = pos32.device
        if self._torch is None or self._torch["dev"] != dev:
            t = lambda a, dt: torch.as_tensor(a, dtype=dt, device=dev)
            self._torch = dict(dev=dev, bi=t(self.bi, torch.long), bj=t(self.bj, torch.long), rest=t(self.rest, torch.float64),
                               ri=t(self.ri, torch.long), rj=t(self.rj, torch.long), pm=t(self.pair_mol, torch.long),
                               bm=t(self.bond_mol, torch.long))
        c = self._torch
        pos = pos32.double()
        v = pos[c["bi"]] - pos[c["bj"]]
        d = v.norm(dim=1)
        e = torch.zeros(len(self.sizes), dtype=torch.float64, device=dev)
        e.index_add_(0, c["bm"], 0.5 * self.k * (d - c["rest"]) ** 2)
        g = (self.k * (d - c["rest"]) / d)[:, None] * v
        grad = torch.zeros_like(pos)
        grad.index_add_(0, c["bi"], g); grad.index_add_(0, c["bj"], -g)
        w = pos[c["ri"]] - pos[c["rj"]]
        r2 = (w * w).sum(1)
        e.index_add_(0, c["pm"], self.c6 / r2 ** 3)
        gr = (-6.0 * self.c6 / r2 ** 4)[:, None] * w
        grad.index_add_(0, c["ri"], gr); grad.index_add_(0, c["rj"], -gr)
        return e, (-grad).float()
