"""CPU restatement of the reference's batch-wise L-BFGS geometry optimiser.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/nablaDFT/optimization/optimizers.py:
    ASEBatchwiseLBFGS.step            :436-548   (two-loop recursion per molecule, converged molecules frozen)
    ASEBatchwiseLBFGS.determine_step  :550-571   (per-molecule rescale to `maxstep`)
    ASEBatchwiseLBFGS.update          :573-598   (history append / pop, rho = 1/(y.s) if y.s > 1e-8 else 1)
    BatchwiseOptimizer.converged      :242-247   (global max |f|^2 < fmax^2)
    BatchwiseDynamics.irun            :84-112    (loop: while not converged and nsteps < max_steps)
    BatchwiseCalculator.get_forces    calculator.py:76-89 (fixed atoms -> zero force)
including its mixed precision: positions / s / a / rho in float64; forces, y, q, z, p, dr in float32 with the float64
intermediates numpy produces (`q -= a*y` is evaluated in float64 and rounded to float32, `y*z` products are float32, ...).
Line search (`use_line_search=True`, line_search.py) is not restated: the shipped config sets it False
(config/optimizer/batchwise_lbfgs.yaml:3).

Pin status: PINNED -- tests/test_oracle_lbfgs.py compares every step of four scenarios with trajectories written by the
reference's own class (tests/golden/make_golden_lbfgs.py -> tests/golden/lbfgs_ref.npz).
"""
import numpy as np


class BatchLBFGS:
    """force_fn(pos [N,3] float64) -> (energy [B], forces [N,3] float32).  `sizes` = atoms per molecule."""

    def __init__(self, force_fn, sizes, memory=100, maxstep=0.2, damping=1.0, alpha=1.0, fixed_atoms_mask=None):
        if maxstep > 1.0:
            raise ValueError("You are using a much too large value for the maximum step size: %.1f Angstrom" % maxstep)
        self.force_fn, self.sizes = force_fn, np.asarray(sizes, dtype=np.int64)
        self.ptr = np.concatenate([[0], np.cumsum(self.sizes)])
        self.batch = np.repeat(np.arange(len(self.sizes)), self.sizes)
        self.memory, self.maxstep, self.damping, self.H0 = memory, maxstep, damping, 1.0 / alpha
        self.fixed = fixed_atoms_mask
        self.initialize()

    def initialize(self):  # optimizers.py:405-421
        self.nsteps = self.iteration = 0
        self.s, self.y, self.rho = [], [], []
        self.r0 = self.f0 = None
        self.n_normalizations = self.n_calls = 0
        self._cache = None

    def get_forces(self, pos):  # calculator.py:76-89 (cached while the geometry is unchanged)
        if self._cache is None or not np.array_equal(self._cache[0], pos):
            e, f = self.force_fn(pos)
            self.n_calls += 1
            self._cache = (pos.copy(), np.asarray(e), np.asarray(f, dtype=np.float32))
        f = self._cache[2]
        if self.fixed is not None:
            f[self.fixed] = 0.0
        return f

    def converged(self, pos):  # optimizers.py:242-247
        f = self.get_forces(pos)
        return (f ** 2).sum(axis=1).max() < self.fmax ** 2

    def _mol_sum(self, per_coord):  # np_scatter_add(...).sum(axis=1), opt_utils.py:6-9
        target = np.zeros((len(self.sizes), 3), dtype=per_coord.dtype)
        np.add.at(target, self.batch, per_coord)
        return target.sum(axis=1)

    def step(self, pos):  # optimizers.py:436-548
        f = self.get_forces(pos)
        r = pos.astype(np.float64)
        n_at = len(r)
        sq = (f.astype(np.float32) ** 2).sum(axis=-1)
        frozen = np.array([sq[self.ptr[m]:self.ptr[m + 1]].max() < self.fmax ** 2 for m in range(len(self.sizes))])
        self.update(r, f)
        loopmax = min(self.memory, self.iteration)
        a = np.empty((loopmax, n_at, 1), dtype=np.float64)
        q = -f
        for i in range(loopmax - 1, -1, -1):
            ai = self.rho[i] * self._mol_sum((self.s[i].reshape(-1, 1) * q.reshape(-1, 1)).reshape(-1, 3))
            a[i] = np.repeat(ai, self.sizes, axis=0).reshape(-1, 1)
            q -= a[i] * self.y[i]
        z = self.H0 * q
        for i in range(loopmax):
            b = self.rho[i] * self._mol_sum((self.y[i].reshape(-1, 1) * z.reshape(-1, 1)).reshape(-1, 3))
            b = np.repeat(b, self.sizes, axis=0).reshape(-1, 1)
            z += self.s[i] * (a[i] - b)
        p = -z.reshape((-1, 3))
        p = np.where(np.repeat(frozen, self.sizes)[:, None], np.zeros_like(p), p)
        dr = self.determine_step(p) * self.damping
        self.iteration += 1
        self.r0, self.f0 = r, f.copy()
        return r + dr

    def determine_step(self, dr):  # optimizers.py:550-571
        steplengths = (dr ** 2).sum(-1) ** 0.5
        if np.max(steplengths) >= self.maxstep:
            for m in range(len(self.sizes)):
                a, b = self.ptr[m], self.ptr[m + 1]
                longest = np.max(steplengths[a:b])
                if longest >= self.maxstep:
                    self.n_normalizations += 1
                    dr[a:b] *= self.maxstep / longest
        return dr

    def update(self, r, f):  # optimizers.py:573-598
        if self.iteration > 0:
            s0 = r - self.r0
            y0 = self.f0 - f
            rho0 = np.ones(len(self.sizes), dtype=np.float64)
            for m in range(len(self.sizes)):
                a, b = self.ptr[m], self.ptr[m + 1]
                ys0 = np.dot(y0[a:b].reshape(-1), s0[a:b].reshape(-1))
                if ys0 > 1e-8:
                    rho0[m] = 1.0 / ys0
            self.s.append(s0); self.y.append(y0); self.rho.append(rho0)
        if self.iteration > self.memory:
            self.s.pop(0); self.y.pop(0); self.rho.pop(0)

    def run(self, pos0, fmax=0.05, steps=None, record=True):  # optimizers.py:84-123, 216-240
        self.fmax = fmax
        max_steps = steps if steps else 100000000
        pos = np.asarray(pos0, dtype=np.float64).copy()
        traj = [pos.copy()]
        self.get_forces(pos)
        while not self.converged(pos) and self.nsteps < max_steps:
            pos = self.step(pos)
            self.nsteps += 1
            if record:
                traj.append(pos.copy())
        self.final_energy, self.final_forces = self._refresh(pos)
        return pos, bool(self.converged(pos)), (np.stack(traj) if record else None)

    def _refresh(self, pos):
        self.get_forces(pos)
        return self._cache[1], self._cache[2]
