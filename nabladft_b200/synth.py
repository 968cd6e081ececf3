"""Seeded synthetic drug-like conformations (SURVEY.md section 8d "Synthetic inputs").

Per molecule: n_heavy ~ U{heavy_min..heavy_max}; elements with MOSES-like frequencies; a
self-avoiding random tree of heavy atoms (bond 1.35-1.55 A, non-bonded >= 2.2 A), then H atoms
(bond 1.0-1.1 A, >= 1.6 A from everything else) until H:heavy ~ 0.9.  Host-side numpy; it is
input generation, not part of the timed path.
"""
import numpy as np

ELEMENTS = np.array([6, 7, 8, 9, 16, 17, 35])
FREQ = np.array([0.72, 0.12, 0.11, 0.015, 0.015, 0.015, 0.005])
MAX_VALENCE = {6: 4, 7: 3, 8: 2, 9: 1, 16: 2, 17: 1, 35: 1}


def _rand_dir(rng):
    v = rng.standard_normal(3)
    return v / np.linalg.norm(v)


def synth_molecule(rng, heavy_min=10, heavy_max=30, h_ratio=0.9):
    n_heavy = int(rng.integers(heavy_min, heavy_max + 1))
    z = [int(rng.choice(ELEMENTS, p=FREQ / FREQ.sum()))]
    if MAX_VALENCE[z[0]] < 2:
        z[0] = 6
    pos = [np.zeros(3)]
    val = [0]
    tries = 0
    while len(z) < n_heavy and tries < 20000:
        tries += 1
        parent = int(rng.integers(0, len(z)))
        if val[parent] >= min(MAX_VALENCE[z[parent]], 3):
            continue
        cand = pos[parent] + _rand_dir(rng) * rng.uniform(1.35, 1.55)
        d = np.linalg.norm(np.asarray(pos) - cand, axis=1)
        d[parent] = np.inf
        if d.min() < 2.2:
            continue
        zn = int(rng.choice(ELEMENTS, p=FREQ / FREQ.sum()))
        z.append(zn); pos.append(cand); val.append(1); val[parent] += 1
    n_h_target = int(round(h_ratio * len(z)))
    n_h, tries = 0, 0
    heavy_count = len(z)
    while n_h < n_h_target and tries < 20000:
        tries += 1
        parent = int(rng.integers(0, heavy_count))
        if val[parent] >= MAX_VALENCE[z[parent]]:
            continue
        cand = pos[parent] + _rand_dir(rng) * rng.uniform(1.0, 1.1)
        d = np.linalg.norm(np.asarray(pos) - cand, axis=1)
        d[parent] = np.inf
        if d.min() < 1.6:
            continue
        z.append(1); pos.append(cand); val.append(1); val[parent] += 1
        n_h += 1
    return np.asarray(z, dtype=np.int32), np.asarray(pos, dtype=np.float32)


def synth_batch(seed: int, n_mol: int, heavy_min=10, heavy_max=30):
    """Returns dict(z int32[N], pos float32[N,3], mol_ptr int32[B+1], batch int64[N])."""
    rng = np.random.default_rng(seed)
    zs, ps, ptr = [], [], [0]
    for _ in range(n_mol):
        z, p = synth_molecule(rng, heavy_min, heavy_max)
        # random rigid motion so nothing is axis-aligned
        q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        p = (p @ q.T + rng.uniform(-1, 1, 3)).astype(np.float32)
        zs.append(z); ps.append(p); ptr.append(ptr[-1] + len(z))
    ptr = np.asarray(ptr, dtype=np.int32)
    return dict(z=np.concatenate(zs), pos=np.concatenate(ps), mol_ptr=ptr,
                batch=np.repeat(np.arange(n_mol), np.diff(ptr)).astype(np.int64))
