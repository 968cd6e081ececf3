"""Extract the 100 real drug-like conformations of the reference's energy fixture DB
(`/root/reference/tests/data/raw/test_database.db`, ASE-sqlite v9; layout in SURVEY.md
Appendix B) into a small, travel-safe npz (`tests/golden/fixture_molecules.npz`).

Runs ONLY in the build container (it reads /root/reference); the GPU box uses the npz.
Reader semantics follow `nablaDFT/dataset/pyg_datasets.py:101-109` (numbers -> z,
positions -> float32 pos, data["energy"] -> y, data["forces"] -> float32 forces).

    python tests/golden/extract_fixture_molecules.py
"""
import json
import os
import sqlite3
import struct

import numpy as np

SRC = "/root/reference/tests/data/raw/test_database.db"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixture_molecules.npz")


def decode_data_blob(blob: bytes):
    """ASE 'bytes' container: int64 offset of trailing JSON, then raw arrays."""
    off = struct.unpack("<q", blob[:8])[0]
    meta = json.loads(blob[off:].decode())
    out = {}
    for key, val in meta.items():
        if isinstance(val, dict) and "__ndarray__" in val:
            shape, dtype, start = val["__ndarray__"]
            n = int(np.prod(shape))
            out[key] = np.frombuffer(blob, dtype=dtype, count=n, offset=start).reshape(shape)
        else:
            out[key] = np.asarray(val)
    return out


def main():
    con = sqlite3.connect(f"file:{SRC}?mode=ro", uri=True)
    rows = con.execute("select id, numbers, positions, natoms, data from systems order by id").fetchall()
    z, pos, forces, energy, ptr = [], [], [], [], [0]
    for _id, numbers, positions, natoms, data in rows:
        zz = np.frombuffer(numbers, dtype=np.int32)
        pp = np.frombuffer(positions, dtype=np.float64).reshape(-1, 3)
        d = decode_data_blob(data)
        assert len(zz) == natoms == len(pp) == len(d["forces"])
        z.append(zz)
        pos.append(pp)
        forces.append(d["forces"])
        energy.append(float(d["energy"][0]))
        ptr.append(ptr[-1] + natoms)
    np.savez_compressed(
        DST,
        z=np.concatenate(z).astype(np.int32),
        pos=np.concatenate(pos).astype(np.float64),
        forces=np.concatenate(forces).astype(np.float64),
        energy=np.asarray(energy, dtype=np.float64),
        ptr=np.asarray(ptr, dtype=np.int64),
    )
    print("wrote", DST, "molecules", len(energy), "atoms", ptr[-1])


if __name__ == "__main__":
    main()
