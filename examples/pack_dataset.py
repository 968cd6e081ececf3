#!/usr/bin/env python
"""ASE energy database (dataset_*.db) -> packed cache directory;  --hamiltonian for Hamiltonian databases (prints a summary only: the packed
arrays are kept in memory by PackedHamiltonianDataset)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nabladft_b200.data import PackedEnergyDataset, read_hamiltonian_db  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("db")
    ap.add_argument("out", nargs="?", help="cache directory (energy databases)")
    ap.add_argument("--hamiltonian", action="store_true")
    a = ap.parse_args()
    if a.hamiltonian:
        d = read_hamiltonian_db(a.db)
        print(f"{len(d['energy'])} molecules, {len(d['z'])} atoms, {len(d['H']) * 4 / 1e6:.1f} MB of packed Hamiltonians, Norb max {int(d['norb'].max())}")
        return
    ds = PackedEnergyDataset.from_ase_db(a.db)
    ds.save(a.out)
    print(f"{len(ds)} molecules, {len(ds.z)} atoms -> {a.out}")


if __name__ == "__main__":
    main()
