#!/usr/bin/env python
"""Batched geometry optimisation (reference job `job_type: optimize`, config/schnet_optim.yaml): relax one
batch of synthetic molecules with PaiNN (config/model/painn.yaml) for a fixed number of L-BFGS steps through the public API
(`nabladft_b200.optimization.ASEBatchwiseLBFGS.run`), host Atoms in -> host Atoms out.  Reports optimiser steps/s for the whole
batch and molecule-steps/s; `--cpu` times the oracle loop (oracle L-BFGS + oracle PaiNN) on a few molecules for comparison.
Secondary benchmark (the driver's headline is bench.py); prints one JSON line."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--memory", type=int, default=100)
    ap.add_argument("--check-every", type=int, default=10)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--cpu-mols", type=int, default=8)
    ap.add_argument("--cpu-steps", type=int, default=3)
    args = ap.parse_args()
    import numpy as np
    import torch

    from bench import build_model
    from nabladft_b200.optimization import ASEBatchwiseLBFGS, SimpleAtoms, SpkBatchwiseCalculator
    from nabladft_b200.synth import synth_batch

    dev = torch.device("cuda:0")
    model = build_model("painn", dev)
    b = synth_batch(1, args.batch)
    ptr = b["mol_ptr"]
    atoms = [SimpleAtoms(b["pos"][ptr[m]:ptr[m + 1]], b["z"][ptr[m]:ptr[m + 1]]) for m in range(args.batch)]
    calc = SpkBatchwiseCalculator(model, device=dev, energy_unit="Hartree", position_unit="Ang")
    opt = ASEBatchwiseLBFGS(calc, logfile=None, memory=args.memory, check_every=args.check_every)
    opt.run(atoms, fmax=1e-9, steps=5)  # warm-up (allocations, cuBLAS handles)
    torch.cuda.synchronize()
    opt.initialize()
    t0 = time.perf_counter()
    opt.run(atoms, fmax=1e-9, steps=args.steps)  # fmax unreachable: exactly `steps` E+F + L-BFGS steps, like the reference would run
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    f = calc.results["forces"]
    out = {"metric": "L-BFGS steps/sec (PaiNN E+F + batched L-BFGS, B molecules per step)", "value": opt.nsteps / dt,
           "molecule_steps_per_s": opt.nsteps * args.batch / dt, "ms_per_step": dt / opt.nsteps * 1e3, "batch": args.batch, "steps": opt.nsteps,
           "memory": args.memory, "check_every": args.check_every, "atoms": int(ptr[-1]), "n_normalizations": opt.n_normalizations,
           "timing": "host wall clock around ASEBatchwiseLBFGS.run (includes packing Atoms, H2D, final D2H)",
           "dtype": "f32 model / f64 positions", "data": "synthetic"}
    if args.cpu:
        from bench import build_oracle, oracle_pass
        from oracle.lbfgs import BatchLBFGS

        ref = build_oracle("painn", model)
        nm = args.cpu_mols
        best = None
        for nt in (8, 16, 32):
            torch.set_num_threads(min(nt, os.cpu_count()))

            def ff(pos):
                bb = dict(b)
                bb["pos"] = np.concatenate([np.asarray(pos, dtype=np.float32), b["pos"][ptr[nm]:]])
                e, fo = oracle_pass("painn", ref, bb, nm)
                return e.detach().numpy(), fo.detach().numpy()

            o = BatchLBFGS(ff, np.diff(ptr[:nm + 1]), memory=args.memory)
            t0 = time.perf_counter()
            o.run(b["pos"][:ptr[nm]].astype(np.float64), fmax=1e-9, steps=args.cpu_steps, record=False)
            d = time.perf_counter() - t0
            rate = o.nsteps * nm / d
            if best is None or rate > best[0]:
                best = (rate, nt)
        out["cpu_baseline"] = {"value": best[0], "unit": "molecule-steps/s", "cores": best[1], "kind": "port",
                               "sample": f"{args.cpu_steps} steps on the first {nm} molecules: oracle L-BFGS (reference-pinned) + oracle PaiNN, neighbour list per step"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
