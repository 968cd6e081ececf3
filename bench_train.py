#!/usr/bin/env python
"""PaiNN training step throughput (BASELINE.json configs[2] shape: per-GPU batch of 256 synthetic conformations, data parallel):
forward (E + F) -> MSE(E) + MSE(F) loss as the reference trains (config/model/painn.yaml:30-46) -> backward through the engine
(analytic parameter gradients incl. the force-loss double backward, nabladft_b200/training.py) -> ONE flat gradient all-reduce
(NCCL) -> AdamW step.  fp32 (no bf16 storage in the CUDA path; the reference is fp32-only too).  Secondary number.  Launch like bench.py
(`python bench_train.py` or torchrun --nproc-per-node N); rank 0 prints one JSON line; timing = CUDA events, max over ranks."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--loss", choices=["ef", "e"], default="ef", help="ef: MSE(E) + MSE(F) (reference); e: energy only")
    ap.add_argument("--model", choices=["painn", "schnet"], default="painn",
                    help="schnet: config/model/schnet.yaml through csrc/schnet_train.cu (first correct path, DESIGN.md 3.10; not yet measured)")
    ap.add_argument("--storage", choices=["f32", "bf16"], default="f32",
                    help="bf16: per-edge arrays (filter rows, dW/dd, per-edge filter gradients) stored as bf16, fp32 arithmetic (BASELINE configs[2])")
    ap.add_argument("--epoch-molecules", type=int, default=0,
                    help="instead of cycling 4 resident batches: one shuffled epoch over a synthetic packed dataset of this many conformations per rank "
                         "through nabladft_b200.data.DeviceBatcher (host gather + pinned H2D inside the timed region)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from bench import build_model
    from nabladft_b200.parallel import allreduce_gradients, max_over_ranks
    from nabladft_b200.synth import synth_batch

    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = build_model(args.model, dev).train()
    if args.model == "painn":
        model.train_edge_storage = args.storage
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5)
    pool = []
    for k in range(4):
        b = synth_batch(1 + rank * 4 + k, args.batch)
        n_atoms = torch.from_numpy(b["mol_ptr"][1:] - b["mol_ptr"][:-1]).to(dev)
        inputs = {"_atomic_numbers": torch.from_numpy(b["z"]).to(dev), "_positions": torch.from_numpy(b["pos"]).to(dev),
                  "_idx_m": torch.from_numpy(b["batch"]).to(dev), "_n_atoms": n_atoms}
        pool.append((inputs, torch.randn(args.batch, device=dev), 0.1 * torch.randn(b["pos"].shape[0], 3, device=dev)))

    def step(k):
        inputs, target, f_target = pool[k % len(pool)]
        opt.zero_grad(set_to_none=True)
        out = model(inputs)
        loss = ((out["energy"] - target) ** 2).mean()
        if args.loss == "ef":
            loss = loss + ((out["forces"] - f_target) ** 2).mean()
        loss.backward()
        n = allreduce_gradients(model.parameters())
        opt.step()
        return n

    if args.epoch_molecules:
        import numpy as np
        import time

        from nabladft_b200.data import DeviceBatcher, PackedEnergyDataset

        # 2048 distinct synthetic conformations tiled up to the requested size (targets are noise: throughput run)
        base = synth_batch(100 + rank, 2048)
        reps = max(1, (args.epoch_molecules + 2047) // 2048)
        n_at = np.diff(base["mol_ptr"])
        rng = np.random.default_rng(rank)
        ds = PackedEnergyDataset(np.tile(base["z"], reps), np.tile(base["pos"], (reps, 1)), (0.1 * rng.standard_normal((len(base["z"]) * reps, 3))).astype(np.float32),
                                 rng.standard_normal(2048 * reps).astype(np.float32), np.concatenate([[0], np.cumsum(np.tile(n_at, reps))]).astype(np.int64))
        loader = DeviceBatcher(ds, args.batch, device=dev, shuffle=True, seed=1, drop_last=True)

        def epoch():
            n = 0
            for b in loader:
                inputs = b.as_spk()
                opt.zero_grad(set_to_none=True)
                out = model(inputs)
                loss = ((out["energy"] - b.energy) ** 2).mean()
                if args.loss == "ef":
                    loss = loss + ((out["forces"] - b.forces) ** 2).mean()
                loss.backward()
                allreduce_gradients(model.parameters())
                opt.step()
                n += b.n_mol
            return n

        for k in range(3):
            step(k)  # warm-up of allocations / handles on the resident pool
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        n_done = epoch()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0, dev)
        if rank == 0:
            print(json.dumps({"metric": "molecules/sec (" + args.model + " training epoch incl. data path, " + ("MSE(E)+MSE(F)" if args.loss == "ef" else "MSE(E)") + ")",
                              "value": world * n_done / dt, "n_gpus": world, "molecules_per_rank": n_done, "seconds": dt, "batch": args.batch,
                              "timing": "host wall clock around the epoch loop (DeviceBatcher gather + pinned H2D + step), max over ranks", "dtype": "f32",
                              "data": "synthetic (2048 distinct conformations tiled)"}))
        if world > 1:
            dist.destroy_process_group()
        return
    for k in range(args.warmup):
        n_grad = step(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        n_grad = step(k)
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps, dev)
    if rank == 0:
        print(json.dumps({"metric": "molecules/sec (" + args.model + " training step, " + ("MSE(E)+MSE(F)" if args.loss == "ef" else "MSE(E)") + " loss, AdamW, data parallel)", "value": world * args.batch / (ms / 1e3),
                          "ms_per_step": ms, "n_gpus": world, "global_batch": world * args.batch, "steps": args.steps, "warmup": args.warmup,
                          "allreduce_elements": n_grad, "dtype": "f32" if args.storage == "f32" else "bf16 edge storage / f32 arithmetic",
                          "data": "synthetic", "scaling": "weak", "loss": args.loss}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
