"""CPU oracle for the nablaDFT model-forward hot path (TEST INFRASTRUCTURE ONLY).

Pure-PyTorch, CPU, fp32/fp64 restatement of the reference's `forward(batch) -> {E, F}`
arithmetic.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this package; the product (`nabladft_b200/`) never
does and fails loudly when its CUDA library is missing.

Parity pin status:
  * `oracle.painn_oc`  -- PINNED against the reference's own in-repo code
    (`nablaDFT/painn_pyg/painn.py`, `layers.py`) executed in the build container with
    minimal third-party shims (torch_geometric / torch_scatter / pytorch_lightning are
    absent; `tests/golden/_refshim` restates only `MessagePassing.propagate`, `scatter`,
    `segment_coo/csr`, `radius_graph`, `GaussianSmearing`).  Golden vectors:
    `tests/golden/painn_oc_*.npz`, generator `tests/golden/make_golden_painn_oc.py`.
  * `oracle.spk`       -- schnetpack==2.0.4 is an un-vendored dependency
    (`/root/reference/setup.py:32`) that is not installable here: restated from its
    published algorithm; "parity unpinned" by the reference's own tests (they assert shapes
    only, `tests/model/test_torch_models.py:31-40`).  Cross-pinned against `painn_oc`
    through the documented weight-role permutation (same mathematical layer).
  * `oracle.qhnet`, `oracle.e3` -- QHNet over a restated e3nn 0.5.1 subset; PINNED: `tests/golden/qhnet_f64.npz` comes from the reference's
    own QHNet classes (`tests/golden/make_golden_qhnet.py`; SH / Wigner-3j checked against the reference's vendored Jd.pt).
  * `oracle.lbfgs`     -- batch-wise L-BFGS; PINNED to trajectories of the reference's own `ASEBatchwiseLBFGS`
    (`tests/golden/make_golden_lbfgs.py`, ASE shimmed).
  * `oracle.gemnet_graph`, `oracle.gemnet_oc` -- GemNet-OC (graphs and index structures; the whole network); PINNED to the reference's own
    classes (`tests/golden/make_golden_gemnet_oc.py`): indices bit-exact, E / F / per-block intermediates to 2e-4 relative in float32.
    Checker of `csrc/gemnet_oc.cu` / `gemnet_oc_train.inc` (forward and parameter gradients, through the host-emulation build on the CPU and
    on the device in `tests/test_zz_gpu_first_runs.py`).
"""
