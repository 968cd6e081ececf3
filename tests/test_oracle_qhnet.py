"""CPU tests: e3nn-primitive restatement (oracle/e3.py) conventions + the QHNet oracle against golden
outputs of the reference's own QHNet classes (tests/golden/make_golden_qhnet.py)."""
import math
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, load_golden_weights, random_rotation
from oracle import e3
from oracle.qhnet import QHNetOracle

ORBITALS = {1: [0, 0, 1], 6: [0, 0, 0, 1, 1, 2], 7: [0, 0, 0, 1, 1, 2], 8: [0, 0, 0, 1, 1, 2], 9: [0, 0, 0, 1, 1, 2],
            16: [0, 0, 0, 0, 1, 1, 1, 2], 17: [0, 0, 0, 0, 1, 1, 1, 2], 35: [0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2]}


def test_e3_conventions():
    torch.set_default_dtype(torch.float64)
    try:
        eps = torch.zeros(3, 3, 3)
        for i, j, k, s in [(0, 1, 2, 1), (1, 2, 0, 1), (2, 0, 1, 1), (0, 2, 1, -1), (2, 1, 0, -1), (1, 0, 2, -1)]:
            eps[i, j, k] = s
        assert (e3.wigner_3j(1, 1, 1) - eps / math.sqrt(6)).abs().max() < 1e-14  # e3nn: w3j(1,1,1) = +eps/sqrt6
        for l in range(5):
            assert (e3.wigner_3j(0, l, l)[0] - torch.eye(2 * l + 1) / math.sqrt(2 * l + 1)).abs().max() < 1e-14
        g = torch.Generator().manual_seed(0)
        v = torch.randn(9, 3, generator=g)
        Y = e3.spherical_harmonics(4, v)
        u = v / v.norm(dim=-1, keepdim=True)
        assert (Y[:, 0] - 1).abs().max() < 1e-14 and (Y[:, 1:4] - math.sqrt(3) * u).abs().max() < 1e-14  # Y1 = sqrt3 (x,y,z)
        for l in range(5):
            assert (Y[:, l * l:(l + 1) ** 2].pow(2).sum(-1) - (2 * l + 1)).abs().max() < 1e-12  # component normalisation
        for l in range(4):  # e3nn builds Y_{l+1} from Y_l (x) Y_1 with a positive coefficient
            r = torch.einsum("ijk,zi,zj->zk", e3.wigner_3j(l, 1, l + 1), Y[:, l * l:(l + 1) ** 2], Y[:, 1:4])
            t = Y[:, (l + 1) ** 2:(l + 2) ** 2]
            c = (r * t).sum(-1) / (t * t).sum(-1)
            assert (c > 0).all() and (r - c[:, None] * t).abs().max() < 1e-12
        # equivariance: w3j invariant under the Wigner-D induced by the SH
        R = random_rotation(3)
        p = torch.randn(300, 3, generator=g)
        Yp, YR = e3.spherical_harmonics(4, p), e3.spherical_harmonics(4, p @ R.T)
        D = [torch.linalg.lstsq(Yp[:, l * l:(l + 1) ** 2], YR[:, l * l:(l + 1) ** 2]).solution.T for l in range(5)]
        for l1, l2, l3 in [(1, 1, 2), (2, 2, 2), (1, 2, 3), (4, 4, 2), (3, 4, 1), (4, 4, 4)]:
            C = e3.wigner_3j(l1, l2, l3)
            assert (torch.einsum("ia,jb,kc,abc->ijk", D[l1], D[l2], D[l3], C) - C).abs().max() < 1e-10
    finally:
        torch.set_default_dtype(torch.float32)


def test_tensor_product_and_linear_normalisation():
    torch.set_default_dtype(torch.float64)
    try:
        irr = e3.Irreps("4x0e+4x1o+4x2e")
        sh = e3.Irreps.spherical_harmonics(2)
        from oracle.qhnet import feasible_irrep
        mid, ins = feasible_irrep(irr, sh, irr, "uvu")
        assert str(mid) == "4x0e+4x1o+4x2e" and len(ins) == 11  # 1o x 1o -> 1e is pruned: 15 - 4 parity-forbidden paths
        tp = e3.TensorProduct(irr, sh, mid, ins, shared_weights=False, internal_weights=False)
        assert tp.weight_numel == 11 * 4
        lin = e3.Linear(irr, irr)
        assert lin.weight.numel() == 3 * 16 and lin.bias.numel() == 4
        x = torch.randn(5, irr.dim)
        y = lin(x)
        W = lin.weight[:16].reshape(4, 4)
        assert torch.allclose(y[:, :4], x[:, :4] @ W / 2.0 + lin.bias)  # 1/sqrt(fan_in = 4)
    finally:
        torch.set_default_dtype(torch.float32)


@pytest.fixture(scope="module")
def oracle_f64():
    torch.set_default_dtype(torch.float64)
    try:
        net = load_golden_weights(QHNetOracle(orbitals=ORBITALS), torch.float64, style="e3")
    finally:
        torch.set_default_dtype(torch.float32)
    return net.eval()


@pytest.mark.timeout(600)
def test_qhnet_oracle_matches_reference_golden(oracle_f64):
    g = np.load(os.path.join(GOLDEN, "qhnet_f64.npz"))
    with torch.no_grad():
        H = oracle_f64(torch.from_numpy(g["a.z"]), torch.from_numpy(g["a.pos"]), torch.from_numpy(g["a.batch"]))
    assert H.shape == g["a.H"].shape
    assert np.abs(H.numpy() - g["a.H"]).max() < 1e-10
    assert float((H - H.T).abs().max()) == 0.0
    with torch.no_grad():
        Hb = oracle_f64(torch.from_numpy(g["b.z"]), torch.from_numpy(g["b.pos"]), torch.from_numpy(g["b.batch"]))
    assert np.abs(Hb.sum(1).numpy() - g["b.H_rowsum"]).max() < 1e-9 and abs(float(Hb.norm()) - float(g["b.H_fro"])) < 1e-9
    n = H.shape[0]
    assert float(Hb[:n, n:].abs().max()) == 0.0 and np.abs(Hb[:n, :n].numpy() - g["a.H"]).max() < 1e-10  # block diagonal over molecules


@pytest.mark.timeout(600)
def test_qhnet_oracle_blocks_are_equivariant(oracle_f64):
    """Rotating the molecule rotates every (l1,l2) tile of the H blocks with D^{l1} (.) D^{l2}."""
    g = np.load(os.path.join(GOLDEN, "qhnet_f64.npz"))
    z, pos, batch = torch.from_numpy(g["a.z"])[:12], torch.from_numpy(g["a.pos"])[:12], torch.from_numpy(g["a.batch"])[:12]
    R = random_rotation(7)
    with torch.no_grad():
        d0, o0, *_ = oracle_f64.blocks(z, pos, batch)
        d1, o1, *_ = oracle_f64.blocks(z, pos @ R.T, batch)
    # QHNet feeds vec[:, [1,2,0]] to the SH, so the block basis rotates with the permuted rotation
    P = torch.zeros(3, 3, dtype=torch.float64)
    P[0, 1] = P[1, 2] = P[2, 0] = 1.0
    Rp = P @ R @ P.T
    gen = torch.Generator().manual_seed(0)
    p = torch.randn(200, 3, dtype=torch.float64, generator=gen)
    Yp, YR = e3.spherical_harmonics(2, p), e3.spherical_harmonics(2, p @ Rp.T)
    D = [torch.linalg.lstsq(Yp[:, l * l:(l + 1) ** 2], YR[:, l * l:(l + 1) ** 2]).solution.T for l in range(3)]
    big = torch.block_diag(*([D[0]] * 5 + [D[1]] * 4 + [D[2]] * 3))  # 5s 4p 3d -> 32 x 32
    assert (big @ d0 @ big.T - d1).abs().max() < 1e-9
    assert (big @ o0 @ big.T - o1).abs().max() < 1e-9
