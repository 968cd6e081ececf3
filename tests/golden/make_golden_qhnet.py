"""Golden vectors for QHNet: the REFERENCE'S OWN CLASSES (`/root/reference/nablaDFT/qhnet/
{qhnet,layers}.py`, unmodified) executed in the build container.  Third-party wheels are absent,
so `e3nn` is provided by `oracle.e3` (our restatement of the e3nn 0.5.1 primitives), and
torch_cluster / torch_scatter / torch_geometric.data / pytorch_lightning by the same few-line
shims as the PaiNN generator.  This pins everything QHNet-specific (graph construction, the
doubly-applied path weights of get_feasible_irrep, NormGate, Conv/Pair/Self layers, Expansion,
build_final_matrix); e3nn semantics themselves stay [3P-memory] except what is checked here
against the reference's vendored e3nn-convention Wigner-D (`equiformer_v2/Jd.pt`):
  * our spherical harmonics transform with that Wigner-D (l <= 4),
  * our wigner_3j tensors are invariant under it.

    python tests/golden/make_golden_qhnet.py     # writes tests/golden/qhnet_f64.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden_painn_oc import _scatter, load_fixture  # noqa: E402
from weights import golden_state_dict  # noqa: E402

REF = "/root/reference"
ORBITALS = {1: [0, 0, 1], 6: [0, 0, 0, 1, 1, 2], 7: [0, 0, 0, 1, 1, 2], 8: [0, 0, 0, 1, 1, 2], 9: [0, 0, 0, 1, 1, 2],
            16: [0, 0, 0, 0, 1, 1, 1, 2], 17: [0, 0, 0, 0, 1, 1, 1, 2], 35: [0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2]}
BOHR = 1.8897261


def install_shims():
    from oracle import e3
    from oracle.graph import radius_graph as _rg

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    o3 = mod("e3nn.o3", Irreps=e3.Irreps, Irrep=e3.Irrep, TensorProduct=e3.TensorProduct, Linear=e3.Linear, Norm=e3.Norm,
             ElementwiseTensorProduct=e3.ElementwiseTensorProduct, spherical_harmonics=e3.spherical_harmonics, wigner_3j=e3.wigner_3j)

    class FCN(e3.FullyConnectedNet):
        def __init__(self, hs, act):
            super().__init__(hs, act, "ssp" if "Soft" in getattr(act, "__name__", "") else "silu")

    enn = mod("e3nn.nn", FullyConnectedNet=FCN)
    mod("e3nn", o3=o3, nn=enn)
    mod("torch_scatter", scatter=_scatter)
    mod("torch_cluster", radius_graph=lambda x, r, batch=None, max_num_neighbors=32, **kw: _rg(x, r, batch, max_num_neighbors))
    tg = mod("torch_geometric")
    tg.data = mod("torch_geometric.data", Data=object)

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

    mod("pytorch_lightning", LightningModule=LightningModule)
    pkg = types.ModuleType("nablaDFT")
    pkg.__path__ = [os.path.join(REF, "nablaDFT")]
    sys.modules["nablaDFT"] = pkg
    # nablaDFT/qhnet/__init__.py imports loss / metric modules needing torchmetrics: expose the subpackage dir only
    sub = types.ModuleType("nablaDFT.qhnet")
    sub.__path__ = [os.path.join(REF, "nablaDFT", "qhnet")]
    sys.modules["nablaDFT.qhnet"] = sub


def check_e3_against_reference_wigner():
    """oracle.e3 SH / w3j vs the e3nn-convention Wigner-D vendored by the reference (Jd.pt)."""
    import math

    from oracle import e3

    Jd = torch.load(os.path.join(REF, "nablaDFT", "equiformer_v2", "Jd.pt"))

    def zrot(angle, l):
        M = torch.zeros(2 * l + 1, 2 * l + 1, dtype=torch.float64)
        inds, rinds = torch.arange(2 * l + 1), torch.arange(2 * l, -1, -1)
        fr = torch.arange(l, -l - 1, -1, dtype=torch.float64)
        M[inds, rinds] = torch.sin(fr * angle)
        M[inds, inds] = torch.cos(fr * angle)
        return M

    def D(l, a, b, c):
        J = Jd[l].double()
        return zrot(a, l) @ J @ zrot(b, l) @ J @ zrot(c, l)

    a, b, c = 0.3, 1.1, -0.7
    Ry = lambda t: torch.tensor([[math.cos(t), 0, math.sin(t)], [0, 1, 0], [-math.sin(t), 0, math.cos(t)]], dtype=torch.float64)
    Rx = lambda t: torch.tensor([[1, 0, 0], [0, math.cos(t), -math.sin(t)], [0, math.sin(t), math.cos(t)]], dtype=torch.float64)
    R = Ry(a) @ Rx(b) @ Ry(c)
    p = torch.randn(50, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    Y, YR = e3.spherical_harmonics(4, p), e3.spherical_harmonics(4, p @ R.T)
    for l in range(5):
        err = (YR[:, l * l:(l + 1) ** 2] - Y[:, l * l:(l + 1) ** 2] @ D(l, a, b, c).T).abs().max().item()
        assert err < 1e-12, (l, err)
    for l1, l2, l3 in [(1, 1, 1), (1, 1, 2), (2, 2, 2), (1, 2, 3), (4, 4, 2), (3, 4, 1), (2, 2, 4), (4, 4, 4), (0, 3, 3)]:
        C = e3.wigner_3j(l1, l2, l3)
        C2 = torch.einsum("ia,jb,kc,abc->ijk", D(l1, a, b, c), D(l2, a, b, c), D(l3, a, b, c), C)
        assert (C - C2).abs().max() < 1e-12, (l1, l2, l3)
    print("oracle.e3 SH and wigner_3j agree with the reference's vendored e3nn Wigner-D (Jd.pt)")


class Data:
    def __init__(self, z, pos, batch):
        self.z, self.pos, self.batch = z, pos, batch
        self.num_nodes = z.shape[0]
        counts = torch.bincount(batch)
        self.ptr = torch.zeros(counts.numel() + 1, dtype=torch.long)
        self.ptr[1:] = torch.cumsum(counts, 0)


def main():
    install_shims()
    check_e3_against_reference_wigner()
    import importlib

    qh = importlib.import_module("nablaDFT.qhnet.qhnet")
    torch.set_default_dtype(torch.float64)
    net = qh.QHNet(sh_lmax=4, hidden_size=128, bottle_hidden_size=32, num_gnn_layers=5, max_radius=12, num_nodes=83,
                   radius_embed_dim=32, orbitals=ORBITALS)
    sd = net.state_dict()
    for k, v in golden_state_dict(sd, style="e3").items():
        sd[k] = torch.from_numpy(v)
    net.load_state_dict(sd, strict=True)
    net.eval()
    from oracle.qhnet import QHNetOracle

    ora = QHNetOracle(orbitals=ORBITALS)
    assert set(ora.state_dict().keys()) == set(sd.keys()), set(ora.state_dict().keys()) ^ set(sd.keys())
    out = {}
    for tag, mols in (("a", [64]), ("b", [64, 3])):
        z, pos, batch = load_fixture(mols)
        pos = pos * BOHR  # the Hamiltonian DBs are in bohr
        with torch.no_grad():
            H = net(Data(z, pos.clone(), batch))
        out.update({f"{tag}.mols": np.asarray(mols), f"{tag}.z": z.numpy(), f"{tag}.pos": pos.numpy(), f"{tag}.batch": batch.numpy()})
        if tag == "a":
            out["a.H"] = H.numpy()
        out[f"{tag}.H_rowsum"] = H.sum(1).numpy()
        out[f"{tag}.H_fro"] = np.asarray(float(H.norm()))
        out[f"{tag}.H_diag"] = torch.diagonal(H).numpy()
        print(tag, "H", tuple(H.shape), "fro", float(H.norm()), "max", float(H.abs().max()), "sym err", float((H - H.T).abs().max()))
    np.savez_compressed(os.path.join(HERE, "qhnet_f64.npz"), **out)
    torch.set_default_dtype(torch.float32)


if __name__ == "__main__":
    main()
