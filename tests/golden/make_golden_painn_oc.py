"""Generate golden vectors for PaiNN-OC by running the REFERENCE'S OWN CLASSES
(`/root/reference/nablaDFT/painn_pyg/{painn,layers,utils}.py`, unmodified, imported where
they lie) on the fixture molecules, in the build container.

The reference's third-party dependencies are not installable here (no network), so the
handful of primitives it imports are provided as *semantic shims* (below), each a few lines
restating the published behaviour of the pinned wheel:

    torch_scatter.scatter(reduce='sum'|'min'), segment_coo, segment_csr   (2.1.2)
    torch_geometric.nn.MessagePassing.propagate (aggr='add', flow source_to_target) (2.4.0)
    torch_geometric.nn.radius_graph == torch_cluster.radius_graph (1.6.3)  -> oracle.graph
    torch_geometric.nn.models.schnet.GaussianSmearing
    pytorch_lightning.LightningModule (only subclassed; never stepped)

Everything else -- graph post-processing (`symmetrize_edges`, `repeat_blocks`), radial
basis, message, update, readout, autograd forces -- is the reference's code.

    python tests/golden/make_golden_painn_oc.py     # writes tests/golden/painn_oc_*.npz
"""
import inspect
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
sys.path.insert(0, HERE)
from weights import golden_state_dict  # noqa: E402


# ----------------------------------------------------------------------------- shims
def _scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    dim = dim % src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    idx = index
    if idx.dim() != src.dim():
        view = [1] * src.dim()
        view[dim] = -1
        idx = idx.view(view).expand_as(src)
    if reduce in ("sum", "add"):
        return torch.zeros(shape, dtype=src.dtype).scatter_add_(dim, idx, src)
    if reduce == "min":
        big = torch.full(shape, torch.iinfo(src.dtype).max if not src.is_floating_point() else float("inf"), dtype=src.dtype)
        return big.scatter_reduce_(dim, idx, src, reduce="amin", include_self=True)
    raise NotImplementedError(reduce)


def _segment_coo(src, index, out=None, dim_size=None, reduce="sum"):
    assert reduce == "sum"
    return _scatter(src, index, dim=0, dim_size=dim_size, reduce="sum")


def _segment_csr(src, indptr, out=None, reduce="sum"):
    assert reduce == "sum"
    csum = torch.zeros(src.shape[0] + 1, dtype=src.dtype)
    csum[1:] = torch.cumsum(src, 0)
    return csum[indptr[1:]] - csum[indptr[:-1]]


class _MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2):
        super().__init__()
        assert aggr == "add" and flow == "source_to_target"
        self.node_dim = node_dim

    def jittable(self):
        return self

    def propagate(self, edge_index, size=None, **kwargs):
        params = list(inspect.signature(self.message).parameters)
        msg_kwargs, n_nodes = {}, None
        for p in params:
            if p.endswith("_j") or p.endswith("_i"):
                data = kwargs[p[:-2]]
                n_nodes = data.size(self.node_dim)
                msg_kwargs[p] = data.index_select(self.node_dim, edge_index[0 if p.endswith("_j") else 1])
            else:
                msg_kwargs[p] = kwargs[p]
        out = self.message(**msg_kwargs)
        out = self.aggregate(out, edge_index[1], None, n_nodes)
        return self.update(out)


class _GaussianSmearing(torch.nn.Module):
    def __init__(self, start=0.0, stop=5.0, num_gaussians=50):
        super().__init__()
        offset = torch.linspace(start, stop, num_gaussians)
        self.coeff = -0.5 / (offset[1] - offset[0]).item() ** 2
        self.register_buffer("offset", offset)

    def forward(self, dist):
        dist = dist.view(-1, 1) - self.offset.view(1, -1)
        return torch.exp(self.coeff * torch.pow(dist, 2))


def install_shims():
    from oracle.graph import radius_graph as _rg

    def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow="source_to_target", **kw):
        assert not loop and flow == "source_to_target"
        return _rg(x, r, batch, max_num_neighbors)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("torch_scatter", scatter=_scatter, segment_coo=_segment_coo, segment_csr=_segment_csr)
    tg = mod("torch_geometric")
    tg.nn = mod("torch_geometric.nn", MessagePassing=_MessagePassing, radius_graph=radius_graph)
    tg.nn.models = mod("torch_geometric.nn.models")
    tg.nn.models.schnet = mod("torch_geometric.nn.models.schnet", GaussianSmearing=_GaussianSmearing)

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

    mod("pytorch_lightning", LightningModule=LightningModule)
    # bypass nablaDFT/__init__.py (imports every model family); expose the package dir only
    pkg = types.ModuleType("nablaDFT")
    pkg.__path__ = [os.path.join(REF, "nablaDFT")]
    sys.modules["nablaDFT"] = pkg


class Data:
    """Duck-typed PyG Batch: the attributes `PaiNN.forward` reads (painn.py:90-104,423)."""

    def __init__(self, z, pos, batch):
        self.z, self.pos, self.batch = z, pos, batch
        self.num_nodes = z.shape[0]


def load_fixture(mol_ids):
    fx = np.load(os.path.join(HERE, "fixture_molecules.npz"))
    z, pos, b = [], [], []
    for k, m in enumerate(mol_ids):
        a, e = fx["ptr"][m], fx["ptr"][m + 1]
        z.append(fx["z"][a:e])
        pos.append(fx["pos"][a:e])
        b.append(np.full(e - a, k))
    return (
        torch.from_numpy(np.concatenate(z)).long(),
        torch.from_numpy(np.concatenate(pos)),
        torch.from_numpy(np.concatenate(b)).long(),
    )


def main():
    install_shims()
    from nablaDFT.painn_pyg.painn import PaiNN  # the reference class, unmodified

    for tag, dtype, mols in (("f64", torch.float64, [0, 1, 2, 3, 17, 42, 99]), ("f32", torch.float32, [5, 6, 7, 8])):
        torch.manual_seed(23)  # config/painn-oc.yaml:37 seed
        torch.set_default_dtype(dtype)
        net = PaiNN(
            hidden_channels=128, num_layers=6, num_rbf=100, cutoff=5.0, max_neighbors=100,
            rbf={"name": "gaussian"}, envelope={"name": "polynomial", "exponent": 5},
            regress_forces=True, direct_forces=False, use_pbc=False, otf_graph=True, num_elements=100,
        )
        # deterministic name-keyed weights (tests/golden/weights.py); the tests rebuild the same
        sd = net.state_dict()
        for k, v in golden_state_dict(sd).items():
            sd[k] = torch.from_numpy(v).to(dtype)
        net.load_state_dict(sd, strict=True)
        net.eval()
        z, pos, batch = load_fixture(mols)
        pos = pos.to(dtype)
        energy, forces = net(Data(z, pos.clone(), batch))
        out = {
            "mol_ids": np.asarray(mols),
            "z": z.numpy(), "pos": pos.numpy(), "batch": batch.numpy(),
            "energy": energy.detach().numpy(), "forces": forces.detach().numpy(),
        }
        path = os.path.join(HERE, f"painn_oc_{tag}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, "E", energy.detach().numpy()[:3], "|F|max", float(forces.abs().max()))
    torch.set_default_dtype(torch.float32)


if __name__ == "__main__":
    main()
