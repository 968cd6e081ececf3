"""The PhiSNet mixing-layer oracle (oracle/phisnet.py) against outputs of the REFERENCE'S OWN modules (tests/golden/phisnet_mixing.npz,
written by tests/golden/make_golden_phisnet.py) and against properties of the Clebsch-Gordan table it uses."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN
from weights import golden_state_dict

from oracle import phisnet as op

G = np.load(os.path.join(GOLDEN, "phisnet_mixing.npz"))


def load_named(mod, tag, dtype=torch.float64):
    """Weights are not stored in the golden file: they are rebuilt from their names (weights.py), as the generator did."""
    sd = mod.state_dict()
    new = golden_state_dict({tag + "." + k: v for k, v in sd.items()}, bias_std=0.05, weight_scale=1.0)
    for k in sd:
        sd[k] = torch.from_numpy(new[tag + "." + k]).to(dtype).reshape(sd[k].shape)
    mod.load_state_dict(sd, strict=True)
    return mod


def feats(tag, name, order):
    return [torch.from_numpy(G[f"{tag}/{name}/{L}"]) for L in range(order + 1)]


def check(ys, tag, order_out, tol=1e-11):
    for L in range(order_out + 1):
        ref = G[f"{tag}/y/{L}"]
        assert tuple(ys[L].shape) == ref.shape
        assert np.abs(ys[L].detach().numpy() - ref).max() < tol * max(1.0, np.abs(ref).max()), (tag, L)


@pytest.mark.parametrize("tag,orders", [("pair444", (4, 4, 4)), ("pair214", (2, 1, 4))])
def test_pair_mixing_matches_reference_module(tag, orders):
    o1, o2, oo = orders
    pm = load_named(op.PairMixing(o1, o2, oo, 128, 128, op.ClebschGordan()).double(), tag)
    ys = pm(feats(tag, "x1", o1), feats(tag, "x2", o2), torch.from_numpy(G[f"{tag}/rbf"]))
    check(ys, tag, oo)


@pytest.mark.parametrize("tag,orders", [("self44", (4, 4)), ("self42", (4, 2)), ("self24", (2, 4))])
def test_self_mixing_matches_reference_module(tag, orders):
    oi, oo = orders
    sm = load_named(op.SelfMixing(oi, oo, 128, op.ClebschGordan()).double(), tag)
    check(sm(feats(tag, "x", oi)), tag, oo)


@pytest.mark.parametrize("tag,cfg", [("lin44", (4, 128, 4, 128, True)), ("lin40", (4, 128, 0, 64, True)), ("lin22n", (2, 128, 2, 64, False))])
def test_spherical_linear_matches_reference_module(tag, cfg):
    oi, fi, oo, fo, mix = cfg
    sl = load_named(op.SphericalLinear(oi, fi, oo, fo, op.ClebschGordan(), mix_orders=mix, bias=True).double(), tag)
    check(sl(feats(tag, "x", oi)), tag, oo)


def test_cg_table_properties():
    """Independent of the goldens: (1,1,1) is the Levi-Civita tensor / sqrt 6, (l,0,l)-type tensors are diagonal,
    the table is symmetric under exchanging the first two indices up to the sign (-1)^(l1+l2+L) of real CG tensors, and each (l1,l2,L)
    slice is orthogonal over (m1,m2) (unitarity of the coupling)."""
    cg = op.ClebschGordan()
    eps = np.zeros((3, 3, 3))
    for a, b, c, s in ((0, 1, 2, 1), (1, 2, 0, 1), (2, 0, 1, 1), (0, 2, 1, -1), (2, 1, 0, -1), (1, 0, 2, -1)):
        eps[a, b, c] = s
    t111 = cg(1, 1, 1).numpy()
    assert np.allclose(np.abs(t111), np.abs(eps) / np.sqrt(6.0), atol=1e-12)
    for l in range(5):
        t = cg(0, l, l).numpy()[0]
        assert np.allclose(t, np.diag(np.diag(t)), atol=1e-12) and np.allclose(np.abs(np.diag(t)), np.abs(t[0, 0]), atol=1e-12)
    for l1, l2, L in op.paths(4, 4, 4):
        a, b = cg(l1, l2, L).numpy(), cg(l2, l1, L).numpy()
        assert np.allclose(a, ((-1) ** (l1 + l2 + L)) * b.transpose(1, 0, 2), atol=1e-12) or np.allclose(a, b.transpose(1, 0, 2), atol=1e-12)
        gram = np.einsum("abc,abd->cd", a, a)
        assert np.allclose(gram, gram[0, 0] * np.eye(2 * L + 1), atol=1e-12)
